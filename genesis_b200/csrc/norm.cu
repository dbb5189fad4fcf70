// genesis_b200 -- normalisation + gate / ReLU kernels (HBM-bound, NHWC fp32, 128-bit accesses).
//
// One family covers BatchNorm2d (train / eval), InstanceNorm2d(affine), GroupNorm and "no norm", each
// followed by either the sylvester gate  out = hn * sigmoid(gn)  (y carries 2C channels: h | g;
// reference third_party/sylvester/layers.py:42-54) or ReLU (modules/blocks.py:151-165).
//
// forward : stats (per-sample per-channel sum / sum-of-squares, double accumulators)
//           -> finalize (mean, rstd, folded scale/shift per (n,c); BN running-stat update)
//           -> apply   (out = post(scale*y + shift))
// backward: bwd_stats (per (n,c): sum d_n, sum d_n*yhat) -> bwd_finalize (m1, m2 per (n,c); dgamma,
//           dbeta) -> bwd_apply (dy = rstd*(gamma*d_n - m1 - yhat*m2))
// where d_n is the gradient w.r.t. the normalised, affine-transformed value.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------- forward stats
// y [N, HW, C]; sums [N, C, 2] (double, pre-zeroed).  grid = (chunks, N), 256 threads.  A thread owns one channel quad and
// strides over rows; the loads of NORM_U rows are issued before any of them is used (bytes in flight per thread).
constexpr int NORM_U = 4;

__global__ void __launch_bounds__(256) norm_stats_kernel(const float* __restrict__ y, double* __restrict__ sums,
                                                         int HW, int C, int rows_per_block) {
    const int q = C >> 2;
    const int lanes = 256 / q;
    const int t = threadIdx.x;
    const int lane = t / q, quad = t - lane * q;
    const int n = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(HW, r0 + rows_per_block);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    if (lane < lanes) {
        const float* base = y + ((long)n * HW) * C + quad * 4;
        for (int r = r0 + lane; r < r1; r += lanes * NORM_U) {
            float4 v[NORM_U];
#pragma unroll
            for (int u = 0; u < NORM_U; ++u) {
                const int rr = r + u * lanes;
                v[u] = rr < r1 ? g2_ldg4(base + (long)rr * C) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < NORM_U; ++u) {
                s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
                ss[0] += v[u].x * v[u].x; ss[1] += v[u].y * v[u].y; ss[2] += v[u].z * v[u].z; ss[3] += v[u].w * v[u].w;
            }
        }
    }
    __shared__ float sm[256 * 8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { sm[t * 8 + j] = s[j]; sm[t * 8 + 4 + j] = ss[j]; }
    __syncthreads();
    // thread (quad, j) for j<8 sums over lanes
    for (int o = t; o < q * 8; o += 256) {
        const int qq = o >> 3, j = o & 7;
        double acc = 0.0;
        for (int l = 0; l < lanes; ++l) acc += (double)sm[(l * q + qq) * 8 + j];
        const int c = qq * 4 + (j & 3);
        atomicAdd(sums + ((long)n * C + c) * 2 + (j >> 2), acc);
    }
}

struct NormFinP {
    const double* sums;                 // [N, C, 2]
    const float *g0, *b0, *g1, *b1;     // affine of first / second half (g1,b1 null when not split)
    float *rm0, *rv0, *rm1, *rv1;       // BN running stats (may be null)
    float *mean, *rstd, *scale, *shift; // outputs [Ns, C]  (Ns = 1 for batch mode, N otherwise)
    int N, HW, C, half, mode, groups, training;
    float eps, momentum;
};

// Batch mode: one WARP per channel (lanes stride over the N per-sample partial sums, then shuffle-reduce);
// instance / group mode: one thread per (n, c).
__global__ void norm_finalize_kernel(const NormFinP p) {
    const bool batch = p.mode == G2_NORM_BATCH;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int idx = batch ? gid >> 5 : gid;
    const int lane = threadIdx.x & 31;
    const int Ns = batch ? 1 : p.N;
    if (idx >= Ns * p.C) return;
    const int n = idx / p.C, c = idx - n * p.C;
    const bool second = c >= p.half;
    const int ch = second ? c - p.half : c;
    const float gamma = second ? (p.g1 ? p.g1[ch] : 1.f) : (p.g0 ? p.g0[ch] : 1.f);
    const float beta = second ? (p.b1 ? p.b1[ch] : 0.f) : (p.b0 ? p.b0[ch] : 0.f);
    double mean, var;
    if (batch) {
        float* rm = second ? p.rm1 : p.rm0;
        float* rv = second ? p.rv1 : p.rv0;
        if (p.training) {
            double s = 0.0, ss = 0.0;
            for (int i = lane; i < p.N; i += 32) { s += p.sums[((long)i * p.C + c) * 2]; ss += p.sums[((long)i * p.C + c) * 2 + 1]; }
            s = g2_warp_sum_d(s); ss = g2_warp_sum_d(ss);
            const double cnt = (double)p.N * p.HW;
            mean = s / cnt; var = ss / cnt - mean * mean; if (var < 0.0) var = 0.0;
            if (rm && lane == 0) {
                rm[ch] = (1.f - p.momentum) * rm[ch] + p.momentum * (float)mean;
                rv[ch] = (1.f - p.momentum) * rv[ch] + p.momentum * (float)(var * cnt / (cnt > 1.0 ? cnt - 1.0 : 1.0));
            }
        } else { mean = rm[ch]; var = rv[ch]; }
        if (lane != 0) return;
    } else if (p.mode == G2_NORM_INSTANCE) {
        const double s = p.sums[((long)n * p.C + c) * 2], ss = p.sums[((long)n * p.C + c) * 2 + 1];
        mean = s / p.HW; var = ss / p.HW - mean * mean; if (var < 0.0) var = 0.0;
    } else {  // group
        const int cpg = p.C / p.groups, g = c / cpg;
        double s = 0.0, ss = 0.0;
        for (int j = 0; j < cpg; ++j) { s += p.sums[((long)n * p.C + g * cpg + j) * 2]; ss += p.sums[((long)n * p.C + g * cpg + j) * 2 + 1]; }
        const double cnt = (double)p.HW * cpg;
        mean = s / cnt; var = ss / cnt - mean * mean; if (var < 0.0) var = 0.0;
    }
    const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
    p.mean[idx] = (float)mean;
    p.rstd[idx] = rstd;
    p.scale[idx] = gamma * rstd;
    p.shift[idx] = beta - (float)mean * gamma * rstd;
}

// ---------------------------------------------------------------------------------- forward apply
// POST_GATE: y [N,HW,2C] -> out [N,HW,C];  POST_RELU: y [N,HW,C] -> out [N,HW,C].
// scale/shift [Ns, Cy] (null -> identity), sn = stride over n (0 for batch mode).  grid = (chunks, N): a thread owns one channel
// quad of one image, so its scale / shift live in registers; rows are strided with NORM_U loads in flight.
__device__ __forceinline__ float4 affine4(const float4 a, const float4 v, const float4 b) {
    return make_float4(fmaf(a.x, v.x, b.x), fmaf(a.y, v.y, b.y), fmaf(a.z, v.z, b.z), fmaf(a.w, v.w, b.w));
}

template <int POST>
__global__ void __launch_bounds__(256) norm_apply_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, float* __restrict__ out,
                                                         int HW, int C, int sn, int rows_per_block) {
    const int q = C >> 2;
    const int Cy = POST == G2_POST_GATE ? 2 * C : C;
    const int lanes = 256 / q;
    const int lane = threadIdx.x / q, quad = threadIdx.x - lane * q;
    if (lane >= lanes) return;
    const int n = blockIdx.y;
    const int c = quad * 4;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(HW, r0 + rows_per_block);
    float4 a = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f), ag = a, bg = b;
    if (scale) {
        a = g2_ldg4(scale + (long)n * sn + c); b = g2_ldg4(shift + (long)n * sn + c);
        if (POST == G2_POST_GATE) { ag = g2_ldg4(scale + (long)n * sn + C + c); bg = g2_ldg4(shift + (long)n * sn + C + c); }
    }
    const float* yb = y + (long)n * HW * Cy + c;
    float* ob = out + (long)n * HW * C + c;
    for (int r = r0 + lane; r < r1; r += lanes * NORM_U) {
        float4 hv[NORM_U], gv[NORM_U];
#pragma unroll
        for (int u = 0; u < NORM_U; ++u) {
            const int rr = r + u * lanes;
            if (rr < r1) {
                hv[u] = g2_ldg4(yb + (long)rr * Cy);
                if (POST == G2_POST_GATE) gv[u] = g2_ldg4(yb + (long)rr * Cy + C);
            }
        }
#pragma unroll
        for (int u = 0; u < NORM_U; ++u) {
            const int rr = r + u * lanes;
            if (rr >= r1) break;
            const float4 h = affine4(a, hv[u], b);
            float4 o;
            if (POST == G2_POST_GATE) {
                const float4 g = affine4(ag, gv[u], bg);
                o.x = h.x * g2_sigmoidf(g.x); o.y = h.y * g2_sigmoidf(g.y);
                o.z = h.z * g2_sigmoidf(g.z); o.w = h.w * g2_sigmoidf(g.w);
            } else if (POST == G2_POST_RELU) {
                o.x = fmaxf(h.x, 0.f); o.y = fmaxf(h.y, 0.f); o.z = fmaxf(h.z, 0.f); o.w = fmaxf(h.w, 0.f);
            } else {
                o = h;
            }
            *reinterpret_cast<float4*>(ob + (long)rr * C) = o;
        }
    }
}

// gradient w.r.t. the normalised+affine values for 4 channels, from the RAW conv outputs hv (features) / gv (gate) and the
// per-thread affine (a,b) / (ag,bg): dh for the feature half, dg for the gate half
template <int POST>
__device__ __forceinline__ void post_grad(const float4 hv, const float4 gv, const float4 a, const float4 b, const float4 ag,
                                          const float4 bg, const float4 d, float4& dh, float4& dg) {
    const float4 h = affine4(a, hv, b);
    if (POST == G2_POST_GATE) {
        const float s0 = g2_sigmoidf(fmaf(ag.x, gv.x, bg.x)), s1 = g2_sigmoidf(fmaf(ag.y, gv.y, bg.y));
        const float s2 = g2_sigmoidf(fmaf(ag.z, gv.z, bg.z)), s3 = g2_sigmoidf(fmaf(ag.w, gv.w, bg.w));
        dh = make_float4(d.x * s0, d.y * s1, d.z * s2, d.w * s3);
        dg = make_float4(d.x * h.x * s0 * (1.f - s0), d.y * h.y * s1 * (1.f - s1),
                         d.z * h.z * s2 * (1.f - s2), d.w * h.w * s3 * (1.f - s3));
    } else if (POST == G2_POST_RELU) {
        dh = make_float4(h.x > 0.f ? d.x : 0.f, h.y > 0.f ? d.y : 0.f, h.z > 0.f ? d.z : 0.f, h.w > 0.f ? d.w : 0.f);
        dg = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        dh = d;
        dg = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ---------------------------------------------------------------------------------- backward stats
// sums2 [N, Cy, 2] (double, pre-zeroed): (sum d_n, sum d_n * yhat) per sample and normalised channel.
template <int POST>
__global__ void __launch_bounds__(256, 2) norm_bwd_stats_kernel(const float* __restrict__ y, const float* __restrict__ dout,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             double* __restrict__ sums2, int HW, int C, int sn, int rows_per_block) {
    const int q = C >> 2;
    const int Cy = POST == G2_POST_GATE ? 2 * C : C;
    const int lanes = 256 / q;
    const int t = threadIdx.x;
    const int lane = t / q, quad = t - lane * q;
    const int n = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(HW, r0 + rows_per_block);
    const int c = quad * 4;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    if (lane < lanes) {
        const long sbase = (long)n * sn;
        const float4 mh = g2_ldg4(mean + sbase + c), rh = g2_ldg4(rstd + sbase + c);
        const float4 a = g2_ldg4(scale + sbase + c), b = g2_ldg4(shift + sbase + c);
        float4 mg = mh, rg = rh, ag = a, bg = b;
        if (POST == G2_POST_GATE) {
            mg = g2_ldg4(mean + sbase + C + c); rg = g2_ldg4(rstd + sbase + C + c);
            ag = g2_ldg4(scale + sbase + C + c); bg = g2_ldg4(shift + sbase + C + c);
        }
        const float* yb = y + (long)n * HW * Cy + c;
        const float* db = dout + (long)n * HW * C + c;
        for (int r = r0 + lane; r < r1; r += lanes * NORM_U) {
            float4 hv[NORM_U], gv[NORM_U], dv[NORM_U];
#pragma unroll
            for (int u = 0; u < NORM_U; ++u) {
                const int rr = r + u * lanes;
                if (rr < r1) {
                    hv[u] = g2_ldg4(yb + (long)rr * Cy);
                    if (POST == G2_POST_GATE) gv[u] = g2_ldg4(yb + (long)rr * Cy + C);
                    dv[u] = g2_ldg4(db + (long)rr * C);
                }
            }
#pragma unroll
            for (int u = 0; u < NORM_U; ++u) {
                if (r + u * lanes >= r1) break;
                float4 dh, dg;
                post_grad<POST>(hv[u], POST == G2_POST_GATE ? gv[u] : hv[u], a, b, ag, bg, dv[u], dh, dg);
                const float4 yh = hv[u];
                acc[0] += dh.x; acc[1] += dh.y; acc[2] += dh.z; acc[3] += dh.w;
                acc[4] += dh.x * (yh.x - mh.x) * rh.x; acc[5] += dh.y * (yh.y - mh.y) * rh.y;
                acc[6] += dh.z * (yh.z - mh.z) * rh.z; acc[7] += dh.w * (yh.w - mh.w) * rh.w;
                if (POST == G2_POST_GATE) {
                    const float4 yg = gv[u];
                    acc[8] += dg.x; acc[9] += dg.y; acc[10] += dg.z; acc[11] += dg.w;
                    acc[12] += dg.x * (yg.x - mg.x) * rg.x; acc[13] += dg.y * (yg.y - mg.y) * rg.y;
                    acc[14] += dg.z * (yg.z - mg.z) * rg.z; acc[15] += dg.w * (yg.w - mg.w) * rg.w;
                }
            }
        }
    }
    constexpr int NJ = POST == G2_POST_GATE ? 16 : 8;
    __shared__ float sm[256 * NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) sm[t * NJ + j] = acc[j];
    __syncthreads();
    for (int o = t; o < q * NJ; o += 256) {
        const int qq = o / NJ, j = o - qq * NJ;
        double a = 0.0;
        for (int l = 0; l < lanes; ++l) a += (double)sm[(l * q + qq) * NJ + j];
        const int cc = qq * 4 + (j & 3) + (j >= 8 ? C : 0);
        atomicAdd(sums2 + ((long)n * Cy + cc) * 2 + ((j >> 2) & 1), a);
    }
}

struct NormBwdFinP {
    const double* sums2;                // [N, Cy, 2]
    const float *g0, *g1;               // gamma halves (null -> 1)
    float *m1, *m2;                     // [Ns, Cy]
    float *dg0, *db0, *dg1, *db1;       // parameter grads (may be null)
    int N, HW, Cy, half, mode, groups;
    int accumulate;                     // 1: add the parameter gradients to dg* / db* (direct-gradient mode: they point into param.grad)
};

// Blocks [0, main_blocks): one thread per (n, c) -> m1, m2 (instance / group mode).
// Blocks [main_blocks, ..): one WARP per channel sums the per-sample partials over N (lanes stride, shuffle-reduce)
// -> dgamma, dbeta, and in batch mode also m1, m2 (main_blocks = 0 there).
__global__ void norm_bwd_finalize_kernel(const NormBwdFinP p, int main_blocks) {
    auto gamma_of = [&](int cc) {
        const bool sec = cc >= p.half; const int ch = sec ? cc - p.half : cc;
        return sec ? (p.g1 ? p.g1[ch] : 1.f) : (p.g0 ? p.g0[ch] : 1.f);
    };
    if ((int)blockIdx.x >= main_blocks) {
        const int c = (((int)blockIdx.x - main_blocks) * blockDim.x + threadIdx.x) >> 5;
        const int lane = threadIdx.x & 31;
        if (c >= p.Cy) return;
        double s1 = 0.0, s2 = 0.0;
        for (int i = lane; i < p.N; i += 32) { s1 += p.sums2[((long)i * p.Cy + c) * 2]; s2 += p.sums2[((long)i * p.Cy + c) * 2 + 1]; }
        s1 = g2_warp_sum_d(s1); s2 = g2_warp_sum_d(s2);
        if (lane != 0) return;
        if (p.mode == G2_NORM_BATCH) {
            const double cnt = (double)p.N * p.HW, g = gamma_of(c);
            p.m1[c] = (float)(g * s1 / cnt);
            p.m2[c] = (float)(g * s2 / cnt);
        }
        const bool sec = c >= p.half; const int ch = sec ? c - p.half : c;
        float* dg = sec ? p.dg1 : p.dg0; float* db = sec ? p.db1 : p.db0;
        if (dg) dg[ch] = p.accumulate ? dg[ch] + (float)s2 : (float)s2;
        if (db) db[ch] = p.accumulate ? db[ch] + (float)s1 : (float)s1;
        return;
    }
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.N * p.Cy) return;
    const int n = idx / p.Cy, c = idx - n * p.Cy;
    double m1, m2;
    if (p.mode == G2_NORM_INSTANCE) {
        const double g = gamma_of(c);
        m1 = g * p.sums2[((long)n * p.Cy + c) * 2] / p.HW;
        m2 = g * p.sums2[((long)n * p.Cy + c) * 2 + 1] / p.HW;
    } else {
        const int cpg = p.Cy / p.groups, g0 = (c / cpg) * cpg;
        double s1 = 0.0, s2 = 0.0;
        for (int j = 0; j < cpg; ++j) {
            const double g = gamma_of(g0 + j);
            s1 += g * p.sums2[((long)n * p.Cy + g0 + j) * 2];
            s2 += g * p.sums2[((long)n * p.Cy + g0 + j) * 2 + 1];
        }
        const double cnt = (double)p.HW * cpg;
        m1 = s1 / cnt; m2 = s2 / cnt;
    }
    p.m1[idx] = (float)m1;
    p.m2[idx] = (float)m2;
}

// ---------------------------------------------------------------------------------- backward apply
// dy [N,HW,Cy] = rstd * (gamma * d_n - m1 - yhat * m2)   (mode NONE: dy = d_n).  grid = (chunks, N); per-thread constants in
// registers; NORM_U rows in flight.  dbias (optional, [Cy]): the gradient of the bias of the convolution that produced y,
// dbias[c] += sum over rows of dy -- fused here so that no separate column-sum pass re-reads dy (float atomics, one per
// channel and CTA).
template <int POST>
__global__ void __launch_bounds__(256, 2) norm_bwd_apply_kernel(const float* __restrict__ y, const float* __restrict__ dout,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const float* __restrict__ m1, const float* __restrict__ m2,
                                                             float* __restrict__ dy, float* __restrict__ dbias, int HW, int C, int sn,
                                                             int rows_per_block) {
    constexpr int NH = POST == G2_POST_GATE ? 2 : 1;
    const int q = C >> 2;
    const int Cy = NH * C;
    const int lanes = 256 / q;
    const int t = threadIdx.x;
    const int lane = t / q, quad = t - lane * q;
    const int n = blockIdx.y;
    const int c = quad * 4;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(HW, r0 + rows_per_block);
    float bsum[4 * NH];
#pragma unroll
    for (int j = 0; j < 4 * NH; ++j) bsum[j] = 0.f;
    if (lane < lanes) {
        const long sbase = (long)n * sn;
        const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
        // dy = gamma rstd d_n - rstd (m1 + (y - mean) rstd m2) = a d_n + k0 + k1 (y - mean)   with  k0 = -rstd m1,  k1 = -rstd^2 m2
        float4 a[2] = {one, one}, b[2] = {zero, zero}, k0[2] = {zero, zero}, k1[2] = {zero, zero}, mu[2] = {zero, zero};
        if (scale) {
#pragma unroll
            for (int hf = 0; hf < NH; ++hf) {
                const long o = sbase + c + hf * C;
                a[hf] = g2_ldg4(scale + o); b[hf] = g2_ldg4(shift + o);
                mu[hf] = g2_ldg4(mean + o);
                const float4 rs = g2_ldg4(rstd + o), a1 = g2_ldg4(m1 + o), a2 = g2_ldg4(m2 + o);
                k1[hf] = make_float4(-rs.x * rs.x * a2.x, -rs.y * rs.y * a2.y, -rs.z * rs.z * a2.z, -rs.w * rs.w * a2.w);
                k0[hf] = make_float4(-rs.x * a1.x, -rs.y * a1.y, -rs.z * a1.z, -rs.w * a1.w);
            }
        }
        const float* yb = y + (long)n * HW * Cy + c;
        const float* db = dout + (long)n * HW * C + c;
        float* ob = dy + (long)n * HW * Cy + c;
        for (int r = r0 + lane; r < r1; r += lanes * NORM_U) {
            float4 hv[NORM_U], gv[NORM_U], dv[NORM_U];
#pragma unroll
            for (int u = 0; u < NORM_U; ++u) {
                const int rr = r + u * lanes;
                if (rr < r1) {
                    hv[u] = g2_ldg4(yb + (long)rr * Cy);
                    if (POST == G2_POST_GATE) gv[u] = g2_ldg4(yb + (long)rr * Cy + C);
                    dv[u] = g2_ldg4(db + (long)rr * C);
                }
            }
#pragma unroll
            for (int u = 0; u < NORM_U; ++u) {
                const int rr = r + u * lanes;
                if (rr >= r1) break;
                float4 dn[2];
                post_grad<POST>(hv[u], POST == G2_POST_GATE ? gv[u] : hv[u], a[0], b[0], a[NH - 1], b[NH - 1], dv[u], dn[0], dn[1]);
#pragma unroll
                for (int hf = 0; hf < NH; ++hf) {
                    float4 o = dn[hf];
                    if (scale) {
                        const float4 yv = hf ? gv[u] : hv[u];
                        o.x = fmaf(a[hf].x, dn[hf].x, fmaf(k1[hf].x, yv.x - mu[hf].x, k0[hf].x));
                        o.y = fmaf(a[hf].y, dn[hf].y, fmaf(k1[hf].y, yv.y - mu[hf].y, k0[hf].y));
                        o.z = fmaf(a[hf].z, dn[hf].z, fmaf(k1[hf].z, yv.z - mu[hf].z, k0[hf].z));
                        o.w = fmaf(a[hf].w, dn[hf].w, fmaf(k1[hf].w, yv.w - mu[hf].w, k0[hf].w));
                    }
                    *reinterpret_cast<float4*>(ob + (long)rr * Cy + hf * C) = o;
                    bsum[4 * hf] += o.x; bsum[4 * hf + 1] += o.y; bsum[4 * hf + 2] += o.z; bsum[4 * hf + 3] += o.w;
                }
            }
        }
    }
    if (dbias == nullptr) return;          // uniform over the grid
    __shared__ float sm[256 * 4 * NH];
#pragma unroll
    for (int j = 0; j < 4 * NH; ++j) sm[t * 4 * NH + j] = bsum[j];
    __syncthreads();
    for (int o = t; o < q * 4 * NH; o += 256) {
        const int qq = o / (4 * NH), j = o - qq * (4 * NH);
        float acc = 0.f;
        for (int l = 0; l < lanes; ++l) acc += sm[(l * q + qq) * 4 * NH + j];
        atomicAdd(dbias + qq * 4 + (j & 3) + (j >= 4 ? C : 0), acc);
    }
}

// rows per block of the two statistics kernels (grid = (chunks, N)): about 12 CTAs per SM over the whole launch -- large maps get
// long row runs per CTA (the block reduction + atomics tail is paid once per CTA), small maps still fill the machine (round 1
// used >= 32 passes per CTA: 320 CTAs for a 32 x 32 map, 2.3 TB/s)
inline int stats_rows_per_block(int N, int HW, int C) {
    const int lanes = 256 / (C / 4);
    const int unit = lanes * NORM_U;
    long rpb = ((long)HW * N + 148L * 12 - 1) / (148L * 12);
    rpb = (rpb + unit - 1) / unit * unit;
    if (rpb < unit) rpb = unit;
    if (rpb > HW) rpb = (HW + unit - 1) / unit * unit;
    return (int)rpb;
}

// rows per block for the (chunks, N) streaming kernels: a few waves of 148 SMs, at least one unrolled pass per thread
inline int norm_rows_per_block(int N, int HW, int C) {
    const int lanes = 256 / (C / 4);
    int rpb = lanes * NORM_U * 2;
    while ((long)g2_cdiv(HW, rpb) * N > 148L * 24 && rpb < HW) rpb *= 2;
    return rpb;
}

}  // namespace

extern "C" {

int g2_norm_stats_f32(const float* y, double* sums, int N, int HW, int Cy, cudaStream_t stream) {
    const int C = Cy;       // channels of y as stored (2C' for a gated layer), named as in the header
    G2_CHECK_ARG(y && sums && N > 0 && HW > 0 && C >= 4 && (C % 4) == 0 && C <= 1024);
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)N * C, stream);
    if (e != cudaSuccess) return (int)e;
    const int rpb = stats_rows_per_block(N, HW, C);
    dim3 grid(g2_cdiv(HW, rpb), N);
    norm_stats_kernel<<<grid, 256, 0, stream>>>(y, sums, HW, C, rpb);
    G2_LAUNCH_RET();
}

int g2_norm_finalize_f32(const double* sums, const float* g0, const float* b0, const float* g1, const float* b1,
                         float* rm0, float* rv0, float* rm1, float* rv1, float* mean, float* rstd, float* scale,
                         float* shift, int N, int HW, int Cy, int half, int mode, int groups, int training,
                         float eps, float momentum, cudaStream_t stream) {
    const int C = Cy;
    G2_CHECK_ARG(mean && rstd && scale && shift && N > 0 && C > 0 && half > 0 && half <= C);
    G2_CHECK_ARG(mode == G2_NORM_BATCH || mode == G2_NORM_INSTANCE || mode == G2_NORM_GROUP);
    if (mode == G2_NORM_GROUP) G2_CHECK_ARG(groups > 0 && C % groups == 0);
    if (mode == G2_NORM_BATCH && !training) G2_CHECK_ARG(rm0 && rv0 && (half == C || (rm1 && rv1)));
    else G2_CHECK_ARG(sums != nullptr);
    NormFinP p{sums, g0, b0, g1, b1, rm0, rv0, rm1, rv1, mean, rstd, scale, shift, N, HW, C, half, mode, groups, training, eps, momentum};
    const int Ns = mode == G2_NORM_BATCH ? 1 : N;
    norm_finalize_kernel<<<g2_cdiv((long)Ns * C * (mode == G2_NORM_BATCH ? 32 : 1), 128), 128, 0, stream>>>(p);
    G2_LAUNCH_RET();
}

int g2_norm_apply_f32(const float* y, const float* scale, const float* shift, float* out, int N, int HW, int C,
                      int sn, int post, cudaStream_t stream) {
    G2_CHECK_ARG(y && out && N > 0 && HW > 0 && C >= 4 && (C % 4) == 0 && C <= 1024 && (scale == nullptr) == (shift == nullptr));
    const int rpb = norm_rows_per_block(N, HW, C);
    dim3 grid(g2_cdiv(HW, rpb), N);
    if (post == G2_POST_GATE) norm_apply_kernel<G2_POST_GATE><<<grid, 256, 0, stream>>>(y, scale, shift, out, HW, C, sn, rpb);
    else if (post == G2_POST_RELU) norm_apply_kernel<G2_POST_RELU><<<grid, 256, 0, stream>>>(y, scale, shift, out, HW, C, sn, rpb);
    else if (post == G2_POST_NONE) norm_apply_kernel<G2_POST_NONE><<<grid, 256, 0, stream>>>(y, scale, shift, out, HW, C, sn, rpb);
    else return G2_ERR_ARG;
    G2_LAUNCH_RET();
}

int g2_norm_bwd_stats_f32(const float* y, const float* dout, const float* scale, const float* shift, const float* mean,
                          const float* rstd, double* sums2, int N, int HW, int C, int sn, int post, cudaStream_t stream) {
    G2_CHECK_ARG(y && dout && scale && shift && mean && rstd && sums2 && N > 0 && HW > 0 && C >= 4 && (C % 4) == 0 && C <= 1024);
    const int Cy = post == G2_POST_GATE ? 2 * C : C;
    cudaError_t e = cudaMemsetAsync(sums2, 0, sizeof(double) * 2 * (size_t)N * Cy, stream);
    if (e != cudaSuccess) return (int)e;
    const int rpb = stats_rows_per_block(N, HW, C);
    dim3 grid(g2_cdiv(HW, rpb), N);
    if (post == G2_POST_GATE) norm_bwd_stats_kernel<G2_POST_GATE><<<grid, 256, 0, stream>>>(y, dout, scale, shift, mean, rstd, sums2, HW, C, sn, rpb);
    else if (post == G2_POST_RELU) norm_bwd_stats_kernel<G2_POST_RELU><<<grid, 256, 0, stream>>>(y, dout, scale, shift, mean, rstd, sums2, HW, C, sn, rpb);
    else if (post == G2_POST_NONE) norm_bwd_stats_kernel<G2_POST_NONE><<<grid, 256, 0, stream>>>(y, dout, scale, shift, mean, rstd, sums2, HW, C, sn, rpb);
    else return G2_ERR_ARG;
    G2_LAUNCH_RET();
}

static int norm_bwd_finalize_impl(const double* sums2, const float* g0, const float* g1, float* m1, float* m2, float* dg0,
                                  float* db0, float* dg1, float* db1, int N, int HW, int Cy, int half, int mode, int groups,
                                  int accumulate, cudaStream_t stream) {
    G2_CHECK_ARG(sums2 && m1 && m2 && N > 0 && Cy > 0 && half > 0 && half <= Cy);
    G2_CHECK_ARG(mode == G2_NORM_BATCH || mode == G2_NORM_INSTANCE || mode == G2_NORM_GROUP);
    if (mode == G2_NORM_GROUP) G2_CHECK_ARG(groups > 0 && Cy % groups == 0);
    NormBwdFinP p{sums2, g0, g1, m1, m2, dg0, db0, dg1, db1, N, HW, Cy, half, mode, groups, accumulate ? 1 : 0};
    const int main_blocks = mode == G2_NORM_BATCH ? 0 : g2_cdiv((long)N * Cy, 128);
    norm_bwd_finalize_kernel<<<main_blocks + g2_cdiv((long)Cy * 32, 128), 128, 0, stream>>>(p, main_blocks);
    G2_LAUNCH_RET();
}

int g2_norm_bwd_finalize_f32(const double* sums2, const float* g0, const float* g1, float* m1, float* m2, float* dg0,
                             float* db0, float* dg1, float* db1, int N, int HW, int Cy, int half, int mode, int groups,
                             cudaStream_t stream) {
    return norm_bwd_finalize_impl(sums2, g0, g1, m1, m2, dg0, db0, dg1, db1, N, HW, Cy, half, mode, groups, 0, stream);
}

// As above with `accumulate`: the parameter gradients are ADDED to dg* / db* (which then point into param.grad), one thread per
// element, so the caller orders it after other writers of those tensors on the stream.
int g2_norm_bwd_finalize_acc_f32(const double* sums2, const float* g0, const float* g1, float* m1, float* m2, float* dg0,
                                 float* db0, float* dg1, float* db1, int N, int HW, int Cy, int half, int mode, int groups,
                                 int accumulate, cudaStream_t stream) {
    return norm_bwd_finalize_impl(sums2, g0, g1, m1, m2, dg0, db0, dg1, db1, N, HW, Cy, half, mode, groups, accumulate, stream);
}

static int norm_bwd_apply_impl(const float* y, const float* dout, const float* scale, const float* shift, const float* mean,
                               const float* rstd, const float* m1, const float* m2, float* dy, float* dbias, int N, int HW, int C,
                               int sn, int post, cudaStream_t stream) {
    G2_CHECK_ARG(y && dout && dy && N > 0 && HW > 0 && C >= 4 && (C % 4) == 0 && C <= 1024);
    if (scale) G2_CHECK_ARG(shift && mean && rstd && m1 && m2);
    const int rpb = norm_rows_per_block(N, HW, C);
    dim3 grid(g2_cdiv(HW, rpb), N);
    if (post == G2_POST_GATE) norm_bwd_apply_kernel<G2_POST_GATE><<<grid, 256, 0, stream>>>(y, dout, scale, shift, mean, rstd, m1, m2, dy, dbias, HW, C, sn, rpb);
    else if (post == G2_POST_RELU) norm_bwd_apply_kernel<G2_POST_RELU><<<grid, 256, 0, stream>>>(y, dout, scale, shift, mean, rstd, m1, m2, dy, dbias, HW, C, sn, rpb);
    else if (post == G2_POST_NONE) norm_bwd_apply_kernel<G2_POST_NONE><<<grid, 256, 0, stream>>>(y, dout, scale, shift, mean, rstd, m1, m2, dy, dbias, HW, C, sn, rpb);
    else return G2_ERR_ARG;
    G2_LAUNCH_RET();
}

int g2_norm_bwd_apply_f32(const float* y, const float* dout, const float* scale, const float* shift, const float* mean,
                          const float* rstd, const float* m1, const float* m2, float* dy, int N, int HW, int C, int sn,
                          int post, cudaStream_t stream) {
    return norm_bwd_apply_impl(y, dout, scale, shift, mean, rstd, m1, m2, dy, nullptr, N, HW, C, sn, post, stream);
}

// As above, and dbias[c] += sum over all rows of dy[.., c] for the Cy channels of y: the bias gradient of the convolution that
// produced y (third_party/sylvester/layers.py:19-20,65-67: gated convs carry a bias in front of their norm), accumulated into
// a caller-initialised buffer (param.grad in direct-gradient mode, a zeroed tensor otherwise).
int g2_norm_bwd_apply_bias_f32(const float* y, const float* dout, const float* scale, const float* shift, const float* mean,
                               const float* rstd, const float* m1, const float* m2, float* dy, float* dbias, int N, int HW, int C,
                               int sn, int post, cudaStream_t stream) {
    G2_CHECK_ARG(dbias != nullptr);
    return norm_bwd_apply_impl(y, dout, scale, shift, mean, rstd, m1, m2, dy, dbias, N, HW, C, sn, post, stream);
}

}  // extern "C"
