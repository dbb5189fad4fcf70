// genesis_b200 -- fp32 SIMT implicit-GEMM family (exact-fp32 path).
//
// NHWC fp32 activations.  Three gather patterns cover every conv / conv-transpose of the reference's
// hot path (forward, data-gradient and weight-gradient) plus a plain GEMM for the Linear / LSTM / FC
// layers.  These kernels are the exact-fp32 path: they serve the layers whose shape does not fit the
// tcgen05 tiles (Cin = 3/4, Cout = 1/3/4, tiny GEMMs) and are the on-device fp32 cross-check of the
// TF32 tensor-core kernels in igemm_tc.cu.
//
//   mode 0 "correlate"  out[n,oh,ow,co] = b[co] + sum_{r,s,c} in[n, oh*S+r-P, ow*S+s-P, c] * W[r,s,c,co]
//       = nn.Conv2d forward; nn.ConvTranspose2d data-gradient
//   mode 1 "scatter-gather" out[n,h,w,co] = b[co] + sum_{r,s,c : (h+P-r)%S==0, (w+P-s)%S==0}
//                                              in[n,(h+P-r)/S,(w+P-s)/S,c] * W[r,s,c,co]
//       = nn.ConvTranspose2d forward; nn.Conv2d data-gradient
//   wgrad               dW[r,s,a,b] = sum_{n,oh,ow} G[n, oh*S+r-P, ow*S+s-P, a] * T[n,oh,ow,b]
// Weights are packed [R,S,Cred,Cout] (wT=0) or [R,S,Cout,Cred] (wT=1).
#include "common.cuh"

namespace {

struct ConvP {
    const float* in; const float* w; const float* bias; const float* aux; float* out;
    int N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, wT, act, mode;
};

constexpr int BK = 16;

template <int MODE>
__device__ __forceinline__ void decode_row(const ConvP& p, int m, int& n, int& h, int& w) {
    if (MODE == 1 && p.stride == 2) {
        const int Wh = p.Wo >> 1, Hh = p.Ho >> 1;
        const int per = p.N * Hh * Wh;
        const int cls = m / per;
        int rem = m - cls * per;
        const int wq = rem % Wh; rem /= Wh;
        const int hq = rem % Hh;
        n = rem / Hh;
        h = 2 * hq + (cls >> 1);
        w = 2 * wq + (cls & 1);
    } else {
        w = m % p.Wo; int t = m / p.Wo;
        h = t % p.Ho; n = t / p.Ho;
    }
}

// BM x BN output tile, 256 threads, each TM x TN outputs (BM = 16*TM or 32*TM, BN = TN * threads-in-n).
template <int MODE, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const ConvP p) {
    constexpr int TXN = BN / TN;            // threads along n
    constexpr int AS = BM + 4;
    static_assert(256 / TXN * TM == BM, "tile shape");
    constexpr int A_PER_THREAD = BM * BK / 4 / 256;   // float4 loads per thread
    static_assert(A_PER_THREAD >= 1, "A tile");
    __shared__ __align__(16) float As[BK][AS];
    __shared__ __align__(16) float Bs[BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const long M = (long)p.N * p.Ho * p.Wo;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const bool vecA = (p.Ci & 3) == 0;
    const bool vecB = (p.Co & 3) == 0 && !p.wT;

    // rows this thread loads for the A tile
    int a_n[A_PER_THREAD], a_h[A_PER_THREAD], a_w[A_PER_THREAD];
    bool a_ok[A_PER_THREAD];
    const int kq = tid & 3;
#pragma unroll
    for (int i = 0; i < A_PER_THREAD; ++i) {
        const int row = (tid >> 2) + i * 64;
        const long m = (long)m0 + row;
        a_ok[i] = m < M;
        int n = 0, h = 0, w = 0;
        if (a_ok[i]) decode_row<MODE>(p, (int)m, n, h, w);
        a_n[i] = n;
        if (MODE == 0) { a_h[i] = h * p.stride - p.pad; a_w[i] = w * p.stride - p.pad; }
        else { a_h[i] = h + p.pad; a_w[i] = w + p.pad; }
    }
    // CTA-uniform parity class (mode 1, stride 2): skip taps that no row of the tile can use
    int cls_h = -1, cls_w = -1;
    if (MODE == 1 && p.stride == 2) {
        const long mlast = (m0 + BM - 1 < M ? m0 + BM - 1 : M - 1);
        const int per = p.N * (p.Ho >> 1) * (p.Wo >> 1);
        const int c0 = m0 / per, c1 = (int)(mlast / per);
        if (c0 == c1) { cls_h = c0 >> 1; cls_w = c0 & 1; }
    }

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int taps = p.R * p.S;
    const int kchunks = (p.Ci + BK - 1) / BK;
    for (int tap = 0; tap < taps; ++tap) {
        const int r = tap / p.S, s = tap - r * p.S;
        if (MODE == 1 && cls_h >= 0) {
            if (((cls_h + p.pad - r) & 1) || ((cls_w + p.pad - s) & 1)) continue;
        }
        // per-row source pixel for this tap
        const float* a_ptr[A_PER_THREAD];
#pragma unroll
        for (int i = 0; i < A_PER_THREAD; ++i) {
            int ih, iw; bool ok = a_ok[i];
            if (MODE == 0) { ih = a_h[i] + r; iw = a_w[i] + s; }
            else {
                const int th = a_h[i] - r, tw = a_w[i] - s;
                if (p.stride == 2) { ok = ok && !((th | tw) & 1) && th >= 0 && tw >= 0; ih = th >> 1; iw = tw >> 1; }
                else { ih = th; iw = tw; }
            }
            ok = ok && ih >= 0 && iw >= 0 && ih < p.Hi && iw < p.Wi;
            a_ptr[i] = ok ? p.in + (((long)a_n[i] * p.Hi + ih) * p.Wi + iw) * p.Ci : nullptr;
        }
        const float* wtap = p.w + (long)tap * p.Ci * p.Co;
        for (int kc = 0; kc < kchunks; ++kc) {
            const int c0 = kc * BK;
            // ---- global -> registers
            float4 av[A_PER_THREAD];
#pragma unroll
            for (int i = 0; i < A_PER_THREAD; ++i) {
                const int c = c0 + kq * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a_ptr[i] != nullptr) {
                    if (vecA) { if (c < p.Ci) v = g2_ldg4(a_ptr[i] + c); }
                    else {
                        if (c + 0 < p.Ci) v.x = __ldg(a_ptr[i] + c + 0);
                        if (c + 1 < p.Ci) v.y = __ldg(a_ptr[i] + c + 1);
                        if (c + 2 < p.Ci) v.z = __ldg(a_ptr[i] + c + 2);
                        if (c + 3 < p.Ci) v.w = __ldg(a_ptr[i] + c + 3);
                    }
                }
                av[i] = v;
            }
            __syncthreads();   // previous compute done before overwriting smem
#pragma unroll
            for (int i = 0; i < A_PER_THREAD; ++i) {
                const int row = (tid >> 2) + i * 64;
                As[kq * 4 + 0][row] = av[i].x; As[kq * 4 + 1][row] = av[i].y;
                As[kq * 4 + 2][row] = av[i].z; As[kq * 4 + 3][row] = av[i].w;
            }
            if (vecB) {
                constexpr int NQ = BN / 4;
                for (int idx = tid; idx < BK * NQ; idx += 256) {
                    const int k = idx / NQ, nq = idx - k * NQ;
                    const int c = c0 + k, nn = n0 + nq * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c < p.Ci && nn < p.Co) v = g2_ldg4(wtap + (long)c * p.Co + nn);
                    *reinterpret_cast<float4*>(&Bs[k][nq * 4]) = v;
                }
            } else {
                for (int idx = tid; idx < BK * BN; idx += 256) {
                    int k, nn;
                    if (p.wT) { k = idx % BK; nn = idx / BK; } else { nn = idx % BN; k = idx / BN; }
                    const int c = c0 + k, co = n0 + nn;
                    float v = 0.f;
                    if (c < p.Ci && co < p.Co)
                        v = __ldg(p.wT ? wtap + (long)co * p.Ci + c : wtap + (long)c * p.Co + co);
                    Bs[k][nn] = v;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[TM], b[TN];
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
                    a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
                }
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + j]);
                    b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
    }
    // ---- epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long m = (long)m0 + ty * TM + i;
        if (m >= M) continue;
        int n, h, w;
        decode_row<MODE>(p, (int)m, n, h, w);
        const long off = (((long)n * p.Ho + h) * p.Wo + w) * p.Co;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int co = n0 + tx * TN + j;
            if (co >= p.Co) continue;
            float v = acc[i][j];
            if (p.bias) v += __ldg(p.bias + co);
            const float ax = p.aux ? __ldg(p.aux + off + co) : 0.f;
            p.out[off + co] = g2_apply_act(v, p.act, ax);
        }
    }
}

struct WgradP {
    const float* g; const float* t; float* dw;
    int N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride, pad, outT;
    long pix_per_split;
};

// dW tile 64(a) x 64(b) per CTA; K = pixels of this CTA's split; atomicAdd into dW (pre-zeroed).
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradP p) {
    constexpr int BM = 64, BN = 64;
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int tiles_b = (p.Ct + BN - 1) / BN;
    const int a0 = (blockIdx.x / tiles_b) * BM, b0 = (blockIdx.x % tiles_b) * BN;
    const int tap = blockIdx.y;
    const int r = tap / p.S, s = tap - r * p.S;
    const long npix = (long)p.N * p.Ht * p.Wt;
    const long p_begin = (long)blockIdx.z * p.pix_per_split;
    const long p_end = p_begin + p.pix_per_split < npix ? p_begin + p.pix_per_split : npix;
    const bool vecA = (p.Cg & 3) == 0, vecB = (p.Ct & 3) == 0;
    const int tx = tid & 15, ty = tid >> 4;
    const int lk = tid >> 4, lq = tid & 15;    // loader: pixel lk, channel quad lq

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long pb = p_begin; pb < p_end; pb += BK) {
        const long pix = pb + lk;
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
        if (pix < p_end) {
            const int ow = (int)(pix % p.Wt); long tt = pix / p.Wt;
            const int oh = (int)(tt % p.Ht); const int n = (int)(tt / p.Ht);
            const int ih = oh * p.stride + r - p.pad, iw = ow * p.stride + s - p.pad;
            if (ih >= 0 && iw >= 0 && ih < p.Hg && iw < p.Wg) {
                const float* gp = p.g + (((long)n * p.Hg + ih) * p.Wg + iw) * p.Cg;
                const int c = a0 + lq * 4;
                if (vecA) { if (c < p.Cg) av = g2_ldg4(gp + c); }
                else {
                    if (c + 0 < p.Cg) av.x = __ldg(gp + c + 0);
                    if (c + 1 < p.Cg) av.y = __ldg(gp + c + 1);
                    if (c + 2 < p.Cg) av.z = __ldg(gp + c + 2);
                    if (c + 3 < p.Cg) av.w = __ldg(gp + c + 3);
                }
            }
            const float* tp = p.t + pix * p.Ct;
            const int c = b0 + lq * 4;
            if (vecB) { if (c < p.Ct) bv = g2_ldg4(tp + c); }
            else {
                if (c + 0 < p.Ct) bv.x = __ldg(tp + c + 0);
                if (c + 1 < p.Ct) bv.y = __ldg(tp + c + 1);
                if (c + 2 < p.Ct) bv.z = __ldg(tp + c + 2);
                if (c + 3 < p.Ct) bv.w = __ldg(tp + c + 3);
            }
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&As[lk][lq * 4]) = av;
        *reinterpret_cast<float4*>(&Bs[lk][lq * 4]) = bv;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a = a0 + ty * 4 + i;
        if (a >= p.Cg) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = b0 + tx * 4 + j;
            if (b >= p.Ct) continue;
            const long off = p.outT ? ((long)tap * p.Ct + b) * p.Cg + a : ((long)tap * p.Cg + a) * p.Ct + b;
            atomicAdd(p.dw + off, acc[i][j]);
        }
    }
}

// ------------------------------------------------------------------------------------------ GEMM
// C[M,N] (+)= op(A)[M,K] * op(B)[K,N] + bias[N], row-major storage with leading dims; split-K over
// gridDim.z accumulates with atomicAdd (C must be pre-initialised by the caller when splits > 1 or
// accumulate != 0; bias is added by split 0).
struct GemmP {
    const float* A; const float* B; const float* bias; float* C;
    int M, N, K, lda, ldb, ldc, tA, tB, act, atomic;
    int k_per_split;
};

__global__ void __launch_bounds__(256) sgemm_kernel(const GemmP p) {
    constexpr int BM = 64, BN = 64;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int k_begin = blockIdx.z * p.k_per_split;
    const int k_end = min(p.K, k_begin + p.k_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        __syncthreads();
        for (int idx = tid; idx < BM * BK; idx += 256) {
            int m, k;
            if (p.tA) { m = idx % BM; k = idx / BM; } else { k = idx % BK; m = idx / BK; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < p.M && gk < k_end) v = __ldg(p.tA ? p.A + (long)gk * p.lda + gm : p.A + (long)gm * p.lda + gk);
            As[k][m] = v;
        }
        for (int idx = tid; idx < BN * BK; idx += 256) {
            int n, k;
            if (p.tB) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < p.N && gk < k_end) v = __ldg(p.tB ? p.B + (long)gn * p.ldb + gk : p.B + (long)gk * p.ldb + gn);
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            if (p.bias && blockIdx.z == 0) v += __ldg(p.bias + n);
            float* c = p.C + (long)m * p.ldc + n;
            if (p.atomic) atomicAdd(c, v);
            else *c = g2_apply_act(v, p.act, 0.f);
        }
    }
}

// out[c] += sum_m x[m, c]   (C % 4 == 0): float4 loads, 256/(C/4) row lanes per block, smem reduce, atomicAdd
__global__ void __launch_bounds__(256) colsum4_kernel(const float* __restrict__ x, float* __restrict__ out, long M, int C,
                                                      long rows_per_block) {
    const int q = C >> 2, lanes = 256 / q;
    const int t = threadIdx.x, lane = t / q, quad = t - lane * q;
    const long r0 = (long)blockIdx.x * rows_per_block;
    const long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < lanes) {
        long r = r0 + lane;
        for (; r + 3L * lanes < r1; r += 4L * lanes) {          // 4 independent loads in flight
            const float4 a = g2_ldg4(x + r * C + quad * 4), b = g2_ldg4(x + (r + lanes) * C + quad * 4);
            const float4 c = g2_ldg4(x + (r + 2L * lanes) * C + quad * 4), d = g2_ldg4(x + (r + 3L * lanes) * C + quad * 4);
            s.x += (a.x + b.x) + (c.x + d.x); s.y += (a.y + b.y) + (c.y + d.y);
            s.z += (a.z + b.z) + (c.z + d.z); s.w += (a.w + b.w) + (c.w + d.w);
        }
        for (; r < r1; r += lanes) {
            const float4 a = g2_ldg4(x + r * C + quad * 4);
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
    }
    __shared__ float4 sm[256];
    sm[t] = s;
    __syncthreads();
    if (t < q) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < lanes; ++l) { const float4 v = sm[l * q + t]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
        float* o = out + t * 4;
        atomicAdd(o, a.x); atomicAdd(o + 1, a.y); atomicAdd(o + 2, a.z); atomicAdd(o + 3, a.w);
    }
}

__global__ void colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long M, int C, long rows_per_block) {
    // generic C: out[c] += sum_m x[m, c];  block handles rows [blockIdx.x*rpb, ...), threads stride over (row-lane, c)
    extern __shared__ float sm[];
    const int lanes = blockDim.x / C > 0 ? blockDim.x / C : 1;   // row lanes per block (C <= blockDim.x)
    const int c = threadIdx.x % C, lane = threadIdx.x / C;
    const long r0 = (long)blockIdx.x * rows_per_block;
    const long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    float s = 0.f;
    if (lane < lanes)
        for (long r = r0 + lane; r < r1; r += lanes) s += __ldg(x + r * C + c);
    sm[threadIdx.x] = (lane < lanes) ? s : 0.f;
    __syncthreads();
    if (threadIdx.x < C) {
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += sm[l * C + threadIdx.x];
        atomicAdd(out + threadIdx.x, t);
    }
}

// 1x1 output-head weight gradient: dw[o][c] += sum_pix d[pix][o] * h[pix][c]   (d has 4 channels, Cin % 32 == 0).
// One warp per pixel run: lane = channel (coalesced 128-byte row), d broadcast; block partials -> atomicAdd.
__global__ void __launch_bounds__(256) head_wgrad_kernel(const float* __restrict__ h, const float* __restrict__ d4,
                                                         float* __restrict__ dw, long NP, int Cin, long pix_per_block) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long p0 = (long)blockIdx.x * pix_per_block;
    const long p1 = p0 + pix_per_block < NP ? p0 + pix_per_block : NP;
    __shared__ float red[8][4][32];
    for (int cb = 0; cb < Cin; cb += 32) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (long p = p0 + warp; p < p1; p += 8) {
            const float hv = __ldg(h + p * Cin + cb + lane);
            const float4 d = g2_ldg4(d4 + p * 4);
            a0 = fmaf(d.x, hv, a0); a1 = fmaf(d.y, hv, a1); a2 = fmaf(d.z, hv, a2); a3 = fmaf(d.w, hv, a3);
        }
        __syncthreads();
        red[warp][0][lane] = a0; red[warp][1][lane] = a1; red[warp][2][lane] = a2; red[warp][3][lane] = a3;
        __syncthreads();
        if (warp < 4) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w][warp][lane];
            atomicAdd(dw + (long)warp * Cin + cb + lane, s);
        }
    }
}

}  // namespace

extern "C" {

// See include/genesis_b200.h for the contract of each entry point.
int g2_conv_igemm_f32(const float* in, const float* w, const float* bias, const float* aux, float* out,
                      int N, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int R, int S, int stride,
                      int pad, int mode, int wT, int act, cudaStream_t stream) {
    G2_CHECK_ARG(in && w && out && N > 0 && Hi > 0 && Wi > 0 && Ci > 0 && Ho > 0 && Wo > 0 && Co > 0);
    G2_CHECK_ARG(R > 0 && S > 0 && (mode == 0 || mode == 1) && stride >= 1 && pad >= 0);
    if (mode == 1) {
        G2_CHECK_ARG(stride == 1 || stride == 2);
        if (stride == 2) G2_CHECK_ARG((Ho % 2 == 0) && (Wo % 2 == 0));
    }
    if (act == G2_ACT_MUL_RELU_GRAD || act == G2_ACT_MUL_ELU_GRAD) G2_CHECK_ARG(aux != nullptr);
    const long M = (long)N * Ho * Wo;
    G2_CHECK_ARG(M < (1L << 31));
    ConvP p{in, w, bias, aux, out, N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, wT, act, mode};
    if (Co > 32) {
        dim3 grid(g2_cdiv(M, 128), g2_cdiv(Co, 64));
        if (mode == 0) conv_igemm_kernel<0, 128, 64, 8, 4><<<grid, 256, 0, stream>>>(p);
        else conv_igemm_kernel<1, 128, 64, 8, 4><<<grid, 256, 0, stream>>>(p);
    } else {
        dim3 grid(g2_cdiv(M, 128), g2_cdiv(Co, 32));
        if (mode == 0) conv_igemm_kernel<0, 128, 32, 4, 4><<<grid, 256, 0, stream>>>(p);
        else conv_igemm_kernel<1, 128, 32, 4, 4><<<grid, 256, 0, stream>>>(p);
    }
    G2_LAUNCH_RET();
}

int g2_conv_wgrad_f32(const float* g, const float* t, float* dw, int N, int Hg, int Wg, int Cg, int Ht,
                      int Wt, int Ct, int R, int S, int stride, int pad, int outT, cudaStream_t stream) {
    G2_CHECK_ARG(g && t && dw && N > 0 && Cg > 0 && Ct > 0 && R > 0 && S > 0 && stride >= 1);
    const long npix = (long)N * Ht * Wt;
    const int tiles = g2_cdiv(Cg, 64) * g2_cdiv(Ct, 64);
    const int taps = R * S;
    long splits = (148L * 8 + (long)tiles * taps - 1) / ((long)tiles * taps);
    long max_splits = (npix + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    long pps = (npix + splits - 1) / splits;
    pps = (pps + BK - 1) / BK * BK;
    splits = (npix + pps - 1) / pps;
    cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)taps * Cg * Ct, stream);
    if (e != cudaSuccess) return (int)e;
    WgradP p{g, t, dw, N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride, pad, outT, pps};
    dim3 grid(tiles, taps, (unsigned)splits);
    conv_wgrad_kernel<<<grid, 256, 0, stream>>>(p);
    G2_LAUNCH_RET();
}

int g2_gemm_f32(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda,
                int ldb, int ldc, int transA, int transB, int act, int accumulate, cudaStream_t stream) {
    G2_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0);
    const int tiles = g2_cdiv(M, 64) * g2_cdiv(N, 64);
    int splits = 1;
    if (tiles < 148 && K >= 512 && act == G2_ACT_NONE) {
        splits = (296 + tiles - 1) / tiles;
        const int maxs = K / 128;
        if (splits > maxs) splits = maxs;
        if (splits < 1) splits = 1;
    }
    int kps = (K + splits - 1) / splits;
    kps = (kps + BK - 1) / BK * BK;
    splits = (K + kps - 1) / kps;
    const int atomic = (splits > 1 || accumulate) ? 1 : 0;
    if (atomic) G2_CHECK_ARG(act == G2_ACT_NONE);
    if (atomic && !accumulate) {
        cudaError_t e = cudaMemset2DAsync(C, sizeof(float) * (size_t)ldc, 0, sizeof(float) * (size_t)N, (size_t)M, stream);
        if (e != cudaSuccess) return (int)e;
    }
    GemmP p{A, B, bias, C, M, N, K, lda, ldb, ldc, transA, transB, act, atomic, kps};
    dim3 grid(g2_cdiv(M, 64), g2_cdiv(N, 64), splits);
    sgemm_kernel<<<grid, 256, 0, stream>>>(p);
    G2_LAUNCH_RET();
}

int g2_colsum_f32(const float* x, float* out, long M, int C, int accumulate, cudaStream_t stream) {
    G2_CHECK_ARG(x && out && M > 0 && C > 0 && C <= 1024);
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C, stream);
        if (e != cudaSuccess) return (int)e;
    }
    if ((C & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int lanes = 256 / (C / 4);
        long rpb = (long)lanes * 16;
        while (g2_cdiv(M, rpb) > 148L * 8) rpb *= 2;
        colsum4_kernel<<<g2_cdiv(M, rpb), 256, 0, stream>>>(x, out, M, C, rpb);
        G2_LAUNCH_RET();
    }
    const int threads = C >= 256 ? ((C + 31) / 32 * 32) : 256;
    long rpb = (M + 148L * 4 - 1) / (148L * 4);
    if (rpb < 64) rpb = 64;
    const int blocks = g2_cdiv(M, rpb);
    colsum_kernel<<<blocks, threads, threads * sizeof(float), stream>>>(x, out, M, C, rpb);
    G2_LAUNCH_RET();
}

// dw[4][Cin] = d4^T h  (rows >= nout of d4 are zero): weight gradient of the 1x1 output head.
int g2_head_wgrad_f32(const float* h, const float* d4, float* dw, long NP, int Cin, cudaStream_t stream) {
    G2_CHECK_ARG(h && d4 && dw && NP > 0 && Cin >= 32 && (Cin % 32) == 0);
    cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * 4 * (size_t)Cin, stream);
    if (e != cudaSuccess) return (int)e;
    long ppb = 512;
    while (g2_cdiv(NP, ppb) > 148L * 8) ppb *= 2;
    head_wgrad_kernel<<<g2_cdiv(NP, ppb), 256, 0, stream>>>(h, d4, dw, NP, Cin, ppb);
    G2_LAUNCH_RET();
}

}  // extern "C"
