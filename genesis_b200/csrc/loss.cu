// genesis_b200 -- decoder output head and the fused mixture-likelihood kernels (HBM-bound, fp32).
//
// out1x1: per-pixel Cin -> nout (<= 4) projection that reads NHWC activations and writes the NCHW planes the
//         reference returns (x_r_k, mask logits), with an optional fused sigmoid (pixel_bound,
//         modules/component_vae.py:89-93).
// mixture: Genesis.x_loss (models/genesis_config.py:273-286) fused with the reconstruction
//         recon = sum_k m_k * x_r_k (:188-190) and, optionally, the log-softmax over the K reconstructed
//         mask logits (models/monet_config.py:136-140, genesisv2_config.py:213-219):
//         err_b = - sum_{c,p} log sum_k exp(log m_k + log N(x; x_r_k, std_k)), evaluated with a running
//         log-sum-exp.  One thread = 4 consecutive pixels x 3 channels, 128-bit loads.
#include "common.cuh"

namespace {

inline int ew_blocks(long work, int threads = 256) {
    long b = (work + threads - 1) / threads;
    const long cap = 148L * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

constexpr int MAX_CIN = 128;

__global__ void __launch_bounds__(256) out1x1_fwd_kernel(const float* __restrict__ h, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                         long NP, int P, int Cin, int nout, int nsig) {
    __shared__ float ws[4 * MAX_CIN];
    __shared__ float bs[4];
    for (int i = threadIdx.x; i < nout * Cin; i += blockDim.x) ws[i] = w[i];
    if (threadIdx.x < 4) bs[threadIdx.x] = (threadIdx.x < nout && bias) ? bias[threadIdx.x] : 0.f;
    __syncthreads();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < NP; i += (long)gridDim.x * blockDim.x) {
        const float* hp = h + i * Cin;
        float acc[4] = {bs[0], bs[1], bs[2], bs[3]};
        for (int c = 0; c < Cin; c += 4) {
            const float4 v = g2_ldg4(hp + c);
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (o < nout)
                    acc[o] += v.x * ws[o * Cin + c] + v.y * ws[o * Cin + c + 1] + v.z * ws[o * Cin + c + 2] + v.w * ws[o * Cin + c + 3];
        }
        const long n = i / P; const int p = (int)(i - n * P);
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (o < nout) out[(n * nout + o) * P + p] = o < nsig ? 1.f / (1.f + expf(-acc[o])) : acc[o];
    }
}

// dh [N,P,Cin] = sum_o dpre_o * W[o,:];  dpre4 [N,P,4] (channels >= nout are zero)
__global__ void __launch_bounds__(256) out1x1_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                         const float* __restrict__ w, float* __restrict__ dh,
                                                         float* __restrict__ dpre4, long NP, int P, int Cin, int nout, int nsig) {
    __shared__ float ws[4 * MAX_CIN];
    for (int i = threadIdx.x; i < 4 * Cin; i += blockDim.x) ws[i] = i < nout * Cin ? w[i] : 0.f;
    __syncthreads();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < NP; i += (long)gridDim.x * blockDim.x) {
        const long n = i / P; const int p = (int)(i - n * P);
        float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (o < nout) {
                float g = __ldg(dout + (n * nout + o) * P + p);
                if (o < nsig) { const float y = __ldg(out + (n * nout + o) * P + p); g *= y * (1.f - y); }
                d[o] = g;
            }
        *reinterpret_cast<float4*>(dpre4 + i * 4) = make_float4(d[0], d[1], d[2], d[3]);
        if (dh) {
            float* hp = dh + i * Cin;
            for (int c = 0; c < Cin; c += 4) {
                float4 v;
                v.x = d[0] * ws[c] + d[1] * ws[Cin + c] + d[2] * ws[2 * Cin + c] + d[3] * ws[3 * Cin + c];
                v.y = d[0] * ws[c + 1] + d[1] * ws[Cin + c + 1] + d[2] * ws[2 * Cin + c + 1] + d[3] * ws[3 * Cin + c + 1];
                v.z = d[0] * ws[c + 2] + d[1] * ws[Cin + c + 2] + d[2] * ws[2 * Cin + c + 2] + d[3] * ws[3 * Cin + c + 2];
                v.w = d[0] * ws[c + 3] + d[1] * ws[Cin + c + 3] + d[2] * ws[2 * Cin + c + 3] + d[3] * ws[3 * Cin + c + 3];
                *reinterpret_cast<float4*>(hp + c) = v;
            }
        }
    }
}

struct MixP {
    const float* x;        // [B,3,P]
    const float* xr;       // [K,B,3,P]
    const float* lm;       // [K,B,P]   log masks, or mask LOGITS when softmax != 0
    const float* stdv;     // [K]
    float* err;            // [B]  (pre-zeroed, atomicAdd)
    float* recon;          // [B,3,P]
    float* lse;            // [B,3,P]  saved log sum_k exp(.) for the backward
    float* lm_out;         // [K,B,P]  log-softmax masks (softmax != 0), else unused
    int K, B, P, softmax;
    int xr_cs, lm_cs;      // channels per (k,b) slot in xr (3 dense, 4 when packed with the mask logit) / in lm (1 or 4)
};

__device__ __forceinline__ void lse_push(float& m, float& s, float a) {
    if (a > m) { s = s * expf(m - a) + 1.f; m = a; } else { s += expf(a - m); }
}

// grid (P/4/256, B)
__global__ void __launch_bounds__(256) mixture_fwd_kernel(const MixP p) {
    const int b = blockIdx.y;
    const int i4 = blockIdx.x * blockDim.x + threadIdx.x;     // quad of pixels
    const int P4 = p.P >> 2;
    float part = 0.f;
    if (i4 < P4) {
        const long KBP = (long)p.B * p.P;
        const long LKB = KBP * p.lm_cs;                       // lm stride between slots k
        const float* lmb = p.lm + (long)b * p.P * p.lm_cs + i4 * 4;
        float4 mnorm = make_float4(0.f, 0.f, 0.f, 0.f);      // log sum_k exp(logit_k) when softmax
        if (p.softmax) {
            float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, s[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = 0; k < p.K; ++k) {
                const float4 l = g2_ldg4(lmb + k * LKB);
                lse_push(m[0], s[0], l.x); lse_push(m[1], s[1], l.y); lse_push(m[2], s[2], l.z); lse_push(m[3], s[3], l.w);
            }
            mnorm = make_float4(m[0] + logf(s[0]), m[1] + logf(s[1]), m[2] + logf(s[2]), m[3] + logf(s[3]));
        }
        float4 xv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) xv[c] = g2_ldg4(p.x + ((long)b * 3 + c) * p.P + i4 * 4);
        float mx[3][4], sm[3][4], rc[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) { mx[c][j] = -INFINITY; sm[c][j] = 0.f; rc[c][j] = 0.f; }
        for (int k = 0; k < p.K; ++k) {
            float4 l = g2_ldg4(lmb + k * LKB);
            if (p.softmax) {
                l.x -= mnorm.x; l.y -= mnorm.y; l.z -= mnorm.z; l.w -= mnorm.w;
                *reinterpret_cast<float4*>(p.lm_out + k * KBP + (long)b * p.P + i4 * 4) = l;
            }
            const float sd = __ldg(p.stdv + k);
            const float inv2v = 0.5f / (sd * sd), cst = -logf(sd) - 0.9189385332046727f;
            const float lv[4] = {l.x, l.y, l.z, l.w};
            float mk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) mk[j] = expf(lv[j]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 r4 = g2_ldg4(p.xr + (((long)k * p.B + b) * p.xr_cs + c) * p.P + i4 * 4);
                const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
                const float xx[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float d = xx[j] - rv[j];
                    lse_push(mx[c][j], sm[c][j], lv[j] + cst - d * d * inv2v);
                    rc[c][j] = fmaf(mk[j], rv[j], rc[c][j]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float4 L;
            L.x = mx[c][0] + logf(sm[c][0]); L.y = mx[c][1] + logf(sm[c][1]);
            L.z = mx[c][2] + logf(sm[c][2]); L.w = mx[c][3] + logf(sm[c][3]);
            part -= (L.x + L.y) + (L.z + L.w);
            const long o = ((long)b * 3 + c) * p.P + i4 * 4;
            *reinterpret_cast<float4*>(p.lse + o) = L;
            *reinterpret_cast<float4*>(p.recon + o) = make_float4(rc[c][0], rc[c][1], rc[c][2], rc[c][3]);
        }
    }
    __shared__ float red[32];
    part = g2_block_sum(part, red);
    if (threadIdx.x == 0) atomicAdd(p.err + b, part);
}

struct MixBP {
    const float* x; const float* xr; const float* lm /* log masks (post-softmax when softmax) */; const float* stdv;
    const float* lse; const float* gerr;   // [B] upstream gradient of err (PIX: [B,3,P] upstream gradient of the per-pixel, per-channel loss)
    float* dxr;            // [K,B,3,P]
    float* dlm;            // [K,B,P]   gradient w.r.t. log masks, or w.r.t. mask LOGITS when softmax != 0
    int K, B, P, softmax;
    int xr_cs, lm_cs, dlm_cs;   // slot strides (channels) of xr/dxr, of lm, of dlm
};

// r_kc = exp(a_kc - lse_c);  d err / d log m_k = - sum_c r_kc;  d err / d xr_kc = - r_kc (x_c - xr_kc)/std_k^2
// softmax: d/d logit_k = G_k - m_k * sum_j G_j   with sum_j G_j = -3 (responsibilities sum to one per channel)
template <bool PIX>
__global__ void __launch_bounds__(256) mixture_bwd_kernel(const MixBP p) {
    const int b = blockIdx.y;
    const int i4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int P4 = p.P >> 2;
    if (i4 >= P4) return;
    const long KBP = (long)p.B * p.P;
    const float g = PIX ? 1.f : __ldg(p.gerr + b);
    float4 xv[3], Lv[3], Gv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const long o = ((long)b * 3 + c) * p.P + i4 * 4;
        xv[c] = g2_ldg4(p.x + o); Lv[c] = g2_ldg4(p.lse + o);
        if (PIX) Gv[c] = g2_ldg4(p.gerr + o);
    }
    for (int k = 0; k < p.K; ++k) {
        const float4 l = g2_ldg4(p.lm + ((long)k * p.B + b) * p.lm_cs * p.P + i4 * 4);
        const float sd = __ldg(p.stdv + k);
        const float inv2v = 0.5f / (sd * sd), cst = -logf(sd) - 0.9189385332046727f;
        const float lv[4] = {l.x, l.y, l.z, l.w};
        float gm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const long o = (((long)k * p.B + b) * p.xr_cs + c) * p.P + i4 * 4;
            const float4 r4 = g2_ldg4(p.xr + o);
            const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
            const float xx[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w};
            const float LL[4] = {Lv[c].x, Lv[c].y, Lv[c].z, Lv[c].w};
            const float GG[4] = {PIX ? Gv[c].x : 1.f, PIX ? Gv[c].y : 1.f, PIX ? Gv[c].z : 1.f, PIX ? Gv[c].w : 1.f};
            float dr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = xx[j] - rv[j];
                const float r = expf(lv[j] + cst - d * d * inv2v - LL[j]);
                if (PIX) { gm[j] += GG[j] * r; dr[j] = -GG[j] * r * d * 2.f * inv2v; }
                else { gm[j] += r; dr[j] = -g * r * d * 2.f * inv2v; }
            }
            *reinterpret_cast<float4*>(p.dxr + o) = make_float4(dr[0], dr[1], dr[2], dr[3]);
        }
        float4 o4;
        if (p.softmax) {
            o4.x = -g * (gm[0] - 3.f * expf(lv[0])); o4.y = -g * (gm[1] - 3.f * expf(lv[1]));
            o4.z = -g * (gm[2] - 3.f * expf(lv[2])); o4.w = -g * (gm[3] - 3.f * expf(lv[3]));
        } else {
            o4 = make_float4(-g * gm[0], -g * gm[1], -g * gm[2], -g * gm[3]);
        }
        *reinterpret_cast<float4*>(p.dlm + ((long)k * p.B + b) * p.dlm_cs * p.P + i4 * 4) = o4;
    }
}


// ------------------------------------------------------------------------------------------ MONet mask KL
// MONet.kl_m_loss (reference models/monet_config.py:157-170) fused with get_mask_recon_stack(softmax, log=True)
// (:136-140):   q_k = max(exp(lm_k), 1e-5) / sum_j max(exp(lm_j), 1e-5)            (torch Categorical renormalises)
//               lr_k = logit_k - logsumexp_j logit_j  (written to lmr);  p_k = max(exp(lr_k), 1e-5) / sum_j (...)
//               kl_b = sum_pixels sum_k q_k (log q_k - log p_k)
// lm [K,B,lm_cs,P] (plane 0), logits [K,B,lg_cs,P] (plane 0 of the given base pointer).  One thread = one pixel quad.
struct MklP {
    const float* lm; const float* lg; float* lmr; float* kl; const float* gkl; float* dlm; float* dlg;
    int K, B, P, lm_cs, lg_cs, dlm_cs, dlg_cs, accumulate_dlm;
};
constexpr float MKL_FLOOR = 1e-5f;
constexpr float MKL_EPS = 1.1920928955078125e-07f;   // torch.finfo(float32).eps used by probs_to_logits

template <bool BWD>
__global__ void __launch_bounds__(256) mask_kl_kernel(const MklP p) {
    const int b = blockIdx.y;
    const int i4 = blockIdx.x * blockDim.x + threadIdx.x;
    float part = 0.f;
    if (i4 < p.P / 4) {
        const long off = (long)i4 * 4;
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, se[4] = {0.f, 0.f, 0.f, 0.f}, Q[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < p.K; ++k) {
            const float4 l4 = g2_ldg4(p.lg + ((long)k * p.B + b) * p.lg_cs * p.P + off);
            const float4 m4 = g2_ldg4(p.lm + ((long)k * p.B + b) * p.lm_cs * p.P + off);
            const float l[4] = {l4.x, l4.y, l4.z, l4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (l[j] > mx[j]) { se[j] = se[j] * expf(mx[j] - l[j]) + 1.f; mx[j] = l[j]; } else se[j] += expf(l[j] - mx[j]);
                Q[j] += fmaxf(expf(m[j]), MKL_FLOOR);
            }
        }
        float lse[4], Pn[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) lse[j] = mx[j] + logf(se[j]);
        for (int k = 0; k < p.K; ++k) {
            const float4 l4 = g2_ldg4(p.lg + ((long)k * p.B + b) * p.lg_cs * p.P + off);
            const float lr[4] = {l4.x - lse[0], l4.y - lse[1], l4.z - lse[2], l4.w - lse[3]};
#pragma unroll
            for (int j = 0; j < 4; ++j) Pn[j] += fmaxf(expf(lr[j]), MKL_FLOOR);
            if (!BWD) *reinterpret_cast<float4*>(p.lmr + ((long)k * p.B + b) * p.P + off) = make_float4(lr[0], lr[1], lr[2], lr[3]);
        }
        float f[4] = {0.f, 0.f, 0.f, 0.f}, gs[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < p.K; ++k) {
            const float4 l4 = g2_ldg4(p.lg + ((long)k * p.B + b) * p.lg_cs * p.P + off);
            const float4 m4 = g2_ldg4(p.lm + ((long)k * p.B + b) * p.lm_cs * p.P + off);
            const float l[4] = {l4.x, l4.y, l4.z, l4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float qn = fmaxf(expf(m[j]), MKL_FLOOR) / Q[j];
                const float er = expf(l[j] - lse[j]);
                const float pn = fmaxf(er, MKL_FLOOR) / Pn[j];
                const float d = logf(fminf(fmaxf(qn, MKL_EPS), 1.f - MKL_EPS)) - logf(fminf(fmaxf(pn, MKL_EPS), 1.f - MKL_EPS));
                f[j] += qn * d;
                if (BWD) gs[j] += er > MKL_FLOOR ? er * (1.f - qn / pn) / Pn[j] : 0.f;     // sum_j d f / d lr_j
            }
        }
        if (!BWD) {
            part = (f[0] + f[1]) + (f[2] + f[3]);
        } else {
            const float g = __ldg(p.gkl + b);
            for (int k = 0; k < p.K; ++k) {
                const float4 l4 = g2_ldg4(p.lg + ((long)k * p.B + b) * p.lg_cs * p.P + off);
                const float4 m4 = g2_ldg4(p.lm + ((long)k * p.B + b) * p.lm_cs * p.P + off);
                const float l[4] = {l4.x, l4.y, l4.z, l4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w};
                float dm[4], dl[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float em = expf(m[j]);
                    const float qn = fmaxf(em, MKL_FLOOR) / Q[j];
                    const float er = expf(l[j] - lse[j]);
                    const float pn = fmaxf(er, MKL_FLOOR) / Pn[j];
                    const float d = logf(fminf(fmaxf(qn, MKL_EPS), 1.f - MKL_EPS)) - logf(fminf(fmaxf(pn, MKL_EPS), 1.f - MKL_EPS));
                    dm[j] = em > MKL_FLOOR ? g * em * (d - f[j]) / Q[j] : 0.f;
                    const float glr = er > MKL_FLOOR ? er * (1.f - qn / pn) / Pn[j] : 0.f;
                    dl[j] = g * (glr - er * gs[j]);
                }
                float* dmp = p.dlm + ((long)k * p.B + b) * p.dlm_cs * p.P + off;
                float4 o = make_float4(dm[0], dm[1], dm[2], dm[3]);
                if (p.accumulate_dlm) { const float4 a = *reinterpret_cast<const float4*>(dmp); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
                *reinterpret_cast<float4*>(dmp) = o;
                *reinterpret_cast<float4*>(p.dlg + ((long)k * p.B + b) * p.dlg_cs * p.P + off) = make_float4(dl[0], dl[1], dl[2], dl[3]);
            }
        }
    }
    if (!BWD) {
        __shared__ float red[32];
        part = g2_block_sum(part, red);
        if (threadIdx.x == 0) atomicAdd(p.kl + b, part);
    }
}

}  // namespace

extern "C" {

int g2_out1x1_fwd_f32(const float* h, const float* w, const float* bias, float* out, long N, int P, int Cin, int nout,
                      int nsig, cudaStream_t stream) {
    G2_CHECK_ARG(h && w && out && N > 0 && P > 0 && Cin >= 4 && (Cin % 4) == 0 && Cin <= MAX_CIN && nout >= 1 && nout <= 4);
    const long NP = N * P;
    out1x1_fwd_kernel<<<ew_blocks(NP), 256, 0, stream>>>(h, w, bias, out, NP, P, Cin, nout, nsig);
    G2_LAUNCH_RET();
}

int g2_out1x1_bwd_f32(const float* dout, const float* out, const float* w, float* dh, float* dpre4, long N, int P, int Cin,
                      int nout, int nsig, cudaStream_t stream) {
    G2_CHECK_ARG(dout && out && w && dpre4 && N > 0 && P > 0 && Cin >= 4 && (Cin % 4) == 0 && Cin <= MAX_CIN && nout >= 1 && nout <= 4);
    const long NP = N * P;
    out1x1_bwd_kernel<<<ew_blocks(NP), 256, 0, stream>>>(dout, out, w, dh, dpre4, NP, P, Cin, nout, nsig);
    G2_LAUNCH_RET();
}

int g2_mixture_fwd_f32(const float* x, const float* xr, const float* lm, const float* stdv, float* err, float* recon,
                       float* lse, float* lm_out, int K, int B, int P, int softmax, int xr_cs, int lm_cs, cudaStream_t stream) {
    G2_CHECK_ARG(x && xr && lm && stdv && err && recon && lse && K >= 1 && B > 0 && P > 0 && (P % 4) == 0);
    G2_CHECK_ARG(xr_cs >= 3 && lm_cs >= 1);
    if (softmax) G2_CHECK_ARG(lm_out != nullptr);
    cudaError_t e = cudaMemsetAsync(err, 0, sizeof(float) * (size_t)B, stream);
    if (e != cudaSuccess) return (int)e;
    MixP p{x, xr, lm, stdv, err, recon, lse, lm_out, K, B, P, softmax, xr_cs, lm_cs};
    dim3 grid(g2_cdiv(P / 4, 256), B);
    mixture_fwd_kernel<<<grid, 256, 0, stream>>>(p);
    G2_LAUNCH_RET();
}

int g2_mixture_bwd_f32(const float* x, const float* xr, const float* lm, const float* stdv, const float* lse,
                       const float* gerr, float* dxr, float* dlm, int K, int B, int P, int softmax, int xr_cs, int lm_cs,
                       int dlm_cs, cudaStream_t stream) {
    G2_CHECK_ARG(x && xr && lm && stdv && lse && gerr && dxr && dlm && K >= 1 && B > 0 && P > 0 && (P % 4) == 0);
    MixBP p{x, xr, lm, stdv, lse, gerr, dxr, dlm, K, B, P, softmax, xr_cs, lm_cs, dlm_cs};
    dim3 grid(g2_cdiv(P / 4, 256), B);
    mixture_bwd_kernel<false><<<grid, 256, 0, stream>>>(p);
    G2_LAUNCH_RET();
}

// Backward of the PIXEL-WISE loss (Genesis.x_loss(pixel_wise=True), genesis_config.py:283-284): err_ppc = -lse [B,3,P] with
// upstream gradient gpix [B,3,P]; given masks only (softmax = 0).
int g2_mixture_bwd_pix_f32(const float* x, const float* xr, const float* lm, const float* stdv, const float* lse,
                           const float* gpix, float* dxr, float* dlm, int K, int B, int P, int xr_cs, int lm_cs, int dlm_cs,
                           cudaStream_t stream) {
    G2_CHECK_ARG(x && xr && lm && stdv && lse && gpix && dxr && dlm && K >= 1 && B > 0 && P > 0 && (P % 4) == 0);
    MixBP p{x, xr, lm, stdv, lse, gpix, dxr, dlm, K, B, P, 0, xr_cs, lm_cs, dlm_cs};
    dim3 grid(g2_cdiv(P / 4, 256), B);
    mixture_bwd_kernel<true><<<grid, 256, 0, stream>>>(p);
    G2_LAUNCH_RET();
}

int g2_mask_kl_fwd_f32(const float* lm, const float* logits, float* lmr, float* kl, int K, int B, int P, int lm_cs, int lg_cs,
                       cudaStream_t stream) {
    G2_CHECK_ARG(lm && logits && lmr && kl && K >= 1 && B > 0 && P > 0 && (P % 4) == 0 && lm_cs >= 1 && lg_cs >= 1);
    cudaError_t e = cudaMemsetAsync(kl, 0, sizeof(float) * (size_t)B, stream);
    if (e != cudaSuccess) return (int)e;
    MklP p{lm, logits, lmr, kl, nullptr, nullptr, nullptr, K, B, P, lm_cs, lg_cs, 1, 1, 0};
    dim3 grid(g2_cdiv(P / 4, 256), B);
    mask_kl_kernel<false><<<grid, 256, 0, stream>>>(p);
    G2_LAUNCH_RET();
}

int g2_mask_kl_bwd_f32(const float* lm, const float* logits, const float* gkl, float* dlm, float* dlogits, int K, int B, int P,
                       int lm_cs, int lg_cs, int dlm_cs, int dlg_cs, int accumulate_dlm, cudaStream_t stream) {
    G2_CHECK_ARG(lm && logits && gkl && dlm && dlogits && K >= 1 && B > 0 && P > 0 && (P % 4) == 0);
    G2_CHECK_ARG(lm_cs >= 1 && lg_cs >= 1 && dlm_cs >= 1 && dlg_cs >= 1);
    MklP p{lm, logits, nullptr, nullptr, gkl, dlm, dlogits, K, B, P, lm_cs, lg_cs, dlm_cs, dlg_cs, accumulate_dlm};
    dim3 grid(g2_cdiv(P / 4, 256), B);
    mask_kl_kernel<true><<<grid, 256, 0, stream>>>(p);
    G2_LAUNCH_RET();
}

}  // extern "C"
