// genesis_b200 -- kernels specific to GENESIS-V2 / MONet: nearest resampling for the UNet, the instance-colouring
// stick-breaking process (IC-SBP) and the masked feature pooling.  fp32, NHWC.
#include "common.cuh"

namespace {

inline int ew_blocks(long work, int threads = 256) {
    long b = (work + threads - 1) / threads;
    const long cap = 148L * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// ------------------------------------------------------------------------------------------ nearest x0.5 / x2
// mode 0: y[n,h,w,:] = x[n,2h,2w,:]            (F.interpolate(scale_factor=0.5,'nearest'), reference unet.py:77-78)
// mode 1: y[n,h,w,:] = x[n,h/2,w/2,:]          (scale_factor=2.0, unet.py:88-89)
// mode 2: y (HxW, zeros except y[n,2h,2w,:] = x[n,h,w,:])       = backward of mode 0   (x is H/2 x W/2)
// mode 3: y[n,h,w,:] = sum_{a,b<2} x[n,2h+a,2w+b,:]              = backward of mode 1   (x is 2H x 2W)
// Ho, Wo are the OUTPUT sizes; C % 4 == 0.
template <typename I>
__global__ void resample_kernel(const float* __restrict__ x, float* __restrict__ y, I total_quads, int Ho_, int Wo_, int C, int mode) {
    // I = unsigned when the quad count fits 31 bits: the four 64-bit divisions per float4 made this an ALU-bound kernel
    const I q = (I)(C >> 2), Ho = (I)Ho_, Wo = (I)Wo_;
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total_quads; i += (I)gridDim.x * blockDim.x) {
        const I quad = i % q; I t = i / q;
        const I w = t % Wo; t /= Wo;
        const I h = t % Ho; const long n = (long)(t / Ho);
        float4 v;
        if (mode == 0) {
            v = g2_ldg4(x + ((n * (2 * Ho) + 2 * h) * (2L * Wo) + 2 * w) * C + quad * 4);
        } else if (mode == 1) {
            v = g2_ldg4(x + ((n * (Ho / 2) + h / 2) * (long)(Wo / 2) + w / 2) * C + quad * 4);
        } else if (mode == 2) {
            v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!((h | w) & 1)) v = g2_ldg4(x + ((n * (Ho / 2) + h / 2) * (long)(Wo / 2) + w / 2) * C + quad * 4);
        } else {
            const float* b = x + ((n * (2 * Ho) + 2 * h) * (2L * Wo) + 2 * w) * C + quad * 4;
            const float4 a0 = g2_ldg4(b), a1 = g2_ldg4(b + C), a2 = g2_ldg4(b + 2L * Wo * C), a3 = g2_ldg4(b + 2L * Wo * C + C);
            v = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z),
                            (a0.w + a1.w) + (a2.w + a3.w));
        }
        *reinterpret_cast<float4*>(y + (long)i * 4) = v;
    }
}

// ------------------------------------------------------------------------------------------ IC-SBP
// InstanceColouringSBP.forward, gaussian kernel (reference modules/attention.py:177-223).  One CTA per image.
// colour [B,P,CD] (NHWC, CD = 8), u [B,P] uniform draws, log_sigma scalar.  steps = K-1.
//   per step k: idx = argmax_p u_p * exp(log_s_k,p) (lowest index wins ties); seed = colour[idx];
//               alpha = exp(-|colour_p - seed|^2 / exp(log_sigma)); ac = clamp(alpha, .01, .99)
//               log_m_k = log_s_k + log ac;  log_s_{k+1} = log_s_k + log(1 - ac)
//   log_m_{K-1} = log_s_{K-1}.   Outputs log_m [K,B,P], log_s [K,B,P], seed_idx [K-1,B] (int32).
constexpr int CD = 8;
constexpr int IC_THREADS = 1024;
constexpr int IC_MAXPPT = 16;       // pixels per thread (P <= 16384)

// KT: 0 gaussian  alpha = exp(-|c - seed|^2 / sigma);  1 laplacian  alpha = exp(-sqrt(clamp(|c - seed|^2, 1e-10, 1e10)) / sigma)
// (blocks.py:49-61);  2 epanechnikov  alpha = relu(1 - |c - seed|^2 / sigma)   (attention.py:195-203).
template <int KT>
__device__ __forceinline__ float icsbp_alpha(float dist, float inv_sigma) {
    if constexpr (KT == 0) return expf(-dist * inv_sigma);
    else if constexpr (KT == 1) return expf(-sqrtf(fminf(fmaxf(dist, 1e-10f), 1e10f)) * inv_sigma);
    else return fmaxf(1.f - dist * inv_sigma, 0.f);
}

// DYN (`dynamic_K`, attention.py:168-169, 218-219): a step whose mask would hold fewer than 20 pixels ends the image's loop
// BEFORE the mask is appended; the current scope becomes the last mask.  n_masks[b] = number of masks (<= K); the unused
// slots k >= n_masks[b] are filled with -1e10 as GenesisV2.forward does for batches (genesisv2_config.py:126-131).
template <int KT, bool DYN>
__global__ void __launch_bounds__(IC_THREADS) icsbp_fwd_kernel(const float* __restrict__ colour, const float* __restrict__ u,
                                                               const float* __restrict__ log_sigma, float* __restrict__ log_m,
                                                               float* __restrict__ log_s, int* __restrict__ seed_idx,
                                                               int* __restrict__ n_masks, int B, int P, int K) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int ppt = (P + IC_THREADS - 1) / IC_THREADS;
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    __shared__ float s_seed[CD];
    __shared__ int s_best;
    const float inv_sigma = 1.f / expf(__ldg(log_sigma));
    const float* col = colour + (long)b * P * CD;
    float ls[IC_MAXPPT];
#pragma unroll
    for (int j = 0; j < IC_MAXPPT; ++j) ls[j] = 0.f;
    const long KB = (long)B * P;
    __shared__ float s_sum[32];
    __shared__ float s_total;
    int kf = K - 1;                       // index of the final mask (= scope)
    for (int k = 0; k < K - 1; ++k) {
        // ---- block argmax of u * scope (first maximum, as torch.argmax)
        float best = -1.f; int bi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < IC_MAXPPT; ++j) {
            const int p = tid + j * IC_THREADS;
            if (j < ppt && p < P) {
                const float v = __ldg(u + (long)b * P + p) * expf(ls[j]);
                if (v > best) { best = v; bi = p; }       // p increases with j: keeps the lowest index on ties
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_idx[tid >> 5] = bi; }
        __syncthreads();
        if (tid < 32) {
            best = s_val[tid]; bi = s_idx[tid];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (tid == 0) { s_best = bi; seed_idx[(long)k * B + b] = bi; }
        }
        __syncthreads();
        if (tid < CD) s_seed[tid] = __ldg(col + (long)s_best * CD + tid);
        __syncthreads();
        if constexpr (DYN) {
            // ---- would-be mask mass sum_p exp(log_s + log alpha); fewer than 20 pixels ends the loop (attention.py:218-219)
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < IC_MAXPPT; ++j) {
                const int p = tid + j * IC_THREADS;
                if (j < ppt && p < P) {
                    const float4 c0 = g2_ldg4(col + (long)p * CD), c1 = g2_ldg4(col + (long)p * CD + 4);
                    const float d0 = c0.x - s_seed[0], d1 = c0.y - s_seed[1], d2 = c0.z - s_seed[2], d3 = c0.w - s_seed[3];
                    const float d4 = c1.x - s_seed[4], d5 = c1.y - s_seed[5], d6 = c1.z - s_seed[6], d7 = c1.w - s_seed[7];
                    const float dist = ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3)) + ((d4 * d4 + d5 * d5) + (d6 * d6 + d7 * d7));
                    const float a = fminf(fmaxf(icsbp_alpha<KT>(dist, inv_sigma), 0.01f), 0.99f);
                    part += expf(ls[j] + logf(a));
                }
            }
            part = g2_block_sum(part, s_sum);
            if (tid == 0) s_total = part;
            __syncthreads();
            if (s_total < 20.f) { kf = k; break; }
        }
        // ---- masks
#pragma unroll
        for (int j = 0; j < IC_MAXPPT; ++j) {
            const int p = tid + j * IC_THREADS;
            if (j < ppt && p < P) {
                const float4 c0 = g2_ldg4(col + (long)p * CD), c1 = g2_ldg4(col + (long)p * CD + 4);
                const float d0 = c0.x - s_seed[0], d1 = c0.y - s_seed[1], d2 = c0.z - s_seed[2], d3 = c0.w - s_seed[3];
                const float d4 = c1.x - s_seed[4], d5 = c1.y - s_seed[5], d6 = c1.z - s_seed[6], d7 = c1.w - s_seed[7];
                const float dist = ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3)) + ((d4 * d4 + d5 * d5) + (d6 * d6 + d7 * d7));
                float a = icsbp_alpha<KT>(dist, inv_sigma);
                a = fminf(fmaxf(a, 0.01f), 0.99f);
                log_s[k * KB + (long)b * P + p] = ls[j];
                log_m[k * KB + (long)b * P + p] = ls[j] + logf(a);
                ls[j] += logf(1.f - a);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < IC_MAXPPT; ++j) {
        const int p = tid + j * IC_THREADS;
        if (j < ppt && p < P) {
            log_s[(long)kf * KB + (long)b * P + p] = ls[j];
            log_m[(long)kf * KB + (long)b * P + p] = ls[j];
            if constexpr (DYN) {
                for (int k2 = kf + 1; k2 < K; ++k2) {
                    log_s[(long)k2 * KB + (long)b * P + p] = ls[j];
                    log_m[(long)k2 * KB + (long)b * P + p] = -1e10f;
                }
            }
        }
    }
    if constexpr (DYN) {
        if (tid == 0) {
            n_masks[b] = kf + 1;
            for (int k2 = kf + 1; k2 < K - 1; ++k2) seed_idx[(long)k2 * B + b] = -1;
        }
    }
}

// Backward: dcolour [B,P,CD], dlog_sigma partial per image [B] (summed by the caller), from dlog_m [K,B,P].
// alpha does not depend on the scope, so only the seed indices of the forward are needed.  Reverse scan:
//   R = G_{K-1};  for k = K-2..0: d(alpha_c) = G_k / ac - R / (1 - ac)   (straight-through clamp);  R += G_k
//   d(dist) = -d(alpha) * alpha / sigma;  d log_sigma += d(alpha) * alpha * dist / sigma
//   d colour_p += 2 (colour_p - seed) d(dist);  d seed -= sum_p 2 (colour_p - seed) d(dist)  -> colour[idx_k]
// Other kernels (KT): laplacian  d = sqrt(clamp_ste(dist)):  d(dist) = -d(alpha) alpha / (2 d sigma),  d log_sigma += d(alpha) alpha d / sigma;
// epanechnikov, where 1 - dist / sigma > 0:  d(dist) = -d(alpha) / sigma,  d log_sigma += d(alpha) dist / sigma  (else both 0).
template <int KT>
__global__ void __launch_bounds__(IC_THREADS) icsbp_bwd_kernel(const float* __restrict__ colour, const float* __restrict__ log_sigma,
                                                               const int* __restrict__ seed_idx, const float* __restrict__ dlog_m,
                                                               float* __restrict__ dcolour, float* __restrict__ dlog_sigma_b,
                                                               const int* __restrict__ n_masks, int B, int P, int Kmax) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int K = n_masks ? __ldg(n_masks + b) : Kmax;         // dynamic_K: masks of this image; slots beyond it carry no gradient
    const int ppt = (P + IC_THREADS - 1) / IC_THREADS;
    __shared__ float s_seed[CD];
    __shared__ float s_red[32][CD + 1];
    const float inv_sigma = 1.f / expf(__ldg(log_sigma));
    const float* col = colour + (long)b * P * CD;
    float* dcol = dcolour + (long)b * P * CD;
    const long KB = (long)B * P;
    // each thread owns pixels tid + j*1024: their dcolour rows are accumulated in global memory (L2-resident,
    // 128 KB per image) instead of 128 registers per thread
    float R[IC_MAXPPT];
#pragma unroll
    for (int j = 0; j < IC_MAXPPT; ++j) {
        const int p = tid + j * IC_THREADS;
        const bool ok = j < ppt && p < P;
        R[j] = ok ? __ldg(dlog_m + (long)(K - 1) * KB + (long)b * P + p) : 0.f;
        if (ok) {
            *reinterpret_cast<float4*>(dcol + (long)p * CD) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(dcol + (long)p * CD + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float dls = 0.f;
    for (int k = K - 2; k >= 0; --k) {
        const int sidx = __ldg(seed_idx + (long)k * B + b);
        __syncthreads();
        if (tid < CD) s_seed[tid] = __ldg(col + (long)sidx * CD + tid);
        __syncthreads();
        float dseed[CD];
#pragma unroll
        for (int c = 0; c < CD; ++c) dseed[c] = 0.f;
#pragma unroll
        for (int j = 0; j < IC_MAXPPT; ++j) {
            const int p = tid + j * IC_THREADS;
            if (j < ppt && p < P) {
                const float4 c0 = g2_ldg4(col + (long)p * CD), c1 = g2_ldg4(col + (long)p * CD + 4);
                const float df[CD] = {c0.x - s_seed[0], c0.y - s_seed[1], c0.z - s_seed[2], c0.w - s_seed[3],
                                      c1.x - s_seed[4], c1.y - s_seed[5], c1.z - s_seed[6], c1.w - s_seed[7]};
                float dist = 0.f;
#pragma unroll
                for (int c = 0; c < CD; ++c) dist += df[c] * df[c];
                const float a = icsbp_alpha<KT>(dist, inv_sigma);
                const float ac = fminf(fmaxf(a, 0.01f), 0.99f);
                const float G = __ldg(dlog_m + (long)k * KB + (long)b * P + p);
                const float da = G / ac - R[j] / (1.f - ac);
                R[j] += G;
                float dd;
                if constexpr (KT == 0) {
                    dd = -da * a * inv_sigma;
                    dls += da * a * dist * inv_sigma;
                } else if constexpr (KT == 1) {
                    const float d = sqrtf(fminf(fmaxf(dist, 1e-10f), 1e10f));
                    dd = -da * a * inv_sigma * (0.5f / d);
                    dls += da * a * d * inv_sigma;
                } else {
                    const bool on = 1.f - dist * inv_sigma > 0.f;
                    dd = on ? -da * inv_sigma : 0.f;
                    dls += on ? da * dist * inv_sigma : 0.f;
                }
                float g[CD];
#pragma unroll
                for (int c = 0; c < CD; ++c) { g[c] = 2.f * df[c] * dd; dseed[c] -= g[c]; }
                float4* dp = reinterpret_cast<float4*>(dcol + (long)p * CD);
                float4 o0 = dp[0], o1 = dp[1];
                o0.x += g[0]; o0.y += g[1]; o0.z += g[2]; o0.w += g[3];
                o1.x += g[4]; o1.y += g[5]; o1.z += g[6]; o1.w += g[7];
                dp[0] = o0; dp[1] = o1;
            }
        }
        // block-reduce dseed (8 values) and add it to the seed pixel's gradient (held by exactly one thread)
#pragma unroll
        for (int c = 0; c < CD; ++c) dseed[c] = g2_warp_sum(dseed[c]);
        if ((tid & 31) == 0)
#pragma unroll
            for (int c = 0; c < CD; ++c) s_red[tid >> 5][c] = dseed[c];
        __syncthreads();
        if (tid == sidx % IC_THREADS) {          // the thread that owns the seed pixel adds d(seed)
#pragma unroll
            for (int c = 0; c < CD; ++c) {
                float t = 0.f;
                for (int w = 0; w < IC_THREADS / 32; ++w) t += s_red[w][c];
                dcol[(long)sidx * CD + c] += t;
            }
        }
    }
    __shared__ float red2[32];
    dls = g2_block_sum(dls, red2);
    if (tid == 0) dlog_sigma_b[b] = dls;
}

// ------------------------------------------------------------------------------------------ masked pooling
// num[k,b,c] = sum_p exp(log_m[k,b,p]) f[b,p,c];  msum[k,b] = sum_p exp(log_m[k,b,p])
// (reference models/genesisv2_config.py:147-152; the feature head is evaluated once, not K times).
// grid (chunks, B); block = C threads (C = 128); K <= 16.
constexpr int MP_MAXK = 16;
constexpr int MP_CHUNK = 64;
__global__ void masked_pool_fwd_kernel(const float* __restrict__ f, const float* __restrict__ log_m, float* __restrict__ num,
                                       float* __restrict__ msum, int B, int P, int C, int K, int pix_per_block) {
    const int b = blockIdx.y, c = threadIdx.x;
    const int p0 = blockIdx.x * pix_per_block, p1 = min(P, p0 + pix_per_block);
    __shared__ float sm[MP_MAXK][MP_CHUNK];
    float acc[MP_MAXK], ms[MP_MAXK];
#pragma unroll
    for (int k = 0; k < MP_MAXK; ++k) { acc[k] = 0.f; ms[k] = 0.f; }
    for (int pb = p0; pb < p1; pb += MP_CHUNK) {
        const int np = min(MP_CHUNK, p1 - pb);
        __syncthreads();
        for (int i = threadIdx.x; i < K * MP_CHUNK; i += blockDim.x) {
            const int k = i / MP_CHUNK, j = i - k * MP_CHUNK;
            sm[k][j] = j < np ? expf(__ldg(log_m + ((long)k * B + b) * P + pb + j)) : 0.f;
        }
        __syncthreads();
        for (int j = 0; j < np; ++j) {
            const float fv = __ldg(f + ((long)b * P + pb + j) * C + c);
#pragma unroll
            for (int k = 0; k < MP_MAXK; ++k)
                if (k < K) acc[k] = fmaf(sm[k][j], fv, acc[k]);
        }
        if (c < K) for (int j = 0; j < np; ++j) ms[0] += sm[c][j];
    }
    for (int k = 0; k < K; ++k) atomicAdd(num + ((long)k * B + b) * C + c, acc[k]);
    if (c < K) atomicAdd(msum + (long)c * B + b, ms[0]);
}

// df[b,p,c] = sum_k m_kp dnum[k,b,c];   dlog_m[k,b,p] = m_kp (sum_c f[b,p,c] dnum[k,b,c] + dmsum[k,b])
// grid (chunks, B); block = 256 threads = 8 warps.  A warp works on FOUR pixels at a time: lane = (pixel lane/8, channel
// group lane%8), each lane owns C/8 channels as float4s, so every lane has C/32 independent 16-byte loads in flight (the
// first version -- one pixel per warp, 178 registers, one block per SM -- was latency-bound at 2 ms for B=128).
// KMAX bounds the per-lane arrays (m, dot); CPL = C/32 float4s per lane.
template <int KMAX, int CPL>
__global__ void __launch_bounds__(256) masked_pool_bwd_kernel(const float* __restrict__ f, const float* __restrict__ log_m,
                                                              const float* __restrict__ dnum, const float* __restrict__ dmsum,
                                                              float* __restrict__ df, float* __restrict__ dlog_m, int B, int P, int C,
                                                              int K, int pix_per_block) {
    extern __shared__ float sd[];          // dnum [K][C]
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
        const int k = i / C, c = i - k * C;
        sd[i] = __ldg(dnum + ((long)k * B + b) * C + c);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane >> 3, cg = lane & 7;
    const int c0 = cg * 4;                 // float4 j of this lane covers channels c0 + 32*j .. +3
    const int p0 = blockIdx.x * pix_per_block, p1 = min(P, p0 + pix_per_block);
    for (int pb = p0 + warp * 4; pb < p1; pb += 32) {
        const int p = pb + sub;
        const bool ok = p < p1;
        const int pc = ok ? p : p1 - 1;
        float4 fv[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) fv[j] = g2_ldg4(f + ((long)b * P + pc) * C + c0 + 32 * j);
        float m[KMAX], dot[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) { m[k] = k < K ? expf(__ldg(log_m + ((long)k * B + b) * P + pc)) : 0.f; dot[k] = 0.f; }
        float4 g[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) {
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    const float4 d = *reinterpret_cast<const float4*>(sd + k * C + c0 + 32 * j);
                    g[j].x = fmaf(m[k], d.x, g[j].x); g[j].y = fmaf(m[k], d.y, g[j].y);
                    g[j].z = fmaf(m[k], d.z, g[j].z); g[j].w = fmaf(m[k], d.w, g[j].w);
                    dot[k] = fmaf(fv[j].x, d.x, fmaf(fv[j].y, d.y, fmaf(fv[j].z, d.z, fmaf(fv[j].w, d.w, dot[k]))));
                }
            }
        if (ok) {
#pragma unroll
            for (int j = 0; j < CPL; ++j) *reinterpret_cast<float4*>(df + ((long)b * P + p) * C + c0 + 32 * j) = g[j];
        }
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) {
                float t = dot[k];
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                t += __shfl_xor_sync(0xffffffffu, t, 4);
                if (ok && cg == 0) dlog_m[((long)k * B + b) * P + p] = m[k] * (t + __ldg(dmsum + (long)k * B + b));
            }
    }
}

}  // namespace

extern "C" {

int g2_resample_f32(const float* x, float* y, long N, int Ho, int Wo, int C, int mode, cudaStream_t stream) {
    G2_CHECK_ARG(x && y && N > 0 && Ho > 0 && Wo > 0 && C >= 4 && (C % 4) == 0 && mode >= 0 && mode <= 3);
    if (mode == 1 || mode == 2) G2_CHECK_ARG((Ho % 2) == 0 && (Wo % 2) == 0);
    const long quads = N * Ho * Wo * (C / 4);
    if (quads < (1L << 31)) resample_kernel<unsigned><<<ew_blocks(quads), 256, 0, stream>>>(x, y, (unsigned)quads, Ho, Wo, C, mode);
    else resample_kernel<long><<<ew_blocks(quads), 256, 0, stream>>>(x, y, quads, Ho, Wo, C, mode);
    G2_LAUNCH_RET();
}

int g2_icsbp_fwd_f32(const float* colour, const float* u, const float* log_sigma, float* log_m, float* log_s, int* seed_idx,
                     int B, int P, int K, int colour_dim, cudaStream_t stream) {
    G2_CHECK_ARG(colour && u && log_sigma && log_m && log_s && seed_idx && B > 0 && K >= 2);
    G2_CHECK_ARG(colour_dim == CD && P > 0 && P <= IC_THREADS * IC_MAXPPT);
    icsbp_fwd_kernel<0, false><<<B, IC_THREADS, 0, stream>>>(colour, u, log_sigma, log_m, log_s, seed_idx, nullptr, B, P, K);
    G2_LAUNCH_RET();
}

// Same contract with the kernel of InstanceColouringSBP selectable: 0 gaussian, 1 laplacian, 2 epanechnikov.
int g2_icsbp_kernel_fwd_f32(const float* colour, const float* u, const float* log_sigma, float* log_m, float* log_s, int* seed_idx,
                            int B, int P, int K, int colour_dim, int kernel_type, cudaStream_t stream) {
    G2_CHECK_ARG(colour && u && log_sigma && log_m && log_s && seed_idx && B > 0 && K >= 2);
    G2_CHECK_ARG(colour_dim == CD && P > 0 && P <= IC_THREADS * IC_MAXPPT && kernel_type >= 0 && kernel_type <= 2);
    if (kernel_type == 0) icsbp_fwd_kernel<0, false><<<B, IC_THREADS, 0, stream>>>(colour, u, log_sigma, log_m, log_s, seed_idx, nullptr, B, P, K);
    else if (kernel_type == 1) icsbp_fwd_kernel<1, false><<<B, IC_THREADS, 0, stream>>>(colour, u, log_sigma, log_m, log_s, seed_idx, nullptr, B, P, K);
    else icsbp_fwd_kernel<2, false><<<B, IC_THREADS, 0, stream>>>(colour, u, log_sigma, log_m, log_s, seed_idx, nullptr, B, P, K);
    G2_LAUNCH_RET();
}

int g2_icsbp_bwd_f32(const float* colour, const float* log_sigma, const int* seed_idx, const float* dlog_m, float* dcolour,
                     float* dlog_sigma_b, int B, int P, int K, int colour_dim, cudaStream_t stream) {
    G2_CHECK_ARG(colour && log_sigma && seed_idx && dlog_m && dcolour && dlog_sigma_b && B > 0 && K >= 2);
    G2_CHECK_ARG(colour_dim == CD && P > 0 && P <= IC_THREADS * IC_MAXPPT);
    icsbp_bwd_kernel<0><<<B, IC_THREADS, 0, stream>>>(colour, log_sigma, seed_idx, dlog_m, dcolour, dlog_sigma_b, nullptr, B, P, K);
    G2_LAUNCH_RET();
}

int g2_icsbp_kernel_bwd_f32(const float* colour, const float* log_sigma, const int* seed_idx, const float* dlog_m, float* dcolour,
                            float* dlog_sigma_b, int B, int P, int K, int colour_dim, int kernel_type, cudaStream_t stream) {
    G2_CHECK_ARG(colour && log_sigma && seed_idx && dlog_m && dcolour && dlog_sigma_b && B > 0 && K >= 2);
    G2_CHECK_ARG(colour_dim == CD && P > 0 && P <= IC_THREADS * IC_MAXPPT && kernel_type >= 0 && kernel_type <= 2);
    if (kernel_type == 0) icsbp_bwd_kernel<0><<<B, IC_THREADS, 0, stream>>>(colour, log_sigma, seed_idx, dlog_m, dcolour, dlog_sigma_b, nullptr, B, P, K);
    else if (kernel_type == 1) icsbp_bwd_kernel<1><<<B, IC_THREADS, 0, stream>>>(colour, log_sigma, seed_idx, dlog_m, dcolour, dlog_sigma_b, nullptr, B, P, K);
    else icsbp_bwd_kernel<2><<<B, IC_THREADS, 0, stream>>>(colour, log_sigma, seed_idx, dlog_m, dcolour, dlog_sigma_b, nullptr, B, P, K);
    G2_LAUNCH_RET();
}

// dynamic_K variant (models/genesisv2_config.py:118-137, modules/attention.py:168-169, 218-219): as g2_icsbp_kernel_*_f32 plus
// n_masks [B] (int32): masks produced for each image; log_m slots k >= n_masks[b] hold -1e10, seed_idx entries beyond the
// last step hold -1.
int g2_icsbp_dynamic_fwd_f32(const float* colour, const float* u, const float* log_sigma, float* log_m, float* log_s, int* seed_idx,
                             int* n_masks, int B, int P, int K, int colour_dim, int kernel_type, cudaStream_t stream) {
    G2_CHECK_ARG(colour && u && log_sigma && log_m && log_s && seed_idx && n_masks && B > 0 && K >= 2);
    G2_CHECK_ARG(colour_dim == CD && P > 0 && P <= IC_THREADS * IC_MAXPPT && kernel_type >= 0 && kernel_type <= 2);
    if (kernel_type == 0) icsbp_fwd_kernel<0, true><<<B, IC_THREADS, 0, stream>>>(colour, u, log_sigma, log_m, log_s, seed_idx, n_masks, B, P, K);
    else if (kernel_type == 1) icsbp_fwd_kernel<1, true><<<B, IC_THREADS, 0, stream>>>(colour, u, log_sigma, log_m, log_s, seed_idx, n_masks, B, P, K);
    else icsbp_fwd_kernel<2, true><<<B, IC_THREADS, 0, stream>>>(colour, u, log_sigma, log_m, log_s, seed_idx, n_masks, B, P, K);
    G2_LAUNCH_RET();
}

int g2_icsbp_dynamic_bwd_f32(const float* colour, const float* log_sigma, const int* seed_idx, const int* n_masks, const float* dlog_m,
                             float* dcolour, float* dlog_sigma_b, int B, int P, int K, int colour_dim, int kernel_type,
                             cudaStream_t stream) {
    G2_CHECK_ARG(colour && log_sigma && seed_idx && n_masks && dlog_m && dcolour && dlog_sigma_b && B > 0 && K >= 2);
    G2_CHECK_ARG(colour_dim == CD && P > 0 && P <= IC_THREADS * IC_MAXPPT && kernel_type >= 0 && kernel_type <= 2);
    if (kernel_type == 0) icsbp_bwd_kernel<0><<<B, IC_THREADS, 0, stream>>>(colour, log_sigma, seed_idx, dlog_m, dcolour, dlog_sigma_b, n_masks, B, P, K);
    else if (kernel_type == 1) icsbp_bwd_kernel<1><<<B, IC_THREADS, 0, stream>>>(colour, log_sigma, seed_idx, dlog_m, dcolour, dlog_sigma_b, n_masks, B, P, K);
    else icsbp_bwd_kernel<2><<<B, IC_THREADS, 0, stream>>>(colour, log_sigma, seed_idx, dlog_m, dcolour, dlog_sigma_b, n_masks, B, P, K);
    G2_LAUNCH_RET();
}

int g2_masked_pool_fwd_f32(const float* f, const float* log_m, float* num, float* msum, int B, int P, int C, int K, cudaStream_t stream) {
    G2_CHECK_ARG(f && log_m && num && msum && B > 0 && P > 0 && C >= 32 && C <= 1024 && (C % 32) == 0 && K >= 1 && K <= MP_MAXK);
    cudaError_t e = cudaMemsetAsync(num, 0, sizeof(float) * (size_t)K * B * C, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(msum, 0, sizeof(float) * (size_t)K * B, stream);
    if (e != cudaSuccess) return (int)e;
    int ppb = 512;
    while ((long)g2_cdiv(P, ppb) * B < 296 && ppb > MP_CHUNK) ppb /= 2;
    dim3 grid(g2_cdiv(P, ppb), B);
    masked_pool_fwd_kernel<<<grid, C, 0, stream>>>(f, log_m, num, msum, B, P, C, K, ppb);
    G2_LAUNCH_RET();
}

int g2_masked_pool_bwd_f32(const float* f, const float* log_m, const float* dnum, const float* dmsum, float* df, float* dlog_m,
                           int B, int P, int C, int K, cudaStream_t stream) {
    G2_CHECK_ARG(f && log_m && dnum && dmsum && df && dlog_m && B > 0 && P > 0 && C >= 32 && (C % 32) == 0 && K >= 1 && K <= MP_MAXK);
    G2_CHECK_ARG((size_t)K * C * sizeof(float) <= 48 * 1024);
    G2_CHECK_ARG((reinterpret_cast<uintptr_t>(f) & 15) == 0 && (reinterpret_cast<uintptr_t>(df) & 15) == 0);
    int ppb = 256;
    while ((long)g2_cdiv(P, ppb) * B < 296 && ppb > 32) ppb /= 2;
    dim3 grid(g2_cdiv(P, ppb), B);
    const size_t smem = (size_t)K * C * sizeof(float);
#define MP_LAUNCH(KM, CP) masked_pool_bwd_kernel<KM, CP><<<grid, 256, smem, stream>>>(f, log_m, dnum, dmsum, df, dlog_m, B, P, C, K, ppb)
    const int cpl = C / 32;
    if (cpl == 4) { if (K <= 8) MP_LAUNCH(8, 4); else MP_LAUNCH(16, 4); }
    else if (cpl == 2) { if (K <= 8) MP_LAUNCH(8, 2); else MP_LAUNCH(16, 2); }
    else if (cpl == 1) { if (K <= 8) MP_LAUNCH(8, 1); else MP_LAUNCH(16, 1); }
    else if (cpl == 8) { if (K <= 8) MP_LAUNCH(8, 8); else MP_LAUNCH(16, 8); }
    else return G2_ERR_UNSUPPORTED;
#undef MP_LAUNCH
    G2_LAUNCH_RET();
}

}  // extern "C"
