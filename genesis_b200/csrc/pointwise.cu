// genesis_b200 -- small HBM-bound kernels: layout changes, stick-breaking scans, component-VAE input
// packing, broadcast-add, activation gradients and simple reductions.  fp32, 128-bit accesses where the
// shapes allow.
#include "common.cuh"

namespace {

inline int ew_blocks(long work, int threads = 256) {
    long b = (work + threads - 1) / threads;
    const long cap = 148L * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// x [N,C,P] -> y [N,P,C]   (dir 0)   or   x [N,P,C] -> y [N,C,P]   (dir 1)
// Index type I: unsigned when the element count fits 31 bits (every shape of the BASELINE configs), else long -- the 64-bit
// divisions of the index decomposition cost more than the memory access they address.
template <typename I>
__global__ void layout_kernel(const float* __restrict__ x, float* __restrict__ y, I total, int C, int P, int dir) {
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        // i indexes the OUTPUT (coalesced writes)
        if (dir == 0) {
            const I c = i % (I)C, t = i / (I)C, p = t % (I)P, n = t / (I)P;
            y[i] = __ldg(x + ((long)n * C + c) * P + p);
        } else {
            const I p = i % (I)P, t = i / (I)P, c = t % (I)C, n = t / (I)C;
            y[i] = __ldg(x + ((long)n * P + p) * C + c);
        }
    }
}

// Stick-breaking scan over K slots (reference modules/attention.py:40-50,114-130 and the fix-up at
// models/genesis_config.py:169-171).  logits [nl, BP] with nl >= K-1 (GENESIS decodes nl = K logit maps,
// the last one is unused for the masks); log_m [K, BP]; log_s [nl+1, BP].
//   log_s_0 = 0;  log_m_k = log_s_k + logsig(a_k), log_s_{k+1} = log_s_k + logsig(-a_k)  (k < K-1)
//   log_m_{K-1} = log_s_{K-1}
__global__ void sbp_scan_fwd_kernel(const float* __restrict__ logits, float* __restrict__ log_m, float* __restrict__ log_s,
                                    long BP4, int K, int nl) {
    const long BP = BP4 * 4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < BP4; i += (long)gridDim.x * blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(log_s + i * 4) = s;
        for (int k = 0; k < nl; ++k) {
            const float4 a = g2_ldg4(logits + k * BP + i * 4);
            if (k < K - 1) {
                float4 m;
                m.x = s.x + g2_logsigmoid(a.x); m.y = s.y + g2_logsigmoid(a.y);
                m.z = s.z + g2_logsigmoid(a.z); m.w = s.w + g2_logsigmoid(a.w);
                *reinterpret_cast<float4*>(log_m + k * BP + i * 4) = m;
            } else if (k == K - 1) {
                *reinterpret_cast<float4*>(log_m + k * BP + i * 4) = s;
            }
            s.x += g2_logsigmoid(-a.x); s.y += g2_logsigmoid(-a.y);
            s.z += g2_logsigmoid(-a.z); s.w += g2_logsigmoid(-a.w);
            *reinterpret_cast<float4*>(log_s + (k + 1) * BP + i * 4) = s;
        }
        if (nl == K - 1) *reinterpret_cast<float4*>(log_m + (long)(K - 1) * BP + i * 4) = s;
    }
}

// d logits from d log_m (log_s is a non-differentiated statistic).  Reverse scan:
//   R_{K-1} = G_{K-1};  for k = K-2..0:  da_k = G_k * sig(-a_k) - R_{k+1} * sig(a_k);  R_k = G_k + R_{k+1}
__global__ void sbp_scan_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ dlog_m,
                                    float* __restrict__ dlogits, long BP, int K, int nl) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < BP; i += (long)gridDim.x * blockDim.x) {
        float R = __ldg(dlog_m + (long)(K - 1) * BP + i);
        for (int k = nl - 1; k >= K - 1; --k) dlogits[k * BP + i] = 0.f;
        for (int k = K - 2; k >= 0; --k) {
            const float a = __ldg(logits + k * BP + i);
            const float G = __ldg(dlog_m + k * BP + i);
            const float sg = 1.f / (1.f + expf(-a));
            dlogits[k * BP + i] = G * (1.f - sg) - R * sg;
            R += G;
        }
    }
}

// component-VAE encoder input (reference modules/component_vae.py:58-63): for slot k, image b:
// out[(k*B+b), p, 0] = log_m[k,b,p];  out[(k*B+b), p, 1..3] = x[b, 0..2, p]      (x NCHW, out NHWC4)
// `Cp` (multiple of 4) output channels: channels >= 4 are zero padding so the encoder's first conv fits the
// 32-channel k-blocks of the tensor-core kernels.
template <typename I>
__global__ void comp_pack_kernel(const float* __restrict__ x, const float* __restrict__ log_m, float* __restrict__ out,
                                 I total, int B, int P, int Cp) {
    const I q = (I)(Cp >> 2);
    for (I j = (I)blockIdx.x * blockDim.x + threadIdx.x; j < total * q; j += (I)gridDim.x * blockDim.x) {
        const I quad = j % q;
        const I i = j / q;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (quad == 0) {
            const I p = i % (I)P, kb = i / (I)P, b = kb % (I)B;
            o.x = __ldg(log_m + i);
            const float* xb = x + (long)b * 3 * P + p;
            o.y = __ldg(xb); o.z = __ldg(xb + P); o.w = __ldg(xb + 2 * P);
        }
        *reinterpret_cast<float4*>(out + (long)j * 4) = o;
    }
}

// x [N,C,P] (NCHW) -> y [N,P,Cp] (NHWC, channels >= C zero)
template <typename I>
__global__ void nhwc_pad_kernel(const float* __restrict__ x, float* __restrict__ y, I total, int C, int P, int Cp) {
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        const I c = i % (I)Cp, t = i / (I)Cp, p = t % (I)P, n = t / (I)P;
        y[i] = c < (I)C ? __ldg(x + ((long)n * C + c) * P + p) : 0.f;
    }
}

// out[n,p,c] = act(a[n,c] + m[p,c])      (broadcast decoder first layer, see decoder ops)
template <typename I>
__global__ void bcast_add_act_kernel(const float* __restrict__ a, const float* __restrict__ m, float* __restrict__ out,
                                     I total_quads, int P, int C, int act) {
    const I q = (I)(C >> 2);
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total_quads; i += (I)gridDim.x * blockDim.x) {
        const I quad = i % q, row = i / q, p = row % (I)P, n = row / (I)P;
        const float4 av = g2_ldg4(a + (long)n * C + quad * 4), mv = g2_ldg4(m + (long)p * C + quad * 4);
        float4 o;
        o.x = g2_apply_act(av.x + mv.x, act, 0.f); o.y = g2_apply_act(av.y + mv.y, act, 0.f);
        o.z = g2_apply_act(av.z + mv.z, act, 0.f); o.w = g2_apply_act(av.w + mv.w, act, 0.f);
        *reinterpret_cast<float4*>(out + (long)i * 4) = o;
    }
}

// dpre = dout * act'(out)   (act given as G2_ACT_MUL_*_GRAD), elementwise
__global__ void act_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* __restrict__ dpre,
                               long total4, int act) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
        const float4 d = g2_ldg4(dout + i * 4), o = g2_ldg4(out + i * 4);
        float4 r;
        r.x = g2_apply_act(d.x, act, o.x); r.y = g2_apply_act(d.y, act, o.y);
        r.z = g2_apply_act(d.z, act, o.z); r.w = g2_apply_act(d.w, act, o.w);
        *reinterpret_cast<float4*>(dpre + i * 4) = r;
    }
}

// dpre = dout * act'(out) and, in the same pass, dbias[c] += sum over rows of dpre[row][c]  (rows x C, C % 4 == 0): the
// activation backward of a conv + bias + activation layer fused with its bias gradient (no separate column-sum pass over
// dpre).  A thread owns one channel quad and strides over rows, 4 rows in flight; block partials -> float atomics.
__global__ void __launch_bounds__(256) act_bwd_bias_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                           float* __restrict__ dpre, float* __restrict__ dbias, long M, int C,
                                                           long rows_per_block, int act) {
    const int q = C >> 2, lanes = 256 / q;
    const int t = threadIdx.x, lane = t / q, quad = t - lane * q;
    const long r0 = (long)blockIdx.x * rows_per_block;
    const long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < lanes) {
        for (long r = r0 + lane; r < r1; r += 4L * lanes) {
            float4 d[4], o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long rr = r + (long)u * lanes;
                if (rr < r1) { d[u] = g2_ldg4(dout + rr * C + quad * 4); o[u] = g2_ldg4(out + rr * C + quad * 4); }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long rr = r + (long)u * lanes;
                if (rr >= r1) break;
                float4 v;
                v.x = g2_apply_act(d[u].x, act, o[u].x); v.y = g2_apply_act(d[u].y, act, o[u].y);
                v.z = g2_apply_act(d[u].z, act, o[u].z); v.w = g2_apply_act(d[u].w, act, o[u].w);
                *reinterpret_cast<float4*>(dpre + rr * C + quad * 4) = v;
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        }
    }
    __shared__ float4 sm[256];
    sm[t] = s;
    __syncthreads();
    if (t < q) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < lanes; ++l) { const float4 v = sm[l * q + t]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
        float* o = dbias + t * 4;
        atomicAdd(o, a.x); atomicAdd(o + 1, a.y); atomicAdd(o + 2, a.z); atomicAdd(o + 3, a.w);
    }
}

// out[n, c] = sum_p x[n, p, c]      grid (chunks, N), atomicAdd into pre-zeroed out
__global__ void __launch_bounds__(256) seg_colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int P, int C,
                                                         int rows_per_block) {
    const int q = C >> 2, lanes = 256 / q;
    const int t = threadIdx.x, lane = t / q, quad = t - lane * q, n = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(P, r0 + rows_per_block);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < lanes)
        for (int r = r0 + lane; r < r1; r += lanes) {
            const float4 v = g2_ldg4(x + ((long)n * P + r) * C + quad * 4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    __shared__ float4 sm[256];
    sm[t] = s;
    __syncthreads();
    if (t < q) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < lanes; ++l) { const float4 v = sm[l * q + t]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
        float* o = out + (long)n * C + t * 4;
        atomicAdd(o, a.x); atomicAdd(o + 1, a.y); atomicAdd(o + 2, a.z); atomicAdd(o + 3, a.w);
    }
}

// out[j] = sum_n x[n, j]     (j < J), coalesced over j
// Sixteen rows are in flight per thread (the grid is only J/4 threads wide: without them the loads of a thread serialise on the add
// and the kernel ran at 2.4 TB/s); rows are added in order, so the result does not depend on the launch geometry.
__global__ void sum_dim0_kernel(const float* __restrict__ x, float* __restrict__ out, int N, long J4) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < J4; i += (long)gridDim.x * blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* base = x + i * 4;
        int n = 0;
        for (; n + 16 <= N; n += 16) {
            float4 v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = g2_ldg4(base + (long)(n + u) * J4 * 4);
#pragma unroll
            for (int u = 0; u < 16; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
        }
        for (; n < N; ++n) {
            const float4 v = g2_ldg4(base + (long)n * J4 * 4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        *reinterpret_cast<float4*>(out + i * 4) = s;
    }
}

}  // namespace

extern "C" {

int g2_layout_f32(const float* x, float* y, long N, int C, int P, int to_nchw, cudaStream_t stream) {
    G2_CHECK_ARG(x && y && N > 0 && C > 0 && P > 0);
    const long total = N * C * P;
    if (total < (1L << 31)) layout_kernel<unsigned><<<ew_blocks(total), 256, 0, stream>>>(x, y, (unsigned)total, C, P, to_nchw ? 1 : 0);
    else layout_kernel<long><<<ew_blocks(total), 256, 0, stream>>>(x, y, total, C, P, to_nchw ? 1 : 0);
    G2_LAUNCH_RET();
}

int g2_sbp_scan_fwd_f32(const float* logits, float* log_m, float* log_s, long BP, int K, int nl, cudaStream_t stream) {
    G2_CHECK_ARG(logits && log_m && log_s && BP > 0 && (BP % 4) == 0 && K >= 2 && (nl == K || nl == K - 1));
    sbp_scan_fwd_kernel<<<ew_blocks(BP / 4), 256, 0, stream>>>(logits, log_m, log_s, BP / 4, K, nl);
    G2_LAUNCH_RET();
}

int g2_sbp_scan_bwd_f32(const float* logits, const float* dlog_m, float* dlogits, long BP, int K, int nl, cudaStream_t stream) {
    G2_CHECK_ARG(logits && dlog_m && dlogits && BP > 0 && K >= 2 && (nl == K || nl == K - 1));
    sbp_scan_bwd_kernel<<<ew_blocks(BP), 256, 0, stream>>>(logits, dlog_m, dlogits, BP, K, nl);
    G2_LAUNCH_RET();
}

int g2_comp_pack_f32(const float* x, const float* log_m, float* out, int K, int B, int P, int Cp, cudaStream_t stream) {
    G2_CHECK_ARG(x && log_m && out && K > 0 && B > 0 && P > 0 && Cp >= 4 && (Cp % 4) == 0);
    const long total = (long)K * B * P;
    if (total * (Cp / 4) < (1L << 31)) comp_pack_kernel<unsigned><<<ew_blocks(total * (Cp / 4)), 256, 0, stream>>>(x, log_m, out, (unsigned)total, B, P, Cp);
    else comp_pack_kernel<long><<<ew_blocks(total * (Cp / 4)), 256, 0, stream>>>(x, log_m, out, total, B, P, Cp);
    G2_LAUNCH_RET();
}

int g2_nhwc_pad_f32(const float* x, float* y, long N, int C, int P, int Cp, cudaStream_t stream) {
    G2_CHECK_ARG(x && y && N > 0 && C > 0 && P > 0 && Cp >= C);
    const long total = N * P * Cp;
    if (total < (1L << 31)) nhwc_pad_kernel<unsigned><<<ew_blocks(total), 256, 0, stream>>>(x, y, (unsigned)total, C, P, Cp);
    else nhwc_pad_kernel<long><<<ew_blocks(total), 256, 0, stream>>>(x, y, total, C, P, Cp);
    G2_LAUNCH_RET();
}

int g2_bcast_add_act_f32(const float* a, const float* m, float* out, long N, int P, int C, int act, cudaStream_t stream) {
    G2_CHECK_ARG(a && m && out && N > 0 && P > 0 && C >= 4 && (C % 4) == 0);
    const long quads = N * P * (C / 4);
    if (quads < (1L << 31)) bcast_add_act_kernel<unsigned><<<ew_blocks(quads), 256, 0, stream>>>(a, m, out, (unsigned)quads, P, C, act);
    else bcast_add_act_kernel<long><<<ew_blocks(quads), 256, 0, stream>>>(a, m, out, quads, P, C, act);
    G2_LAUNCH_RET();
}

int g2_act_bwd_f32(const float* dout, const float* out, float* dpre, long total, int act, cudaStream_t stream) {
    G2_CHECK_ARG(dout && out && dpre && total > 0 && (total % 4) == 0);
    G2_CHECK_ARG(act == G2_ACT_RELU || act == G2_ACT_ELU);
    const int code = act == G2_ACT_RELU ? G2_ACT_MUL_RELU_GRAD : G2_ACT_MUL_ELU_GRAD;
    act_bwd_kernel<<<ew_blocks(total / 4), 256, 0, stream>>>(dout, out, dpre, total / 4, code);
    G2_LAUNCH_RET();
}

// dpre = dout * act'(out);  dbias[c] += column sums of dpre (accumulated into a caller-initialised buffer)
int g2_act_bwd_bias_f32(const float* dout, const float* out, float* dpre, float* dbias, long M, int C, int act,
                        cudaStream_t stream) {
    G2_CHECK_ARG(dout && out && dpre && dbias && M > 0 && C >= 4 && (C % 4) == 0 && C <= 1024);
    G2_CHECK_ARG(act == G2_ACT_RELU || act == G2_ACT_ELU);
    const int code = act == G2_ACT_RELU ? G2_ACT_MUL_RELU_GRAD : G2_ACT_MUL_ELU_GRAD;
    const int lanes = 256 / (C / 4);
    long rpb = (long)lanes * 8;
    while (g2_cdiv(M, rpb) > 148L * 24) rpb *= 2;
    act_bwd_bias_kernel<<<g2_cdiv(M, rpb), 256, 0, stream>>>(dout, out, dpre, dbias, M, C, rpb, code);
    G2_LAUNCH_RET();
}

int g2_seg_colsum_f32(const float* x, float* out, int N, int P, int C, cudaStream_t stream) {
    G2_CHECK_ARG(x && out && N > 0 && P > 0 && C >= 4 && (C % 4) == 0 && C <= 1024);
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N * C, stream);
    if (e != cudaSuccess) return (int)e;
    const int lanes = 256 / (C / 4);
    int rpb = lanes * 32;
    while ((long)g2_cdiv(P, rpb) * N > 148L * 32 && rpb < P) rpb *= 2;
    dim3 grid(g2_cdiv(P, rpb), N);
    seg_colsum_kernel<<<grid, 256, 0, stream>>>(x, out, P, C, rpb);
    G2_LAUNCH_RET();
}

int g2_sum_dim0_f32(const float* x, float* out, int N, long J, cudaStream_t stream) {
    G2_CHECK_ARG(x && out && N > 0 && J > 0 && (J % 4) == 0);
    sum_dim0_kernel<<<ew_blocks(J / 4, 128), 128, 0, stream>>>(x, out, N, J / 4);
    G2_LAUNCH_RET();
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ weight packs
// One pass over a torch-layout conv weight writes both tensor-core operand packs:
//   packA[tap][co][ci_pad] (reduction over Ci: conv forward / conv-transpose forward)
//   packB[tap][ci_pad][co] (reduction over Co: the matching data gradient); channels ci >= Ci are zero.
// transposed = 0: w is Conv2d [Co,Ci,R,S]; 1: ConvTranspose2d [Ci,Co,R,S].
namespace {
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ pa, float* __restrict__ pb, int Co, int Ci,
                                        int Cip, int RS, int transposed) {
    const long total = (long)RS * Co * Cip;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cip); const long t = i / Cip; const int co = (int)(t % Co); const int tap = (int)(t / Co);
        float v = 0.f;
        if (ci < Ci) v = __ldg(w + (transposed ? ((long)ci * Co + co) : ((long)co * Ci + ci)) * RS + tap);
        pa[i] = v;
        if (pb) pb[((long)tap * Cip + ci) * Co + co] = v;
    }
}
}  // namespace

extern "C" int g2_pack_conv_weight_f32(const float* w, float* packA, float* packB, int Co, int Ci, int Ci_pad, int RS,
                                       int transposed, cudaStream_t stream) {
    G2_CHECK_ARG(w && packA && Co > 0 && Ci > 0 && Ci_pad >= Ci && RS > 0);
    const long total = (long)RS * Co * Ci_pad;
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 8) blocks = 148L * 8;
    pack_conv_weight_kernel<<<(int)blocks, 256, 0, stream>>>(w, packA, packB, Co, Ci, Ci_pad, RS, transposed);
    G2_LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------ optimiser
// Fused Adam over a flat fp32 arena (reference train.py:175,263 uses torch.optim.Adam: lr, betas (0.9, 0.999),
// eps 1e-8, no weight decay).  `step` is a device counter (float) so the kernel is CUDA-graph replayable;
// grads are multiplied by grad_scale (1/world_size after an NCCL sum) and zeroed for the next step.
namespace {
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long n4, float lr, float b1, float b2, float eps, const float* __restrict__ step, float gscale,
                            int zero_grad) {
    const float t = __ldg(step);
    const float c1 = 1.f - powf(b1, t), c2 = 1.f - powf(b2, t);
    const float step_size = lr / c1, inv_c2 = rsqrtf(c2);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 pp = *reinterpret_cast<float4*>(p + i * 4), gg = *reinterpret_cast<float4*>(g + i * 4);
        float4 mm = *reinterpret_cast<float4*>(m + i * 4), vv = *reinterpret_cast<float4*>(v + i * 4);
        float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = ga[j] * gscale;
            ma[j] = b1 * ma[j] + (1.f - b1) * gr;
            va[j] = b2 * va[j] + (1.f - b2) * gr * gr;
            pa[j] -= step_size * ma[j] / (sqrtf(va[j]) * inv_c2 + eps);
        }
        *reinterpret_cast<float4*>(p + i * 4) = pp;
        *reinterpret_cast<float4*>(m + i * 4) = mm;
        *reinterpret_cast<float4*>(v + i * 4) = vv;
        if (zero_grad) *reinterpret_cast<float4*>(g + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
}  // namespace

extern "C" int g2_adam_f32(float* p, float* g, float* m, float* v, long n, float lr, float b1, float b2, float eps,
                           const float* step, float grad_scale, int zero_grad, cudaStream_t stream) {
    G2_CHECK_ARG(p && g && m && v && step && n > 0 && (n % 4) == 0);
    long b = (n / 4 + 255) / 256;
    if (b > 148L * 16) b = 148L * 16;
    adam_kernel<<<(int)b, 256, 0, stream>>>(p, g, m, v, n / 4, lr, b1, b2, eps, step, grad_scale, zero_grad);
    G2_LAUNCH_RET();
}

// ---- fused RMSprop / SGD over the flat arena: the other two optimisers train.py offers (train.py:171-176:
// optim.RMSprop(params, lr) -> alpha 0.99, eps 1e-8, no momentum, not centred; optim.SGD(params, lr, 0.9) -> momentum 0.9,
// no dampening, no Nesterov, the first step initialises the buffer with the gradient).  Same conventions as adam_kernel:
// gradients are scaled by grad_scale (1 / world after an NCCL sum) and zeroed for the next step; `step` is the 1-based
// device counter (SGD needs it to recognise the first step).
namespace {
__global__ void rmsprop_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ sq, long n4, float lr, float alpha,
                               float eps, float gscale, int zero_grad) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 pp = *reinterpret_cast<float4*>(p + i * 4), gg = *reinterpret_cast<float4*>(g + i * 4);
        float4 ss = *reinterpret_cast<float4*>(sq + i * 4);
        float* pa = &pp.x; float* ga = &gg.x; float* sa = &ss.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = ga[j] * gscale;
            sa[j] = alpha * sa[j] + (1.f - alpha) * gr * gr;
            pa[j] -= lr * gr / (sqrtf(sa[j]) + eps);
        }
        *reinterpret_cast<float4*>(p + i * 4) = pp;
        *reinterpret_cast<float4*>(sq + i * 4) = ss;
        if (zero_grad) *reinterpret_cast<float4*>(g + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__global__ void sgd_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ buf, long n4, float lr, float momentum,
                           const float* __restrict__ step, float gscale, int zero_grad) {
    const bool first = __ldg(step) <= 1.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 pp = *reinterpret_cast<float4*>(p + i * 4), gg = *reinterpret_cast<float4*>(g + i * 4);
        float4 bb = *reinterpret_cast<float4*>(buf + i * 4);
        float* pa = &pp.x; float* ga = &gg.x; float* ba = &bb.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = ga[j] * gscale;
            ba[j] = first ? gr : momentum * ba[j] + gr;
            pa[j] -= lr * ba[j];
        }
        *reinterpret_cast<float4*>(p + i * 4) = pp;
        *reinterpret_cast<float4*>(buf + i * 4) = bb;
        if (zero_grad) *reinterpret_cast<float4*>(g + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// GECO (utils/geco.py:35-51) as ONE single-thread kernel on device scalars -- no host sync, graph-replayable:
//   err_ema = started ? (1-alpha) err + alpha err_ema : err;  constraint = goal - err_ema;
//   beta = clamp(beta * exp(rate * constraint), beta_min, beta_max),  rate = step_size * (constraint > 0 ? speedup : 1).
// state = {beta, err_ema, started}; also steps the optimiser's step counter (+1) and emits elbo = err + kl.
__global__ void geco_kernel(float* __restrict__ state, const float* __restrict__ err_kl, float* __restrict__ step_count,
                            float* __restrict__ elbo, float inv_world, float goal, float step_size, float alpha, float speedup,
                            float beta_min, float beta_max, int update) {
    const float err = err_kl[0] * inv_world, kl = err_kl[1] * inv_world;
    if (update) {
        const float ema = state[2] > 0.f ? (1.f - alpha) * err + alpha * state[1] : err;
        state[1] = ema;
        state[2] = 1.f;
        const float c = goal - ema;
        const float rate = (speedup > 0.f && c > 0.f) ? speedup * step_size : step_size;
        state[0] = fminf(fmaxf(state[0] * expf(rate * c), beta_min), beta_max);
    }
    if (step_count) step_count[0] += 1.f;
    if (elbo) elbo[0] = err + kl;
}
}  // namespace

extern "C" int g2_rmsprop_f32(float* p, float* g, float* sq, long n, float lr, float alpha, float eps, float grad_scale,
                              int zero_grad, cudaStream_t stream) {
    G2_CHECK_ARG(p && g && sq && n > 0 && (n % 4) == 0);
    long b = (n / 4 + 255) / 256;
    if (b > 148L * 16) b = 148L * 16;
    rmsprop_kernel<<<(int)b, 256, 0, stream>>>(p, g, sq, n / 4, lr, alpha, eps, grad_scale, zero_grad);
    G2_LAUNCH_RET();
}

extern "C" int g2_sgd_f32(float* p, float* g, float* buf, long n, float lr, float momentum, const float* step, float grad_scale,
                          int zero_grad, cudaStream_t stream) {
    G2_CHECK_ARG(p && g && buf && step && n > 0 && (n % 4) == 0);
    long b = (n / 4 + 255) / 256;
    if (b > 148L * 16) b = 148L * 16;
    sgd_kernel<<<(int)b, 256, 0, stream>>>(p, g, buf, n / 4, lr, momentum, step, grad_scale, zero_grad);
    G2_LAUNCH_RET();
}

extern "C" int g2_geco_step_f32(float* state, const float* err_kl, float* step_count, float* elbo, float inv_world, float goal,
                                float step_size, float alpha, float speedup, float beta_min, float beta_max, int update,
                                cudaStream_t stream) {
    G2_CHECK_ARG(state && err_kl);
    geco_kernel<<<1, 1, 0, stream>>>(state, err_kl, step_count, elbo, inv_world, goal, step_size, alpha, speedup, beta_min, beta_max,
                                     update);
    G2_LAUNCH_RET();
}

// ---- 3xTF32 operand split ("tf32x3": a * b ~= a_hi b_hi + a_hi b_lo + a_lo b_hi, error ~2^-21 instead of 2^-11).
// hi = x rounded to TF32 (10-bit mantissa; exactly representable, so the TMA / tensor-core operand rounding leaves it
// unchanged), lo = x - hi (exact in fp32; its own TF32 rounding is a 2^-22 relative error of x).  The three products run as ONE
// tensor-core contraction over a 3x longer reduction: the activation is written as [hi | hi | lo] channel blocks and the weight as
// [w_hi | w_lo | w_hi], so no kernel needs an accumulate mode.  mode 0: out [rows, 3C] = [hi | hi | lo];  1: out [rows, 2C] =
// [hi | lo] (weight-gradient operands: all four blocks come out, three are summed);  2: out [rows, C] = hi;  3: out = lo.
namespace {
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, float* __restrict__ out, long rows, int C,
                                                         int mode) {
    const int q = C >> 2;
    const long total = rows * q;
    const int oc = mode == 0 ? 3 * C : mode == 1 ? 2 * C : C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / q;
        const int c = (int)(i - r * q) * 4;
        const float4 v = g2_ldg4(x + r * C + c);
        const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        float* o = out + r * oc + c;
        if (mode == 0) {
            *reinterpret_cast<float4*>(o) = h; *reinterpret_cast<float4*>(o + C) = h; *reinterpret_cast<float4*>(o + 2 * C) = l;
        } else if (mode == 1) {
            *reinterpret_cast<float4*>(o) = h; *reinterpret_cast<float4*>(o + C) = l;
        } else {
            *reinterpret_cast<float4*>(o) = mode == 2 ? h : l;
        }
    }
}
}  // namespace

extern "C" int g2_split_tf32_f32(const float* x, float* out, long rows, int C, int mode, cudaStream_t stream) {
    G2_CHECK_ARG(x && out && rows > 0 && C >= 4 && (C % 4) == 0 && mode >= 0 && mode <= 3);
    long b = (rows * (C / 4) + 255) / 256;
    if (b > 148L * 16) b = 148L * 16;
    split_tf32_kernel<<<(int)b, 256, 0, stream>>>(x, out, rows, C, mode);
    G2_LAUNCH_RET();
}
