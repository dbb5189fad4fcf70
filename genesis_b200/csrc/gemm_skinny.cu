// genesis_b200 -- exact-fp32 GEMM for the small, latency-bound products of the latent path (LSTM steps, latent heads, prior
// MLP: 64-320 rows, a few to ~150 MFLOP; reference modules/attention.py:94-103, models/genesis_config.py:230-239, 301-314,
// modules/encoders.py:35-37).  EXPERIMENTAL (ops: G2_SKINNY_GEMM=1; off by default, not yet run on a B200).
//
//   C[M,N] (+)= op(A)[M,K] op(B)[K,N] + bias[N], then the activation           (same contract as g2_gemm_f32)
//
// Why another GEMM: for ~20 MFLOP the tensor-core tile kernel is all prologue (TMEM allocation, barrier setup, tensor-map
// fetch, one 128-row tile on a handful of SMs: ~10 us) and needs its operands K-major, which costs the backward pass three
// transpose copies and an extra add per linear layer; the 64x64-tile SIMT kernel covers M = 64 with N/64 CTAs and no
// prefetch (20-70 us).  Here: 32x32 output tiles (M = 64, N = 512 -> 32 CTAs; a weight gradient of 1024 x 320 -> 320 CTAs),
// BK-deep K tiles fetched into registers while the previous tile is multiplied out of shared memory, operands read in place
// in either orientation, bias / activation / accumulation fused, one writer per element (no atomics, no workspace:
// deterministic and bit-reproducible).
#include "common.cuh"

namespace {

struct SkP {
    const float* A; const float* B; const float* bias; float* C;
    int M, N, K, lda, ldb, ldc, tA, tB, act, accumulate;
};

constexpr int SK_BM = 32, SK_BN = 32, SK_THREADS = 128;

// element (m, k) of op(A), (k, n) of op(B); zero outside the problem
__device__ __forceinline__ float sk_load_a(const SkP& p, int m, int k) {
    if (m >= p.M || k >= p.K) return 0.f;
    return __ldg(p.tA ? p.A + (long)k * p.lda + m : p.A + (long)m * p.lda + k);
}
__device__ __forceinline__ float sk_load_b(const SkP& p, int k, int n) {
    if (n >= p.N || k >= p.K) return 0.f;
    return __ldg(p.tB ? p.B + (long)n * p.ldb + k : p.B + (long)k * p.ldb + n);
}

template <int BK>
__global__ void __launch_bounds__(SK_THREADS) skinny_gemm_kernel(const SkP p) {
    constexpr int PER_A = SK_BM * BK / SK_THREADS, PER_B = SK_BN * BK / SK_THREADS;      // elements per thread and K tile
    __shared__ float As[BK][SK_BM + 1];                 // [k][m]; +1: conflict-free transposing stores
    __shared__ __align__(16) float Bs[BK][SK_BN + 4];   // [k][n]; rows 16-byte aligned for float4 reads
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;       // 8 x 4 columns, 16 x 2 rows
    const int m0 = blockIdx.x * SK_BM, n0 = blockIdx.y * SK_BN;
    float ra[PER_A], rb[PER_B];
    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // the index that is contiguous in memory runs fastest over the threads (coalesced loads)
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < PER_A; ++i) {
            const int idx = tid + i * SK_THREADS;
            const int m = p.tA ? idx % SK_BM : idx / BK, k = p.tA ? idx / SK_BM : idx % BK;
            ra[i] = sk_load_a(p, m0 + m, k0 + k);
        }
#pragma unroll
        for (int i = 0; i < PER_B; ++i) {
            const int idx = tid + i * SK_THREADS;
            const int n = p.tB ? idx / BK : idx % SK_BN, k = p.tB ? idx % BK : idx / SK_BN;
            rb[i] = sk_load_b(p, k0 + k, n0 + n);
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < PER_A; ++i) {
            const int idx = tid + i * SK_THREADS;
            const int m = p.tA ? idx % SK_BM : idx / BK, k = p.tA ? idx / SK_BM : idx % BK;
            As[k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < PER_B; ++i) {
            const int idx = tid + i * SK_THREADS;
            const int n = p.tB ? idx / BK : idx % SK_BN, k = p.tB ? idx % BK : idx / SK_BN;
            Bs[k][n] = rb[i];
        }
    };

    fetch(0);
    for (int k0 = 0; k0 < p.K; k0 += BK) {
        __syncthreads();                    // the previous tile has been consumed
        stash();
        __syncthreads();
        if (k0 + BK < p.K) fetch(k0 + BK);  // next tile in flight while this one is multiplied
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1];
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
            acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
            acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
            acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty * 2 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            if (p.bias) v += __ldg(p.bias + n);
            v = g2_apply_act(v, p.act, 0.f);
            float* c = p.C + (long)m * p.ldc + n;
            *c = p.accumulate ? *c + v : v;
        }
    }
}

}  // namespace

extern "C" {

// Same contract as g2_gemm_f32 (row-major; transA: A is stored [K,M]; transB: B is stored [N,K]); with `accumulate` the
// result is added to C by the single thread that owns the element (no atomics), so the caller must order it after other
// writers of C on the stream, as for any other kernel.
int g2_gemm_skinny_f32(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int lda, int ldb, int ldc,
                       int transA, int transB, int act, int accumulate, cudaStream_t stream) {
    G2_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0 && lda > 0 && ldb > 0 && ldc >= N);
    G2_CHECK_ARG(act == G2_ACT_NONE || act == G2_ACT_RELU || act == G2_ACT_ELU || act == G2_ACT_SIGMOID);
    SkP p{A, B, bias, C, M, N, K, lda, ldb, ldc, transA ? 1 : 0, transB ? 1 : 0, act, accumulate ? 1 : 0};
    dim3 grid(g2_cdiv(M, SK_BM), g2_cdiv(N, SK_BN));
    if (K >= 128) skinny_gemm_kernel<64><<<grid, SK_THREADS, 0, stream>>>(p);
    else skinny_gemm_kernel<32><<<grid, SK_THREADS, 0, stream>>>(p);
    G2_LAUNCH_RET();
}

}  // extern "C"
