// genesis_b200 -- shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define G2_OK 0
#define G2_ERR_ARG (-1)        // bad argument (shape / alignment / null pointer)
#define G2_ERR_UNSUPPORTED (-2)

#define G2_CHECK_ARG(cond) do { if (!(cond)) return G2_ERR_ARG; } while (0)
#define G2_LAUNCH_RET() do { cudaError_t e__ = cudaGetLastError(); return e__ == cudaSuccess ? G2_OK : (int)e__; } while (0)

// epilogue / activation codes shared by conv, gemm and decoder kernels
enum {
    G2_ACT_NONE = 0,
    G2_ACT_RELU = 1,
    G2_ACT_ELU = 2,
    G2_ACT_MUL_RELU_GRAD = 3,   // out = acc * d relu(aux)/d pre,  aux = saved post-activation
    G2_ACT_MUL_ELU_GRAD = 4,    // out = acc * d elu(aux)/d pre
    G2_ACT_SIGMOID = 5,
};

// norm modes
enum { G2_NORM_NONE = 0, G2_NORM_BATCH = 1, G2_NORM_INSTANCE = 2, G2_NORM_GROUP = 3 };
// norm post-ops
enum { G2_POST_GATE = 0, G2_POST_RELU = 1, G2_POST_NONE = 2 };

static inline int g2_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: a process-wide `static bool` would leave every
// device after the first without the opt-in (one process driving several GPUs).  One bit per device ordinal in a 64-bit mask
// (set-attribute is idempotent, so a race between two threads is benign).  `want` lets a call site raise the size later.
#include <atomic>
struct G2DevOnce {
    std::atomic<unsigned long long> mask{0};
    std::atomic<int> size{0};
    bool needed(int want = 1) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (want > size.load(std::memory_order_acquire)) { size.store(want, std::memory_order_release); mask.store(0, std::memory_order_release); }
        return (mask.load(std::memory_order_acquire) & (1ull << (dev & 63))) == 0;
    }
    void done() {
        int dev = 0;
        cudaGetDevice(&dev);
        mask.fetch_or(1ull << (dev & 63), std::memory_order_acq_rel);
    }
};

__device__ __forceinline__ float g2_sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float g2_apply_act(float v, int act, float aux) {
    switch (act) {
        case G2_ACT_RELU: return v > 0.f ? v : 0.f;
        case G2_ACT_ELU: return v > 0.f ? v : expm1f(v);
        case G2_ACT_MUL_RELU_GRAD: return aux > 0.f ? v : 0.f;
        case G2_ACT_MUL_ELU_GRAD: return aux > 0.f ? v : v * (aux + 1.f);
        case G2_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

// numerically stable log(sigmoid(x)) matching torch's logsigmoid: min(x,0) - log1p(exp(-|x|))
__device__ __forceinline__ float g2_logsigmoid(float x) {
    return fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
}

__device__ __forceinline__ float g2_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double g2_warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; result valid in thread 0 (and all threads of warp 0). `red` holds >= 32 floats.
__device__ __forceinline__ float g2_block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = g2_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (w == 0) v = g2_warp_sum(v);
    return v;
}

__device__ __forceinline__ float4 g2_ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
