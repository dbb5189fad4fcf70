// genesis_b200 -- UMMA descriptor self-test: run a few tcgen05.mma.kind::tf32 instructions on caller-provided
// shared-memory IMAGES of the A and B operands with caller-provided descriptor templates, and dump the
// 128 x N accumulator.  Used by tests/test_umma_layouts_gpu.py to pin the smem-descriptor conventions
// (K-major / MN-major, SWIZZLE_128B, base offset) the production kernels rely on.
#include "common.cuh"

namespace dbg {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) umma_probe_kernel(const uint4* __restrict__ a_img, const uint4* __restrict__ b_img,
                                                         float* __restrict__ D, int a_bytes, int b_bytes,
                                                         unsigned long long adesc_t, unsigned long long bdesc_t, uint32_t idesc,
                                                         int N, int nk, int a_kstep, int b_kstep, int a_off, int b_off,
                                                         int a_base_off_auto) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sA = sm;
    uint8_t* sB = sm + ((a_bytes + 1023) & ~1023);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + ((b_bytes + 1023) & ~1023));
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < a_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = a_img[i];
    for (int i = threadIdx.x; i < b_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = b_img[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA (async proxy)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        for (int k = 0; k < nk; ++k) {
            const uint32_t aa = smem_u32(sA) + a_off + k * a_kstep, bb = smem_u32(sB) + b_off + k * b_kstep;
            unsigned long long ad = adesc_t | (unsigned long long)((aa >> 4) & 0x3FFF);
            unsigned long long bd = bdesc_t | (unsigned long long)((bb >> 4) & 0x3FFF);
            if (a_base_off_auto) {
                ad |= (unsigned long long)((aa >> 7) & 7) << 49;
                bd |= (unsigned long long)((bb >> 7) & 7) << 49;
            }
            const uint32_t acc = k > 0 ? 1u : 0u;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    {   // everyone waits for the accumulator
        uint32_t ok = 0;
        const long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
            if (clock64() - t0 > 2000000000LL) __trap();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[(long)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}
}  // namespace dbg

extern "C" int g2_debug_umma_probe(const float* a_img, const float* b_img, float* D, int a_bytes, int b_bytes, long adesc_t,
                                   long bdesc_t, int idesc, int N, int nk, int a_kstep, int b_kstep, int a_off, int b_off,
                                   int base_off_auto, cudaStream_t stream) {
    G2_CHECK_ARG(a_img && b_img && D && a_bytes > 0 && b_bytes > 0 && (a_bytes % 16) == 0 && (b_bytes % 16) == 0);
    G2_CHECK_ARG(N >= 8 && N <= 256 && (N % 8) == 0 && a_bytes + b_bytes <= 180 * 1024);
    const int smem = ((a_bytes + 1023) & ~1023) + ((b_bytes + 1023) & ~1023) + 64 + 1024;
    cudaError_t e = cudaFuncSetAttribute(dbg::umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    dbg::umma_probe_kernel<<<1, 128, smem, stream>>>(reinterpret_cast<const uint4*>(a_img), reinterpret_cast<const uint4*>(b_img), D,
                                                     a_bytes, b_bytes, (unsigned long long)adesc_t, (unsigned long long)bdesc_t,
                                                     (uint32_t)idesc, N, nk, a_kstep, b_kstep, a_off, b_off, base_off_auto);
    G2_LAUNCH_RET();
}

// ---------------------------------------------------------------------------------------------- UMMA issue-rate probe
// How many clocks one tcgen05.mma.kind::tf32 (M = 128, N, K = 8) costs when its operands come from shared memory, as a
// function of N, of the alignment of the A descriptor's start row (the halo kernels start A at ANY 128-byte row: a filter tap is
// a row shift of the resident window) and of a second CTA on the same SM.  One thread issues `n_mma` instructions back to back
// in the production order (4 K-steps of 32 B per (tile, tap), accumulators rotating over n_acc tiles), commits and waits.
// out[cta] = clocks from the first issue to the completion of the last.  scripts/umma_rate.py prints the table
// (profiles/r02_umma_rate.txt); DESIGN.md section 3.1 uses it as the cost model of the convolution kernels.
namespace dbg {
__global__ void __launch_bounds__(128) umma_rate_kernel(long long* __restrict__ out, int N, int n_mma, int a_shift_rows, int n_acc,
                                                        int a_rows, int tmem_cols, int b_tiles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sA = sm;
    const int a_bytes = a_rows * 128, b_bytes = N * 128 * b_tiles;
    uint8_t* sB = sm + a_bytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + b_bytes);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int issuers = n_acc < 0 ? 2 : 1;              // n_acc < 0: TWO issuing threads (warps 0 and 1), |n_acc| accumulators each
    n_acc = n_acc < 0 ? -n_acc : n_acc;
    for (int i = threadIdx.x; i < (a_bytes + b_bytes) / 16; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (lane == 0 && warp < issuers) {
        unsigned long long hi = 0;
        hi |= (unsigned long long)1 << 16;
        hi |= (unsigned long long)(1024 >> 4) << 32;
        hi |= (unsigned long long)1 << 46;
        hi |= (unsigned long long)2 << 61;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        n_mma /= issuers;
        bar += warp;
        const uint32_t tmem_w = tmem + (uint32_t)(warp * n_acc * N);
        const int span = a_rows - 128 - 8;             // the start row stays inside the image
        const long long t0 = clock64();
        int row = 0, bt = 0;
        for (int i = 0; i < n_mma; i += 4) {
            const uint32_t aa = a0 + (uint32_t)row * 128u, bb = b0 + (uint32_t)bt * (uint32_t)(N * 128);
            const unsigned long long ad = hi | (unsigned long long)((aa >> 4) & 0x3FFF);
            const unsigned long long bd = hi | (unsigned long long)((bb >> 4) & 0x3FFF);
            const uint32_t d = tmem_w + (uint32_t)(((i >> 2) % n_acc) * N);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                             ::"r"(d), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(1u) : "memory");
            row += a_shift_rows;
            if (row > span) row -= span;
            if (((i >> 2) + 1) % n_acc == 0 && ++bt == b_tiles) bt = 0;       // next weight tile after a sweep over the accumulators
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
            if (clock64() - t0 > 2000000000LL) __trap();
        }
        if (warp == 0) out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}
}  // namespace dbg

/* out[ctas] = clocks for n_mma back-to-back MMAs per CTA.  ctas_per_sm in {1, 2}: grid = 148 * ctas_per_sm with the shared
 * memory sized so that exactly that many CTAs are resident per SM. */
extern "C" int g2_debug_umma_rate(int64_t* out, int N, int n_mma, int a_shift_rows, int n_acc, int a_rows, int ctas_per_sm,
                                  int b_tiles, cudaStream_t stream) {
    G2_CHECK_ARG(out && N >= 8 && N <= 256 && (N % 8) == 0 && n_mma > 0 && (n_mma % 8) == 0 && n_acc != 0 &&
                 (n_acc > 0 ? n_acc : -2 * n_acc) * N <= 256);
    G2_CHECK_ARG(a_rows >= 256 && (a_rows % 8) == 0 && (ctas_per_sm == 1 || ctas_per_sm == 2) && b_tiles >= 1);
    const int need = a_rows * 128 + N * 128 * b_tiles + 64 + 1024;
    const int smem = ctas_per_sm == 1 ? 120 * 1024 : 100 * 1024;       // 1 CTA/SM: > half of the SM's 228 KB; 2: two fit
    G2_CHECK_ARG(need <= smem);
    cudaError_t e = cudaFuncSetAttribute(dbg::umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    dbg::umma_rate_kernel<<<sms * ctas_per_sm, 128, smem, stream>>>(reinterpret_cast<long long*>(out), N, n_mma, a_shift_rows, n_acc, a_rows, 256, b_tiles);
    G2_LAUNCH_RET();
}
