// genesis_b200 -- TF32 tensor-core implicit GEMM for sm_100a: tcgen05.mma with TMEM accumulators, operands
// staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle), mbarrier producer/consumer ring.
//
// One kernel covers conv forward, conv data-gradient, conv-transpose forward, conv-transpose data-gradient
// (strides 1 and 2) and plain GEMMs.  The host turns each of them into the same problem:
//
//     out[pix(n,h',w')][co] = bias[co] + sum_{tap} sum_{c} X_plane(tap)[n, h'+dh(tap), w'+dw(tap), c] * W[widx(tap)][co][c]
//
// over a "virtual" output grid (h',w') that is tiled in 128-pixel tiles = the M dimension of one UMMA
// (M=128, N=BN=Cout tile, K=8 tf32 per instruction, 4 instructions per 32-channel k-block).
//   * stride-1 conv / conv-transpose: one plane (the input), (dh,dw) = +-(r-P, s-P); zero padding comes from
//     TMA out-of-bounds fill.
//   * stride-2 conv forward (and conv-transpose data-gradient): the input is viewed as 4 parity planes
//     (tensor maps with doubled strides), tap (r,s) reads plane ((r-P)&1,(s-P)&1) at offset floor((r-P)/2).
//   * stride-2 conv-transpose forward (and conv data-gradient): 4 launches, one per output parity class
//     (sub-pixel decomposition); each class uses only its valid taps and writes out[2h'+ph, 2w'+pw].
//   * VALID convs on awkward widths use a flat pixel run per image (tile = 128 consecutive pixels,
//     tap offset r*Wi+s) so no 2-D tile quantisation is paid.
// A (pixels x 32 channels) and B (Cout x 32 channels) tiles are K-major, 128 B per row, SWIZZLE_128B.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue
// (tcgen05.ld -> bias/activation -> global).
#include "umma.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>

namespace tc {

constexpr int BM = 128;
constexpr int A_BYTES = BM * 128;
constexpr int MAX_TAPS = 32;

struct Taps { short plane[MAX_TAPS], dh[MAX_TAPS], dw[MAX_TAPS], widx[MAX_TAPS]; };

struct Maps { CUtensorMap a[4]; CUtensorMap b; };

struct P {
    float* out; const float* bias;
    int N, Hv, Wv, TH, TW, TNB, tiles_h, tiles_w;
    int Ho, Wo, Co, os, ph, pw, flat_wi;
    int ntaps, cblocks, kb_total, kb_per_split, act, atomic;
    long split_stride;      // split-K: floats between the partial outputs of consecutive splits (workspace mode)
    Taps taps;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must not hang the GPU -- trap after ~2 s instead.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) break;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// K-major, SWIZZLE_128B operand tile (rows of 128 B, 8-row groups 1024 B apart); see cute/arch/mma_sm100_desc.hpp
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int ACT> __device__ __forceinline__ float act_apply(float v) {
    if (ACT == G2_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == G2_ACT_ELU) return v > 0.f ? v : __expf(v) - 1.f;     // TF32 path: |error| < 2e-7 absolute
    if (ACT == G2_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}

// One epilogue warp: its 32 TMEM lanes (one output pixel each, `pix` = flat output pixel or -1) -> +bias -> act -> global.
template <int BN, int ACT>
__device__ __forceinline__ void store_tile(float* out, int Co, int n0c, uint32_t tmem_base, const float* sBias, float* stage, long pix) {
    const int lane = threadIdx.x & 31, q = (threadIdx.x >> 5) & 3;
    const int sub = lane >> 3, chunk = lane & 7;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = *reinterpret_cast<const float4*>(sBias + c0 + 4 * j);
            float4 o;
            o.x = act_apply<ACT>(__uint_as_float(v[4 * j]) + b.x);
            o.y = act_apply<ACT>(__uint_as_float(v[4 * j + 1]) + b.y);
            o.z = act_apply<ACT>(__uint_as_float(v[4 * j + 2]) + b.z);
            o.w = act_apply<ACT>(__uint_as_float(v[4 * j + 3]) + b.w);
            *reinterpret_cast<float4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) = o;
        }
        __syncwarp();
#pragma unroll
        for (int r4 = 0; r4 < 8; ++r4) {
            const int row = r4 * 4 + sub;
            const long rp = __shfl_sync(0xffffffffu, pix, row);
            const float4 o = *reinterpret_cast<const float4*>(stage + row * 32 + ((chunk ^ (row & 7)) << 2));
            if (rp >= 0) *reinterpret_cast<float4*>(out + rp * Co + n0c + c0 + chunk * 4) = o;
        }
        __syncwarp();
    }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192) conv_tc_kernel(const __grid_constant__ Maps maps, const __grid_constant__ P p) {
    constexpr int B_BYTES = BN * 128;
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sA = sm;
    uint8_t* sB = sm + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* accf = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accf + 1);
    float* sBias = reinterpret_cast<float*>(full) + 32;           // 128 B past the barrier block

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile coordinates
    const int t = blockIdx.x;
    const int tw_i = t % p.tiles_w;
    const int th_i = (t / p.tiles_w) % p.tiles_h;
    const int n = (t / (p.tiles_w * p.tiles_h)) * p.TNB;      // first image of the tile (TNB images per tile)
    const int h0 = th_i * p.TH, w0 = tw_i * p.TW;
    const int n0c = blockIdx.y * BN;
    const int kb0 = blockIdx.z * p.kb_per_split;
    const int nkb = min(p.kb_per_split, p.kb_total - kb0);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a[0]);
        prefetch_tmap(&maps.b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
            mbar_init(accf, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + BN) sBias[threadIdx.x - 64] = p.bias ? p.bias[n0c + threadIdx.x - 64] : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // whole warp walks the schedule (uniform control flow); one elected lane issues the TMA loads
        for (int i = 0; i < nkb; ++i) {
            const int kb = kb0 + i;
            const int s = i % STAGES;
            const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            if (umma::elect_one()) {
                mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
                tma_load_4d(sA + s * A_BYTES, &maps.a[p.taps.plane[tap]], &full[s], cb * 32, w0 + p.taps.dw[tap],
                            h0 + p.taps.dh[tap], n);
                tma_load_3d(sB + s * B_BYTES, &maps.b, &full[s], cb * 32, n0c, p.taps.widx[tap]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % STAGES;
            const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
            mbar_wait(&full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t adesc = make_desc_k_sw128(smem_u32(sA + s * A_BYTES));
            const uint64_t bdesc = make_desc_k_sw128(smem_u32(sB + s * B_BYTES));
            if (umma::elect_one()) {
                umma_tf32(tmem_base, adesc, bdesc, idesc, i > 0 ? 1u : 0u);      // 4 x (K = 8 tf32 = 32 B) per 128-byte row
                umma_tf32(tmem_base, adesc + 2, bdesc + 2, idesc, 1u);
                umma_tf32(tmem_base, adesc + 4, bdesc + 4, idesc, 1u);
                umma_tf32(tmem_base, adesc + 6, bdesc + 6, idesc, 1u);
                umma_commit(&empty[s]);          // frees the smem stage when these MMAs retire
            }
            __syncwarp();
        }
        if (umma::elect_one()) umma_commit(accf);    // accumulator complete
        __syncwarp();
    } else {
        // ---- epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31
        const int q = warp & 3;
        const int row = q * 32 + lane;
        mbar_wait(accf, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int per_img = p.TH * p.TW;
        const int tn = row / per_img, rem = row - tn * per_img;
        const int th = rem / p.TW, tw = rem - th * p.TW;
        const int hv = h0 + th, wv = w0 + tw;
        bool valid = hv < p.Hv && wv < p.Wv && (n + tn) < p.N;
        int oh, ow;
        if (p.flat_wi > 0) { oh = wv / p.flat_wi; ow = wv - oh * p.flat_wi; }
        else { oh = hv * p.os + p.ph; ow = wv * p.os + p.pw; }
        valid = valid && oh < p.Ho && ow < p.Wo;
        const long pix = valid ? ((long)(n + tn) * p.Ho + oh) * p.Wo + ow : -1;
        {
            // split-K: split z writes its partial tile to its own slab of the workspace (plain stores, no bias); the ordered
            // reduction kernel adds the slabs in split order, so the result is bitwise reproducible (no float atomics)
            float* const outz = p.out + (long)blockIdx.z * p.split_stride;
            // all MMAs have retired, so stage 0 of the A ring is free: transpose each 32 x 128-byte block through a swizzled
            // 4 KB staging tile so that every store instruction writes four complete 128-byte rows
            float* stage = reinterpret_cast<float*>(sA) + q * 1024;
            switch (p.act) {
                case G2_ACT_RELU: store_tile<BN, G2_ACT_RELU>(outz, p.Co, n0c, tmem_base, sBias, stage, pix); break;
                case G2_ACT_ELU: store_tile<BN, G2_ACT_ELU>(outz, p.Co, n0c, tmem_base, sBias, stage, pix); break;
                case G2_ACT_SIGMOID: store_tile<BN, G2_ACT_SIGMOID>(outz, p.Co, n0c, tmem_base, sBias, stage, pix); break;
                default: store_tile<BN, G2_ACT_NONE>(outz, p.Co, n0c, tmem_base, sBias, stage, pix); break;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled get_encode() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
    });
    return fn;
}

// rank-r fp32 (loaded as tf32, round-to-nearest) tensor map, 128B swizzle, zero OOB fill.
// dims / strides innermost first; strides in BYTES for dims 1..r-1.
bool encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
            const cuuint32_t* box) {
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BN, int STAGES>
int launch(const Maps& maps, const P& p, dim3 grid, cudaStream_t stream) {
    constexpr int smem = STAGES * (A_BYTES + BN * 128) + 128 + BN * 4 + 1024;
    static G2DevOnce once;
    if (once.needed()) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        once.done();
    }
    conv_tc_kernel<BN, STAGES><<<grid, 192, smem, stream>>>(maps, p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? G2_OK : (int)e;
}

int dispatch(const Maps& maps, const P& p, int BN, dim3 grid, cudaStream_t stream) {
    switch (BN) {
        case 32: return launch<32, 4>(maps, p, grid, stream);
        case 64: return launch<64, 4>(maps, p, grid, stream);
        case 128: return launch<128, 3>(maps, p, grid, stream);
        default: return G2_ERR_UNSUPPORTED;
    }
}

int pick_bn(int Co) {
    if (Co == 32 || Co == 64 || Co == 128) return Co;
    if (Co % 128 == 0) return 128;
    if (Co % 64 == 0) return 64;
    return 0;
}

inline int floordiv2(int t) { return (t - (t & 1)) / 2; }

// 128-pixel tile of the virtual output grid: full rows for power-of-two widths, 8 x 16 otherwise (edges masked);
// maps smaller than 128 pixels (power-of-two sides) put TNB whole images into one tile.
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
inline void pick_tile(int Hv, int Wv, int* TH, int* TW, int* TNB) {
    *TNB = 1;
    if ((long)Hv * Wv < 128 && is_pow2(Hv) && is_pow2(Wv)) { *TW = Wv; *TH = Hv; *TNB = 128 / (Hv * Wv); }
    else if (is_pow2(Wv) && Wv >= 8) { *TW = Wv < 128 ? Wv : 128; *TH = 128 / *TW; }
    else { *TW = 16; *TH = 8; }
}

}  // namespace tc

// Ordered split-K reduction: C = act(bias + slab_0 + slab_1 + ...), float4 per thread; fixed order -> bitwise reproducible.
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, const float* __restrict__ bias, float* __restrict__ C,
                                     long n4, int N4, int splits, int act) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const int c4 = (int)(i % N4) * 4;
    float4 a = bias ? make_float4(bias[c4], bias[c4 + 1], bias[c4 + 2], bias[c4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {
        const float4 v = reinterpret_cast<const float4*>(ws)[(long)z * n4 + i];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    if (act == G2_ACT_RELU) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
    else if (act == G2_ACT_ELU) {
        a.x = a.x > 0.f ? a.x : __expf(a.x) - 1.f; a.y = a.y > 0.f ? a.y : __expf(a.y) - 1.f;
        a.z = a.z > 0.f ? a.z : __expf(a.z) - 1.f; a.w = a.w > 0.f ? a.w : __expf(a.w) - 1.f;
    }
    reinterpret_cast<float4*>(C)[i] = a;
}

extern "C" {

// halo kernel (igemm_halo.cu): stride-1 problems and stride-2 conv-transpose classes with the activation window resident
int g2_conv_halo_supported(int N, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode);
int g2_conv_halo_tf32(const float* in, const float* w, const float* bias, float* out, int N, int Hi, int Wi, int Ci,
                      int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode, int act, cudaStream_t stream);

// 1 if g2_conv_igemm_tf32 handles this problem, else 0 (the caller then uses the exact fp32 kernel).
int g2_conv_tf32_supported(int N, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode) {
    if (Ci % 32 != 0 || tc::pick_bn(Co) == 0) return 0;
    if (R * S > tc::MAX_TAPS || stride < 1 || stride > 2) return 0;
    if (stride == 2 && mode == 0 && ((Hi | Wi) & 1)) return 0;
    if (stride == 2 && mode == 1 && ((Ho | Wo) & 1)) return 0;
    const int Hv = (mode == 1 && stride == 2) ? Ho / 2 : Ho, Wv = (mode == 1 && stride == 2) ? Wo / 2 : Wo;
    if ((long)Hv * Wv < 128 && !(tc::is_pow2(Hv) && tc::is_pow2(Wv) && Hv * Wv >= 16)) return 0;
    const bool flat = (mode == 0 && stride == 1 && pad == 0 && (Wv & (Wv - 1)) != 0);
    if (flat && (long)(R - 1) * Wi + S - 1 > 32000) return 0;
    return 1;
}

// Same contract as g2_conv_igemm_f32 (mode 0 / 1) with TF32 operands and fp32 accumulation, except:
// weights are packed [R*S][Co][Ci] (K-major rows), there is no aux / fused activation-gradient epilogue.
int g2_conv_igemm_tf32(const float* in, const float* w, const float* bias, float* out, int N, int Hi, int Wi, int Ci,
                       int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode, int act, cudaStream_t stream) {
    using namespace tc;
    G2_CHECK_ARG(in && w && out && N > 0);
    if (!g2_conv_tf32_supported(N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode)) return G2_ERR_UNSUPPORTED;
    G2_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (g2_conv_halo_supported(N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode)) {
        const int rc = g2_conv_halo_tf32(in, w, bias, out, N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode, act, stream);
        if (rc != G2_ERR_UNSUPPORTED) return rc;
    }
    const int BN = pick_bn(Co);
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    // weights: {Ci, Co, taps}
    {
        cuuint64_t dims[3] = {(cuuint64_t)Ci, (cuuint64_t)Co, (cuuint64_t)(R * S)};
        cuuint64_t str[2] = {(cuuint64_t)Ci * 4, (cuuint64_t)Ci * Co * 4};
        cuuint32_t box[3] = {32, (uint32_t)BN, 1};
        if (!encode(&maps.b, w, 3, dims, str, box)) return G2_ERR_UNSUPPORTED;
    }
    const int nclass = (mode == 1 && stride == 2) ? 4 : 1;
    for (int cls = 0; cls < nclass; ++cls) {
        P p;
        memset(&p, 0, sizeof(p));
        p.out = out; p.bias = bias; p.N = N; p.Ho = Ho; p.Wo = Wo; p.Co = Co; p.act = act; p.atomic = 0;
        p.os = 1; p.ph = 0; p.pw = 0; p.flat_wi = 0;
        p.cblocks = Ci / 32;
        int nt = 0;
        bool flat = false;
        if (mode == 0 && stride == 1) {
            p.Hv = Ho; p.Wv = Wo;
            flat = (pad == 0 && (Wo & (Wo - 1)) != 0);
            for (int r = 0; r < R; ++r)
                for (int s = 0; s < S; ++s) {
                    p.taps.plane[nt] = 0; p.taps.widx[nt] = (short)(r * S + s);
                    if (flat) { p.taps.dh[nt] = 0; p.taps.dw[nt] = (short)(r * Wi + s); }
                    else { p.taps.dh[nt] = (short)(r - pad); p.taps.dw[nt] = (short)(s - pad); }
                    ++nt;
                }
        } else if (mode == 0 && stride == 2) {
            p.Hv = Ho; p.Wv = Wo;
            for (int r = 0; r < R; ++r)
                for (int s = 0; s < S; ++s) {
                    const int tr = r - pad, ts = s - pad;
                    p.taps.plane[nt] = (short)(((tr & 1) << 1) | (ts & 1));
                    p.taps.dh[nt] = (short)floordiv2(tr); p.taps.dw[nt] = (short)floordiv2(ts);
                    p.taps.widx[nt] = (short)(r * S + s);
                    ++nt;
                }
        } else if (mode == 1 && stride == 1) {
            p.Hv = Ho; p.Wv = Wo;
            for (int r = 0; r < R; ++r)
                for (int s = 0; s < S; ++s) {
                    p.taps.plane[nt] = 0; p.taps.dh[nt] = (short)(pad - r); p.taps.dw[nt] = (short)(pad - s);
                    p.taps.widx[nt] = (short)(r * S + s);
                    ++nt;
                }
        } else {   // mode 1, stride 2: output parity class (ph,pw)
            p.Hv = Ho / 2; p.Wv = Wo / 2; p.os = 2; p.ph = cls >> 1; p.pw = cls & 1;
            for (int r = 0; r < R; ++r)
                for (int s = 0; s < S; ++s) {
                    const int th = p.ph + pad - r, tw = p.pw + pad - s;
                    if ((th & 1) || (tw & 1)) continue;
                    p.taps.plane[nt] = 0; p.taps.dh[nt] = (short)(th / 2); p.taps.dw[nt] = (short)(tw / 2);
                    p.taps.widx[nt] = (short)(r * S + s);
                    ++nt;
                }
            if (nt == 0) continue;   // (cannot happen for R,S >= 2)
        }
        p.ntaps = nt;
        p.kb_total = nt * p.cblocks;
        p.kb_per_split = p.kb_total;
        if (cls == 0) {
            // activation maps (built once; identical for the 4 classes of mode 1 / stride 2)
            if (flat) {
                cuuint64_t dims[4] = {(cuuint64_t)Ci, (cuuint64_t)Hi * Wi, 1, (cuuint64_t)N};
                cuuint64_t str[3] = {(cuuint64_t)Ci * 4, (cuuint64_t)Hi * Wi * Ci * 4, (cuuint64_t)Hi * Wi * Ci * 4};
                cuuint32_t box[4] = {32, 128, 1, 1};
                if (!encode(&maps.a[0], in, 4, dims, str, box)) return G2_ERR_UNSUPPORTED;
            } else if (mode == 0 && stride == 2) {
                int TW, TH, TNB;
                pick_tile(p.Hv, p.Wv, &TH, &TW, &TNB);
                for (int pl = 0; pl < 4; ++pl) {
                    const int pr = pl >> 1, ps = pl & 1;
                    cuuint64_t dims[4] = {(cuuint64_t)Ci, (cuuint64_t)Wi / 2, (cuuint64_t)Hi / 2, (cuuint64_t)N};
                    cuuint64_t str[3] = {(cuuint64_t)2 * Ci * 4, (cuuint64_t)2 * Wi * Ci * 4, (cuuint64_t)Hi * Wi * Ci * 4};
                    cuuint32_t box[4] = {32, (uint32_t)TW, (uint32_t)TH, (uint32_t)TNB};
                    if (!encode(&maps.a[pl], in + ((long)pr * Wi + ps) * Ci, 4, dims, str, box)) return G2_ERR_UNSUPPORTED;
                }
            } else {
                int TW, TH, TNB;
                pick_tile(p.Hv, p.Wv, &TH, &TW, &TNB);
                cuuint64_t dims[4] = {(cuuint64_t)Ci, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)N};
                cuuint64_t str[3] = {(cuuint64_t)Ci * 4, (cuuint64_t)Wi * Ci * 4, (cuuint64_t)Hi * Wi * Ci * 4};
                cuuint32_t box[4] = {32, (uint32_t)TW, (uint32_t)TH, (uint32_t)TNB};
                if (!encode(&maps.a[0], in, 4, dims, str, box)) return G2_ERR_UNSUPPORTED;
            }
        }
        if (flat) {
            p.flat_wi = Wi; p.Hv = 1; p.Wv = (Ho - 1) * Wi + Wo;      // last valid flat position + 1
            p.TH = 1; p.TW = 128; p.TNB = 1;
        } else {
            pick_tile(p.Hv, p.Wv, &p.TH, &p.TW, &p.TNB);
        }
        p.tiles_w = g2_cdiv(p.Wv, p.TW); p.tiles_h = g2_cdiv(p.Hv, p.TH);
        dim3 grid((unsigned)((long)g2_cdiv(N, p.TNB) * p.tiles_h * p.tiles_w), (unsigned)(Co / BN), 1);
        const int rc = dispatch(maps, p, BN, grid, stream);
        if (rc != G2_OK) return rc;
    }
    return G2_OK;
}

// C[M,N] = act(A[M,K] * W[N,K]^T + bias[N])   (TF32 operands, fp32 accumulate).
// Split-K (few output tiles, long reduction) is deterministic: every split writes its partial product to its own slab of
// the caller's workspace and splitk_reduce_kernel adds the slabs in split order, with bias and activation fused.
static int gemm_splits(int M, int N, int K, int BN, int* kb_per_split) {
    const int kb_total = K / 32;
    const int tiles = g2_cdiv(M, 128) * (N / BN);
    int splits = 1;
    if (tiles < 148 && kb_total >= 16) {
        splits = (296 + tiles - 1) / tiles;
        if (splits > kb_total / 4) splits = kb_total / 4;
        if (splits < 1) splits = 1;
    }
    const int per = (kb_total + splits - 1) / splits;
    if (kb_per_split) *kb_per_split = per;
    return (kb_total + per - 1) / per;
}

long g2_gemm_tf32_workspace(int M, int N, int K) {
    const int BN = tc::pick_bn(N);
    if (BN == 0 || K % 32 != 0 || M <= 0) return 0;
    const int splits = gemm_splits(M, N, K, BN, nullptr);
    return splits > 1 ? (long)sizeof(float) * splits * (long)M * (long)N : 0;
}

static int gemm_tf32_impl(const float* A, const float* W, const float* bias, float* C, float* ws, int M, int N, int K, int act,
                          cudaStream_t stream) {
    using namespace tc;
    G2_CHECK_ARG(A && W && C && M > 0 && N > 0 && K > 0);
    const int BN = pick_bn(N);
    if (BN == 0 || K % 32 != 0) return G2_ERR_UNSUPPORTED;
    G2_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0);
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    {
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, 1};
        cuuint64_t str[2] = {(cuuint64_t)K * 4, (cuuint64_t)K * N * 4};
        cuuint32_t box[3] = {32, (uint32_t)BN, 1};
        if (!encode(&maps.b, W, 3, dims, str, box)) return G2_ERR_UNSUPPORTED;
        cuuint64_t dimsa[4] = {(cuuint64_t)K, (cuuint64_t)M, 1, 1};
        cuuint64_t stra[3] = {(cuuint64_t)K * 4, (cuuint64_t)K * M * 4, (cuuint64_t)K * M * 4};
        cuuint32_t boxa[4] = {32, 128, 1, 1};
        if (!encode(&maps.a[0], A, 4, dimsa, stra, boxa)) return G2_ERR_UNSUPPORTED;
    }
    P p;
    memset(&p, 0, sizeof(p));
    p.N = 1; p.Hv = 1; p.Wv = M; p.TH = 1; p.TW = 128; p.TNB = 1; p.tiles_h = 1; p.tiles_w = g2_cdiv(M, 128);
    p.Ho = 1; p.Wo = M; p.Co = N; p.os = 1; p.ntaps = 1; p.cblocks = K / 32; p.kb_total = K / 32;
    int splits = ws ? gemm_splits(M, N, K, BN, &p.kb_per_split) : 1;
    if (splits == 1) {
        p.kb_per_split = p.kb_total;
        p.out = C; p.bias = bias; p.act = act; p.split_stride = 0;
    } else {
        p.out = ws; p.bias = nullptr; p.act = G2_ACT_NONE; p.split_stride = (long)M * N;
    }
    dim3 grid((unsigned)p.tiles_w, (unsigned)(N / BN), (unsigned)splits);
    const int rc = dispatch(maps, p, BN, grid, stream);
    if (rc != G2_OK || splits == 1) return rc;
    const long n4 = (long)M * N / 4;
    const int blocks = (int)((n4 + 255) / 256);
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(ws, bias, C, n4, N / 4, splits, act);
    G2_LAUNCH_RET();
}

int g2_gemm_tf32(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, cudaStream_t stream) {
    return gemm_tf32_impl(A, W, bias, C, nullptr, M, N, K, G2_ACT_NONE, stream);
}

int g2_gemm_tf32_ws(const float* A, const float* W, const float* bias, float* C, float* ws, int M, int N, int K, int act,
                    cudaStream_t stream) {
    return gemm_tf32_impl(A, W, bias, C, ws, M, N, K, act, stream);
}

}  // extern "C"
