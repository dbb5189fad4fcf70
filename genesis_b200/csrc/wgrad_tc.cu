// genesis_b200 -- TF32 tensor-core weight gradient for sm_100a (tcgen05.mma, MN-major operands, TMA, TMEM).
//
//   dW[tap][a][b] = sum_{n,oh,ow} G_plane(tap)[n, oh+dh(tap), ow+dw(tap), a] * T[n, oh, ow, b]
//
// The reduction (UMMA K) runs over pixels, so both operands are consumed "MN-major": a TMA box of
// [pixels][32 channels] (128-byte rows, SWIZZLE_128B_ATOM_32B) IS the canonical MN-major tile -- no transposes.
//   A (M = 128) = four stacked 32-channel blocks, each one (tap, channel-block) of the shifted G window;
//   B (N = Ct)  = the Ct/32 channel blocks of the un-shifted T tile;  K = 8 pixels per instruction.
// One CTA owns a range of 64-pixel chunks and up to 512/Ct accumulators (M-groups) in TMEM, streams the G
// blocks through a 4-stage mbarrier ring while the T tile of the chunk stays resident (double-buffered),
// and finally writes its partial dW to a workspace; a second kernel reduces the partials over CTAs
// (deterministic, no atomics) and writes dW in the requested layout.
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>
#include <mutex>

namespace wg {

constexpr int CH = 64;                 // pixels per chunk
constexpr int TILE_BYTES = CH * 128;   // one [64 pixels][32 channels] block
constexpr int STAGES = 4;
constexpr int MAX_BLK = 112;

struct Maps { CUtensorMap g[4]; CUtensorMap t; };

struct P {
    float* ws;                    // [splits][ntaps][Cg][Ct]
    int N, Ht, Wt, TH, TW, TNB, tiles_h, tiles_w, chunks_total, chunks_per_cta;
    int Cg, Ct, ntaps, nblk, groups_per_cta;
    short tap_plane[32], tap_dh[32], tap_dw[32];
    short blk_tap[MAX_BLK], blk_cb[MAX_BLK];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
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
// MN-major TF32 operands must use the SWIZZLE_128B_BASE32B layout (layout type 1; measured with
// g2_debug_umma_probe, tests/test_umma_layouts_gpu.py): 128-byte rows (32 channels of one pixel) whose 32-byte chunks
// are XORed with (row & 3) -- what TMA's SWIZZLE_128B_ATOM_32B writes; 32-element atoms along M/N are `lbo` bytes
// apart, 4-row K atoms 512 B apart.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
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

// BN = Ct (32, 64 or 128).  ONE issuer warp on purpose: this kernel is bound by its L2 -> shared-memory tile reloads, a second
// issuer (M-groups split between two warps) measured no gain (0.216 ms either way on the stride-2 5x5 layer), and the split is
// easy to get wrong -- a consumer that skips its partner's ring stages is no longer within one phase of their barriers (its
// parity waits pass early or block forever); the first version of it deadlocked in the training step once other kernels ran
// beside it, while every isolated test passed (round 2).
template <int BN>
__global__ void __launch_bounds__(192, 1) wgrad_tc_kernel(const __grid_constant__ Maps maps, const __grid_constant__ P p) {
    constexpr int TB = BN / 32;                 // T channel blocks
    constexpr int T_BYTES = TB * TILE_BYTES;
    constexpr int G_STAGE = 4 * TILE_BYTES;     // one M-group = 4 blocks
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sG = sm;
    uint8_t* sT = sm + STAGES * G_STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(sT + 2 * T_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* accf = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accf + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    const int g_begin = blockIdx.y * p.groups_per_cta;                // first M-group of this CTA
    const int ngroups_all = (p.nblk + 3) >> 2;
    const int ng = min(p.groups_per_cta, ngroups_all - g_begin);
    const int c_begin = split * p.chunks_per_cta;
    const int nchunks = max(0, min(p.chunks_per_cta, p.chunks_total - c_begin));

    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 1); }
            mbar_init(accf, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // whole warp walks the schedule (uniform control flow); one elected lane issues the TMA loads
        int it = 0;                                   // global stage counter
        for (int c = 0; c < nchunks; ++c) {
            const int chunk = c_begin + c;
            const int tw_i = chunk % p.tiles_w;
            const int th_i = (chunk / p.tiles_w) % p.tiles_h;
            const int n = (chunk / (p.tiles_w * p.tiles_h)) * p.TNB;
            const int h0 = th_i * p.TH, w0 = tw_i * p.TW;
            const int tb = c & 1;
            mbar_wait(&tempty[tb], ((uint32_t)(c >> 1) & 1u) ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(&tfull[tb], T_BYTES);
#pragma unroll
                for (int j = 0; j < TB; ++j)
                    tma_load_4d(sT + tb * T_BYTES + j * TILE_BYTES, &maps.t, &tfull[tb], j * 32, w0, h0, n);
            }
            __syncwarp();
            for (int g = 0; g < ng; ++g, ++it) {
                const int s = it % STAGES;
                mbar_wait(&empty[s], ((uint32_t)(it / STAGES) & 1u) ^ 1u);
                if (elect_one()) {
                    const int b0 = (g_begin + g) * 4;
                    const int nb = min(4, p.nblk - b0);
                    mbar_expect_tx(&full[s], nb * TILE_BYTES);
                    for (int j = 0; j < nb; ++j) {
                        const int tap = p.blk_tap[b0 + j];
                        tma_load_4d(sG + s * G_STAGE + j * TILE_BYTES, &maps.g[p.tap_plane[tap]], &full[s],
                                    p.blk_cb[b0 + j] * 32, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], n);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // a_major = b_major = MN (bits 15, 16); M = 128; N = BN; tf32 operands, fp32 accumulate
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int it = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int tb = c & 1;
            mbar_wait(&tfull[tb], (uint32_t)(c >> 1) & 1u);
            const uint64_t bdesc = make_desc_mn_sw128(smem_u32(sT + tb * T_BYTES), TILE_BYTES);
            for (int g = 0; g < ng; ++g, ++it) {
                const int s = it % STAGES;
                mbar_wait(&full[s], (uint32_t)(it / STAGES) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t adesc = make_desc_mn_sw128(smem_u32(sG + s * G_STAGE), TILE_BYTES);
                const uint32_t d = tmem_base + (uint32_t)(g * BN);
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < CH / 8; ++kk)       // 8 pixels (= 8 rows = 1024 B = 64 descriptor units) per instruction
                        umma_tf32(d, adesc + 64 * kk, bdesc + 64 * kk, idesc, (c > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&empty[s]);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&tempty[tb]);
            __syncwarp();
        }
        if (elect_one()) umma_commit(accf);
        __syncwarp();
    } else {
        const int q = warp & 3;                  // TMEM lane quarter == block index within the M-group
        mbar_wait(accf, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* wsp = p.ws + (long)split * p.ntaps * p.Cg * p.Ct;
        for (int g = 0; g < ng; ++g) {
            const int blk = (g_begin + g) * 4 + q;
            const bool valid = blk < p.nblk;
            const int tap = valid ? p.blk_tap[blk] : 0;
            const int a = valid ? p.blk_cb[blk] * 32 + lane : 0;
            float* dst = wsp + ((long)tap * p.Cg + a) * p.Ct;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BN + c0), v);
                if (valid) {
                    if (nchunks == 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// dw[tap*s_tap + a*s_a + b*s_b] (+)= sum over splits of ws[split][tap][a][b]   for a < a_lim, b < b_lim.
// Packed layouts: [taps][Cg][Ct] = (Cg*Ct, Ct, 1), [taps][Ct][Cg] = (Cg*Ct, 1, Cg); torch Conv2d weight [Co,Ci,R,S] with
// a = Ci, b = Co: (1, RS, Ci*RS); ConvTranspose2d weight [Ci,Co,R,S] with a = Co, b = Ci: (1, RS, Co*RS).
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int splits, int taps, int Cg, int Ct,
                                    long s_tap, long s_a, long s_b, int a_lim, int b_lim, int accumulate) {
    const long total = (long)taps * Cg * Ct;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int b = (int)(i % Ct); const long t = i / Ct; const int a = (int)(t % Cg); const long tap = t / Cg;
        if (a >= a_lim || b >= b_lim) continue;
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += __ldg(ws + (long)k * total + i);
        float* dst = dw + tap * s_tap + a * s_a + b * s_b;
        *dst = accumulate ? *dst + s : s;
    }
}

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

bool encode4(CUtensorMap* m, const void* base, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, const_cast<void*>(base), dims, strides_bytes, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int floordiv2(int t) { return (t - (t & 1)) / 2; }

struct Plan { int TH, TW, TNB, tiles_h, tiles_w, chunks, ngroups, gpc, grid_y, splits, cpc; };

bool make_plan(int N, int Hg, int Wg, int Cg, int Ht, int Wt, int Ct, int R, int S, int stride, Plan* pl) {
    if (Cg % 32 != 0 || !(Ct == 32 || Ct == 64 || Ct == 128)) return false;
    if (R * S > 32 || stride < 1 || stride > 2) return false;
    if (stride == 2 && ((Hg | Wg) & 1)) return false;
    const int nblk = R * S * (Cg / 32);
    if (nblk > MAX_BLK) return false;
    const bool p2 = Ht > 0 && Wt > 0 && (Ht & (Ht - 1)) == 0 && (Wt & (Wt - 1)) == 0;
    pl->TNB = 1;
    if ((long)Ht * Wt < 64) {
        if (!p2 || Ht * Wt < 8) return false;
        pl->TW = Wt; pl->TH = Ht; pl->TNB = CH / (Ht * Wt);          // several whole images per 64-pixel chunk
    } else if ((Wt & (Wt - 1)) == 0 && Wt >= 8) { pl->TW = Wt < CH ? Wt : CH; pl->TH = CH / pl->TW; }
    else { pl->TW = 8; pl->TH = 8; }
    pl->tiles_w = g2_cdiv(Wt, pl->TW); pl->tiles_h = g2_cdiv(Ht, pl->TH);
    pl->chunks = g2_cdiv(N, pl->TNB) * pl->tiles_h * pl->tiles_w;
    pl->ngroups = (nblk + 3) / 4;
    const int max_gpc = 512 / Ct;
    pl->grid_y = g2_cdiv(pl->ngroups, max_gpc);
    pl->gpc = g2_cdiv(pl->ngroups, pl->grid_y);
    int splits = 148 / pl->grid_y;
    if (splits < 1) splits = 1;
    if (splits > pl->chunks) splits = pl->chunks;
    pl->cpc = g2_cdiv(pl->chunks, splits);
    pl->splits = g2_cdiv(pl->chunks, pl->cpc);
    return true;
}

}  // namespace wg


// =====================================================================================================================
// Default path for stride-1 filters with >= 4 taps (validated on B200 in round 2: 1.7-2.0x the tile kernel; G2_WGRAD_HALO=0
// switches it off): the same weight gradient on a "halo" layout.
// The kernel above re-loads one shifted G tile per tap (R*S loads of the same pixels through L2).  Here one CTA item is
// (window of TH rows of one image, one 32-channel block of G): a single TMA box lands the zero-padded G window
// [(TH+R-1) * Wp pixel rows][32 ch] and one box per 32-channel block of T lands [TH * Wp][32 ch] with the SAME row pitch
// Wp = Wt + S - 1 (the Wp - Wt pad columns are out of bounds of T and arrive as zeros).  With flat pixel index p as the
// MMA's K dimension every tap is a row shift of the resident window (tests/test_wgrad_halo_math.py):
//     dW[tap] += G[toff : toff + F]^T  T[0 : F],     toff = (dh - dh_min) * Wp + (dw - dw_min),  F = TH * Wp,
// and the four 32-row atoms of the M = 128 operand are four TAPS of one channel block: atoms are LBO bytes apart, so
// LBO = 128 B stacks taps (dh, dw..dw+3) and LBO = Wp * 128 B stacks (dh..dh+3, dw) -- no data is duplicated.
// Requires that the MN-major operand fetch applies the 128-byte swizzle to absolute shared-memory addresses (the K-major
// path of igemm_halo.cu relies on the same property; tests/test_pending_next_round.py probes it for MN-major).
namespace wgh {
using namespace wg;

constexpr int MAX_GRP = 12;          // M-groups per channel block (5x5: 7, 3x3: 3)

struct Maps { CUtensorMap g; CUtensorMap t; };

struct P {
    float* ws;                       // [splits][ntaps][Cg][Ct]
    int N, Ht, TH, Wp, win_per_img, windows, win_per_cta;
    int Cg, Ct, ntaps, ngroups, cbs_per_cta, ncb;
    int g_rows, dh_min, dw_min, ksteps;
    int g_box_bytes, t_box_bytes;    // TMA transaction bytes per box
    int g_buf, t_blk, t_buf;         // smem bytes: one G window (incl. slack), one T channel block, all T blocks
    int slack_off, slack_bytes;      // zero-filled tail of each G buffer (the last taps' shifts and unused atoms read into it)
    short grp_off[MAX_GRP], grp_lbo[MAX_GRP], grp_n[MAX_GRP], grp_tap[MAX_GRP][4];   // rows, rows, valid atoms, tap index per atom
};

constexpr int NISSUE = 2;           // MMA issuer warps: one thread retires one tcgen05.mma per ~80 clk (N <= 128) whatever N is; two
                                    // instruction streams on different accumulators reach the operand bound (profiles/r02_umma_rate.txt)

template <int BN>
__global__ void __launch_bounds__(224, 1) wgrad_halo_kernel(const __grid_constant__ Maps maps, const __grid_constant__ P p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sG = sm;                                   // 2 x g_buf
    uint8_t* sT = sm + 2 * p.g_buf;                     // 2 x t_buf
    uint64_t* full = reinterpret_cast<uint64_t*>(sT + 2 * p.t_buf);
    uint64_t* empty = full + 2;
    uint64_t* accf = empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accf + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    const int cb_begin = blockIdx.y * p.cbs_per_cta;
    const int ncb = min(p.cbs_per_cta, p.ncb - cb_begin);
    const int w_begin = split * p.win_per_cta;
    const int nwin = max(0, min(p.win_per_cta, p.windows - w_begin));
    const int nitems = nwin * ncb;

    // zero the slack rows behind both G windows (never written by TMA; read by shifted / unused atoms: must be finite)
    for (int i = threadIdx.x * 16; i < p.slack_bytes; i += blockDim.x * 16) {
        *reinterpret_cast<uint4*>(sG + p.slack_off + i) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(sG + p.g_buf + p.slack_off + i) = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NISSUE); }
            mbar_init(accf, NISSUE);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int it = 0;
        for (int w = 0; w < nwin; ++w) {
            const int win = w_begin + w;
            const int n = win / p.win_per_img;
            const int h0 = (win - n * p.win_per_img) * p.TH;
            for (int c = 0; c < ncb; ++c, ++it) {
                const int buf = it & 1;
                mbar_wait(&empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(&full[buf], (uint32_t)(p.g_box_bytes + (BN / 32) * p.t_box_bytes));
                    tma_load_4d(sG + buf * p.g_buf, &maps.g, &full[buf], (cb_begin + c) * 32, p.dw_min, h0 + p.dh_min, n);
#pragma unroll
                    for (int j = 0; j < BN / 32; ++j)
                        tma_load_4d(sT + buf * p.t_buf + j * p.t_blk, &maps.t, &full[buf], j * 32, 0, h0, n);
                }
                __syncwarp();
            }
        }
    } else if (warp <= NISSUE) {
        // issuer iw owns the tap groups iw, iw + 2, ... (each group has its own accumulator)
        const int iw = warp - 1;
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int it = 0;
        for (int w = 0; w < nwin; ++w) {
            for (int c = 0; c < ncb; ++c, ++it) {
                const int buf = it & 1;
                mbar_wait(&full[buf], (uint32_t)(it >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bdesc = make_desc_mn_sw128(smem_u32(sT + buf * p.t_buf), (uint32_t)p.t_blk);
                const uint32_t gbase = smem_u32(sG + buf * p.g_buf);
                if (elect_one()) {
                    for (int g = iw; g < p.ngroups; g += NISSUE) {
                        const uint64_t adesc = make_desc_mn_sw128(gbase + (uint32_t)p.grp_off[g] * 128u, (uint32_t)p.grp_lbo[g] * 128u);
                        const uint32_t d = tmem_base + (uint32_t)((c * p.ngroups + g) * BN);
                        for (int kk = 0; kk < p.ksteps; ++kk)        // 8 pixel rows (1024 B = 64 descriptor units) per instruction
                            umma_tf32(d, adesc + 64 * kk, bdesc + 64 * kk, idesc, (w > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty[buf]);
                }
                __syncwarp();
            }
        }
        if (elect_one()) umma_commit(accf);
        __syncwarp();
    } else {
        const int q = warp & 3;                  // TMEM lane quarter == atom (tap) index within the M-group
        mbar_wait(accf, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* wsp = p.ws + (long)split * p.ntaps * p.Cg * p.Ct;
        for (int c = 0; c < ncb; ++c) {
            for (int g = 0; g < p.ngroups; ++g) {
                const bool valid = q < p.grp_n[g];
                const int tap = valid ? p.grp_tap[g][q] : 0;
                const int a = (cb_begin + c) * 32 + lane;
                float* dst = wsp + ((long)tap * p.Cg + a) * p.Ct;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((c * p.ngroups + g) * BN + c0), v);
                    if (valid) {
                        if (nitems == 0) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = 0u;
                        }
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

struct HPlan {
    int TH, Wp, g_rows, win_per_img, windows, ngroups, cbs_per_cta, grid_y, splits, wpc, slack_rows, ksteps;
    int g_buf, t_blk, t_buf, smem;
    short grp_off[MAX_GRP], grp_lbo[MAX_GRP], grp_n[MAX_GRP], grp_tap[MAX_GRP][4];
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Tap groups of one channel block: rows of four consecutive dw (LBO = 1 pixel row); what is left of each filter row is
// stacked along dh (LBO = Wp rows) when that needs fewer groups than short row groups; a group of one tap uses LBO = 1 row.
int make_groups(int R, int S, int Wp, HPlan* pl) {
    int ng = 0;
    auto add = [&](int off, int lbo, int n, const int* taps) {
        if (ng >= MAX_GRP) { ng = MAX_GRP + 1; return; }
        pl->grp_off[ng] = (short)off; pl->grp_lbo[ng] = (short)lbo; pl->grp_n[ng] = (short)n;
        for (int j = 0; j < 4; ++j) pl->grp_tap[ng][j] = (short)(j < n ? taps[j] : 0);
        ++ng;
    };
    const int rem = S % 4;
    for (int r = 0; r < R && ng <= MAX_GRP; ++r)
        for (int s0 = 0; s0 + 4 <= S; s0 += 4) {
            const int taps[4] = {r * S + s0, r * S + s0 + 1, r * S + s0 + 2, r * S + s0 + 3};
            add(r * Wp + s0, 1, 4, taps);
        }
    if (rem) {
        const int col_groups = rem * ((R + 3) / 4);
        if (col_groups < R) {
            for (int s = S - rem; s < S; ++s)
                for (int r0 = 0; r0 < R && ng <= MAX_GRP; r0 += 4) {
                    const int n = R - r0 < 4 ? R - r0 : 4;
                    int taps[4] = {0, 0, 0, 0};
                    for (int j = 0; j < n; ++j) taps[j] = (r0 + j) * S + s;
                    add(r0 * Wp + s, n == 1 ? 1 : Wp, n, taps);
                }
        } else {
            for (int r = 0; r < R && ng <= MAX_GRP; ++r) {
                int taps[4] = {0, 0, 0, 0};
                for (int j = 0; j < rem; ++j) taps[j] = r * S + (S - rem) + j;
                add(r * Wp + (S - rem), 1, rem, taps);
            }
        }
    }
    return ng;
}

bool make_hplan(int N, int Hg, int Wg, int Cg, int Ht, int Wt, int Ct, int R, int S, int stride, HPlan* pl) {
    if (stride != 1 || Cg % 32 != 0 || !(Ct == 32 || Ct == 64 || Ct == 128)) return false;
    if (R * S < 4 || R * S > 32) return false;                   // 1x1 / tiny filters: the tile kernel packs channel blocks instead
    if ((long)Ht * Wt < 256) return false;
    const int budget = 212 * 1024;
    long best_cost = -1;
    HPlan cand;
    // pitch candidates: the exact window width, or rounded up to 4 / 8 pixels (any TH then gives whole 8-row K steps)
    for (int align = 1; align <= 8; align *= 2) {
        const int Wp = round_up(Wt + S - 1, align);
        if (align == 2 || Wp > 256) continue;
        cand = *pl;
        cand.Wp = Wp;
        cand.ngroups = make_groups(R, S, Wp, &cand);
        if (cand.ngroups > MAX_GRP || cand.ngroups * Ct > 512) return false;
        int max_row = 0;            // furthest first row of any atom, valid or not
        for (int g = 0; g < cand.ngroups; ++g) {
            const int r = cand.grp_off[g] + 3 * cand.grp_lbo[g];
            if (r > max_row) max_row = r;
        }
        for (int TH = (Ht < 64 ? Ht : 64); TH >= 2; --TH) {
            if ((TH * Wp) % 8) continue;
            const int g_rows = TH + R - 1;
            if (g_rows > 256) continue;
            const int slack = max_row + TH * Wp - g_rows * Wp;      // rows read past the window by the furthest atom
            const int slack_rows = round_up(slack > 0 ? slack : 0, 8) + 8;
            const int g_buf = round_up((g_rows * Wp + slack_rows) * 128, 1024);
            const int t_blk = round_up(TH * Wp * 128, 1024);
            const int t_buf = (Ct / 32) * t_blk;
            const int smem = 2 * (g_buf + t_buf) + 1024 + 256;
            if (smem > budget) continue;
            // cost ~ MMA K rows issued per image and group, plus the halo rows of the G loads at a quarter weight
            const long cost = (long)g2_cdiv(Ht, TH) * (4L * TH * Wp + (long)(R - 1) * Wp);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                cand.TH = TH; cand.g_rows = g_rows; cand.slack_rows = slack_rows; cand.g_buf = g_buf; cand.t_blk = t_blk;
                cand.t_buf = t_buf; cand.smem = smem;
                *pl = cand;
            }
        }
    }
    if (best_cost < 0) return false;
    const int Wp = pl->Wp;
    pl->ksteps = pl->TH * Wp / 8;
    pl->win_per_img = g2_cdiv(Ht, pl->TH);
    pl->windows = N * pl->win_per_img;
    const int ncb = Cg / 32;
    const int max_cbs = 512 / (pl->ngroups * Ct);
    pl->grid_y = g2_cdiv(ncb, max_cbs);
    pl->cbs_per_cta = g2_cdiv(ncb, pl->grid_y);
    static const int ctas = [] { const char* e = getenv("G2_WGRAD_HALO_CTAS"); const int v = e && *e ? atoi(e) : 148; return v < 1 ? 1 : v; }();
    int splits = ctas / pl->grid_y;         // one CTA per SM; the CPU emulation lowers it so that a CTA accumulates several windows
    if (splits < 1) splits = 1;
    if (splits > pl->windows) splits = pl->windows;
    pl->wpc = g2_cdiv(pl->windows, splits);
    pl->splits = g2_cdiv(pl->windows, pl->wpc);
    return true;
}

bool enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("G2_WGRAD_HALO"); on = (e && e[0] == '0') ? 0 : 1; }      // default on (G2_WGRAD_HALO=0: tile kernel only)
    return on == 1;
}

}  // namespace wgh

extern "C" {

// Bytes of workspace g2_conv_wgrad_tf32 needs for this problem, or 0 if the shape is not supported.
long g2_conv_wgrad_tf32_workspace(int N, int Hg, int Wg, int Cg, int Ht, int Wt, int Ct, int R, int S, int stride) {
    wg::Plan pl;
    if (!wg::make_plan(N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride, &pl)) return 0;
    int splits = pl.splits;
    wgh::HPlan hp;
    if (wgh::enabled() && wgh::make_hplan(N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride, &hp) && hp.splits > splits) splits = hp.splits;
    return (long)splits * R * S * Cg * Ct * (long)sizeof(float);
}

// Host-only: the halo plan of a problem (for the CPU replay test).  out[0..15] = supported, TH, Wp, g_rows, win_per_img,
// windows, ngroups, cbs_per_cta, grid_y, splits, windows per CTA, slack rows, ksteps, g_buf, t_blk, smem; then per group
// 7 ints: first row, LBO in rows, valid atoms, tap of atom 0..3.  Returns the number of ints written.
int g2_conv_wgrad_halo_plan(int N, int Hg, int Wg, int Cg, int Ht, int Wt, int Ct, int R, int S, int stride, int* out) {
    wgh::HPlan hp;
    memset(&hp, 0, sizeof(hp));
    const bool ok = wgh::make_hplan(N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride, &hp);
    out[0] = ok ? 1 : 0;
    if (!ok) return 1;
    const int head[15] = {hp.TH, hp.Wp, hp.g_rows, hp.win_per_img, hp.windows, hp.ngroups, hp.cbs_per_cta, hp.grid_y, hp.splits,
                          hp.wpc, hp.slack_rows, hp.ksteps, hp.g_buf, hp.t_blk, hp.smem};
    for (int i = 0; i < 15; ++i) out[1 + i] = head[i];
    int n = 16;
    for (int g = 0; g < hp.ngroups; ++g) {
        out[n++] = hp.grp_off[g]; out[n++] = hp.grp_lbo[g]; out[n++] = hp.grp_n[g];
        for (int j = 0; j < 4; ++j) out[n++] = hp.grp_tap[g][j];
    }
    return n;
}

static int wgrad_halo_launch(const wgh::HPlan& hp, const float* g, const float* t, float* dw, float* ws, int N, int Hg, int Wg,
                             int Cg, int Ht, int Wt, int Ct, int R, int S, int pad, long s_tap, long s_a, long s_b, int a_lim,
                             int b_lim, int accumulate, cudaStream_t stream) {
    wgh::Maps maps;
    memset(&maps, 0, sizeof(maps));
    wgh::P p;
    memset(&p, 0, sizeof(p));
    p.ws = ws; p.N = N; p.Ht = Ht; p.TH = hp.TH; p.Wp = hp.Wp; p.win_per_img = hp.win_per_img; p.windows = hp.windows;
    p.win_per_cta = hp.wpc; p.Cg = Cg; p.Ct = Ct; p.ntaps = R * S; p.ngroups = hp.ngroups; p.cbs_per_cta = hp.cbs_per_cta;
    p.ncb = Cg / 32; p.g_rows = hp.g_rows; p.dh_min = -pad; p.dw_min = -pad; p.ksteps = hp.ksteps;
    p.g_box_bytes = hp.g_rows * hp.Wp * 128; p.t_box_bytes = hp.TH * hp.Wp * 128;
    p.g_buf = hp.g_buf; p.t_blk = hp.t_blk; p.t_buf = hp.t_buf;
    p.slack_off = hp.g_rows * hp.Wp * 128; p.slack_bytes = hp.g_buf - p.slack_off;
    for (int i = 0; i < hp.ngroups; ++i) {
        p.grp_off[i] = hp.grp_off[i]; p.grp_lbo[i] = hp.grp_lbo[i]; p.grp_n[i] = hp.grp_n[i];
        for (int j = 0; j < 4; ++j) p.grp_tap[i][j] = hp.grp_tap[i][j];
    }
    {
        const cuuint32_t box[4] = {32, (cuuint32_t)hp.Wp, (cuuint32_t)hp.TH, 1};
        const cuuint64_t dims[4] = {(cuuint64_t)Ct, (cuuint64_t)Wt, (cuuint64_t)Ht, (cuuint64_t)N};
        const cuuint64_t str[3] = {(cuuint64_t)Ct * 4, (cuuint64_t)Wt * Ct * 4, (cuuint64_t)Ht * Wt * Ct * 4};
        if (!wg::encode4(&maps.t, t, dims, str, box)) return G2_ERR_UNSUPPORTED;
    }
    {
        const cuuint32_t box[4] = {32, (cuuint32_t)hp.Wp, (cuuint32_t)hp.g_rows, 1};
        const cuuint64_t dims[4] = {(cuuint64_t)Cg, (cuuint64_t)Wg, (cuuint64_t)Hg, (cuuint64_t)N};
        const cuuint64_t str[3] = {(cuuint64_t)Cg * 4, (cuuint64_t)Wg * Cg * 4, (cuuint64_t)Hg * Wg * Cg * 4};
        if (!wg::encode4(&maps.g, g, dims, str, box)) return G2_ERR_UNSUPPORTED;
    }
    dim3 grid((unsigned)hp.splits, (unsigned)hp.grid_y, 1);
    cudaError_t e = cudaSuccess;
#define WGH_LAUNCH(BN)                                                                                                     \
    {                                                                                                                      \
        static G2DevOnce once;                                                                                             \
        if (once.needed(hp.smem)) {                                                                                        \
            e = cudaFuncSetAttribute(wgh::wgrad_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);  \
            if (e != cudaSuccess) return (int)e;                                                                           \
            once.done();                                                                                                   \
        }                                                                                                                  \
        wgh::wgrad_halo_kernel<BN><<<grid, 224, hp.smem, stream>>>(maps, p);                                                   \
    }
    if (Ct == 32) WGH_LAUNCH(32) else if (Ct == 64) WGH_LAUNCH(64) else WGH_LAUNCH(128)
#undef WGH_LAUNCH
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const long total = (long)R * S * Cg * Ct;
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 8) blocks = 148L * 8;
    wg::wgrad_reduce_kernel<<<(int)blocks, 256, 0, stream>>>(ws, dw, hp.splits, R * S, Cg, Ct, s_tap, s_a, s_b, a_lim, b_lim, accumulate);
    G2_LAUNCH_RET();
}

// Same contract as g2_conv_wgrad_f32 (TF32 operands, fp32 accumulation); `ws` is caller-owned scratch of
// g2_conv_wgrad_tf32_workspace(...) bytes.
static int wgrad_impl(const float* g, const float* t, float* dw, float* ws, int N, int Hg, int Wg, int Cg, int Ht, int Wt,
                      int Ct, int R, int S, int stride, int pad, long s_tap, long s_a, long s_b, int a_lim, int b_lim,
                      int accumulate, cudaStream_t stream) {
    using namespace wg;
    G2_CHECK_ARG(g && t && dw && ws && N > 0);
    Plan pl;
    if (!make_plan(N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride, &pl)) return G2_ERR_UNSUPPORTED;
    G2_CHECK_ARG((reinterpret_cast<uintptr_t>(g) & 15) == 0 && (reinterpret_cast<uintptr_t>(t) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(ws) & 15) == 0);
    if (wgh::enabled()) {
        wgh::HPlan hp;
        if (wgh::make_hplan(N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride, &hp))
            return wgrad_halo_launch(hp, g, t, dw, ws, N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, pad, s_tap, s_a, s_b, a_lim, b_lim,
                                     accumulate, stream);
    }
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    P p;
    memset(&p, 0, sizeof(p));
    p.ws = ws; p.N = N; p.Ht = Ht; p.Wt = Wt; p.TH = pl.TH; p.TW = pl.TW; p.TNB = pl.TNB; p.tiles_h = pl.tiles_h; p.tiles_w = pl.tiles_w;
    p.chunks_total = pl.chunks; p.chunks_per_cta = pl.cpc; p.Cg = Cg; p.Ct = Ct; p.ntaps = R * S;
    p.groups_per_cta = pl.gpc;
    for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s) {
            const int tap = r * S + s, tr = r - pad, ts = s - pad;
            if (stride == 1) { p.tap_plane[tap] = 0; p.tap_dh[tap] = (short)tr; p.tap_dw[tap] = (short)ts; }
            else {
                p.tap_plane[tap] = (short)(((tr & 1) << 1) | (ts & 1));
                p.tap_dh[tap] = (short)floordiv2(tr); p.tap_dw[tap] = (short)floordiv2(ts);
            }
        }
    int nb = 0;
    for (int tap = 0; tap < R * S; ++tap)
        for (int cb = 0; cb < Cg / 32; ++cb) { p.blk_tap[nb] = (short)tap; p.blk_cb[nb] = (short)cb; ++nb; }
    p.nblk = nb;
    const cuuint32_t box[4] = {32, (cuuint32_t)pl.TW, (cuuint32_t)pl.TH, (cuuint32_t)pl.TNB};
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Ct, (cuuint64_t)Wt, (cuuint64_t)Ht, (cuuint64_t)N};
        const cuuint64_t str[3] = {(cuuint64_t)Ct * 4, (cuuint64_t)Wt * Ct * 4, (cuuint64_t)Ht * Wt * Ct * 4};
        if (!encode4(&maps.t, t, dims, str, box)) return G2_ERR_UNSUPPORTED;
    }
    if (stride == 1) {
        const cuuint64_t dims[4] = {(cuuint64_t)Cg, (cuuint64_t)Wg, (cuuint64_t)Hg, (cuuint64_t)N};
        const cuuint64_t str[3] = {(cuuint64_t)Cg * 4, (cuuint64_t)Wg * Cg * 4, (cuuint64_t)Hg * Wg * Cg * 4};
        if (!encode4(&maps.g[0], g, dims, str, box)) return G2_ERR_UNSUPPORTED;
    } else {
        for (int plane = 0; plane < 4; ++plane) {
            const int pr = plane >> 1, ps = plane & 1;
            const cuuint64_t dims[4] = {(cuuint64_t)Cg, (cuuint64_t)Wg / 2, (cuuint64_t)Hg / 2, (cuuint64_t)N};
            const cuuint64_t str[3] = {(cuuint64_t)2 * Cg * 4, (cuuint64_t)2 * Wg * Cg * 4, (cuuint64_t)Hg * Wg * Cg * 4};
            if (!encode4(&maps.g[plane], g + ((long)pr * Wg + ps) * Cg, dims, str, box)) return G2_ERR_UNSUPPORTED;
        }
    }
    dim3 grid((unsigned)pl.splits, (unsigned)pl.grid_y, 1);
    cudaError_t e = cudaSuccess;
#define WG_LAUNCH(BN)                                                                                                     \
    {                                                                                                                     \
        constexpr int smem = STAGES * 4 * TILE_BYTES + 2 * (BN / 32) * TILE_BYTES + (2 * STAGES + 5) * 8 + 16 + 1024;     \
        static G2DevOnce once;                                                                                            \
        if (once.needed()) {                                                                                              \
            e = cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);             \
            if (e != cudaSuccess) return (int)e;                                                                          \
            once.done();                                                                                                  \
        }                                                                                                                 \
        wgrad_tc_kernel<BN><<<grid, 192, smem, stream>>>(maps, p);                                                        \
    }
    if (Ct == 32) WG_LAUNCH(32) else if (Ct == 64) WG_LAUNCH(64) else WG_LAUNCH(128)
#undef WG_LAUNCH
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const long total = (long)R * S * Cg * Ct;
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 8) blocks = 148L * 8;
    wgrad_reduce_kernel<<<(int)blocks, 256, 0, stream>>>(ws, dw, pl.splits, R * S, Cg, Ct, s_tap, s_a, s_b, a_lim, b_lim, accumulate);
    G2_LAUNCH_RET();
}

int g2_conv_wgrad_tf32(const float* g, const float* t, float* dw, float* ws, int N, int Hg, int Wg, int Cg, int Ht, int Wt,
                       int Ct, int R, int S, int stride, int pad, int outT, cudaStream_t stream) {
    if (outT) return wgrad_impl(g, t, dw, ws, N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride, pad, (long)Cg * Ct, 1, Cg, Cg, Ct, 0, stream);
    return wgrad_impl(g, t, dw, ws, N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride, pad, (long)Cg * Ct, Ct, 1, Cg, Ct, 0, stream);
}

// Weight gradient written straight into a tensor of arbitrary (tap, a, b) strides -- e.g. the torch-layout .grad of the
// parameter -- optionally accumulating, restricted to a < a_lim, b < b_lim (zero-padded channels are dropped).
int g2_conv_wgrad_tf32_to(const float* g, const float* t, float* dw, float* ws, int N, int Hg, int Wg, int Cg, int Ht, int Wt,
                          int Ct, int R, int S, int stride, int pad, long s_tap, long s_a, long s_b, int a_lim, int b_lim,
                          int accumulate, cudaStream_t stream) {
    G2_CHECK_ARG(a_lim > 0 && a_lim <= Cg && b_lim > 0 && b_lim <= Ct);
    return wgrad_impl(g, t, dw, ws, N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride, pad, s_tap, s_a, s_b, a_lim, b_lim, accumulate, stream);
}

}  // extern "C"
