// genesis_b200 -- "halo" TF32 implicit GEMM for sm_100a: the stride-1 convolutions (forward, data gradient,
// conv-transpose forward, and each sub-pixel class of a stride-2 conv-transpose) with the activation window
// loaded into shared memory ONCE per CTA and every filter tap issued from it through shifted UMMA descriptors.
//
//     out[n, h*os+ph, w*os+pw, co] = act(bias[co] + sum_tap sum_c X[n, h+dh(tap), w+dw(tap), c] * W[widx(tap)][co][c])
//
// over the virtual output grid (h, w) in [0,Hv) x [0,Wv); X is zero outside the image.
//
// Layout trick: a CTA owns TH output rows of one image (or TNB whole small images).  One TMA box
// {32 channels, Wp = Wv + dw-span pixels, rows, images} lands the zero-padded input window in shared memory as a
// FLAT run of 128-byte pixel rows with pitch Wp (out-of-bounds fill supplies the padding).  In that flat space the
// output position f = h_local*Wp + w reads input row f + toff(tap), toff = (dh-dh_min)*Wp + (dw-dw_min): every tap
// is the SAME K-major SWIZZLE_128B operand shifted by toff rows.  The 128-byte swizzle is a function of the absolute
// shared-memory address (tests/test_umma_layouts_gpu.py), so a descriptor may start at any row.  M tiles are 128
// consecutive flat positions; positions with w >= Wv (the Wp-Wv pad columns) are computed and dropped.
// Versus one TMA load per (tile, tap) this cuts L2->SMEM traffic by ~taps/halo-overhead (3x3: 4-5x, 5x5: 8-14x).
//
// Roles (one-item kernel): warp 0 = TMA producer (activation window in up to 4 row chunks, weight taps through a ring),
// warps 1-2 = MMA issuers (warp 1 also allocates TMEM; tap-outer, tile-inner so each weight tile is loaded once per CTA;
// issuer i takes the M tiles t = i mod 2), warps 3-6 = epilogue (tcgen05.ld -> bias/activation -> global).
// Why two issuers: the tcgen05.mma instructions of ONE thread retire one every ~80 clocks for N <= 128 (M = 128, K = 8 tf32),
// whatever N is; two independent instruction streams interleave down to the shared-memory operand bound, 40 clk (N = 32) /
// 48 clk (N = 64) -- measured with g2_debug_umma_rate, profiles/r02_umma_rate.txt.  Two CTAs per SM overlap load / MMA / epilogue.
#include "umma.cuh"
#include <cudaTypedefs.h>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace halo {
using namespace umma;

constexpr int MAX_TAPS = 32;
constexpr int MAX_CHUNKS = 8;
constexpr int MAX_STAGES = 12;
constexpr int MAX_PSTAGES = 64;     // persistent variant: up to 64 resident (channel block, tap) weight tiles

struct Maps { CUtensorMap a; CUtensorMap b; };

struct P {
    float* out; const float* bias;
    int N, Hv, Wv, TH, TW, TNB, Wp, RH; // RH: window rows per image (TH + dh-span); Wp: window pitch (>= TW + dw-span)
    int ch_rows, nch, ch_pix;           // rows per chunk (TMA box rows), chunks, pixel rows per chunk (= ch_rows*Wp*TNB)
    int tiles_h, tiles_w, dh_min, dw_min, span_h;
    int Ho, Wo, Co, os, ph, pw;
    int ntaps, cblocks, act, tmem_cols, a_bytes, stages, epi;
    int x3, w_lo_off;                   // 3xTF32 mode (see conv_halo_kernel); channel offset of the w_lo half in the weight pack
    long long* dbg;                     // optional per-CTA phase timestamps [grid][8] (scripts/conv_bench.py --timeline)
    int toff[MAX_TAPS];
    short widx[MAX_TAPS];
};

template <int ACT> __device__ __forceinline__ float act_apply(float v) {
    if (ACT == G2_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == G2_ACT_ELU) return v > 0.f ? v : __expf(v) - 1.f;     // TF32 path: |error| < 2e-7 absolute
    if (ACT == G2_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}

// One epilogue warp: TMEM lanes 32q..32q+31 of every M tile -> +bias -> activation -> global.
// (Measured dead end, round 2: prefetching the next TMEM block into a second register set + batching the eight staging reads
// before the eight stores changes nothing in the persistent kernel -- it is shared-memory-bandwidth bound, not epilogue-latency
// bound -- and costs the one-item kernel its second CTA per SM: 0.157 -> 0.216 ms on the 3x3 32->32 layer.)
// Each lane owns one output pixel row in TMEM; the 32 x 128-byte block is transposed through a swizzled 4 KB
// staging tile so that every store instruction writes four complete 128-byte pixel rows (8 lanes x 16 B each).
template <int BN, int ACT>
__device__ __forceinline__ void epilogue(const P& p, uint32_t tmem_base, const float* sBias, float* stage, int m_tiles, int n0,
                                         int h0, int w0, int n0c, int imgs_valid, int rows_valid, int cols_valid,
                                         int t_first = 0, int t_step = 1) {
    const int lane = threadIdx.x & 31, q = (threadIdx.x >> 5) & 3;
    const int img_pix = p.RH * p.Wp;
    const int sub = lane >> 3, chunk = lane & 7;
    for (int t = t_first; t < m_tiles; t += t_step) {
        const int f = t * 128 + q * 32 + lane;
        const int i = f / img_pix, rem = f - i * img_pix;
        const int hl = rem / p.Wp, w = rem - hl * p.Wp;
        const bool valid = i < imgs_valid && hl < rows_valid && w < cols_valid;
        const int oh = (h0 + hl) * p.os + p.ph, ow = (w0 + w) * p.os + p.pw;
        const int pix = valid ? ((n0 + i) * p.Ho + oh) * p.Wo + ow : -1;        // < 2^31 output pixels (checked on the host)
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * BN + c0), v);
            tmem_ld_wait();
            if (p.epi == 2) {
                // direct (G2_HALO_EPI=2, experiment): the lane writes its own 128-byte row with eight 16-byte stores -- no staging
                // traffic in shared memory, but 32 lines per store instruction: 7-12 % slower (profiles/r02_conv_bench_epilogue_direct.txt)
                float* dst = p.out + (size_t)(pix < 0 ? 0 : pix) * p.Co + n0c + c0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = *reinterpret_cast<const float4*>(sBias + c0 + 4 * j);
                    float4 o;
                    o.x = act_apply<ACT>(__uint_as_float(v[4 * j]) + b.x);
                    o.y = act_apply<ACT>(__uint_as_float(v[4 * j + 1]) + b.y);
                    o.z = act_apply<ACT>(__uint_as_float(v[4 * j + 2]) + b.z);
                    o.w = act_apply<ACT>(__uint_as_float(v[4 * j + 3]) + b.w);
                    if (pix >= 0) *reinterpret_cast<float4*>(dst + 4 * j) = o;
                }
                continue;
            }
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
                const int rp = __shfl_sync(0xffffffffu, pix, row);
                const float4 o = *reinterpret_cast<const float4*>(stage + row * 32 + ((chunk ^ (row & 7)) << 2));
                if (rp >= 0) *reinterpret_cast<float4*>(p.out + (size_t)rp * p.Co + n0c + c0 + chunk * 4) = o;
            }
            __syncwarp();
        }
    }
}

// Bulk-copy epilogue: each lane owns one output pixel = one 128-byte row per 32-channel group.  The lane writes its row
// into a private staging row (pitch 144 B: conflict-free float4 stores), makes it visible to the async proxy and hands it
// to the bulk-copy engine (cp.async.bulk shared -> global, 128 B), so the LSU sees 8 STS per lane instead of
// 8 STS + 8 SHFL + 8 LDS + 8 STG.  Two staging rows per lane: the copy of group g drains while group g+1 is computed.
template <int BN, int ACT>
__device__ __forceinline__ void epilogue_bulk(const P& p, uint32_t tmem_base, const float* sBias, uint8_t* stage_base, int nbuf,
                                              int m_tiles, int n0, int h0, int w0, int n0c, int imgs_valid, int rows_valid,
                                              int cols_valid) {
    const int lane = threadIdx.x & 31, q = (threadIdx.x >> 5) & 3;
    const int img_pix = p.RH * p.Wp;
    uint8_t* my = stage_base + (size_t)(q * 32 + lane) * 144;             // this lane's row in buffer 0; buffer 1 is 128*144 B further
    int it = 0;
    for (int t = 0; t < m_tiles; ++t) {
        const int f = t * 128 + q * 32 + lane;
        const int i = f / img_pix, rem = f - i * img_pix;
        const int hl = rem / p.Wp, w = rem - hl * p.Wp;
        const bool valid = i < imgs_valid && hl < rows_valid && w < cols_valid;
        const int oh = (h0 + hl) * p.os + p.ph, ow = (w0 + w) * p.os + p.pw;
        float* orow = p.out + ((size_t)((n0 + i) * p.Ho + oh) * p.Wo + ow) * p.Co + n0c;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32, ++it) {
            uint32_t v[32];
            tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * BN + c0), v);
            float* row = reinterpret_cast<float*>(my + (nbuf == 2 ? (it & 1) * (128 * 144) : 0));
            // the bulk copy that last read this staging row must have finished reading it
            if (nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(sBias + c0 + 4 * j);
                float4 o;
                o.x = act_apply<ACT>(__uint_as_float(v[4 * j]) + b.x);
                o.y = act_apply<ACT>(__uint_as_float(v[4 * j + 1]) + b.y);
                o.z = act_apply<ACT>(__uint_as_float(v[4 * j + 2]) + b.z);
                o.w = act_apply<ACT>(__uint_as_float(v[4 * j + 3]) + b.w);
                *reinterpret_cast<float4*>(row + 4 * j) = o;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (valid)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;"
                             ::"l"(orow + c0), "r"(smem_u32(row)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all rows written before the CTA retires
}

// 3xTF32 low part of a raw fp32 operand whose high part is what the tensor core sees, trunc(x): x - trunc(x) is exact in fp32
// (13 significant bits); it is then rounded to NEAREST TF32 here, because the tensor core would truncate it (a biased 2^-20
// relative error of the product instead of an unbiased 2^-21).
__device__ __forceinline__ float lo_tf32(float x) {
    const float lo = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    return __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xFFFFE000u);
}

constexpr int NISSUE = 2;            // MMA issuer warps of conv_halo_kernel

template <int BN>
__global__ void __launch_bounds__(224) conv_halo_kernel(const __grid_constant__ Maps maps, const __grid_constant__ P p) {
    constexpr int B_BYTES = BN * 128;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sA = sm;
    uint8_t* sB = sm + p.a_bytes;
    const int STAGES = p.stages;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* emptyA = fullA + MAX_CHUNKS;
    uint64_t* fullB = emptyA + 1;
    uint64_t* emptyB = fullB + MAX_STAGES;
    uint64_t* accf = emptyB + MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accf + 1);
    uint64_t* p12 = accf + 2;                                   // 3xTF32: passes 1+2 of a channel block retired (window may be rewritten)
    uint64_t* loready = accf + 3;                               // 3xTF32: the window now holds lo = x - trunc(x)
    float* sBias = reinterpret_cast<float*>(fullA) + 96;        // 384 B past the start of the barrier block

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tw_i = blockIdx.x % p.tiles_w;
    const int th_i = (blockIdx.x / p.tiles_w) % p.tiles_h;
    const int n0 = (blockIdx.x / (p.tiles_w * p.tiles_h)) * p.TNB;
    const int h0 = th_i * p.TH, w0 = tw_i * p.TW;
    const int n0c = blockIdx.y * BN;
    const int rows_valid = min(p.TH, p.Hv - h0);
    const int cols_valid = min(p.TW, p.Wv - w0);
    const int imgs_valid = min(p.TNB, p.N - n0);
    // number of 128-position M tiles that contain a valid output, and window chunks they read
    const int f_last = ((imgs_valid - 1) * p.RH + rows_valid - 1) * p.Wp + cols_valid - 1;
    const int m_tiles = f_last / 128 + 1;
    const int nch = p.TNB > 1 ? 1 : min(p.nch, (rows_valid + p.span_h + p.ch_rows - 1) / p.ch_rows);

    long long* dbg = p.dbg ? p.dbg + ((long)blockIdx.y * gridDim.x + blockIdx.x) * 64 : nullptr;    // [8 phase stamps | 2 per tap]
    if (dbg && threadIdx.x == 32) { dbg[0] = clock64(); unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); dbg[7] = sm; }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a);
        prefetch_tmap(&maps.b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int j = 0; j < MAX_CHUNKS; ++j) mbar_init(&fullA[j], 1);
            mbar_init(emptyA, NISSUE);                   // every issuer commits once per channel block / stage / kernel
            for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], NISSUE); }
            mbar_init(accf, NISSUE);
            mbar_init(p12, NISSUE);
            mbar_init(loready, 128);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    }
    if (threadIdx.x >= 96 && threadIdx.x < 96 + BN) sBias[threadIdx.x - 96] = p.bias ? p.bias[n0c + threadIdx.x - 96] : 0.f;
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (dbg && threadIdx.x == 32) dbg[1] = clock64();          // prologue done
    // 3xTF32 mode (p.x3): the window holds RAW fp32 activations (tensor map of type FLOAT32: no TMA rounding).  The tensor core
    // truncates fp32 operand bits to TF32 (tests/test_umma_layouts_gpu.py::test_tf32_mma_operand_conversion...), so the raw
    // window IS x_hi = trunc(x).  Per channel block: pass 1 (x_hi, w_hi) and pass 2 (x_hi, w_lo) issue from the raw window; the
    // four epilogue warps then rewrite it IN PLACE to x_lo = x - trunc(x) (exact in fp32) and pass 3 (x_lo, w_hi) follows:
    // a * b ~= a_hi b_hi + a_hi b_lo + a_lo b_hi at 2^-21 relative error, with no extra shared memory and no split pass over HBM.
    const int wtiles = p.x3 ? 3 * p.ntaps : p.ntaps;           // weight tiles per channel block, in issue order

    if (warp == 0) {
        // ---- TMA producer: the whole warp walks the schedule (uniform control flow), one elected lane issues
        const uint32_t ch_bytes = (uint32_t)p.ch_pix * 128u;
        int bi = 0;
        uint32_t bph = 0;
        for (int cb = 0; cb < p.cblocks; ++cb) {
            if (cb > 0) mbar_wait(emptyA, (uint32_t)(cb - 1) & 1u);     // all MMAs of the previous channel block retired
            if (elect_one()) {
                for (int j = 0; j < nch; ++j) {
                    mbar_expect_tx(&fullA[j], ch_bytes);
                    tma_load_4d(sA + (size_t)j * ch_bytes, &maps.a, &fullA[j], cb * 32, w0 + p.dw_min, h0 + p.dh_min + j * p.ch_rows, n0);
                }
            }
            __syncwarp();
            for (int wt = 0; wt < wtiles; ++wt) {
                const int tap = wt % p.ntaps;
                const int wc = cb * 32 + (wt / p.ntaps == 1 ? p.w_lo_off : 0);     // x3: taps with w_hi, then w_lo, then w_hi again
                const int s = bi;
                mbar_wait(&emptyB[s], bph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(&fullB[s], B_BYTES);
                    tma_load_3d(sB + s * B_BYTES, &maps.b, &fullB[s], wc, n0c, p.widx[tap]);
                }
                __syncwarp();
                if (++bi == STAGES) { bi = 0; bph ^= 1u; }
            }
        }
    } else if (warp <= NISSUE) {
        // ---- MMA issuers: uniform control flow, one elected lane issues tcgen05.mma / commit; issuer iw owns tiles iw, iw+2, ...
        const int iw = warp - 1;
        if (iw != 0) dbg = nullptr;
        constexpr uint32_t idesc = idesc_tf32(BN);
        const uint64_t hi = desc_k_sw128_hi();
        const uint32_t a0 = smem_u32(sA) >> 4;
        const uint32_t b0 = smem_u32(sB) >> 4;
        // Per-tap constants live in the lanes (lane L = tap L) and reach the issue loop with one shuffle each: the loop body must
        // stay SHORT -- a tap is 8 MMAs per issuer (320 clk of tensor-core time at the operand bound) and the issuer is a single
        // warp running dependent scalar code; with the tap table read from the parameter bank and an integer division per tap
        // the body took ~1 000 clk and the tensor core starved (profiles/r02_ncu_halo_issue_loop.txt).
        const int my_toff = lane < p.ntaps ? p.toff[lane] : 0;
        const uint32_t my_alo = a0 + (uint32_t)my_toff * 8u;
        // activation chunks a tap reads (monotone in the tile index: the last tile's)
        const int my_need = min(nch - 1, (128 * (m_tiles - 1) + 127 + my_toff) / p.ch_pix);
        const int npass = p.x3 ? 3 : 1;
        int bi = 0;
        uint32_t bph = 0;
        for (int cb = 0; cb < p.cblocks; ++cb) {
            int waited = 0;
            const uint32_t aph = (uint32_t)cb & 1u;
            for (int pass = 0; pass < npass; ++pass) {
                if (pass == 2) {
                    // 3xTF32, passes 1 + 2 issued: signal their retirement, then wait until the window has been rewritten to x_lo
                    if (elect_one()) commit(p12);
                    __syncwarp();
                    mbar_wait(loready, aph);
                    fence_after();
                }
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    const uint32_t alo = __shfl_sync(0xffffffffu, my_alo, tap);
                    const int need = __shfl_sync(0xffffffffu, my_need, tap);
                    mbar_wait(&fullB[bi], bph);
                    while (waited <= need) { mbar_wait(&fullA[waited], aph); ++waited; }
                    fence_after();
                    if (dbg && lane == 0 && (cb | pass | tap) == 0) dbg[2] = clock64();     // first weight tile + its activation chunks landed
                    const uint64_t bdesc = hi | (uint64_t)((b0 + (uint32_t)bi * (uint32_t)(B_BYTES >> 4)) & 0x3FFFu);
                    const uint32_t first = (cb | pass | tap) != 0 ? 1u : 0u;
                    if (elect_one()) {
                        for (int t = iw; t < m_tiles; t += NISSUE) {
                            const uint64_t adesc = hi | (uint64_t)((alo + (uint32_t)t * 1024u) & 0x3FFFu);
                            const uint32_t d = tmem_base + (uint32_t)(t * BN);
                            mma_tf32(d, adesc, bdesc, idesc, first);        // 4 x (K = 8 tf32 = 32 B) per 128-byte row
                            mma_tf32(d, adesc + 2, bdesc + 2, idesc, 1u);
                            mma_tf32(d, adesc + 4, bdesc + 4, idesc, 1u);
                            mma_tf32(d, adesc + 6, bdesc + 6, idesc, 1u);
                        }
                        commit(&emptyB[bi]);                // weight stage free when these MMAs retire
                    }
                    __syncwarp();
                    if (++bi == STAGES) { bi = 0; bph ^= 1u; }
                }
            }
            if (elect_one()) commit(emptyA);            // activation window free
            __syncwarp();
        }
        if (elect_one()) commit(accf);                  // all accumulators complete
        __syncwarp();
        if (dbg && lane == 0) dbg[3] = clock64();       // last MMA issued
    } else {
        // ---- epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31
        if (p.x3) {
            // 3xTF32: between pass 2 and pass 3 of every channel block rewrite the window in place, x -> x - trunc(x)
            const int tid = threadIdx.x - 96;
            const int n4 = nch * p.ch_pix * 8;                  // float4s of the loaded window (128 B = 8 float4 per position)
            for (int cb = 0; cb < p.cblocks; ++cb) {
                const uint32_t ph = (uint32_t)cb & 1u;
                for (int j = 0; j < nch; ++j) mbar_wait(&fullA[j], ph);      // the TMA writes of this channel block are visible
                mbar_wait(p12, ph);                                          // ... and no MMA reads the raw window any more
                float4* w4 = reinterpret_cast<float4*>(sA);
                for (int i = tid; i < n4; i += 128) {
                    float4 v = w4[i];
                    v.x = lo_tf32(v.x); v.y = lo_tf32(v.y); v.z = lo_tf32(v.z); v.w = lo_tf32(v.w);
                    w4[i] = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core (async proxy) reads
                mbar_arrive(loready);
            }
        }
        mbar_wait(accf, 0);
        fence_after();
        if (dbg && threadIdx.x == 96) dbg[4] = clock64();       // accumulators complete
        // the activation window is dead once every MMA has retired: reuse its head as store staging
        if (p.epi == 1) {
            const int nbuf = p.a_bytes >= 2 * 128 * 144 ? 2 : 1;
            switch (p.act) {
                case G2_ACT_RELU: epilogue_bulk<BN, G2_ACT_RELU>(p, tmem_base, sBias, sA, nbuf, m_tiles, n0, h0, w0, n0c, imgs_valid, rows_valid, cols_valid); break;
                case G2_ACT_ELU: epilogue_bulk<BN, G2_ACT_ELU>(p, tmem_base, sBias, sA, nbuf, m_tiles, n0, h0, w0, n0c, imgs_valid, rows_valid, cols_valid); break;
                case G2_ACT_SIGMOID: epilogue_bulk<BN, G2_ACT_SIGMOID>(p, tmem_base, sBias, sA, nbuf, m_tiles, n0, h0, w0, n0c, imgs_valid, rows_valid, cols_valid); break;
                default: epilogue_bulk<BN, G2_ACT_NONE>(p, tmem_base, sBias, sA, nbuf, m_tiles, n0, h0, w0, n0c, imgs_valid, rows_valid, cols_valid); break;
            }
        } else {
        float* stage = reinterpret_cast<float*>(sA) + (warp & 3) * 1024;
        switch (p.act) {
            case G2_ACT_RELU: epilogue<BN, G2_ACT_RELU>(p, tmem_base, sBias, stage, m_tiles, n0, h0, w0, n0c, imgs_valid, rows_valid, cols_valid); break;
            case G2_ACT_ELU: epilogue<BN, G2_ACT_ELU>(p, tmem_base, sBias, stage, m_tiles, n0, h0, w0, n0c, imgs_valid, rows_valid, cols_valid); break;
            case G2_ACT_SIGMOID: epilogue<BN, G2_ACT_SIGMOID>(p, tmem_base, sBias, stage, m_tiles, n0, h0, w0, n0c, imgs_valid, rows_valid, cols_valid); break;
            default: epilogue<BN, G2_ACT_NONE>(p, tmem_base, sBias, stage, m_tiles, n0, h0, w0, n0c, imgs_valid, rows_valid, cols_valid); break;
        }
        }
        if (dbg && threadIdx.x == 96) dbg[5] = clock64();       // epilogue of warp 3 done
    }
    fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------ persistent variant
// Default route since round 2 (G2_HALO_PERSISTENT=0 switches it off; see persistent_mode()).
// One CTA per SM walks the work items (item = the tile the kernel above gives to one CTA) with a static stride.  Two
// activation windows, two TMEM accumulator sets and a dedicated store-staging tile let the three roles run ahead of one
// another: the producer loads window i+1 while the MMA warp works on window i, and the epilogue drains accumulator set
// (i & 1) under the MMAs of item i+1 -- the MMA phase, which already runs at the shared-memory operand-read bound
// (DESIGN.md 3.1), then covers the whole lifetime.  Weights stay resident when all taps fit (3x3), else they stream
// through the ring once per item.  With ONE issuing thread this kernel is capped at ~80 clk per MMA (the per-thread retire rate,
// profiles/r02_umma_rate.txt) and only ties with two non-persistent CTAs per SM; with two issuer warps it reaches the
// shared-memory operand bound.
constexpr int NEPI_P = 4;                   // epilogue warps of the persistent kernel (8 = two per TMEM lane quarter on alternate M tiles: measured slower,
                                            // the extra 16 KB of staging shrinks the windows and the epilogue is shared-memory-bandwidth bound anyway)
constexpr int PSTAGE_BYTES = NEPI_P * 4096; // epilogue staging: one 32 rows x 128 B tile per warp
constexpr int NREWRITE_P = 4;               // 3xTF32: warps that rewrite a window in place to x_lo (idle otherwise)

struct PP {
    P p;                                // per-item geometry, as for the kernel above
    int items_x, items_y;               // work items: blockIdx.x / blockIdx.y space of the non-persistent launch
    int acc_cols;                       // TMEM columns of one accumulator set (m_tiles * BN)
    int b_resident;                     // 1: every (cb, tap) weight tile has its own stage and is loaded once per CTA
};

template <int BN>
__global__ void __launch_bounds__(32 * (2 + NISSUE + NEPI_P + NREWRITE_P), 1) conv_halo_persistent_kernel(const __grid_constant__ Maps maps, const __grid_constant__ PP pp) {
    const P& p = pp.p;
    constexpr int B_BYTES = BN * 128;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sA0 = sm;                                   // two windows of p.a_bytes
    uint8_t* sB = sm + 2 * (size_t)p.a_bytes;
    const int STAGES = p.stages;
    uint8_t* sStage = sB + (size_t)STAGES * B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + PSTAGE_BYTES);
    uint64_t* fullA = bars;                              // [2][MAX_CHUNKS]
    uint64_t* emptyA = fullA + 2 * MAX_CHUNKS;           // [2]
    uint64_t* fullB = emptyA + 2;                        // [MAX_PSTAGES]
    uint64_t* emptyB = fullB + MAX_PSTAGES;
    uint64_t* accFull = emptyB + MAX_PSTAGES;            // [2]
    uint64_t* accEmpty = accFull + 2;                    // [2]
    uint64_t* p12 = accEmpty + 2;                        // [2] 3xTF32: passes 1+2 of the unit in window b retired
    uint64_t* loready = p12 + 2;                         // [2] 3xTF32: window b rewritten to x_lo
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(loready + 2);
    float* sBias = reinterpret_cast<float*>(bars) + 384; // 1.5 KB past the start of the barrier block (154 barriers = 1.2 KB)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = pp.items_x * pp.items_y;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a);
        prefetch_tmap(&maps.b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int j = 0; j < 2 * MAX_CHUNKS; ++j) mbar_init(&fullA[j], 1);
            for (int j = 0; j < 2; ++j) { mbar_init(&emptyA[j], NISSUE); mbar_init(&accFull[j], NISSUE); mbar_init(&accEmpty[j], NEPI_P); }
            for (int s = 0; s < MAX_PSTAGES; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], NISSUE); }
            for (int j = 0; j < 2; ++j) { mbar_init(&p12[j], NISSUE); mbar_init(&loready[j], 32 * NREWRITE_P); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // 3xTF32 (p.x3): a UNIT = (item, channel block) = one window.  Units are issued in pairs, P12(u) P12(u+1) P3(u) P3(u+1):
    // P12 = the (x_hi, w_hi) and (x_hi, w_lo) passes from the raw fp32 window (the tensor core truncates = x_hi), P3 = the
    // (x_lo, w_hi) pass after the rewrite warps have turned the window into x_lo in place -- the rewrite of unit u runs under
    // P12(u+1), the one of u+1 under P3(u).  The weight producer walks the same schedule.
    const int my_items = (int)blockIdx.x < n_items ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n_units = my_items * p.cblocks;

    // geometry of work item `item` (same decomposition as conv_halo_kernel's blockIdx)
    auto decode = [&](int item, int& n0, int& h0, int& w0, int& n0c, int& rows_valid, int& cols_valid, int& imgs_valid, int& m_tiles,
                      int& nch) {
        const int bx = item % pp.items_x, by = item / pp.items_x;
        const int tw_i = bx % p.tiles_w, th_i = (bx / p.tiles_w) % p.tiles_h;
        n0 = (bx / (p.tiles_w * p.tiles_h)) * p.TNB;
        h0 = th_i * p.TH; w0 = tw_i * p.TW; n0c = by * BN;
        rows_valid = min(p.TH, p.Hv - h0); cols_valid = min(p.TW, p.Wv - w0); imgs_valid = min(p.TNB, p.N - n0);
        m_tiles = (((imgs_valid - 1) * p.RH + rows_valid - 1) * p.Wp + cols_valid - 1) / 128 + 1;
        // Every item loads ALL chunks of its window, also the ones a bottom-edge tile does not need (they lie outside the
        // image: TMA zero-fills them without touching memory).  Skipping them, as the one-item kernel above does, would
        // leave their mbarriers one phase behind the (u >> 1) & 1 parity the next full item waits with, and that wait
        // would pass on the stale phase (tests/test_persistent_protocol.py::test_skipping_chunks_breaks_the_parity_scheme).
        nch = p.TNB > 1 ? 1 : p.nch;
    };

    if (warp == 0) {
        // ---- TMA producer of the activation windows
        const uint32_t ch_bytes = (uint32_t)p.ch_pix * 128u;
        int u = 0;                                        // activation-window load counter: buffer u & 1, use (u >> 1)
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int n0, h0, w0, n0c, rv, cv, iv, mt, nch;
            decode(item, n0, h0, w0, n0c, rv, cv, iv, mt, nch);
            for (int cb = 0; cb < p.cblocks; ++cb, ++u) {
                const int buf = u & 1;
                mbar_wait(&emptyA[buf], (((uint32_t)(u >> 1)) & 1u) ^ 1u);
                if (elect_one()) {
                    for (int j = 0; j < nch; ++j) {
                        mbar_expect_tx(&fullA[buf * MAX_CHUNKS + j], ch_bytes);
                        tma_load_4d(sA0 + (size_t)buf * p.a_bytes + (size_t)j * ch_bytes, &maps.a, &fullA[buf * MAX_CHUNKS + j], cb * 32,
                                    w0 + p.dw_min, h0 + p.dh_min + j * p.ch_rows, n0);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1 + NISSUE + NEPI_P) {
        // ---- TMA producer of the weight tiles: its own warp, so the ring refills while the window producer waits for a buffer
        if (pp.b_resident) {
            if (elect_one()) {                            // all (half, cb, tap) weight tiles once per CTA (items_y == 1 when resident)
                for (int half = 0; half < (p.x3 ? 2 : 1); ++half)
                    for (int c2 = 0; c2 < p.cblocks; ++c2)
                        for (int tap = 0; tap < p.ntaps; ++tap) {
                            const int s = (half * p.cblocks + c2) * p.ntaps + tap;
                            mbar_expect_tx(&fullB[s], B_BYTES);
                            tma_load_3d(sB + s * B_BYTES, &maps.b, &fullB[s], c2 * 32 + half * p.w_lo_off, 0, p.widx[tap]);
                        }
            }
            __syncwarp();
        } else if (p.x3) {
            int bi = 0; uint32_t bph = 0;
            auto taps = [&](int u, int half) {
                const int n0c = ((int)(blockIdx.x + (u / p.cblocks) * gridDim.x) / pp.items_x) * BN;
                const int cb = u % p.cblocks;
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    mbar_wait(&emptyB[bi], bph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(&fullB[bi], B_BYTES);
                        tma_load_3d(sB + bi * B_BYTES, &maps.b, &fullB[bi], cb * 32 + half * p.w_lo_off, n0c, p.widx[tap]);
                    }
                    __syncwarp();
                    if (++bi == STAGES) { bi = 0; bph ^= 1u; }
                }
            };
            for (int u = 0; u < n_units; u += 2) {
                taps(u, 0); taps(u, 1);
                if (u + 1 < n_units) { taps(u + 1, 0); taps(u + 1, 1); }
                taps(u, 0);
                if (u + 1 < n_units) taps(u + 1, 0);
            }
        } else {
            int bi = 0; uint32_t bph = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int n0c = (item / pp.items_x) * BN;
                for (int cb = 0; cb < p.cblocks; ++cb)
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        const int s = bi;
                        mbar_wait(&emptyB[s], bph ^ 1u);
                        if (elect_one()) {
                            mbar_expect_tx(&fullB[s], B_BYTES);
                            tma_load_3d(sB + s * B_BYTES, &maps.b, &fullB[s], cb * 32, n0c, p.widx[tap]);
                        }
                        __syncwarp();
                        if (++bi == STAGES) { bi = 0; bph ^= 1u; }
                    }
            }
        }
    } else if (warp <= NISSUE) {
        // ---- MMA issuers (two instruction streams: see conv_halo_kernel); issuer iw owns the M tiles iw, iw + 2, ... of every item
        const int iw = warp - 1;
        constexpr uint32_t idesc = idesc_tf32(BN);
        const uint64_t hi = desc_k_sw128_hi();
        const uint32_t b0 = smem_u32(sB) >> 4;
        const bool resident = pp.b_resident != 0;
        const uint32_t my_toff8 = (lane < p.ntaps ? (uint32_t)p.toff[lane] : 0u) * 8u;      // per-tap constants in the lanes (see conv_halo_kernel)
        const int nch_all = p.TNB > 1 ? 1 : p.nch;
        int bi = 0; uint32_t bph = 0;
        int u = 0, it = 0;
        if (resident) {                                       // the weights land once: wait for all of them here, not per tap
            for (int s2 = 0; s2 < (p.x3 ? 2 : 1) * p.cblocks * p.ntaps; ++s2) mbar_wait(&fullB[s2], 0);
        }
        if (p.x3) {
            auto unit_tiles = [&](int u) {
                int n0, h0, w0, n0c, rv, cv, iv, mt, nch;
                decode((int)(blockIdx.x + (u / p.cblocks) * gridDim.x), n0, h0, w0, n0c, rv, cv, iv, mt, nch);
                return mt;
            };
            // one pass over the taps of unit u: half = which weight half (0 = w_hi, 1 = w_lo); pass 0 of channel block 0 overwrites
            auto taps = [&](int u, int half, bool overwrite_first) {
                const int m_tiles = unit_tiles(u);
                const int cb = u % p.cblocks, buf = u & 1;
                const uint32_t tacc = tmem_base + (uint32_t)(((u / p.cblocks) & 1) * pp.acc_cols);
                const uint32_t a0 = smem_u32(sA0 + (size_t)buf * p.a_bytes) >> 4;
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    const uint32_t alo = a0 + __shfl_sync(0xffffffffu, my_toff8, tap);
                    int s2;
                    if (resident) {
                        s2 = (half * p.cblocks + cb) * p.ntaps + tap;
                    } else {
                        s2 = bi;
                        mbar_wait(&fullB[s2], bph);
                        fence_after();
                    }
                    const uint64_t bdesc = hi | (uint64_t)((b0 + (uint32_t)s2 * (uint32_t)(B_BYTES >> 4)) & 0x3FFFu);
                    const uint32_t first = (overwrite_first && tap == 0) ? 0u : 1u;
                    if (elect_one()) {
                        for (int t = iw; t < m_tiles; t += NISSUE) {
                            const uint64_t adesc = hi | (uint64_t)((alo + (uint32_t)t * 1024u) & 0x3FFFu);
                            const uint32_t d = tacc + (uint32_t)(t * BN);
                            mma_tf32(d, adesc, bdesc, idesc, first);
                            mma_tf32(d, adesc + 2, bdesc + 2, idesc, 1u);
                            mma_tf32(d, adesc + 4, bdesc + 4, idesc, 1u);
                            mma_tf32(d, adesc + 6, bdesc + 6, idesc, 1u);
                        }
                        if (!resident) commit(&emptyB[s2]);
                    }
                    __syncwarp();
                    if (!resident && ++bi == STAGES) { bi = 0; bph ^= 1u; }
                }
            };
            auto P12 = [&](int u) {
                const int itl = u / p.cblocks, cb = u % p.cblocks, buf = u & 1;
                const uint32_t aph = ((uint32_t)(u >> 1)) & 1u;
                if (cb == 0) mbar_wait(&accEmpty[itl & 1], (((uint32_t)(itl >> 1)) & 1u) ^ 1u);     // the epilogue has drained this set
                for (int j = 0; j < nch_all; ++j) mbar_wait(&fullA[buf * MAX_CHUNKS + j], aph);
                fence_after();
                taps(u, 0, cb == 0);
                taps(u, 1, false);
                if (elect_one()) commit(&p12[buf]);              // the raw window may be rewritten when these retire
                __syncwarp();
            };
            auto P3 = [&](int u) {
                const int itl = u / p.cblocks, cb = u % p.cblocks, buf = u & 1;
                mbar_wait(&loready[buf], ((uint32_t)(u >> 1)) & 1u);
                fence_after();
                taps(u, 0, false);
                if (elect_one()) {
                    commit(&emptyA[buf]);
                    if (cb == p.cblocks - 1) commit(&accFull[itl & 1]);
                }
                __syncwarp();
            };
            for (int u = 0; u < n_units; u += 2) {
                P12(u);
                if (u + 1 < n_units) P12(u + 1);
                P3(u);
                if (u + 1 < n_units) P3(u + 1);
            }
        } else
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            int n0, h0, w0, n0c, rv, cv, iv, m_tiles, nch;
            decode(item, n0, h0, w0, n0c, rv, cv, iv, m_tiles, nch);
            const int acc = it & 1;
            mbar_wait(&accEmpty[acc], (((uint32_t)(it >> 1)) & 1u) ^ 1u);      // the epilogue has drained this accumulator set
            const uint32_t tacc = tmem_base + (uint32_t)(acc * pp.acc_cols);
            for (int cb = 0; cb < p.cblocks; ++cb, ++u) {
                const int buf = u & 1;
                const uint32_t aph = ((uint32_t)(u >> 1)) & 1u;
                const uint32_t a0 = smem_u32(sA0 + (size_t)buf * p.a_bytes) >> 4;
                // the window was requested one or two items ago: wait for ALL of its chunks up front (no per-tap chunk arithmetic)
                for (int j = 0; j < nch_all; ++j) mbar_wait(&fullA[buf * MAX_CHUNKS + j], aph);
                fence_after();
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    const uint32_t alo = a0 + __shfl_sync(0xffffffffu, my_toff8, tap);
                    int s2;
                    if (resident) {
                        s2 = cb * p.ntaps + tap;
                    } else {
                        s2 = bi;
                        mbar_wait(&fullB[s2], bph);
                        fence_after();
                    }
                    const uint64_t bdesc = hi | (uint64_t)((b0 + (uint32_t)s2 * (uint32_t)(B_BYTES >> 4)) & 0x3FFFu);
                    const uint32_t first = (cb | tap) != 0 ? 1u : 0u;
                    if (elect_one()) {
                        for (int t = iw; t < m_tiles; t += NISSUE) {
                            const uint64_t adesc = hi | (uint64_t)((alo + (uint32_t)t * 1024u) & 0x3FFFu);
                            const uint32_t d = tacc + (uint32_t)(t * BN);
                            mma_tf32(d, adesc, bdesc, idesc, first);
                            mma_tf32(d, adesc + 2, bdesc + 2, idesc, 1u);
                            mma_tf32(d, adesc + 4, bdesc + 4, idesc, 1u);
                            mma_tf32(d, adesc + 6, bdesc + 6, idesc, 1u);
                        }
                        if (!resident) commit(&emptyB[s2]);
                    }
                    __syncwarp();
                    if (!resident && ++bi == STAGES) { bi = 0; bph ^= 1u; }
                }
                if (elect_one()) commit(&emptyA[buf]);           // window free when these MMAs retire
                __syncwarp();
            }
            if (elect_one()) commit(&accFull[acc]);
            __syncwarp();
        }
    } else if (warp >= 2 + NISSUE + NEPI_P) {
        // ---- 3xTF32 rewrite warps: window of unit u -> x_lo = x - trunc(x), rounded to TF32, once its passes 1 + 2 have retired
        if (p.x3) {
            const int tid = threadIdx.x - 32 * (2 + NISSUE + NEPI_P);
            const int nch_all = p.TNB > 1 ? 1 : p.nch;
            const int n4 = nch_all * p.ch_pix * 8;              // float4s of the window (128 B = 8 float4 per position)
            for (int u = 0; u < n_units; ++u) {
                const int buf = u & 1;
                const uint32_t aph = ((uint32_t)(u >> 1)) & 1u;
                for (int j = 0; j < nch_all; ++j) mbar_wait(&fullA[buf * MAX_CHUNKS + j], aph);
                mbar_wait(&p12[buf], aph);
                float4* w4 = reinterpret_cast<float4*>(sA0 + (size_t)buf * p.a_bytes);
                for (int i = tid; i < n4; i += 32 * NREWRITE_P) {
                    float4 v = w4[i];
                    v.x = lo_tf32(v.x); v.y = lo_tf32(v.y); v.z = lo_tf32(v.z); v.w = lo_tf32(v.w);
                    w4[i] = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core (async proxy) reads
                mbar_arrive(&loready[buf]);
            }
        }
    } else {
        // ---- epilogue warps: drain accumulator set (it & 1) while the MMA warp fills the other one
        // warp w may touch TMEM lanes 32*(w%4)..+31: warps 3..6 and 7..10 cover the four quarters twice; the second set takes the odd tiles
        const int ew = warp - 1 - NISSUE, t_first = ew >> 2, t_step = NEPI_P / 4;
        float* stage = reinterpret_cast<float*>(sStage) + ew * 1024;
        int it = 0, last_n0c = -1;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            int n0, h0, w0, n0c, rows_valid, cols_valid, imgs_valid, m_tiles, nch;
            decode(item, n0, h0, w0, n0c, rows_valid, cols_valid, imgs_valid, m_tiles, nch);
            const int acc = it & 1;
            if (n0c != last_n0c) {                               // bias of this output-channel block (4 epilogue warps = 128 threads)
                asm volatile("bar.sync 1, %0;" ::"n"(32 * NEPI_P) : "memory");
                const int tI = threadIdx.x - 32 * (1 + NISSUE);
                if (tI < BN) sBias[tI] = p.bias ? p.bias[n0c + tI] : 0.f;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * NEPI_P) : "memory");
                last_n0c = n0c;
            }
            mbar_wait(&accFull[acc], ((uint32_t)(it >> 1)) & 1u);
            fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(acc * pp.acc_cols);
            switch (p.act) {
                case G2_ACT_RELU: epilogue<BN, G2_ACT_RELU>(p, tacc, sBias, stage, m_tiles, n0, h0, w0, n0c, imgs_valid, rows_valid, cols_valid, t_first, t_step); break;
                case G2_ACT_ELU: epilogue<BN, G2_ACT_ELU>(p, tacc, sBias, stage, m_tiles, n0, h0, w0, n0c, imgs_valid, rows_valid, cols_valid, t_first, t_step); break;
                case G2_ACT_SIGMOID: epilogue<BN, G2_ACT_SIGMOID>(p, tacc, sBias, stage, m_tiles, n0, h0, w0, n0c, imgs_valid, rows_valid, cols_valid, t_first, t_step); break;
                default: epilogue<BN, G2_ACT_NONE>(p, tacc, sBias, stage, m_tiles, n0, h0, w0, n0c, imgs_valid, rows_valid, cols_valid, t_first, t_step); break;
            }
            fence_before();                                      // our tcgen05.ld's are complete (wait::ld inside) before we hand the set back
            __syncwarp();
            if (lane == 0) mbar_arrive(&accEmpty[acc]);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled get_encode() {
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

static bool encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, bool raw_fp32 = false) {
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    // TFLOAT32: the TMA rounds fp32 to TF32 on the way in; FLOAT32 (3xTF32 mode): the raw bits land and the tensor core truncates
    CUresult r = enc(m, raw_fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static int gcd(int a, int b) { return b ? gcd(b, a % b) : a; }

// tunables (environment, read once): shared-memory budget per CTA in KB (113 = two CTAs per SM), TMEM columns per CTA
static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}
static int g_enabled = -1;
static long long* g_dbg = nullptr;
static int budget_bytes() { static int v = env_int("G2_HALO_SMEM_KB", 112) * 1024; return v; }
static int epi_mode() { static int v = env_int("G2_HALO_EPI", 0); return v; }
static int max_cols() { static int v = env_int("G2_HALO_TMEM_COLS", 256); return v; }

// Weight-ring depth (tunable by environment for experiments).  Measured on B200 (gpurun_out/conv_bench_v6.txt): deeper rings
// (6/4/3, 9/5/3, 12/6/4) are 0-25 % SLOWER than 4/2/2 because the shared memory they take shrinks the activation tile.
static int stages_of(int BN, int ntaps, int cblocks) {
    static int c32 = env_int("G2_HALO_STAGES_32", 4), c64 = env_int("G2_HALO_STAGES_64", 2), c128 = env_int("G2_HALO_STAGES_128", 2);
    int cap = BN == 32 ? c32 : BN == 64 ? c64 : c128;
    if (cap > MAX_STAGES) cap = MAX_STAGES;
    if (cap < 1) cap = 1;
    const int tiles = ntaps * cblocks;
    return tiles < cap ? tiles : cap;
}

struct Geo {
    int TH, TW, TNB, RH, Wp, ch_rows, nch, m, a_bytes, tiles_h, tiles_w;
    double cost;     // estimated SM clocks per image
};

static inline int m_of(int imgs, int RH, int rows, int Wp, int cols) { return (((imgs - 1) * RH + rows - 1) * Wp + cols - 1) / 128 + 1; }

// Choose the CTA tile for one (class of a) convolution: TH x TW output pixels of one image (or TNB whole small
// images), window pitch Wp (exact, or rounded up to 8 pixels so that any row count keeps chunks 1024-byte aligned).
// The cost model is the per-SM time of all CTAs of an image: max(MMA clocks, L2->SMEM clocks) + a fixed per-CTA cost.
static bool pick_geo(int N, int Hv, int Wv, int span_h, int span_w, int ntaps, int cblocks, int BN, int stages, Geo* best,
                     int a_budget = -1, int col_limit = -1) {
    const int fixed = stages * BN * 128 + 1024 + 1024;
    const double clk_mma = 4.0 * (BN == 32 ? 40 : BN == 64 ? 48 : 64) * ntaps * cblocks;   // per M tile (SMEM-read bound for small N)
    const double l2_rate = 28.0;           // bytes / clock / SM sustained by TMA tile loads (measured 21-33)
    const double cta_fixed = 1500.0;       // prologue + pipeline fill + epilogue tail not hidden by the sibling CTA
    const double b_bytes = (double)ntaps * cblocks * BN * 128;
    bool found = false;
    for (int nsplit = 1; nsplit <= 4; ++nsplit) {
        const int TW = (Wv + nsplit - 1) / nsplit;
        const int tiles_w = (Wv + TW - 1) / TW;
        if (tiles_w != nsplit) continue;
        const int cols_last = Wv - (tiles_w - 1) * TW;
        for (int variant = 0; variant < 2; ++variant) {
            int Wp = TW + span_w;
            if (variant == 1) {
                if (Wp % 8 == 0) continue;
                Wp = (Wp + 7) / 8 * 8;
            }
            if (Wp > 256) continue;
            const int gran = 8 / gcd(Wp, 8);                     // rows per chunk must keep chunk bases 1024-byte aligned
            const int off_max = span_h * Wp + span_w;
            auto consider = [&](int TH, int TNB) {
                Geo g;
                g.TH = TH; g.TW = TW; g.TNB = TNB; g.RH = TH + span_h; g.Wp = Wp; g.tiles_w = tiles_w;
                g.m = m_of(TNB, g.RH, TH, Wp, TW);
                if (g.m * BN > (col_limit > 0 ? col_limit : max_cols())) return;
                int loaded_pix;
                if (TNB > 1) {
                    if (g.RH > 256) return;
                    g.ch_rows = g.RH; g.nch = 1;
                    loaded_pix = g.RH * Wp * TNB;
                } else {
                    // fewest loaded rows with at most MAX_CHUNKS chunks; ties -> more (smaller) chunks for earlier MMA start
                    int best_rows = 1 << 30;
                    g.ch_rows = 0; g.nch = 0;
                    for (int cr = gran; cr <= 256; cr += gran) {
                        const int nch = (g.RH + cr - 1) / cr;
                        if (nch > MAX_CHUNKS) continue;
                        if (nch * cr < best_rows) { best_rows = nch * cr; g.ch_rows = cr; g.nch = nch; }
                        if (nch == 1) break;
                    }
                    if (g.nch == 0) return;
                    loaded_pix = best_rows * Wp;
                }
                int alloc_pix = loaded_pix;
                if (128 * g.m + off_max > alloc_pix) alloc_pix = 128 * g.m + off_max;
                g.a_bytes = ((alloc_pix * 128 + 1023) / 1024) * 1024;
                if (g.a_bytes < 128 * 144) g.a_bytes = 128 * 144;      // epilogue staging: 128 rows x 144 B
                if (a_budget > 0 ? g.a_bytes > a_budget : g.a_bytes + fixed > budget_bytes()) return;
                double tiles, ctas, bytes;
                if (TNB > 1) {
                    g.tiles_h = 1;
                    const int groups = (N + TNB - 1) / TNB;
                    tiles = (double)groups * g.m / N;
                    ctas = (double)groups / N;
                    bytes = ctas * (cblocks * loaded_pix * 128.0 + b_bytes);
                } else {
                    g.tiles_h = (Hv + TH - 1) / TH;
                    const int rows_last = Hv - (g.tiles_h - 1) * TH;
                    tiles = (double)(g.tiles_h - 1) * ((tiles_w - 1) * m_of(1, g.RH, TH, Wp, TW) + m_of(1, g.RH, TH, Wp, cols_last)) +
                            ((tiles_w - 1) * m_of(1, g.RH, rows_last, Wp, TW) + m_of(1, g.RH, rows_last, Wp, cols_last));
                    ctas = (double)g.tiles_h * tiles_w;
                    bytes = ctas * (cblocks * loaded_pix * 128.0 + b_bytes);
                }
                const double mma = tiles * clk_mma, load = bytes / l2_rate;
                g.cost = (mma > load ? mma : load) + 0.25 * (mma > load ? load : mma) + ctas * cta_fixed;
                if (!found || g.cost < best->cost) {
                    *best = g;
                    found = true;
                }
            };
            for (int TH = 1; TH <= Hv; ++TH) consider(TH, 1);
            if (nsplit == 1)
                for (int TNB = 2; TNB <= 16 && TNB <= N; ++TNB) consider(Hv, TNB);
        }
    }
    return found;
}

template <int BN>
static int launch(const Maps& maps, const P& p, dim3 grid, cudaStream_t stream) {
    const int smem = p.a_bytes + p.stages * BN * 128 + 1024 + 1024;
    static G2DevOnce once;
    if (once.needed()) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        once.done();
    }
    conv_halo_kernel<BN><<<grid, 224, smem, stream>>>(maps, p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? G2_OK : (int)e;
}

// 0: one-item kernel only; 1: persistent kernel wherever its geometry fits; 2 (default): persistent except where the one-item
// kernel measured faster (profiles/r02_conv_bench_persistent_2issuers.txt): weights streamed through the ring (not resident)
// for two or more channel blocks of a filter with more than 9 taps -- 50+ weight tiles per item, a barrier round trip per tap.
static int persistent_mode() { static int v = env_int("G2_HALO_PERSISTENT", 2); return v; }
// 3xTF32 launches on the one-item kernel (0, default) or on the persistent kernel (1).  Measured on a B200
// (profiles/r02_conv_bench_x3_persistent.txt): with three passes per window the one-item kernel's load / epilogue phases are
// already amortised (75 % of its shared-memory bound) and the persistent variant only ties (+5 % / -5 % per layer) when its
// weights stream through the ring, and loses 20-30 % when the [w_hi | w_lo] pack is made resident at the expense of window size.
static int x3_persistent() { static int v = env_int("G2_HALO_X3_PERSISTENT", 0); return v; }
// CTAs of the persistent launch (one per SM by default; the CPU emulation lowers it to give every CTA several items)
static int persistent_ctas() { static int v = env_int("G2_HALO_PERSISTENT_CTAS", 148); return v < 1 ? 1 : v; }

template <int BN>
static int launch_persistent(const Maps& maps, const PP& pp, int n_ctas, cudaStream_t stream) {
    const int smem = 2 * pp.p.a_bytes + pp.p.stages * BN * 128 + PSTAGE_BYTES + 2048 + 512 + 1024;
    if (smem > 227 * 1024) return G2_ERR_UNSUPPORTED;
    static G2DevOnce once;
    if (once.needed()) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_persistent_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        once.done();
    }
    conv_halo_persistent_kernel<BN><<<n_ctas, 32 * (2 + NISSUE + NEPI_P + NREWRITE_P), smem, stream>>>(maps, pp);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? G2_OK : (int)e;
}

struct TapSet;
static bool persistent_geo(int N, const TapSet& t, int Ci, int Co, int BN, int sh, int sw, Geo* g, int* pstages, bool* resident, int x3 = 0);

static int pick_bn(int Co) {
    if (Co == 32 || Co == 64 || Co == 128) return Co;
    if (Co % 128 == 0) return 128;
    if (Co % 64 == 0) return 64;
    return 0;
}

struct TapSet { int n; int dh[MAX_TAPS], dw[MAX_TAPS], widx[MAX_TAPS]; int os, ph, pw, Hv, Wv; };

// Enumerate the tap sets (1, or 4 sub-pixel classes) of a problem; returns the number of classes or 0 if the
// halo kernel does not cover it.
static int tap_sets(int Hi, int Wi, int Ho, int Wo, int R, int S, int stride, int pad, int mode, TapSet* ts) {
    if (R * S > MAX_TAPS) return 0;
    if (stride == 1) {
        TapSet& t = ts[0];
        t.n = 0; t.os = 1; t.ph = 0; t.pw = 0; t.Hv = Ho; t.Wv = Wo;
        for (int r = 0; r < R; ++r)
            for (int s = 0; s < S; ++s) {
                t.dh[t.n] = mode == 0 ? r - pad : pad - r;
                t.dw[t.n] = mode == 0 ? s - pad : pad - s;
                t.widx[t.n] = r * S + s;
                ++t.n;
            }
        return 1;
    }
    if (stride == 2 && mode == 1) {
        if ((Ho | Wo) & 1) return 0;
        int nc = 0;
        for (int cls = 0; cls < 4; ++cls) {
            TapSet& t = ts[nc];
            t.n = 0; t.os = 2; t.ph = cls >> 1; t.pw = cls & 1; t.Hv = Ho / 2; t.Wv = Wo / 2;
            for (int r = 0; r < R; ++r)
                for (int s = 0; s < S; ++s) {
                    const int th = t.ph + pad - r, tw = t.pw + pad - s;
                    if ((th & 1) || (tw & 1)) continue;
                    t.dh[t.n] = th / 2; t.dw[t.n] = tw / 2; t.widx[t.n] = r * S + s;
                    ++t.n;
                }
            if (t.n == 0) return 0;        // a class without taps would need a bias-only fill: leave it to the tile kernel
            ++nc;
        }
        return nc;
    }
    return 0;
}

static void spans(const TapSet& t, int* dh_min, int* dw_min, int* span_h, int* span_w) {
    int a = t.dh[0], b = t.dh[0], c = t.dw[0], d = t.dw[0];
    for (int i = 1; i < t.n; ++i) {
        a = t.dh[i] < a ? t.dh[i] : a; b = t.dh[i] > b ? t.dh[i] : b;
        c = t.dw[i] < c ? t.dw[i] : c; d = t.dw[i] > d ? t.dw[i] : d;
    }
    *dh_min = a; *dw_min = c; *span_h = b - a; *span_w = d - c;
}

// Geometry of the experimental persistent variant: weights resident when all (cb, tap) tiles fit beside two windows, else a
// deeper ring; two windows and two accumulator sets (<= 256 TMEM columns each) per CTA.  false: use the one-item kernel.
static bool persistent_geo(int N, const TapSet& t, int Ci, int Co, int BN, int sh, int sw, Geo* g, int* pstages, bool* resident, int x3) {
    if (!persistent_mode()) return false;
    if (x3 && !x3_persistent()) return false;
    const int b_all = (x3 ? 2 : 1) * t.n * (Ci / 32);          // 3xTF32: [w_hi | w_lo] halves
    *resident = Co == BN && b_all <= MAX_PSTAGES && b_all * BN * 128 <= 72 * 1024;
    static int x3_res = env_int("G2_HALO_X3_RESIDENT", 0);
    if (x3 && !x3_res) *resident = false;
    // streamed weights: ring of `ring_kb` KB (tiles of BN x 128 B).  Measured (profiles/r02_conv_bench_ring.txt): 64 / 96 KB rings are
    // 20-55 % SLOWER than 32 KB on the 5x5 layers -- the shared memory is worth more as activation window (fewer, larger items)
    static int ring_kb = env_int("G2_HALO_RING_KB", 32);
    *pstages = *resident ? b_all : ring_kb * 1024 / (BN * 128);
    if (!*resident && *pstages > MAX_PSTAGES) *pstages = MAX_PSTAGES;
    if (!*resident && *pstages < 2) *pstages = 2;
    if (persistent_mode() == 2 && !*resident && Ci / 32 >= 2 && t.n > 9) return false;
    // experiments: shared memory per CTA (113 KB -> two persistent CTAs per SM) and TMEM columns per accumulator set
    static int smem_kb = env_int("G2_HALO_PERSISTENT_SMEM_KB", 227), cols = env_int("G2_HALO_PERSISTENT_COLS", 256);
    const int a_budget = (smem_kb * 1024 - 1024 - *pstages * BN * 128 - PSTAGE_BYTES - 2048 - 512) / 2;
    return a_budget >= 128 * 144 && pick_geo(N, t.Hv, t.Wv, sh, sw, t.n, Ci / 32, BN, *pstages, g, a_budget, cols);
}

static bool enabled() {
    if (g_enabled < 0) g_enabled = env_int("G2_HALO", 1) ? 1 : 0;
    return g_enabled == 1;
}

}  // namespace halo

extern "C" {

// Runtime switch for A/B measurements (scripts/conv_bench.py): returns the previous setting.
int g2_conv_halo_enable(int on) {
    const int prev = halo::enabled() ? 1 : 0;
    halo::g_enabled = on ? 1 : 0;
    return prev;
}

// Debug only (scripts/conv_bench.py --timeline): device buffer of [CTAs][64] int64 that subsequent halo launches fill with
// clock64() at {start, prologue done, first operands landed, last MMA issued, accumulators complete, epilogue done, -, smid};
// pass NULL to switch it off.  Returns 0.
int g2_conv_halo_debug(int64_t* buf) {
    static_assert(sizeof(long long) == sizeof(int64_t), "");
    halo::g_dbg = reinterpret_cast<long long*>(buf);
    return 0;
}

// 1 if the halo kernel takes this problem (g2_conv_igemm_tf32 then routes to it), else 0.
int g2_conv_halo_supported(int N, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode) {
    using namespace halo;
    if (!enabled()) return 0;
    if (Ci % 32 != 0 || N <= 0 || (long)N * Ho * Wo >= (1L << 31)) return 0;
    const int BN = pick_bn(Co);
    if (BN == 0) return 0;
    TapSet ts[4];
    const int nc = tap_sets(Hi, Wi, Ho, Wo, R, S, stride, pad, mode, ts);
    if (nc == 0) return 0;
    for (int c = 0; c < nc; ++c) {
        if ((long)ts[c].Hv * ts[c].Wv < 256) return 0;       // tiny maps: the dense-tile kernel packs them better
        int dh_min, dw_min, sh, sw;
        spans(ts[c], &dh_min, &dw_min, &sh, &sw);
        Geo g;
        if (!pick_geo(N, ts[c].Hv, ts[c].Wv, sh, sw, ts[c].n, Ci / 32, BN, stages_of(BN, ts[c].n, Ci / 32), &g)) return 0;
    }
    return 1;
}

static int conv_halo_impl(const float* in, const float* w, const float* bias, float* out, int N, int Hi, int Wi, int Ci,
                          int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode, int act, int x3, cudaStream_t stream);

// Same contract as g2_conv_igemm_tf32 for the problems g2_conv_halo_supported accepts.
int g2_conv_halo_tf32(const float* in, const float* w, const float* bias, float* out, int N, int Hi, int Wi, int Ci,
                      int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode, int act, cudaStream_t stream) {
    return conv_halo_impl(in, w, bias, out, N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode, act, 0, stream);
}

// 3xTF32 form of the same convolution (fp32-level accuracy on the tensor cores): `w` is the pack [R*S][Co][2*Ci] whose channel
// halves are w_hi = the weights as exact TF32 values and w_lo = w - w_hi; `in` is the plain fp32 activation.  The kernel issues
// (x_hi, w_hi), (x_hi, w_lo), rewrites its activation window in place to x_lo = x - trunc(x) and issues (x_lo, w_hi).
int g2_conv_halo_x3_tf32(const float* in, const float* w, const float* bias, float* out, int N, int Hi, int Wi, int Ci,
                         int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode, int act, cudaStream_t stream) {
    return conv_halo_impl(in, w, bias, out, N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode, act, 1, stream);
}

static int conv_halo_impl(const float* in, const float* w, const float* bias, float* out, int N, int Hi, int Wi, int Ci,
                          int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode, int act, int x3, cudaStream_t stream) {
    using namespace halo;
    G2_CHECK_ARG(in && w && out && N > 0);
    G2_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (Ci % 32 != 0 || (long)N * Ho * Wo >= (1L << 31)) return G2_ERR_UNSUPPORTED;
    const int BN = pick_bn(Co);
    if (BN == 0) return G2_ERR_UNSUPPORTED;
    TapSet ts[4];
    const int nc = tap_sets(Hi, Wi, Ho, Wo, R, S, stride, pad, mode, ts);
    if (nc == 0) return G2_ERR_UNSUPPORTED;
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    {
        const int Cw = x3 ? 2 * Ci : Ci;                 // channels of the weight pack (3xTF32: [w_hi | w_lo])
        cuuint64_t dims[3] = {(cuuint64_t)Cw, (cuuint64_t)Co, (cuuint64_t)(R * S)};
        cuuint64_t str[2] = {(cuuint64_t)Cw * 4, (cuuint64_t)Cw * Co * 4};
        cuuint32_t box[3] = {32, (uint32_t)BN, 1};
        if (!encode(&maps.b, w, 3, dims, str, box)) return G2_ERR_UNSUPPORTED;
    }
    for (int c = 0; c < nc; ++c) {
        const TapSet& t = ts[c];
        int dh_min, dw_min, sh, sw;
        spans(t, &dh_min, &dw_min, &sh, &sw);
        Geo g;
        bool resident = false;
        int pstages = 0;
        const bool use_p = persistent_geo(N, t, Ci, Co, BN, sh, sw, &g, &pstages, &resident, x3);
        if (!use_p && !pick_geo(N, t.Hv, t.Wv, sh, sw, t.n, Ci / 32, BN, stages_of(BN, t.n, Ci / 32), &g)) return G2_ERR_UNSUPPORTED;
        P p;
        memset(&p, 0, sizeof(p));
        p.out = out; p.bias = bias; p.N = N; p.Hv = t.Hv; p.Wv = t.Wv; p.TH = g.TH; p.TW = g.TW; p.TNB = g.TNB;
        p.Wp = g.Wp; p.RH = g.RH; p.ch_rows = g.ch_rows; p.nch = g.nch; p.ch_pix = g.ch_rows * p.Wp * g.TNB;
        p.tiles_h = g.tiles_h; p.tiles_w = g.tiles_w; p.dh_min = dh_min; p.dw_min = dw_min; p.span_h = sh;
        p.Ho = Ho; p.Wo = Wo; p.Co = Co; p.os = t.os; p.ph = t.ph; p.pw = t.pw;
        p.ntaps = t.n; p.cblocks = Ci / 32; p.act = act; p.a_bytes = g.a_bytes; p.stages = stages_of(BN, t.n, Ci / 32); p.dbg = g_dbg; p.epi = epi_mode();
        p.x3 = x3; p.w_lo_off = Ci;
        int cols = 32;
        while (cols < g.m * BN) cols <<= 1;
        p.tmem_cols = cols;
        for (int i = 0; i < t.n; ++i) {
            p.toff[i] = (t.dh[i] - dh_min) * p.Wp + (t.dw[i] - dw_min);
            p.widx[i] = (short)t.widx[i];
        }
        {
            cuuint64_t dims[4] = {(cuuint64_t)Ci, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)N};
            cuuint64_t str[3] = {(cuuint64_t)Ci * 4, (cuuint64_t)Wi * Ci * 4, (cuuint64_t)Hi * Wi * Ci * 4};
            cuuint32_t box[4] = {32, (uint32_t)p.Wp, (uint32_t)g.ch_rows, (uint32_t)g.TNB};
            if (!encode(&maps.a, in, 4, dims, str, box, x3 != 0)) return G2_ERR_UNSUPPORTED;
        }
        dim3 grid((unsigned)(((N + g.TNB - 1) / g.TNB) * g.tiles_h * g.tiles_w), (unsigned)(Co / BN), 1);
        int rc;
        if (use_p) {
            PP pp;
            pp.p = p;
            pp.p.stages = pstages;
            pp.items_x = (int)grid.x; pp.items_y = (int)grid.y;
            pp.acc_cols = g.m * BN;
            pp.b_resident = resident ? 1 : 0;
            int pc = 32;
            while (pc < 2 * pp.acc_cols) pc <<= 1;
            pp.p.tmem_cols = pc;
            const long items = (long)grid.x * grid.y;
            const int n_ctas = (int)(items < persistent_ctas() ? items : persistent_ctas());
            switch (BN) {
                case 32: rc = launch_persistent<32>(maps, pp, n_ctas, stream); break;
                case 64: rc = launch_persistent<64>(maps, pp, n_ctas, stream); break;
                default: rc = launch_persistent<128>(maps, pp, n_ctas, stream); break;
            }
            if (rc != G2_OK) return rc;
            continue;
        }
        switch (BN) {
            case 32: rc = launch<32>(maps, p, grid, stream); break;
            case 64: rc = launch<64>(maps, p, grid, stream); break;
            default: rc = launch<128>(maps, p, grid, stream); break;
        }
        if (rc != G2_OK) return rc;
    }
    return G2_OK;
}

// Launch plan of class `cls` for tests (tests/test_halo_plan.py replays it in numpy) and DESIGN.md:
// plan[0..21] = {nclasses, TH, TNB, RH, ch_rows, nch, m_tiles, a_bytes, tiles_h, Wp, dh_min, dw_min, os, ph, pw, Hv, Wv,
//                ntaps, BN, smem_bytes, TW, tiles_w}, plan[32+i] = toff[i], plan[64+i] = widx[i].
int g2_conv_halo_plan(int N, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int R, int S, int stride, int pad, int mode,
                      int cls, int* plan) {
    using namespace halo;
    const int BN = pick_bn(Co);
    TapSet ts[4];
    if (BN == 0 || !plan) return G2_ERR_ARG;
    const int nc = tap_sets(Hi, Wi, Ho, Wo, R, S, stride, pad, mode, ts);
    if (nc == 0 || cls < 0 || cls >= nc) return G2_ERR_UNSUPPORTED;
    const TapSet& t = ts[cls];
    int dh_min, dw_min, sh, sw;
    spans(t, &dh_min, &dw_min, &sh, &sw);
    Geo g;
    bool resident = false;
    int pstages = 0;
    // with G2_HALO_PERSISTENT=1 the plan is the persistent variant's (plan[22] = 1, [23] = resident weights, [24] = weight
    // stages, [25] = TMEM columns of the two accumulator sets), so the same numpy replay validates its geometry
    const bool use_p = persistent_geo(N, t, Ci, Co, BN, sh, sw, &g, &pstages, &resident);
    if (!use_p && !pick_geo(N, t.Hv, t.Wv, sh, sw, t.n, Ci / 32, BN, stages_of(BN, t.n, Ci / 32), &g)) return G2_ERR_UNSUPPORTED;
    const int Wp = g.Wp;
    const int smem = use_p ? 2 * g.a_bytes + pstages * BN * 128 + PSTAGE_BYTES + 2048 + 512 + 1024
                           : g.a_bytes + stages_of(BN, t.n, Ci / 32) * BN * 128 + 1024 + 1024;
    const int v[22] = {nc, g.TH, g.TNB, g.RH, g.ch_rows, g.nch, g.m, g.a_bytes, g.tiles_h, Wp, dh_min, dw_min, t.os, t.ph, t.pw,
                       t.Hv, t.Wv, t.n, BN, smem, g.TW, g.tiles_w};
    for (int i = 0; i < 96; ++i) plan[i] = 0;
    for (int i = 0; i < 22; ++i) plan[i] = v[i];
    plan[22] = use_p ? 1 : 0; plan[23] = resident ? 1 : 0; plan[24] = pstages; plan[25] = use_p ? 2 * g.m * BN : g.m * BN;
    for (int i = 0; i < t.n; ++i) {
        plan[32 + i] = (t.dh[i] - dh_min) * Wp + (t.dw[i] - dw_min);
        plan[64 + i] = t.widx[i];
    }
    return G2_OK;
}

}  // extern "C"
