// TEST INFRASTRUCTURE: run plain-SIMT CUDA kernels of genesis_b200/csrc on the CPU, one OS thread per CUDA thread.
//
// The kernel SOURCE is compiled unchanged with g++ (tests/test_cuda_emu.py extracts the kernel namespace of a .cu file and
// wraps it): the CUDA keywords become no-ops, threadIdx / blockIdx are thread-local, __shared__ becomes `static` (blocks run one
// after another, so one copy per kernel instantiation is the block's shared memory), __syncthreads is a barrier over the
// block's threads, warp shuffles exchange through a per-warp slot array between two warp barriers.  This checks what a
// kernel written without a GPU at hand is most likely to get wrong -- index arithmetic, tile edges, barrier placement, the
// order of a scan -- against numpy/torch references.  It says nothing about performance, and it cannot run the tcgen05 /
// TMA kernels.  Nothing in the product uses it.
#pragma once
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

namespace emu {
// reusable barrier (C++17 has no std::barrier)
class Barrier {
public:
    explicit Barrier(int n) : n_(n), count_(0), gen_(0) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        const int g = gen_;
        if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen_ != g; });
    }
private:
    int n_, count_, gen_;
    std::mutex m_;
    std::condition_variable cv_;
};
struct Block {
    Barrier* all;
    std::vector<Barrier*> warps;
    std::vector<std::vector<uint64_t>> slots;   // [warp][lane]
};
extern thread_local Block* block;
extern std::mutex atomic_mutex;
extern unsigned char dyn_smem[256 * 1024];       // 1024-byte aligned: offsets into it ARE the shared-memory addresses
}  // namespace emu

extern thread_local dim3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

#ifdef CUDA_EMU_MAIN
thread_local dim3 threadIdx, blockIdx;
dim3 blockDim, gridDim;
namespace emu {
thread_local Block* block = nullptr;
std::mutex atomic_mutex;
alignas(1024) unsigned char dyn_smem[256 * 1024];
}
#endif

static inline void __syncthreads() { emu::block->all->wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::block->warps[threadIdx.x >> 5]->wait(); }

template <class T> static inline T __ldg(const T* p) { return *p; }
// CUDA's global min / max overloads
template <class T> static inline T min(T a, T b) { return b < a ? b : a; }
template <class T> static inline T max(T a, T b) { return a < b ? b : a; }
static inline long min(long a, int b) { return a < b ? a : b; }
static inline long min(int a, long b) { return a < b ? a : b; }
static inline long max(long a, int b) { return a > b ? a : b; }
static inline long max(int a, long b) { return a > b ? a : b; }

template <class T> static inline T emu_shfl(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 8 bytes");
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    emu::block->slots[w][lane] = raw;
    emu::block->warps[w]->wait();
    const uint64_t got = emu::block->slots[w][src_lane & 31];
    emu::block->warps[w]->wait();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int o) { return emu_shfl(v, (int)(threadIdx.x & 31) ^ o); }
template <class T> static inline T __shfl_sync(unsigned, T v, int lane) { return emu_shfl(v, lane); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
    const int lane = threadIdx.x & 31;
    return emu_shfl(v, lane + d < 32 ? lane + d : lane);
}

static inline float atomicAdd(float* p, float v) {
    std::lock_guard<std::mutex> lk(emu::atomic_mutex);
    const float old = *p; *p = old + v; return old;
}
static inline double atomicAdd(double* p, double v) {
    std::lock_guard<std::mutex> lk(emu::atomic_mutex);
    const double old = *p; *p = old + v; return old;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) {
    std::lock_guard<std::mutex> lk(emu::atomic_mutex);
    const unsigned old = *p; *p = old + v; return old;
}
static inline int atomicAdd(int* p, int v) {
    std::lock_guard<std::mutex> lk(emu::atomic_mutex);
    const int old = *p; *p = old + v; return old;
}
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float rsqrtf(float x) { return 1.f / sqrtf(x); }
#define __expf(x) expf(x)
#define __logf(x) logf(x)
static inline float __fdividef(float a, float b) { return a / b; }

namespace emu {
// Run `body` as a grid of `grid` blocks of `block` threads (1-D blocks; x must be a multiple of 32), blocks one after another.
template <class F> static inline void launch(dim3 grid, dim3 blk, F body) {
    ::blockDim = blk; ::gridDim = grid;
    const int n = (int)(blk.x * blk.y * blk.z);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                Block b;
                Barrier all(n);
                b.all = &all;
                std::vector<Barrier> wb;
                const int nw = (n + 31) / 32;
                std::vector<std::unique_ptr<Barrier>> own;
                for (int w = 0; w < nw; ++w) {
                    const int lanes = (w + 1) * 32 <= n ? 32 : n - w * 32;
                    own.emplace_back(new Barrier(lanes));
                    b.warps.push_back(own.back().get());
                    b.slots.emplace_back(32, 0);
                }
                std::vector<std::thread> ts;
                ts.reserve(n);
                for (int t = 0; t < n; ++t)
                    ts.emplace_back([&, t] {
                        emu::block = &b;
                        ::threadIdx = dim3((unsigned)t % blk.x, ((unsigned)t / blk.x) % blk.y, (unsigned)t / (blk.x * blk.y));
                        ::blockIdx = dim3(bx, by, bz);
                        body();
                    });
                for (auto& th : ts) th.join();
            }
}
}  // namespace emu
