// TEST INFRASTRUCTURE: run the CUDA kernels of genesis_b200/csrc on the CPU, one fiber (ucontext) per CUDA thread.
//
// The kernel SOURCE is compiled unchanged with g++ (tests/test_cuda_emu.py extracts the kernel namespace of a .cu file and
// wraps it): the CUDA keywords become no-ops, threadIdx / blockIdx are thread-local, __shared__ becomes `static` (blocks run one
// after another, so one copy per kernel instantiation is the block's shared memory), __syncthreads is a barrier over the
// block's threads, warp shuffles exchange through a per-warp slot array between two warp barriers.  The threads of a block are
// cooperative fibers on ONE OS thread, scheduled round-robin and switched only at barriers / shuffles / mbarrier waits: runs are
// deterministic, a kernel without barriers costs well under a microsecond per CUDA thread, and a round in which no fiber makes
// progress is reported as a deadlock instead of hanging.  This checks what a
// kernel written without a GPU at hand is most likely to get wrong -- index arithmetic, tile edges, barrier placement, the
// order of a scan -- against numpy/torch references.  It says nothing about performance, and it cannot run the tcgen05 /
// TMA kernels.  Nothing in the product uses it.
#pragma once
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <ucontext.h>
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
struct Fiber {
    ucontext_t ctx;
    dim3 tid;
    bool done;
    const char* waiting;        // what the fiber is blocked on (for the deadlock report)
    unsigned wait_a, wait_b;
};
extern ucontext_t sched_ctx;
extern Fiber* cur;
extern long progress;           // bumped by every state change another fiber could be waiting for
static inline void yield(const char* why = "barrier", unsigned a = 0, unsigned b = 0) {
    cur->waiting = why; cur->wait_a = a; cur->wait_b = b;
    swapcontext(&cur->ctx, &sched_ctx);
}
// reusable barrier over n fibers
class Barrier {
public:
    explicit Barrier(int n) : n_(n), count_(0), gen_(0) {}
    void wait() {
        ++progress;
        if (++count_ == n_) { count_ = 0; ++gen_; }
        else { const long g = gen_; while (gen_ == g) yield("barrier", (unsigned)n_, (unsigned)count_); }
    }
private:
    int n_, count_;
    long gen_;
};
struct Block {
    Barrier* all;
    std::vector<Barrier*> warps;
    std::vector<std::vector<uint64_t>> slots;   // [warp][lane]
};
extern thread_local Block* block;
extern std::mutex atomic_mutex;
extern unsigned char dyn_smem[256 * 1024];       // 1024-byte aligned: offsets into it ARE the shared-memory addresses
extern std::function<void()>* body_fn;
void trampoline();
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
ucontext_t sched_ctx;
Fiber* cur = nullptr;
long progress = 0;
std::function<void()>* body_fn = nullptr;
void trampoline() {
    (*body_fn)();
    cur->done = true;
    swapcontext(&cur->ctx, &sched_ctx);
}
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
// Run `body` as a grid of `grid` blocks of `blk` threads (block size a multiple of 32), blocks one after another.
template <class F> static inline void launch(dim3 grid, dim3 blk, F body) {
    ::blockDim = blk; ::gridDim = grid;
    const int n = (int)(blk.x * blk.y * blk.z);
    constexpr size_t STACK = 256 * 1024;
    static std::vector<char*> stacks;                 // reused across launches, never initialised
    while ((int)stacks.size() < n) stacks.push_back((char*)std::malloc(STACK));
    std::function<void()> fn = body;
    body_fn = &fn;
    std::vector<Fiber> fibers((size_t)n);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                Block b;
                Barrier all(n);
                b.all = &all;
                const int nw = (n + 31) / 32;
                std::vector<std::unique_ptr<Barrier>> own;
                for (int w = 0; w < nw; ++w) {
                    const int lanes = (w + 1) * 32 <= n ? 32 : n - w * 32;
                    own.emplace_back(new Barrier(lanes));
                    b.warps.push_back(own.back().get());
                    b.slots.emplace_back(32, 0);
                }
                emu::block = &b;
                ::blockIdx = dim3(bx, by, bz);
                for (int t = 0; t < n; ++t) {
                    Fiber& f = fibers[(size_t)t];
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = stacks[(size_t)t];
                    f.ctx.uc_stack.ss_size = STACK;
                    f.ctx.uc_link = nullptr;
                    f.tid = dim3((unsigned)t % blk.x, ((unsigned)t / blk.x) % blk.y, (unsigned)t / (blk.x * blk.y));
                    f.done = false; f.waiting = nullptr;
                    makecontext(&f.ctx, (void (*)())trampoline, 0);
                }
                int live = n;
                while (live) {
                    const long before = progress;
                    for (int t = 0; t < n; ++t) {
                        Fiber& f = fibers[(size_t)t];
                        if (f.done) continue;
                        cur = &f;
                        ::threadIdx = f.tid;
                        swapcontext(&sched_ctx, &f.ctx);
                        if (f.done) { --live; ++progress; }
                    }
                    if (live && progress == before) {
                        std::fprintf(stderr, "emu: DEADLOCK in block (%u,%u,%u): %d of %d threads blocked\n", bx, by, bz, live, n);
                        int shown = 0;
                        for (int t = 0; t < n && shown < 12; ++t)
                            if (!fibers[(size_t)t].done && (t % 32) == 0) {
                                std::fprintf(stderr, "  thread %d (warp %d): %s 0x%x %u\n", t, t / 32,
                                             fibers[(size_t)t].waiting ? fibers[(size_t)t].waiting : "?", fibers[(size_t)t].wait_a,
                                             fibers[(size_t)t].wait_b);
                                ++shown;
                            }
                        std::abort();
                    }
                }
            }
    body_fn = nullptr;
}
}  // namespace emu
