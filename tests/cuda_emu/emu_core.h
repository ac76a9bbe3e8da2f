// emu_core.h -- TEST INFRASTRUCTURE ONLY.  A minimal SIMT emulator: enough of the CUDA device
// vocabulary to compile lpm_v2_b200/csrc/{directsum,pairs,sym_kernels}.cuh with g++ and run their
// kernels on CPU threads -- one OS thread per CUDA thread, one CTA at a time, pthread barriers for
// __syncthreads() and for the warp shuffles.  It checks the LOGIC of kernel source that has not run on
// a GPU yet (indexing, pipelines, warp reductions, retry paths); it says nothing about speed, and
// nothing in liblpmgpu.so uses it.
#pragma once
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
// (__noinline__ is rewritten by convert.py: libstdc++ spells the attribute that way itself)
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __constant__
#define __align__(n)

struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }


// ---- one CTA at a time -------------------------------------------------------------------------
struct EmuCta {
    int nthreads = 0;
    pthread_barrier_t cta_bar;
    std::vector<pthread_barrier_t> warp_bar;
    std::vector<uint64_t> xchg;          // [warp][32]
};
inline EmuCta* g_emu_cta = nullptr;

inline void __syncthreads() { pthread_barrier_wait(&g_emu_cta->cta_bar); }
[[noreturn]] inline void __trap()
{
    std::fprintf(stderr, "cuda_emu: __trap() in block %u thread %u\n", blockIdx.x, threadIdx.x);
    std::abort();
}

template <class T>
inline T emu_warp_exchange(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    g_emu_cta->xchg[w * 32 + lane] = bits;
    pthread_barrier_wait(&g_emu_cta->warp_bar[w]);
    const uint64_t got = g_emu_cta->xchg[w * 32 + (src_lane & 31)];
    pthread_barrier_wait(&g_emu_cta->warp_bar[w]);
    T r;
    std::memcpy(&r, &got, sizeof(T));
    return r;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int off) { return emu_warp_exchange(v, (threadIdx.x & 31) ^ off); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int d)
{
    const int lane = threadIdx.x & 31;
    return emu_warp_exchange(v, lane >= d ? lane - d : lane);
}
inline unsigned __ballot_sync(unsigned, bool p)
{
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (emu_warp_exchange<unsigned>(p ? 1u : 0u, l) << l);
    return m;
}
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }

// ---- atomics (the threads of a CTA run concurrently) -------------------------------------------
inline double atomicAdd(double* p, double v)
{
    uint64_t* u = reinterpret_cast<uint64_t*>(p);
    uint64_t old = __atomic_load_n(u, __ATOMIC_RELAXED);
    for (;;) {
        double o;
        std::memcpy(&o, &old, 8);
        const double nv = o + v;
        uint64_t nb;
        std::memcpy(&nb, &nv, 8);
        if (__atomic_compare_exchange_n(u, &old, nb, false, __ATOMIC_ACQ_REL, __ATOMIC_RELAXED)) return o;
    }
}
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v)
{
    return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL);
}
inline int atomicMax(int* p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_ACQ_REL, __ATOMIC_RELAXED)) {}
    return old;
}

// ---- device math intrinsics --------------------------------------------------------------------
using std::fma;
using std::fmax;
using std::sqrt;
using std::cosh;
using std::log;
using std::floor;
using std::sinh;
using std::sin;
using std::cos;
using std::exp;
using std::atan2;
using std::fabs;
using std::tan;
using std::trunc;
using std::ceil;
using std::ldexp;
inline double rsqrt(double v) { return 1.0 / std::sqrt(v); }
inline void sincospi(double v, double* s, double* c) { *s = std::sin(M_PI * v); *c = std::cos(M_PI * v); }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __hiloint2double(int hi, int lo)
{
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    std::memcpy(&d, &b, 8);
    return d;
}
inline int __double2hiint(double d)
{
    uint64_t b;
    std::memcpy(&b, &d, 8);
    return (int)(b >> 32);
}
inline int __double2loint(double d)
{
    uint64_t b;
    std::memcpy(&b, &d, 8);
    return (int)(uint32_t)b;
}
inline long long __double_as_longlong(double d)
{
    long long b;
    std::memcpy(&b, &d, 8);
    return b;
}
inline double __longlong_as_double(long long b)
{
    double d;
    std::memcpy(&d, &b, 8);
    return d;
}
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }

// ---- dynamic shared memory: one CTA runs at a time ---------------------------------------------
inline unsigned char* emu_dynamic_smem()
{
    alignas(128) static unsigned char buf[232 * 1024];
    return buf;
}

// ---- launchers ---------------------------------------------------------------------------------
// Kernels with barriers / shuffles: one OS thread per CUDA thread, CTAs one after the other.  The threads
// of a block size are created once and parked between CTAs (a 256-thread spawn per CTA costs milliseconds).
struct EmuPool {
    unsigned block = 0;
    std::vector<std::thread> threads;
    pthread_barrier_t start, done;
    EmuCta cta;
    const std::function<void()>* kernel = nullptr;
    unsigned bid = 0, grid = 0;
    bool quit = false;
    explicit EmuPool(unsigned b) : block(b)
    {
        pthread_barrier_init(&start, nullptr, b + 1);
        pthread_barrier_init(&done, nullptr, b + 1);
        cta.nthreads = (int)b;
        pthread_barrier_init(&cta.cta_bar, nullptr, b);
        cta.warp_bar.resize(b / 32);
        for (auto& wb : cta.warp_bar) pthread_barrier_init(&wb, nullptr, 32);
        cta.xchg.assign(b, 0);
        for (unsigned t = 0; t < b; ++t)
            threads.emplace_back([this, t]() {
                for (;;) {
                    pthread_barrier_wait(&start);
                    if (quit) return;
                    threadIdx.x = t; blockIdx.x = bid; blockDim.x = block; gridDim.x = grid;
                    (*kernel)();
                    pthread_barrier_wait(&done);
                }
            });
    }
    void run(unsigned g, const std::function<void()>& k)
    {
        kernel = &k; grid = g;
        g_emu_cta = &cta;
        for (unsigned b = 0; b < g; ++b) {
            bid = b;
            pthread_barrier_wait(&start);
            pthread_barrier_wait(&done);
        }
        g_emu_cta = nullptr;
    }
    ~EmuPool()
    {
        quit = true;
        pthread_barrier_wait(&start);
        for (auto& t : threads) t.join();
    }
};
inline void emu_launch(unsigned grid, unsigned block, const std::function<void()>& kernel)
{
    if (block % 32 != 0 || block == 0) { std::fprintf(stderr, "cuda_emu: block size %u\n", block); std::abort(); }
    static std::vector<EmuPool*> pools;          // never destroyed: the threads die with the process
    EmuPool* p = nullptr;
    for (auto* q : pools)
        if (q->block == block) p = q;
    if (!p) { p = new EmuPool(block); pools.push_back(p); }
    p->run(grid, kernel);
}
// kernels whose threads never talk to each other: a plain loop
inline void emu_launch_seq(unsigned grid, unsigned block, const std::function<void()>& kernel)
{
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned t = 0; t < block; ++t) {
            threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
            kernel();
        }
}
