// TEST INFRASTRUCTURE ONLY -- a minimal SIMT emulator so the *same kernel source* that nvcc compiles
// for sm_100a (colosseumrl_b200/csrc/*.cuh) can be executed by CPU unit tests in the build container,
// which has no GPU.  One host thread per CUDA thread, one CTA at a time; warp collectives
// (__ballot_sync / __shfl*_sync / __reduce_add_sync / __syncwarp) and __syncthreads are implemented with
// std::barrier.  It is never linked into, loaded by, or reachable from the product package: the product
// library is built by nvcc only and the Python host layer refuses to run without it.
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define CRL_HOSTSIM 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __constant__ static const
#define __align__(n) __attribute__((aligned(n)))

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct int4 { int x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
typedef void *cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0

namespace hostsim {
struct Ctx {
    std::barrier<> *block_bar;
    std::barrier<> *warp_bar;
    uint64_t *xchg;  // [32] per warp
};
extern thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
extern thread_local Ctx t_ctx;

template <class F>
void launch(dim3 grid, dim3 block, F body) {
    unsigned nthreads = block.x, nwarps = (nthreads + 31) / 32;
    std::barrier<> block_bar(nthreads);
    std::vector<std::unique_ptr<std::barrier<>>> warp_bars;
    for (unsigned w = 0; w < nwarps; w++) {
        unsigned n = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
        warp_bars.emplace_back(new std::barrier<>(n));
    }
    std::vector<uint64_t> xchg(nwarps * 32);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; t++) {
        th.emplace_back([&, t] {
            t_blockDim = block; t_gridDim = grid;
            t_threadIdx = dim3(t, 0, 0);
            t_ctx.block_bar = &block_bar;
            t_ctx.warp_bar = warp_bars[t / 32].get();
            t_ctx.xchg = xchg.data() + (t / 32) * 32;
            for (unsigned b = 0; b < grid.x; b++) {
                t_blockIdx = dim3(b, 0, 0);
                body();
                block_bar.arrive_and_wait();  // static __shared__ storage is reused by the next CTA
            }
        });
    }
    for (auto &x : th) x.join();
}
}  // namespace hostsim

#define threadIdx hostsim::t_threadIdx
#define blockIdx hostsim::t_blockIdx
#define blockDim hostsim::t_blockDim
#define gridDim hostsim::t_gridDim

static inline void __syncthreads() { hostsim::t_ctx.block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { hostsim::t_ctx.warp_bar->arrive_and_wait(); }

template <class T>
static inline T hostsim_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned lane = threadIdx.x & 31;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    hostsim::t_ctx.xchg[lane] = raw;
    hostsim::t_ctx.warp_bar->arrive_and_wait();
    uint64_t got = hostsim::t_ctx.xchg[src_lane & 31];
    hostsim::t_ctx.warp_bar->arrive_and_wait();
    T r;
    std::memcpy(&r, &got, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return hostsim_exchange(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return hostsim_exchange(v, (threadIdx.x & 31) ^ m); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    int lane = threadIdx.x & 31;
    return hostsim_exchange(v, lane >= (int)d ? lane - (int)d : lane);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    int lane = threadIdx.x & 31;
    return hostsim_exchange(v, lane + (int)d < 32 ? lane + (int)d : lane);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned lane = threadIdx.x & 31;
    hostsim::t_ctx.xchg[lane] = pred ? 1 : 0;
    hostsim::t_ctx.warp_bar->arrive_and_wait();
    unsigned m = 0;
    unsigned nlanes = (blockDim.x - (threadIdx.x & ~31u)) >= 32 ? 32 : blockDim.x - (threadIdx.x & ~31u);
    for (unsigned i = 0; i < nlanes; i++) m |= (unsigned)(hostsim::t_ctx.xchg[i] & 1) << i;
    hostsim::t_ctx.warp_bar->arrive_and_wait();
    return m;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    unsigned s = v;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}
static inline int __reduce_add_sync(unsigned, int v) { return (int)__reduce_add_sync(0xffffffffu, (unsigned)v); }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    unsigned s = v;
    for (int o = 16; o; o >>= 1) s |= __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}
static inline unsigned __activemask() { return 0xffffffffu; }

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    // find the offset-th (1-based) set bit at or above `base` (only the form used here: base = 0, offset >= 1)
    for (unsigned i = base; i < 32; i++)
        if ((mask >> i) & 1u) { if (--offset == 0) return i; }
    return 0xffffffffu;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {   // PRMT, default mode (incl. sign replication)
    uint64_t ab = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned n = (s >> (4 * i)) & 15u, byte = (unsigned)(ab >> (8 * (n & 7u))) & 255u;
        if (n & 8u) byte = (byte & 0x80u) ? 0xFFu : 0u;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __vminu2(unsigned a, unsigned b) {
    unsigned al = a & 0xFFFFu, bl = b & 0xFFFFu, ah = a >> 16, bh = b >> 16, lo = al < bl ? al : bl, hi = ah < bh ? ah : bh;
    return lo | hi << 16;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {   // shf.r.wrap.b32
    return (unsigned)(((((uint64_t)hi) << 32) | lo) >> (sh & 31u));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {   // shf.l.wrap.b32
    return (unsigned)((((((uint64_t)hi) << 32) | lo) << (sh & 31u)) >> 32);
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
template <class T> static inline T __ldg(const T *p) { return *p; }

static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }
template <class T> static inline T max(T a, T b) { return a > b ? a : b; }
