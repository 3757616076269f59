// Common device helpers for the colosseum B200 engine (sm_100a).
#pragma once
#include <stdint.h>

#ifndef CRL_HOSTSIM
#include <cuda_runtime.h>
#define CRL_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#else
// tests/hostsim/cuda_shim.h was included first: kernels run on the SIMT emulator (CPU unit tests only)
#define CRL_LAUNCH(kernel, grid, block, stream, ...) \
    hostsim::launch(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); })
#endif

#define CRL_NSTAT 32
// statistics slots (int64 each); identical to oracle/oracle_rollout.c
enum {
    ST_STEPS = 0, ST_EPISODES = 1, ST_EPLEN = 2, ST_WINS = 3 /*..6*/, ST_NOWIN = 7, ST_SCORE = 8 /*..11*/,
    ST_ERRORS = 12, ST_NVALID = 13, ST_RANK = 14 /*..17*/, ST_REWARD = 18
};

// step flags
#define CRL_FLAG_AUTO_RESET 1

typedef unsigned long long crl_u64;

// 128-bit streaming accesses: state is touched exactly once per step, keep it out of L1
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
#ifndef CRL_HOSTSIM
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
#else
    return *p;
#endif
}
__device__ __forceinline__ void st_stream(uint4 *p, uint4 v) {
#ifndef CRL_HOSTSIM
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
    *p = v;
#endif
}

// Episode statistics: per-thread contribution -> warp reduce (redux.sync) -> shared-memory partial per CTA
// -> one global atomic per CTA and non-zero slot.  Values are small ints; reduce in 32-bit, accumulate in 64.
struct BlockStats {
    int *sm;  // __shared__ int[CRL_NSTAT]
    __device__ __forceinline__ void init() const {
        if (threadIdx.x < CRL_NSTAT) sm[threadIdx.x] = 0;
        __syncthreads();
    }
    __device__ __forceinline__ void add(int slot, int v) const {
        int s = __reduce_add_sync(0xffffffffu, v);
        if ((threadIdx.x & 31) == 0 && s != 0) atomicAdd(&sm[slot], s);
    }
    __device__ __forceinline__ void flush(crl_u64 *g) const {
        __syncthreads();
        if (threadIdx.x < CRL_NSTAT && sm[threadIdx.x] != 0) atomicAdd(g + threadIdx.x, (crl_u64)(long long)sm[threadIdx.x]);
    }
};
