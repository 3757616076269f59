// Philox4x32-10 (Random123), counter = (env_lo, env_hi, step, tag), key = (seed_lo, seed_hi).
// Bit-identical to colosseumrl_b200/philox.py (host) -- see SURVEY.md section 8d for the known answers.
#pragma once
#include "crl_common.cuh"

#define CRL_TAG_TRON 1u
#define CRL_TAG_BLOKUS 2u
#define CRL_TAG_TTT 3u

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1;
        c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 env_words(crl_u64 seed, crl_u64 env, uint32_t step, uint32_t tag) {
    return philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), step, tag, (uint32_t)seed, (uint32_t)(seed >> 32));
}

__global__ void philox_words_kernel(uint4 *out, crl_u64 seed, crl_u64 first_env, uint32_t step, uint32_t tag, long long B) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < B) out[e] = env_words(seed, first_env + (crl_u64)e, step, tag);
}
