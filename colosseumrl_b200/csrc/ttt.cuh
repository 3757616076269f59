// Generalised Tic Tac Toe (2p 3x3, 3p 3x5, 4p 3x3x3): batched next_state with the win test done as
// shifted-mask compares, fused with terminal / winners / ranking / valid-action mask (sm_100a).
//
// Replaces (reference file:line):
//   TicTacToe{2,3,4}PlayerEnv.next_state   envs/tictactoe/tictactoe_2p_env.py:240-315, _3p_env.py:241-316,
//                                          _4p_env.py:271-346 (13 scipy.signal.correlate calls per step for 4p)
//   valid_actions / is_valid_action        tictactoe_2p_env.py:317-380
//   new_state                              :139-170
//   state_to_observation                   :382-407 (+ the "% 3" relabel quirk of tictactoe_4p_env.py:50)
//
// HBM layout: ONE 16-byte vector per environment, uint4 state[B]:
//   .x = cells of player 0 (bit = C-order flat cell index, <= 27 bits) | mover << 27 | (winner + 1) << 29
//   .y = cells of player 1 | min(episode steps, 31) << 27
//   .z = cells of player 2,   .w = cells of player 3
// A warp loads/stores 512 contiguous bytes per instruction; one thread owns one environment.
//
// Win test: for every line direction d (4 in 2-D, 13 in 3-D) with flat stride s_d, three in a row exists iff
//   m & (m >> s_d) & (m >> 2 s_d) & START_d != 0,  START_d = cells where a length-3 segment along d fits.
// This equals "3 in scipy.signal.correlate(mask, pattern, 'valid')" for the reference's WINNING_SHAPES
// (checked against the oracle's 8 / 20 / 49 line tables in tests/test_abi.py::test_ttt_line_tables).
#pragma once
#include "crl_common.cuh"
#include "philox.cuh"
#include "ttt_tables.h"

#define TTT_MAX_DIRS 13

struct TTTParams {
    int n;          // players: 2, 3, 4
    int cells;      // 9, 15, 27
    int ndirs;
    uint32_t cellmask;
    uint32_t stride[TTT_MAX_DIRS];
    uint32_t start[TTT_MAX_DIRS];
};

struct TTTEnv {
    uint32_t m[4];
    int mover, winner1;  // winner1 = winner + 1, 0 = None
    uint32_t ep_len;
};

struct TTTOut {
    int reward, terminal, error, placed, winners, nvalid;
    uint32_t valid_after;
};

__device__ __forceinline__ void ttt_decode(TTTEnv &s, uint4 v) {
    s.m[0] = v.x & 0x07ffffffu; s.m[1] = v.y & 0x07ffffffu; s.m[2] = v.z & 0x07ffffffu; s.m[3] = v.w & 0x07ffffffu;
    s.mover = (v.x >> 27) & 3; s.winner1 = (v.x >> 29) & 7;
    s.ep_len = v.y >> 27;
}
__device__ __forceinline__ uint4 ttt_encode(const TTTEnv &s) {
    return make_uint4(s.m[0] | (uint32_t)s.mover << 27 | (uint32_t)s.winner1 << 29,
                      s.m[1] | min(s.ep_len, 31u) << 27, s.m[2], s.m[3]);
}
__device__ __forceinline__ void ttt_new_state(TTTEnv &s) {
    s.m[0] = s.m[1] = s.m[2] = s.m[3] = 0; s.mover = 0; s.winner1 = 0; s.ep_len = 0;
}
template <int NP>
__device__ __forceinline__ bool ttt_is_terminal(const TTTEnv &s) {
    return s.winner1 != 0 || ((s.m[0] | s.m[1] | s.m[2] | s.m[3]) == TTTGeo<NP>::CELLMASK);
}

// next_state (tictactoe_2p_env.py:283-315).  action: C-order flat cell index, negative = '' (pass).
template <int NP>
__device__ __forceinline__ void ttt_step_env(TTTEnv &s, int action, TTTOut &o) {
    constexpr uint32_t CELLMASK = TTTGeo<NP>::CELLMASK;
    uint32_t occ = s.m[0] | s.m[1] | s.m[2] | s.m[3];
    o.nvalid = __popc(~occ & CELLMASK);
    const bool in_range = (unsigned)action < (unsigned)TTTGeo<NP>::CELLS;
    const uint32_t bit = in_range ? (1u << action) : 0u;
    const bool cell_free = in_range && !(occ & bit);                 // is_valid_action (:350-380)
    const bool placed = cell_free && s.winner1 == 0;                 // :293
    o.error = (action >= 0 && !cell_free) ? 1 : 0;                   // invalid action: silent no-op in the reference
    o.placed = placed;
    if (placed) {
        uint32_t mine = 0;
#pragma unroll
        for (int p = 0; p < NP; p++) {
            s.m[p] |= (p == s.mover) ? bit : 0u;                     // :295
            mine |= (p == s.mover) ? s.m[p] : 0u;
        }
        if (TTTGeo<NP>::win(mine)) s.winner1 = s.mover + 1;          // :297-300
        occ |= bit;
    }
    o.reward = 0; o.terminal = 0; o.winners = 0;
    if (s.winner1) {                                                 // :302-308 (winner persists in the state)
        o.reward = (s.winner1 - 1 == s.mover) ? 1 : -1;
        o.winners = 1 << (s.winner1 - 1);
        o.terminal = 1;
    }
    if (occ == CELLMASK) o.terminal = 1;                             // :310-311 draw / full board
    o.valid_after = ~occ & CELLMASK;
    s.mover = (s.mover + 1 == NP) ? 0 : s.mover + 1;                 // :313
    s.ep_len += 1;
}

// result record, 4 bytes: int8 reward | u8 flags (1 terminal, 2 invalid action, 4 placed) | u8 winners mask |
// u8 ranking bits (bit p = rank of player p: winners 0, everybody else 1 -- BaseEnvironment.py:173-195)
template <int NP>
__device__ __forceinline__ uint32_t ttt_pack_result(const TTTOut &o) {
    uint32_t rank = ((1u << NP) - 1u) & ~(uint32_t)o.winners;
    return ((uint32_t)o.reward & 0xffu) | (uint32_t)(o.terminal | o.error << 1 | o.placed << 2) << 8 |
           (uint32_t)o.winners << 16 | rank << 24;
}

// Episode statistics: every thread ACCUMULATES the packed counters of up to TTT_ACC_MAX env-steps in four registers
// (a handful of instructions per env-step), then the warp reduces them once (4 redux.sync), lane l < 15 extracts
// counter l from the warp-uniform sums and adds it to the CTA's shared partial -- ONE atomic instruction per warp and
// flush -- and the CTA adds its partial to its row of the global buffer once.  (The first version reduced and
// flushed after every env-step with two barriers: 40 % of the step kernel's time.)
//   A: steps | episodes << 8 | episodes without winner << 16 | illegal actions << 24       (sums <= 32 * 4 = 128)
//   W: wins per seat, 8 bits each.    The "not a winner" counters follow: rank[q] = episodes - wins[q]
//   D: valid-action count (12 bits: <= 27 * 128) | episode length of finished episodes << 12 (<= 31 * 128)
//   R: (mover + 1) * reward, biased by +4 per env-step (<= 8 * 128)
//   lane constants: word (0..3) | shift << 4 | bits << 10 | slot << 16 | kind << 24 (1: episodes - x, 2: x - 4 * steps)
#define TTT_ACC_MAX 4
#define TTT_SL(word, shift, bits, slot, kind) ((word) | (shift) << 4 | (bits) << 10 | (slot) << 16 | (kind) << 24)
__constant__ uint32_t TTT_STAT_LANE[32] = {
    TTT_SL(0, 0, 8, ST_STEPS, 0), TTT_SL(0, 8, 8, ST_EPISODES, 0), TTT_SL(0, 16, 8, ST_NOWIN, 0), TTT_SL(0, 24, 8, ST_ERRORS, 0),
    TTT_SL(1, 0, 8, ST_WINS + 0, 0), TTT_SL(1, 8, 8, ST_WINS + 1, 0), TTT_SL(1, 16, 8, ST_WINS + 2, 0), TTT_SL(1, 24, 8, ST_WINS + 3, 0),
    TTT_SL(1, 0, 8, ST_RANK + 0, 1), TTT_SL(1, 8, 8, ST_RANK + 1, 1), TTT_SL(1, 16, 8, ST_RANK + 2, 1), TTT_SL(1, 24, 8, ST_RANK + 3, 1),
    TTT_SL(2, 0, 12, ST_NVALID, 0), TTT_SL(2, 12, 12, ST_EPLEN, 0), TTT_SL(3, 0, 12, ST_REWARD, 2),
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define TTT_STAT_LANES 15

struct TTTStatAcc {
    uint32_t A, W, D, R;
    __device__ __forceinline__ void clear() { A = W = D = R = 0u; }
    // one env-step of a live thread
    template <int NP>
    __device__ __forceinline__ void add(const TTTOut &o, int mover, uint32_t ep_len) {
        const uint32_t t = o.terminal ? 1u : 0u;
        const uint32_t w = t ? (uint32_t)o.winners : 0u;
        A += 1u | t << 8 | (t & (uint32_t)(o.winners == 0)) << 16 | (uint32_t)o.error << 24;
        W += (w * 0x00204081u) & 0x01010101u;                               // bit q -> byte q
        D += (uint32_t)o.nvalid | (t ? min(ep_len, 31u) : 0u) << 12;
        R += (uint32_t)((mover + 1) * o.reward + 4);
    }
    // warp-collective: reduce, extract, add to the CTA's shared partial, clear
    template <int NP>
    __device__ __forceinline__ void flush(int *sm_stat, uint32_t lane_const) {
        const uint32_t a = __reduce_add_sync(0xffffffffu, A), w = __reduce_add_sync(0xffffffffu, W);
        const uint32_t d = __reduce_add_sync(0xffffffffu, D), r = __reduce_add_sync(0xffffffffu, R);
        const uint32_t word = lane_const & 15u;
        const uint32_t x = word == 0u ? a : word == 1u ? w : word == 2u ? d : r;
        int val = (int)((x >> ((lane_const >> 4) & 31u)) & ((1u << ((lane_const >> 10) & 31u)) - 1u));
        const uint32_t kind = lane_const >> 24;
        if (kind == 1u) val = ((lane_const >> 7) & 7u) < (uint32_t)NP ? (int)((a >> 8) & 255u) - val : 0;   // not-a-winner, seats < NP
        if (kind == 2u) val -= 4 * (int)(a & 255u);                         // remove the reward bias
        const int slot = (int)((lane_const >> 16) & 255u);
        if ((int)(threadIdx.x & 31) < TTT_STAT_LANES && val != 0) atomicAdd(&sm_stat[slot], val);
        clear();
    }
};

__device__ __forceinline__ void ttt_zero_out(TTTOut &o) {
    o.reward = o.terminal = o.error = o.placed = o.winners = o.nvalid = 0; o.valid_after = 0;
}

// Grid-stride: the launcher sizes the grid so that a thread steps at most TTT_ACC_MAX environments.
template <int NP>
__global__ void __launch_bounds__(256, 8)
ttt_step_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, const int8_t *__restrict__ actions,
                uint32_t *__restrict__ result, uint32_t *__restrict__ valid_after, crl_u64 *stats, long long B,
                int flags) {
    __shared__ int sm_stat[CRL_NSTAT];
    if (stats) { if (threadIdx.x < CRL_NSTAT) sm_stat[threadIdx.x] = 0; __syncthreads(); }
    TTTStatAcc acc;
    acc.clear();
    const long long stride = (long long)gridDim.x * blockDim.x;
    // (PDL launch: the grid may become resident while its predecessor drains; an L2 prefetch prologue like Tron's was
    // measured and dropped -- this kernel is issue-bound and the extra instructions cost more than the overlap buys)
    pdl_wait();
    pdl_launch_dependents();
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < B; e += stride) {
        TTTEnv s;
        TTTOut o;
        ttt_decode(s, ld_stream(in + e));
        if ((flags & CRL_FLAG_AUTO_RESET) && ttt_is_terminal<NP>(s)) ttt_new_state(s);
        const int mover = s.mover;
        ttt_step_env<NP>(s, (int)actions[e], o);
        st_stream(out + e, ttt_encode(s));
        result[e] = ttt_pack_result<NP>(o);
        if (valid_after) valid_after[e] = o.valid_after;
        if (stats) acc.add<NP>(o, mover, s.ep_len);
    }
    if (stats) {
        acc.flush<NP>(sm_stat, TTT_STAT_LANE[threadIdx.x & 31]);
        __syncthreads();
        stats_flush_row(sm_stat, stats);
    }
}

// k-th (0-based) set bit of a <= 27-bit mask: branch-free binary search on popcounts
__device__ __forceinline__ int ttt_kth_bit(uint32_t mask, int k) {
    int pos = 0;
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const int c = __popc(mask & ((1u << w) - 1u));
        const bool up = k >= c;
        k -= up ? c : 0;
        pos += up ? w : 0;
        mask = up ? (mask >> w) : mask;
    }
    return pos;
}

// Exact r % n for n <= 27 without the generic 32-bit division (~20 instructions): lane l holds floor(2^32 / l), a
// shuffle fetches the entry of the lane's own n, q' = umulhi(r, M[n]) is the quotient or one less, one correction.
// (n = 1: M = 2^32 - 1, q' = r - 1, the correction gives 0.)  Warp-collective: call with all lanes.
__device__ const uint32_t TTT_RCP_G[32] = {      // floor(2^32 / l); l < 2: 2^32 - 1   (global: one coalesced load per warp)
    0xffffffffu, 0xffffffffu, 0x80000000u, 0x55555555u, 0x40000000u, 0x33333333u, 0x2aaaaaaau, 0x24924924u,
    0x20000000u, 0x1c71c71cu, 0x19999999u, 0x1745d174u, 0x15555555u, 0x13b13b13u, 0x12492492u, 0x11111111u,
    0x10000000u, 0x0f0f0f0fu, 0x0e38e38eu, 0x0d79435eu, 0x0cccccccu, 0x0c30c30cu, 0x0ba2e8bau, 0x0b21642cu,
    0x0aaaaaaau, 0x0a3d70a3u, 0x09d89d89u, 0x097b425eu, 0x09249249u, 0x08d3dcb0u, 0x08888888u, 0x08421084u};
__device__ __forceinline__ uint32_t ttt_rcp_lane() { return TTT_RCP_G[threadIdx.x & 31u]; }
__device__ __forceinline__ uint32_t ttt_mod_small(uint32_t r, uint32_t n, uint32_t rcp_lane) {
    const uint32_t m = __shfl_sync(0xffffffffu, rcp_lane, (int)n);
    uint32_t rem = r - __umulhi(r, m) * n;
    return rem >= n ? rem - n : rem;
}

// uniform random policy: the (r0 % n_empty)-th empty cell in C order, pass (-1) if the board is full
// (warp-collective because of ttt_mod_small)
template <int NP>
__device__ __forceinline__ int ttt_random_action(const TTTEnv &s, uint32_t r0, uint32_t rcp_lane) {
    uint32_t empty = ~(s.m[0] | s.m[1] | s.m[2] | s.m[3]) & TTTGeo<NP>::CELLMASK;
    const int n = __popc(empty);
    const uint32_t k = ttt_mod_small(r0, (uint32_t)max(n, 1), rcp_lane);
    return n ? ttt_kth_bit(empty, (int)k) : -1;
}

template <int NP>
__global__ void ttt_policy_random_kernel(const uint4 *__restrict__ st, int8_t *__restrict__ actions, long long B,
                                         int flags, crl_u64 seed, crl_u64 first_env, uint32_t step) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = e < B;
    TTTEnv s;
    ttt_new_state(s);
    if (valid) ttt_decode(s, st[e]);
    if ((flags & CRL_FLAG_AUTO_RESET) && ttt_is_terminal<NP>(s)) ttt_new_state(s);
    uint4 r = env_words(seed, first_env + (crl_u64)e, step, CRL_TAG_TTT);
    const int a = ttt_random_action<NP>(s, r.x, ttt_rcp_lane());
    if (valid) actions[e] = (int8_t)a;
}

// One fused random-policy step of a NON-terminal environment (the rollout resets terminal ones first): a non-terminal
// board has an empty cell and no winner, so the chosen cell is always free and always placed -- next_state
// (tictactoe_2p_env.py:283-315) without its validity branches, sharing the empty mask with the policy.
template <int NP>
__device__ __forceinline__ void ttt_policy_step(TTTEnv &s, uint32_t r0, uint32_t rcp_lane, TTTOut &o) {
    constexpr uint32_t CELLMASK = TTTGeo<NP>::CELLMASK;
    const uint32_t occ = s.m[0] | s.m[1] | s.m[2] | s.m[3], empty = ~occ & CELLMASK;
    const int n = __popc(empty);                                     // >= 1
    const uint32_t bit = 1u << ttt_kth_bit(empty, (int)ttt_mod_small(r0, (uint32_t)max(n, 1), rcp_lane));
    // the mover's cells: a 4:1 multiplexer on the two mover bits (3 selects), then one predicated write-back per seat
    const bool b0 = (s.mover & 1) != 0, b1 = (s.mover & 2) != 0;
    const uint32_t lo = b0 ? s.m[1] : s.m[0], hi = b0 ? s.m[3 % NP] : s.m[2 % NP];
    const uint32_t mine = ((NP > 2 && b1) ? hi : lo) | bit;
#pragma unroll
    for (int p = 0; p < NP; p++) s.m[p] = (p == s.mover) ? mine : s.m[p];
    const bool win = TTTGeo<NP>::win(mine) != 0u;
    s.winner1 = win ? s.mover + 1 : 0;
    o.nvalid = n; o.error = 0; o.placed = 1;
    o.reward = win ? 1 : 0;
    o.winners = win ? 1 << s.mover : 0;
    o.terminal = (win || (occ | bit) == CELLMASK) ? 1 : 0;
    o.valid_after = ~(occ | bit) & CELLMASK;
    s.mover = NP == 4 ? ((s.mover + 1) & 3) : ((s.mover + 1 == NP) ? 0 : s.mover + 1);
    s.ep_len += 1;
}

// K fused random-policy steps with auto-reset; grid-stride like the step kernel (a thread owns <= TTT_ACC_MAX
// environments and keeps one of them in registers for the K steps).
template <int NP>
__global__ void __launch_bounds__(256, 8)
ttt_rollout_kernel(uint4 *__restrict__ state, uint32_t *__restrict__ result, crl_u64 *stats, long long B,
                   const PhiloxKeys keys, crl_u64 first_env, uint32_t step0, int K) {
    __shared__ int sm_stat[CRL_NSTAT];
    if (stats) { if (threadIdx.x < CRL_NSTAT) sm_stat[threadIdx.x] = 0; __syncthreads(); }
    const uint32_t lane_const = TTT_STAT_LANE[threadIdx.x & 31], rcp_lane = ttt_rcp_lane();
    TTTStatAcc acc;
    acc.clear();
    int pending = 0;                                     // env-steps accumulated since the last flush (block-uniform)
    const long long stride = (long long)gridDim.x * blockDim.x, first = (long long)blockIdx.x * blockDim.x;
    pdl_wait();
    pdl_launch_dependents();
    for (long long e0 = first; e0 < B; e0 += stride) {   // block-uniform trip count: the flushes are warp-collective
        const long long e = e0 + threadIdx.x;
        const bool valid = e < B;
        TTTEnv s;
        TTTOut o;
        ttt_zero_out(o);
        ttt_new_state(s);
        if (valid) ttt_decode(s, ld_stream(state + e));
        for (int k = 0; k < K; k++) {                    // (lanes past the end of the batch step a dummy board)
            if (ttt_is_terminal<NP>(s)) ttt_new_state(s);
            const crl_u64 ge = first_env + (crl_u64)e;
            const uint32_t r0 = philox4x32_10_x((uint32_t)ge, (uint32_t)(ge >> 32), step0 + (uint32_t)k, CRL_TAG_TTT, keys);
            const int mover = s.mover;
            ttt_policy_step<NP>(s, r0, rcp_lane, o);
            if (stats && valid) acc.add<NP>(o, mover, s.ep_len);
            if (stats && ++pending == TTT_ACC_MAX) { acc.flush<NP>(sm_stat, lane_const); pending = 0; }
        }
        if (valid) {
            st_stream(state + e, ttt_encode(s));
            if (result) result[e] = ttt_pack_result<NP>(o);
        }
    }
    if (stats) {
        if (pending) acc.flush<NP>(sm_stat, lane_const);
        __syncthreads();
        stats_flush_row(sm_stat, stats);
    }
}

__global__ void ttt_reset_kernel(uint4 *__restrict__ state, const uint8_t *__restrict__ mask, long long B) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B || (mask && !mask[e])) return;
    state[e] = make_uint4(0, 0, 0, 0);
}

// valid_actions (tictactoe_2p_env.py:317-348): bit c set = cell c (C order) is empty; 0 <=> ['']
__global__ void ttt_valid_kernel(const uint4 *__restrict__ st, uint32_t *__restrict__ mask, long long B, TTTParams prm) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B) return;
    uint4 v = st[e];
    mask[e] = ~(v.x | v.y | v.z | v.w) & prm.cellmask;
}

// state_to_observation (2p :382-407).  A warp unpacks 32 environments.  Lane = environment: the viewer-relative label
// of every player's cells (absolute: p; else (p - viewer) mod 2 for 2p, mod 3 for 3p AND 4p -- tictactoe_4p_env.py:50)
// is folded into two label bit-planes + the occupancy plane, four cells at a time are spread to bytes with one
// multiply each ((nibble * 0x00204081) & 0x01010101: bit i -> byte i) and combined into a word of int8 cells
// (-1 = empty), which goes to the lane's CELLS-byte slot of the warp's staging tile.  The tile (32 * CELLS bytes,
// contiguous in the output and 16-byte aligned) then leaves with 128-bit stores.
// player == -1: absolute; -2: each game seen by its current mover (CRL_PLAYER_MOVER).
// board int8[B][cells] (-1 empty); winner int8[B] (-1 None); mover int8[B].
template <int NP>
__global__ void __launch_bounds__(256)
ttt_observe_kernel(const uint4 *__restrict__ st, long long B, int player, int8_t *__restrict__ board,
                   int8_t *__restrict__ winner, int8_t *__restrict__ mover) {
    constexpr int CELLS = TTTGeo<NP>::CELLS, GROUPS = (CELLS + 3) / 4, MOD = NP == 2 ? 2 : 3;
    constexpr int TILE = 32 * CELLS;                         // 288 / 480 / 864: multiples of 16
    __shared__ __align__(16) uint8_t stage[8][TILE + 16];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long e0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll;      // first environment of the warp
    if (e0 >= B) return;
    const long long e = e0 + lane;
    TTTEnv s;
    ttt_new_state(s);
    if (e < B) {
        ttt_decode(s, ld_stream(st + e));
        if (winner) winner[e] = (int8_t)(s.winner1 - 1);
        if (mover) mover[e] = (int8_t)s.mover;
    }
    const int viewer = player == -2 ? s.mover : player;
    uint32_t b0 = 0u, b1 = 0u;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        const int code = viewer < 0 ? p : ((p - viewer) % MOD + MOD) % MOD;
        b0 |= (code & 1) ? s.m[p] : 0u;
        b1 |= (code & 2) ? s.m[p] : 0u;
    }
    const uint32_t emp = ~(s.m[0] | s.m[1] | s.m[2] | s.m[3]);
    uint8_t *mine = stage[wid] + lane * CELLS;
#pragma unroll
    for (int j = 0; j < GROUPS; j++) {
        const uint32_t x0 = (((b0 >> (4 * j)) & 15u) * 0x00204081u) & 0x01010101u;
        const uint32_t x1 = (((b1 >> (4 * j)) & 15u) * 0x00204081u) & 0x01010101u;
        const uint32_t xe = (((emp >> (4 * j)) & 15u) * 0x00204081u) & 0x01010101u;
        const uint32_t word = x0 + 2u * x1 + 255u * xe;      // empty cells have no label bits: 0xff = -1
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (4 * j + i < CELLS) mine[4 * j + i] = (uint8_t)(word >> (8 * i));
    }
    __syncwarp();
    int8_t *dst = board + e0 * CELLS;
    if (B - e0 >= 32) {
        const uint4 *src4 = (const uint4 *)stage[wid];
        for (int i = lane; i < TILE / 16; i += 32) ((uint4 *)dst)[i] = src4[i];
    } else {
        const int nbytes = (int)(B - e0) * CELLS;
        for (int i = lane; i < nbytes; i += 32) dst[i] = (int8_t)stage[wid][i];
    }
}

__global__ void ttt_pack_kernel(uint4 *__restrict__ st, long long B, TTTParams prm, const int8_t *__restrict__ board,
                                const int8_t *__restrict__ winner, const int8_t *__restrict__ mover) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B) return;
    TTTEnv s;
    ttt_new_state(s);
    for (int c = 0; c < prm.cells; c++) {
        int v = board[e * prm.cells + c];
#pragma unroll
        for (int p = 0; p < 4; p++) s.m[p] |= (v == p) ? (1u << c) : 0u;
    }
    s.winner1 = winner[e] + 1;
    s.mover = mover[e];
    st[e] = ttt_encode(s);
}
