// Tron: batched grid movement + collision + rewards/terminal/winners/ranking in ONE kernel (sm_100a).
//
// Replaces (reference file:line):
//   next_state_inplace            envs/tron/CyTronGrid.pyx:3-62
//   TronGridEnvironment.next_state envs/tron/TronGridEnvironment.py:265-323  (alive set, rewards, terminal, winners)
//   TronGridEnvironment.compute_ranking :483-508                            (fused; incl. the deaths[-1] wrap-around)
//   TronGridEnvironment.new_state :228-263                                  (tron_reset_kernel / auto-reset)
//   state_to_observation :363-405 + relative_player_inplace CyTronGrid.pyx:65-71 (tron_observe_kernel)
//
// HBM layout (SoA of 16-byte vectors, [13][B] uint4, 208 B per environment):
//   vector v = 3*p + j (p = player plane 0..3, j = 0..2): 64-bit words 2j and 2j+1 of plane p's bitboard,
//             bit (y*N + x) of the 384-bit plane is set iff player p owns cell (x, y)   (N*N <= 384, N <= 19)
//   vector 12 = header: .x = head0 | head1 << 16, .y = head2 | head3 << 16   (head = x | y << 8)
//                       .z = directions (2 bit each, bits 0-7) | deaths (3 bit each, bits 8-19) | terminal << 20
//                       .w = steps taken in the current episode
// Thread t of a warp loads vector v of environment e0 + t: every load/store instruction of a warp is one
// contiguous, 512-byte, 128-bit-per-lane access.  One thread owns one environment (the dynamics are a
// 4-iteration dependent chain; there is nothing to share between lanes).
#pragma once
#include "crl_common.cuh"
#include "philox.cuh"

#define TRON_VEC 13
#define TRON_WORDS 6

struct TronParams {
    int N, P;
    uint32_t start_head[4];             // x | y << 8
    uint32_t start_dirs;                // 2 bits per player
    uint64_t start_pl[4][TRON_WORDS];   // bitboards of new_state(): one bit per player at its spawn
};

struct TronEnv {
    uint64_t pl[4][TRON_WORDS];
    int hx[4], hy[4], dir[4], death[4];
    uint32_t terminal, ep_len;
};

struct TronOut {
    int reward[4];
    int alive, winners, terminal, rank[4], cells[4];
};

// Register arrays are only ever indexed with compile-time constants (a dynamic index would demote them to
// local memory); run-time selection is done with masks.
__device__ __forceinline__ int tron_sel4(const int (&a)[4], int k) {
    int r = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) r |= a[i] & -(int)(k == i);
    return r;
}

__device__ __forceinline__ void tron_load(TronEnv &s, const uint4 *__restrict__ st, long long B, long long e) {
    uint4 v[TRON_VEC];
#pragma unroll
    for (int i = 0; i < TRON_VEC; i++) v[i] = ld_stream(st + (long long)i * B + e);
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            s.pl[p][2 * j] = (uint64_t)v[3 * p + j].x | ((uint64_t)v[3 * p + j].y << 32);
            s.pl[p][2 * j + 1] = (uint64_t)v[3 * p + j].z | ((uint64_t)v[3 * p + j].w << 32);
        }
    uint4 h = v[12];
    uint32_t hd[4] = {h.x & 0xffffu, h.x >> 16, h.y & 0xffffu, h.y >> 16};
#pragma unroll
    for (int p = 0; p < 4; p++) {
        s.hx[p] = hd[p] & 0xff; s.hy[p] = hd[p] >> 8;
        s.dir[p] = (h.z >> (2 * p)) & 3;
        s.death[p] = (h.z >> (8 + 3 * p)) & 7;
    }
    s.terminal = (h.z >> 20) & 1;
    s.ep_len = h.w;
}

__device__ __forceinline__ void tron_store(const TronEnv &s, uint4 *__restrict__ st, long long B, long long e) {
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            uint64_t a = s.pl[p][2 * j], b = s.pl[p][2 * j + 1];
            st_stream(st + (long long)(3 * p + j) * B + e,
                      make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32)));
        }
    uint32_t hd[4], z = s.terminal << 20;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        hd[p] = (uint32_t)s.hx[p] | ((uint32_t)s.hy[p] << 8);
        z |= (uint32_t)s.dir[p] << (2 * p);
        z |= (uint32_t)s.death[p] << (8 + 3 * p);
    }
    st_stream(st + (long long)12 * B + e, make_uint4(hd[0] | hd[1] << 16, hd[2] | hd[3] << 16, z, s.ep_len));
}

// new_state (TronGridEnvironment.py:228-263): empty board, p+1 written at each head (:261)
__device__ __forceinline__ void tron_new_state(TronEnv &s, const TronParams &prm) {
#pragma unroll
    for (int p = 0; p < 4; p++) {
#pragma unroll
        for (int w = 0; w < TRON_WORDS; w++) s.pl[p][w] = prm.start_pl[p][w];
        s.hx[p] = prm.start_head[p] & 0xff; s.hy[p] = prm.start_head[p] >> 8;
        s.dir[p] = (prm.start_dirs >> (2 * p)) & 3;
        s.death[p] = 0;
    }
    s.terminal = 0; s.ep_len = 0;
}

// One env-step: CyTronGrid.pyx:15-62 (players strictly in index order against the already-updated board),
// then TronGridEnvironment.py:309-321 and the ranking of :483-508.
__device__ __forceinline__ void tron_step_env(TronEnv &s, const int (&act)[4], const TronParams &prm, TronOut &o) {
    const int N = prm.N, P = prm.P;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i < P && s.death[i] == 0) {                                  // pyx:16
            int d = (s.dir[i] + act[i] + 4) & 3;                         // pyx:31
            int x = s.hx[i] + (d == 1) - (d == 3);                       // pyx:34-41
            int y = s.hy[i] + (d == 2) - (d == 0);
            s.dir[i] = d;                                                // pyx:44 (also when i dies)
            if ((unsigned)x >= (unsigned)N || (unsigned)y >= (unsigned)N) {
                s.death[i] = i + 1;                                      // pyx:47-48
            } else {
                int c = y * N + x, w = c >> 6;
                uint64_t bit = 1ull << (c & 63), m[TRON_WORDS];
#pragma unroll
                for (int ww = 0; ww < TRON_WORDS; ww++) m[ww] = (ww == w) ? bit : 0ull;
                int owner = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint64_t hit = 0;
#pragma unroll
                    for (int ww = 0; ww < TRON_WORDS; ww++) hit |= s.pl[q][ww] & m[ww];
                    owner = hit ? q + 1 : owner;
                }
                if (owner) {
                    s.death[i] = owner;                                  // pyx:51-53
#pragma unroll
                    for (int q = 0; q < 4; q++)                          // pyx:56-57 (no liveness check: T3)
                        if (owner == q + 1 && s.hx[q] == x && s.hy[q] == y) s.death[q] = i + 1;
                } else {
#pragma unroll
                    for (int ww = 0; ww < TRON_WORDS; ww++) s.pl[i][ww] |= m[ww];   // pyx:60-62
                    s.hx[i] = x; s.hy[i] = y;
                }
            }
        }
    }
    int alive = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) alive |= (p < P && s.death[p] == 0) ? (1 << p) : 0;   // py:310
    o.alive = alive;
    o.terminal = __popc(alive) <= 1;                                     // py:316
    o.winners = o.terminal ? alive : 0;                                  // py:319
#pragma unroll
    for (int p = 0; p < 4; p++) {
        int r = (p < P) ? ((alive >> p & 1) ? 1 : -1) : 0;               // py:313
        if (o.winners >> p & 1) r += 9;                                  // py:320-321
        o.reward[p] = r;
    }
    // ---- compute_ranking (py:483-508)
    int score[4];
#pragma unroll
    for (int p = 0; p < 4; p++) {
        int n = 0;
#pragma unroll
        for (int w = 0; w < TRON_WORDS; w++) n += __popcll(s.pl[p][w]);
        score[p] = n; o.cells[p] = n;
    }
    int tie = 0;  // tie_locations are evaluated up front (py:492); deaths[-1] addresses the LAST player
#pragma unroll
    for (int p = 0; p < 4; p++) {
        int k = s.death[p] ? s.death[p] - 1 : P - 1;
        if (p < P && tron_sel4(s.death, k) == p + 1) tie |= 1 << p;
    }
#pragma unroll
    for (int p = 0; p < 4; p++) {                                        // py:493-495, ascending, in place
        if (tie >> p & 1) {
            int ks = s.death[p] ? tron_sel4(score, s.death[p] - 1) : 0;  // Counter[-1] reads 0
            score[p] = min(score[p], ks);
        }
    }
#pragma unroll
    for (int p = 0; p < 4; p++) {                                        // competition ranking (py:497-506)
        int r = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) r += (q < P && score[q] > score[p]) ? 1 : 0;
        o.rank[p] = (p < P) ? r : 0;
    }
    s.terminal = (uint32_t)o.terminal;
    s.ep_len += 1;
}

// result record, 8 bytes per environment: int8 reward[4], u8 terminal, u8 alive mask, u8 winners mask,
// u8 ranking (2 bits per player)
__device__ __forceinline__ uint2 tron_pack_result(const TronOut &o) {
    uint32_t a = 0, rk = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        a |= ((uint32_t)o.reward[p] & 0xffu) << (8 * p);
        rk |= (uint32_t)o.rank[p] << (2 * p);
    }
    return make_uint2(a, (uint32_t)o.terminal | (uint32_t)o.alive << 8 | (uint32_t)o.winners << 16 | rk << 24);
}

__device__ __forceinline__ void tron_stats(const BlockStats &bs, bool valid, const TronOut &o, uint32_t ep_len) {
    int t = valid && o.terminal;
    bs.add(ST_STEPS, valid ? 1 : 0);
    bs.add(ST_EPISODES, t);
    bs.add(ST_EPLEN, t ? (int)ep_len : 0);
    bs.add(ST_NOWIN, t && o.winners == 0);
    int rw = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        bs.add(ST_WINS + p, t ? (o.winners >> p & 1) : 0);
        bs.add(ST_SCORE + p, t ? o.cells[p] : 0);
        bs.add(ST_RANK + p, t ? o.rank[p] : 0);
        rw += (p + 1) * o.reward[p];
    }
    bs.add(ST_REWARD, valid ? rw : 0);
}

__device__ __forceinline__ void tron_zero_out(TronOut &o) {
#pragma unroll
    for (int p = 0; p < 4; p++) { o.reward[p] = 0; o.rank[p] = 0; o.cells[p] = 0; }
    o.alive = o.winners = o.terminal = 0;
}

// actions: int8[B][4] (0 forward, +1 right, -1 left  == STRING_TO_ACTION, TronGridEnvironment.py:62-67), one
// coalesced 32-bit load per environment.
__global__ void __launch_bounds__(128)
tron_step_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, const uint32_t *__restrict__ actions,
                 uint2 *__restrict__ result, crl_u64 *stats, long long B, TronParams prm, int flags) {
    __shared__ int sm_stat[CRL_NSTAT];
    BlockStats bs{sm_stat};
    if (stats) bs.init();
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = e < B;
    TronOut o;
    tron_zero_out(o);
    uint32_t ep_len = 0;
    if (valid) {
        TronEnv s;
        tron_load(s, in, B, e);
        uint32_t a = actions[e];
        if ((flags & CRL_FLAG_AUTO_RESET) && s.terminal) tron_new_state(s, prm);
        int act[4];
#pragma unroll
        for (int p = 0; p < 4; p++) act[p] = (int)(int8_t)(a >> (8 * p));
        tron_step_env(s, act, prm, o);
        ep_len = s.ep_len;
        tron_store(s, out, B, e);
        result[e] = tron_pack_result(o);
    }
    if (stats) {
        tron_stats(bs, valid, o, ep_len);
        bs.flush(stats);
    }
}

// K fused steps with the in-kernel Philox random policy (action of player p = {0,+1,-1}[r_p % 3]); the state
// stays in registers between steps.  Benchmark / self-play helper; semantics identical to K tron_step calls
// with CRL_FLAG_AUTO_RESET.
__global__ void __launch_bounds__(128)
tron_rollout_kernel(uint4 *__restrict__ state, uint2 *__restrict__ result, crl_u64 *stats, long long B,
                    TronParams prm, crl_u64 seed, crl_u64 first_env, uint32_t step0, int K) {
    __shared__ int sm_stat[CRL_NSTAT];
    BlockStats bs{sm_stat};
    if (stats) bs.init();
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = e < B;
    TronEnv s;
    TronOut o;
    tron_zero_out(o);
    if (valid) tron_load(s, state, B, e); else tron_new_state(s, prm);
    for (int k = 0; k < K; k++) {
        if (valid) {
            if (s.terminal) tron_new_state(s, prm);
            uint4 r = env_words(seed, first_env + (crl_u64)e, step0 + (uint32_t)k, CRL_TAG_TRON);
            uint32_t rr[4] = {r.x, r.y, r.z, r.w};
            int act[4];
#pragma unroll
            for (int p = 0; p < 4; p++) { int m = (int)(rr[p] % 3u); act[p] = (m == 2) ? -1 : m; }
            tron_step_env(s, act, prm, o);
        }
        if (stats) tron_stats(bs, valid, o, s.ep_len);
    }
    if (valid) {
        tron_store(s, state, B, e);
        if (result) result[e] = tron_pack_result(o);
    }
    if (stats) bs.flush(stats);
}

__global__ void tron_policy_random_kernel(uint32_t *__restrict__ actions, long long B, crl_u64 seed,
                                          crl_u64 first_env, uint32_t step) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B) return;
    uint4 r = env_words(seed, first_env + (crl_u64)e, step, CRL_TAG_TRON);
    uint32_t rr[4] = {r.x, r.y, r.z, r.w}, a = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) { uint32_t m = rr[p] % 3u; a |= ((m == 2) ? 0xffu : m) << (8 * p); }
    actions[e] = a;
}

__global__ void __launch_bounds__(128)
tron_reset_kernel(uint4 *__restrict__ state, const uint8_t *__restrict__ mask, long long B, TronParams prm) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B || (mask && !mask[e])) return;
    TronEnv s;
    tron_new_state(s, prm);
    tron_store(s, state, B, e);
}

// state_to_observation (TronGridEnvironment.py:385-405): one thread per board cell.  player < 0 => absolute
// view (plain unpack: board values p+1, vectors unrolled).  board int8[B][N][N]; heads/dirs/deaths int32[B][P]
// with heads as y*N + x like the reference.
__global__ void tron_observe_kernel(const uint4 *__restrict__ st, long long B, TronParams prm, int player,
                                    int8_t *__restrict__ board, int32_t *__restrict__ heads,
                                    int32_t *__restrict__ dirs, int32_t *__restrict__ deaths,
                                    uint8_t *__restrict__ terminal) {
    const int N = prm.N, P = prm.P, NN = N * N;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * NN) return;
    long long e = idx / NN;
    int c = (int)(idx - e * NN);
    int w = c >> 6, b = c & 63;
    int v = 0;
    for (int p = 0; p < P; p++) {
        uint4 q = st[(long long)(3 * p + (w >> 1)) * B + e];
        uint64_t word = (w & 1) ? ((uint64_t)q.z | (uint64_t)q.w << 32) : ((uint64_t)q.x | (uint64_t)q.y << 32);
        if ((word >> b) & 1) v = p + 1;
    }
    if (v > 0 && player >= 0) v = ((v - (player + 1) + P) % P) + 1;      // CyTronGrid.pyx:65-71
    board[idx] = (int8_t)v;
    if (c < P) {
        uint4 h = st[(long long)12 * B + e];
        int src = player >= 0 ? (c + player) % P : c;                    // py:392
        uint32_t hd = (src < 2 ? h.x : h.y) >> (16 * (src & 1)) & 0xffffu;
        if (heads) heads[e * P + c] = (int)(hd >> 8) * N + (int)(hd & 0xff);
        if (dirs) dirs[e * P + c] = (h.z >> (2 * src)) & 3;
        if (deaths) deaths[e * P + c] = (h.z >> (8 + 3 * src)) & 7;
        if (terminal && c == 0) terminal[e] = (h.z >> 20) & 1;
    }
}

// import a reference-layout state (board int8[B][N][N], heads (y*N+x) / directions / deaths int32[B][P])
__global__ void tron_pack_kernel(uint4 *__restrict__ st, long long B, TronParams prm,
                                 const int8_t *__restrict__ board, const int32_t *__restrict__ heads,
                                 const int32_t *__restrict__ dirs, const int32_t *__restrict__ deaths) {
    const int N = prm.N, P = prm.P, NN = N * N;
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B) return;
    TronEnv s;
    tron_new_state(s, prm);
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int w = 0; w < TRON_WORDS; w++) s.pl[p][w] = 0;
    for (int c = 0; c < NN; c++) {
        int v = board[e * NN + c];
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
            for (int w = 0; w < TRON_WORDS; w++)
                s.pl[p][w] |= (v == p + 1 && w == (c >> 6)) ? (1ull << (c & 63)) : 0ull;
    }
    int alive = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        if (p < P) {
            int h = heads[e * P + p];
            s.hx[p] = h % N; s.hy[p] = h / N;
            s.dir[p] = dirs[e * P + p] & 3;
            s.death[p] = deaths[e * P + p] & 7;
            alive += s.death[p] == 0;
        }
    }
    s.terminal = alive <= 1;
    s.ep_len = 0;
    tron_store(s, st, B, e);
}
