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
//   vector 12 = header:
//     .x = head0 | head1 << 10 | head2 << 20                      (head = x | y << 5)
//     .y = head3 | directions << 10 (2 bit each) | deaths << 18 (3 bit each) | terminal << 30
//     .z = cells0 | cells1 << 9 | cells2 << 18                    (#cells owned = compute_ranking's score)
//     .w = cells3 | episode steps << 9
//
// Data movement: vector v of the TILE environments of a CTA is ONE contiguous global segment (TILE*16 bytes), so
// the whole tile travels as 13 bulk-async copies (TMA, cp.async.bulk -> SASS UBLKCP) issued by one thread into
// shared memory and 13 back; the planes are never unpacked into registers.  Each thread owns one environment and
// only touches the <= 4 bitboard words its players move into (dynamic word index = a shared-memory address, not
// a register-select chain) plus the 16-byte header.  The #cells per player is carried in the header so the
// ranking needs no popcounts.
#pragma once
#include "crl_common.cuh"
#include "philox.cuh"

#define TRON_VEC 13
#define TRON_WORDS 6
#define TRON_TILE 64          // environments (= threads) per CTA: 13 KB of shared memory, ~7 CTAs per SM at B = 65,536

struct TronParams {
    int N, P;
    uint32_t start_hdr[4];              // header of new_state()
    uint64_t start_pl[4][TRON_WORDS];   // bitboards of new_state(): one bit per player at its spawn
};

struct TronHdr {
    int hx[4], hy[4], dir[4], death[4], cells[4];
    uint32_t terminal, ep_len;
};

struct TronOut {
    int reward[4];
    int alive, winners, terminal, rank[4], cells[4];
};

// Register arrays are only ever indexed with compile-time constants; run-time selection is done with masks.
__device__ __forceinline__ int tron_sel4(const int (&a)[4], int k) {
    int r = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) r |= a[i] & -(int)(k == i);
    return r;
}

__device__ __forceinline__ void tron_hdr_decode(TronHdr &s, uint4 h) {
    uint32_t hd[4] = {h.x & 1023u, (h.x >> 10) & 1023u, (h.x >> 20) & 1023u, h.y & 1023u};
    uint32_t ce[4] = {h.z & 511u, (h.z >> 9) & 511u, (h.z >> 18) & 511u, h.w & 511u};
#pragma unroll
    for (int p = 0; p < 4; p++) {
        s.hx[p] = hd[p] & 31; s.hy[p] = hd[p] >> 5;
        s.dir[p] = (h.y >> (10 + 2 * p)) & 3;
        s.death[p] = (h.y >> (18 + 3 * p)) & 7;
        s.cells[p] = (int)ce[p];
    }
    s.terminal = (h.y >> 30) & 1;
    s.ep_len = h.w >> 9;
}

__host__ __device__ __forceinline__ uint4 tron_hdr_encode(const TronHdr &s) {
    uint32_t hd[4], y = s.terminal << 30;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        hd[p] = (uint32_t)s.hx[p] | ((uint32_t)s.hy[p] << 5);
        y |= (uint32_t)s.dir[p] << (10 + 2 * p);
        y |= (uint32_t)s.death[p] << (18 + 3 * p);
    }
    return make_uint4(hd[0] | hd[1] << 10 | hd[2] << 20, hd[3] | y,
                      (uint32_t)s.cells[0] | (uint32_t)s.cells[1] << 9 | (uint32_t)s.cells[2] << 18,
                      (uint32_t)s.cells[3] | s.ep_len << 9);
}

// ---- the shared-memory tile ---------------------------------------------------------------------------------
struct TronTile {
    uint4 v[TRON_VEC][TRON_TILE];
};

// 64-bit word w of plane p of the environment owned by thread t
__device__ __forceinline__ uint64_t *tron_word(TronTile &tile, int t, int p, int w) {
    return reinterpret_cast<uint64_t *>(&tile.v[3 * p + (w >> 1)][t]) + (w & 1);
}

// global -> shared (n <= TRON_TILE environments starting at e0)
__device__ __forceinline__ void tron_tile_load(TronTile &tile, uint64_t *bar, const uint4 *__restrict__ st,
                                               long long B, long long e0, int n) {
#ifndef CRL_HOSTSIM
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)(TRON_VEC * n * 16));
#pragma unroll
        for (int v = 0; v < TRON_VEC; v++) bulk_g2s(&tile.v[v][0], st + (long long)v * B + e0, (uint32_t)(n * 16), bar);
    }
    __syncthreads();           // barrier initialisation visible to the waiters
    mbar_wait(bar, 0);
#else
    for (int v = 0; v < TRON_VEC; v++)
        if ((int)threadIdx.x < n) tile.v[v][threadIdx.x] = st[(long long)v * B + e0 + threadIdx.x];
    __syncthreads();
#endif
}

// shared -> global
__device__ __forceinline__ void tron_tile_store(TronTile &tile, uint4 *__restrict__ st, long long B, long long e0, int n) {
#ifndef CRL_HOSTSIM
    fence_async_smem();        // generic-proxy writes to the tile -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int v = 0; v < TRON_VEC; v++) bulk_s2g(st + (long long)v * B + e0, &tile.v[v][0], (uint32_t)(n * 16));
        bulk_commit_wait_read();   // the tile may be released once the copies have read it
    }
#else
    __syncthreads();
    for (int v = 0; v < TRON_VEC; v++)
        if ((int)threadIdx.x < n) st[(long long)v * B + e0 + threadIdx.x] = tile.v[v][threadIdx.x];
    __syncthreads();
#endif
}

// new_state (TronGridEnvironment.py:228-263): empty board, p+1 written at each head (:261)
__device__ __forceinline__ void tron_new_state(TronTile &tile, int t, TronHdr &s, const TronParams &prm) {
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            uint64_t a = prm.start_pl[p][2 * j], b = prm.start_pl[p][2 * j + 1];
            tile.v[3 * p + j][t] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
        }
    tron_hdr_decode(s, make_uint4(prm.start_hdr[0], prm.start_hdr[1], prm.start_hdr[2], prm.start_hdr[3]));
}

// One env-step: CyTronGrid.pyx:15-62 (players strictly in index order against the already-updated board),
// then TronGridEnvironment.py:309-321 and the ranking of :483-508.
//
// The reference's loop is sequential in the player index, but the only things a later player can observe from an
// earlier one in the same step are (a) the single cell it just claimed, (b) its new head and (c) a head-on kill.
// So all shared-memory lookups are done up front against the OLD board (16 independent loads in flight), the
// sequential part runs on registers only, and the <= 4 bit sets are independent stores (plane i is written by
// player i alone).
__device__ __forceinline__ void tron_step_env(TronTile &tile, int t, TronHdr &s, const int (&act)[4],
                                              const TronParams &prm, TronOut &o) {
    const int N = prm.N, P = prm.P;
    int nd[4], nx[4], ny[4], own_old[4];
    bool inb[4], moved[4];
    uint64_t *wp[4], wold[4], bitm[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        nd[i] = (s.dir[i] + act[i] + 4) & 3;                             // pyx:31
        nx[i] = s.hx[i] + (nd[i] == 1) - (nd[i] == 3);                   // pyx:34-41
        ny[i] = s.hy[i] + (nd[i] == 2) - (nd[i] == 0);
        inb[i] = (unsigned)nx[i] < (unsigned)N && (unsigned)ny[i] < (unsigned)N;
        const int c = inb[i] ? ny[i] * N + nx[i] : 0, w = c >> 6;
        bitm[i] = 1ull << (c & 63);
        own_old[i] = 0;
        moved[i] = false;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint64_t *ptr = tron_word(tile, t, q, w);
            const uint64_t v = *ptr;
            own_old[i] = (v & bitm[i]) ? q + 1 : own_old[i];
            if (q == i) { wp[i] = ptr; wold[i] = v; }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i < P && s.death[i] == 0) {                                  // pyx:16 (a head-on kill by j < i counts)
            s.dir[i] = nd[i];                                            // pyx:44 (also when i dies)
            if (!inb[i]) {
                s.death[i] = i + 1;                                      // pyx:47-48
            } else {
                int owner = own_old[i];
#pragma unroll
                for (int j = 0; j < i; j++)                              // the cell player j claimed this step
                    owner = (moved[j] && nx[j] == nx[i] && ny[j] == ny[i]) ? j + 1 : owner;
                if (owner) {
                    s.death[i] = owner;                                  // pyx:51-53
#pragma unroll
                    for (int q = 0; q < 4; q++)                          // pyx:56-57 (no liveness check: T3)
                        if (owner == q + 1 && s.hx[q] == nx[i] && s.hy[q] == ny[i]) s.death[q] = i + 1;
                } else {
                    moved[i] = true;                                     // pyx:60-62
                    s.hx[i] = nx[i]; s.hy[i] = ny[i];
                    s.cells[i] += 1;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (moved[i]) *wp[i] = wold[i] | bitm[i];
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
    // ---- compute_ranking (py:483-508); score = #cells owned, carried in the header
    int score[4];
#pragma unroll
    for (int p = 0; p < 4; p++) { score[p] = s.cells[p]; o.cells[p] = s.cells[p]; }
    int tie = 0;  // tie_locations are evaluated up front (py:492); deaths[-1] addresses the LAST player
    const uint32_t d12 = (uint32_t)s.death[0] | (uint32_t)s.death[1] << 3 | (uint32_t)s.death[2] << 6 | (uint32_t)s.death[3] << 9;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        int k = s.death[p] ? s.death[p] - 1 : P - 1;
        if (p < P && ((d12 >> (3 * k)) & 7u) == (uint32_t)(p + 1)) tie |= 1 << p;
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

// Episode statistics of one step: 19 counters packed into 6 words so a warp needs 6 redux.sync, not 19.
__device__ __forceinline__ void tron_stats(int *sm_stat, bool valid, const TronOut &o, uint32_t ep_len) {
    const int t = valid && o.terminal;
    int rw = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) rw += (p + 1) * o.reward[p];
    const int w0 = o.winners & -t, nr = t;
    // every field is a sum of <= 32 lane values and stays inside its bit range
    uint32_t A = (valid ? 1u : 0u) | (uint32_t)t << 8 | (uint32_t)(t && o.winners == 0) << 16 | (uint32_t)(w0 & 1) << 24;
    uint32_t Bw = (uint32_t)(w0 >> 1 & 1) | (uint32_t)(w0 >> 2 & 1) << 8 | (uint32_t)(w0 >> 3 & 1) << 16;
    uint32_t C = nr ? ((uint32_t)o.rank[0] | (uint32_t)o.rank[1] << 8 | (uint32_t)o.rank[2] << 16 | (uint32_t)o.rank[3] << 24) : 0u;
    uint32_t D = (t ? ep_len : 0u) | (uint32_t)((valid ? rw : 0) + 16) << 16;          // reward sum biased by +16 per lane
    uint32_t E = t ? ((uint32_t)o.cells[0] | (uint32_t)o.cells[1] << 16) : 0u;
    uint32_t F = t ? ((uint32_t)o.cells[2] | (uint32_t)o.cells[3] << 16) : 0u;
    A = __reduce_add_sync(0xffffffffu, A); Bw = __reduce_add_sync(0xffffffffu, Bw);
    C = __reduce_add_sync(0xffffffffu, C); D = __reduce_add_sync(0xffffffffu, D);
    E = __reduce_add_sync(0xffffffffu, E); F = __reduce_add_sync(0xffffffffu, F);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sm_stat[ST_STEPS], (int)(A & 255u));
        if (A >> 8) {
            atomicAdd(&sm_stat[ST_EPISODES], (int)(A >> 8 & 255u));
            atomicAdd(&sm_stat[ST_NOWIN], (int)(A >> 16 & 255u));
            atomicAdd(&sm_stat[ST_WINS + 0], (int)(A >> 24));
            atomicAdd(&sm_stat[ST_WINS + 1], (int)(Bw & 255u));
            atomicAdd(&sm_stat[ST_WINS + 2], (int)(Bw >> 8 & 255u));
            atomicAdd(&sm_stat[ST_WINS + 3], (int)(Bw >> 16 & 255u));
            atomicAdd(&sm_stat[ST_RANK + 0], (int)(C & 255u));
            atomicAdd(&sm_stat[ST_RANK + 1], (int)(C >> 8 & 255u));
            atomicAdd(&sm_stat[ST_RANK + 2], (int)(C >> 16 & 255u));
            atomicAdd(&sm_stat[ST_RANK + 3], (int)(C >> 24));
            atomicAdd(&sm_stat[ST_EPLEN], (int)(D & 0xffffu));
            atomicAdd(&sm_stat[ST_SCORE + 0], (int)(E & 0xffffu));
            atomicAdd(&sm_stat[ST_SCORE + 1], (int)(E >> 16));
            atomicAdd(&sm_stat[ST_SCORE + 2], (int)(F & 0xffffu));
            atomicAdd(&sm_stat[ST_SCORE + 3], (int)(F >> 16));
        }
        atomicAdd(&sm_stat[ST_REWARD], (int)(D >> 16) - 16 * 32);
    }
}

__device__ __forceinline__ void tron_zero_out(TronOut &o) {
#pragma unroll
    for (int p = 0; p < 4; p++) { o.reward[p] = 0; o.rank[p] = 0; o.cells[p] = 0; }
    o.alive = o.winners = o.terminal = 0;
}

// actions: int8[B][4] (0 forward, +1 right, -1 left  == STRING_TO_ACTION, TronGridEnvironment.py:62-67), one
// coalesced 32-bit load per environment.
__global__ void __launch_bounds__(TRON_TILE)
tron_step_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, const uint32_t *__restrict__ actions,
                 uint2 *__restrict__ result, crl_u64 *stats, long long B, TronParams prm, int flags) {
    __shared__ __align__(128) TronTile tile;
    __shared__ __align__(8) uint64_t bar;
    __shared__ int sm_stat[CRL_NSTAT];
    pdl_launch_dependents();     // let the next launch's CTAs become resident and run their prologue now
    const int t = threadIdx.x;
    const long long e0 = (long long)blockIdx.x * TRON_TILE;
    const int n = (int)min((long long)TRON_TILE, B - e0);
    const bool valid = t < n;
    if (stats && t < CRL_NSTAT) sm_stat[t] = 0;
    pdl_wait();                  // everything above overlapped with the previous kernel's tail
    const uint32_t a = valid ? actions[e0 + t] : 0u;       // overlaps with the tile's flight
    tron_tile_load(tile, &bar, in, B, e0, n);
    TronOut o;
    tron_zero_out(o);
    uint32_t ep_len = 0;
    if (valid) {
        TronHdr s;
        tron_hdr_decode(s, tile.v[12][t]);
        if ((flags & CRL_FLAG_AUTO_RESET) && s.terminal) tron_new_state(tile, t, s, prm);
        int act[4];
#pragma unroll
        for (int p = 0; p < 4; p++) act[p] = (int)(int8_t)(a >> (8 * p));
        if (!(flags & 0x100)) tron_step_env(tile, t, s, act, prm, o);     // 0x100: diagnostics (tools/): data movement only
        ep_len = s.ep_len;
        tile.v[12][t] = tron_hdr_encode(s);
        result[e0 + t] = tron_pack_result(o);
    }
    tron_tile_store(tile, out, B, e0, n);
    if (stats) {
        tron_stats(sm_stat, valid, o, ep_len);
        __syncthreads();
        stats_flush_row(sm_stat, stats);
    }
}

// K fused steps with the in-kernel Philox random policy (action of player p = {0,+1,-1}[r_p % 3]); the tile
// stays in shared memory between steps.  Benchmark / self-play helper; semantics identical to K tron_step calls
// with CRL_FLAG_AUTO_RESET.
__global__ void __launch_bounds__(TRON_TILE)
tron_rollout_kernel(uint4 *__restrict__ state, uint2 *__restrict__ result, crl_u64 *stats, long long B,
                    TronParams prm, crl_u64 seed, crl_u64 first_env, uint32_t step0, int K) {
    __shared__ __align__(128) TronTile tile;
    __shared__ __align__(8) uint64_t bar;
    __shared__ int sm_stat[CRL_NSTAT];
    const int t = threadIdx.x;
    const long long e0 = (long long)blockIdx.x * TRON_TILE;
    const int n = (int)min((long long)TRON_TILE, B - e0);
    const bool valid = t < n;
    if (stats && t < CRL_NSTAT) sm_stat[t] = 0;
    tron_tile_load(tile, &bar, state, B, e0, n);
    TronHdr s;
    TronOut o;
    tron_zero_out(o);
    tron_hdr_decode(s, valid ? tile.v[12][t] : make_uint4(0, 0, 0, 0));
    for (int k = 0; k < K; k++) {
        if (valid) {
            if (s.terminal) tron_new_state(tile, t, s, prm);
            uint4 r = env_words(seed, first_env + (crl_u64)(e0 + t), step0 + (uint32_t)k, CRL_TAG_TRON);
            uint32_t rr[4] = {r.x, r.y, r.z, r.w};
            int act[4];
#pragma unroll
            for (int p = 0; p < 4; p++) { int m = (int)(rr[p] % 3u); act[p] = (m == 2) ? -1 : m; }
            tron_step_env(tile, t, s, act, prm, o);
        }
        if (stats) tron_stats(sm_stat, valid, o, s.ep_len);
    }
    if (valid) {
        tile.v[12][t] = tron_hdr_encode(s);
        if (result) result[e0 + t] = tron_pack_result(o);
    }
    tron_tile_store(tile, state, B, e0, n);
    if (stats) {
        __syncthreads();
        stats_flush_row(sm_stat, stats);
    }
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

// new_state for all / masked environments: one thread per (vector, environment), fully coalesced stores
__global__ void tron_reset_kernel(uint4 *__restrict__ state, const uint8_t *__restrict__ mask, long long B, TronParams prm) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * TRON_VEC) return;
    const int v = (int)(idx / B);
    const long long e = idx - (long long)v * B;
    if (mask && !mask[e]) return;
    uint4 val = make_uint4(prm.start_hdr[0], prm.start_hdr[1], prm.start_hdr[2], prm.start_hdr[3]);
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            if (v == 3 * p + j) {
                uint64_t a = prm.start_pl[p][2 * j], b = prm.start_pl[p][2 * j + 1];
                val = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
            }
    state[idx] = val;
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
        TronHdr s;
        tron_hdr_decode(s, st[(long long)12 * B + e]);
        int src = player >= 0 ? (c + player) % P : c;                    // py:392
        if (heads) heads[e * P + c] = tron_sel4(s.hy, src) * N + tron_sel4(s.hx, src);
        if (dirs) dirs[e * P + c] = tron_sel4(s.dir, src);
        if (deaths) deaths[e * P + c] = tron_sel4(s.death, src);
        if (terminal && c == 0) terminal[e] = (uint8_t)s.terminal;
    }
}

// import a reference-layout state (board int8[B][N][N], heads (y*N+x) / directions / deaths int32[B][P]);
// one thread per environment (import path, not performance critical)
__global__ void tron_pack_kernel(uint4 *__restrict__ st, long long B, TronParams prm,
                                 const int8_t *__restrict__ board, const int32_t *__restrict__ heads,
                                 const int32_t *__restrict__ dirs, const int32_t *__restrict__ deaths) {
    const int N = prm.N, P = prm.P, NN = N * N;
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B) return;
    uint64_t pl[4][TRON_WORDS];
    TronHdr s;
#pragma unroll
    for (int p = 0; p < 4; p++) {
#pragma unroll
        for (int w = 0; w < TRON_WORDS; w++) pl[p][w] = 0;
        s.hx[p] = s.hy[p] = s.dir[p] = s.death[p] = s.cells[p] = 0;
    }
    for (int c = 0; c < NN; c++) {
        int v = board[e * NN + c];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            s.cells[p] += (v == p + 1);
#pragma unroll
            for (int w = 0; w < TRON_WORDS; w++)
                pl[p][w] |= (v == p + 1 && w == (c >> 6)) ? (1ull << (c & 63)) : 0ull;
        }
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
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            uint64_t a = pl[p][2 * j], b = pl[p][2 * j + 1];
            st[(long long)(3 * p + j) * B + e] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
        }
    st[(long long)12 * B + e] = tron_hdr_encode(s);
}
