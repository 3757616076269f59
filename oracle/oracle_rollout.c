/* TEST INFRASTRUCTURE ONLY -- random-policy rollouts over the CPU oracle.
 *
 * Drives oracle_{tron,ttt,blokus}.c with the counter-based Philox4x32-10 action streams of
 * SURVEY.md section 8d so that (a) GPU trajectories can be checked bit-exactly at any batch size
 * through their end states and episode statistics, and (b) bench.py can time the CPU path
 * (pthreads over interleaved environment slices).
 *
 * Policy (identical to csrc/policy.cuh):
 *   words r[0..3] = philox4x32_10(ctr = (env_lo, env_hi, t, tag), key = (seed_lo, seed_hi))
 *   Tron   (tag 1): action of player p = {0, +1, -1}[r[p] % 3]   (players 4..7: [(r[p - 4] / 3) % 3])
 *   Blokus (tag 2): the (r[0] % n)-th entry of the mover's valid list, pass if n == 0
 *   TTT    (tag 3): the (r[0] % n_empty)-th empty cell in C order, pass if none
 * Auto-reset: an environment whose previous step returned terminal is replaced by new_state()
 * before the step is applied, so every step is a real next_state call.
 *
 * Statistics vector (int64[32]), same slots as the device (include/colosseum_b200.h):
 *   [0] env-steps  [1] finished episodes  [2] sum of finished-episode lengths
 *   [3..6] wins per seat  [7] episodes without a winner  [8..11] sum of final scores per seat
 *   [12] invalid-action errors  [13] sum of valid-action counts  [14..17] sum of final ranks per seat
 *   [18] sum over steps of sum_p (p+1)*reward_p
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define NSTAT 32

/* prototypes from the other oracle files */
int orc_tron_new_state(int N, int P, int64_t *board, int64_t *heads, int64_t *directions, int64_t *deaths);
int orc_tron_next_state(int N, int P, int64_t *board, int64_t *heads, int64_t *directions, int64_t *deaths,
                        const int64_t *actions, int64_t *rewards, int *alive_mask, int *winners_mask);
void orc_tron_compute_ranking(int N, int P, const int64_t *board, const int64_t *deaths, int64_t *rank);
int orc_ttt_cells(int n);
int orc_ttt_next_state(int n, int8_t *board, int *winner, int player, int action, int *reward, int *terminal, int *winner_out);
int orc_ttt_valid_actions(int n, const int8_t *board, int32_t *out);
int orc_blokus_valid_moves(const int64_t *board, int round_count, int color, const uint8_t *inventory, int32_t *out_ids, int cap);
int orc_blokus_next_state(int64_t *board, int *round_count, uint8_t *inventory, int64_t *scores, int mover,
                          int action_id, int *reward, int *terminal, int *winners_mask);

/* Philox4x32-10 (Random123; Salmon et al. SC'11). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void env_words(uint64_t seed, uint64_t env, uint32_t t, uint32_t tag, uint32_t r[4]) {
    uint32_t ctr[4] = {(uint32_t)env, (uint32_t)(env >> 32), t, tag};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(ctr, key, r);
}

int orc_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* minimal parallel-for: thread i of T runs fn over envs i, i+T, i+2T, ... (interleaved for balance) */
typedef void (*env_fn)(int64_t e, int64_t *loc, void *arg, void *scratch);
typedef struct { env_fn fn; int64_t B; int tid, T; int64_t loc[NSTAT]; void *arg; size_t scratch_bytes; } job_t;
static void *job_main(void *p) {
    job_t *j = (job_t *)p;
    void *scratch = j->scratch_bytes ? malloc(j->scratch_bytes) : NULL;
    for (int64_t e = j->tid; e < j->B; e += j->T) j->fn(e, j->loc, j->arg, scratch);
    free(scratch);
    return NULL;
}
static void parallel_for(int64_t B, int nthreads, env_fn fn, void *arg, size_t scratch_bytes, int64_t *stats) {
    if (nthreads <= 0) nthreads = orc_num_threads();
    if (nthreads > B) nthreads = (int)(B > 0 ? B : 1);
    job_t *jobs = (job_t *)calloc(nthreads, sizeof(job_t));
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    for (int i = 0; i < nthreads; i++) {
        jobs[i].fn = fn; jobs[i].arg = arg; jobs[i].B = B; jobs[i].tid = i; jobs[i].T = nthreads;
        jobs[i].scratch_bytes = scratch_bytes;
        if (i > 0) pthread_create(&th[i], NULL, job_main, &jobs[i]);
    }
    job_main(&jobs[0]);
    for (int i = 1; i < nthreads; i++) pthread_join(th[i], NULL);
    if (stats) for (int i = 0; i < nthreads; i++) for (int k = 0; k < NSTAT; k++) stats[k] += jobs[i].loc[k];
    free(jobs); free(th);
}

/* ------------------------------------------------------------------ Tron */
typedef struct {
    int N, P; uint64_t seed; int64_t env0; uint32_t t0; int K, fresh;
    int64_t *board, *heads, *directions, *deaths; uint8_t *terminal; int32_t *ep_len;
} tron_args;

static void tron_env(int64_t e, int64_t *loc, void *argp, void *scratch) {
    static const int64_t MOVE[3] = {0, 1, -1};
    tron_args *a = (tron_args *)argp;
    int N = a->N, P = a->P;
    int64_t *bd = a->board + e * N * N, *hd = a->heads + e * P, *dr = a->directions + e * P, *de = a->deaths + e * P;
    if (a->fresh) { orc_tron_new_state(N, P, bd, hd, dr, de); a->terminal[e] = 0; a->ep_len[e] = 0; }
    for (int s = 0; s < a->K; s++) {
        if (a->terminal[e]) { orc_tron_new_state(N, P, bd, hd, dr, de); a->terminal[e] = 0; a->ep_len[e] = 0; }
        uint32_t r[4]; env_words(a->seed, (uint64_t)(a->env0 + e), a->t0 + s, 1, r);
        int64_t act[8] = {0}, rew[8];
        for (int p = 0; p < P && p < 8; p++) act[p] = MOVE[(p < 4 ? r[p] : r[p - 4] / 3) % 3];   /* = oracle/make_golden.py */
        int alive, winners;
        int term = orc_tron_next_state(N, P, bd, hd, dr, de, act, rew, &alive, &winners);
        a->ep_len[e]++; loc[0]++;
        for (int p = 0; p < P; p++) loc[18] += (p + 1) * rew[p];
        if (term) {
            int64_t rank[8];
            orc_tron_compute_ranking(N, P, bd, de, rank);
            loc[1]++; loc[2] += a->ep_len[e];
            if (!winners) loc[7]++;
            for (int p = 0; p < P && p < 4; p++) {
                if (winners >> p & 1) loc[3 + p]++;
                int64_t cells = 0;
                for (int c = 0; c < N * N; c++) cells += bd[c] == p + 1;
                loc[8 + p] += cells;
                loc[14 + p] += rank[p];
            }
            a->terminal[e] = 1;
        }
    }
}

/* Runs K steps (t = t0 .. t0+K-1) for envs [env0, env0+B).  State arrays are [B][...] and are both input
 * and output; terminal[B] (uint8) and ep_len[B] (int32) carry the auto-reset flag / episode step counter.
 * If fresh != 0 the state is initialised with new_state() first. */
void orc_tron_rollout(int N, int P, uint64_t seed, int64_t env0, int64_t B, uint32_t t0, int K, int fresh,
                      int64_t *board, int64_t *heads, int64_t *directions, int64_t *deaths,
                      uint8_t *terminal, int32_t *ep_len, int64_t *stats, int nthreads) {
    tron_args a = {N, P, seed, env0, t0, K, fresh, board, heads, directions, deaths, terminal, ep_len};
    parallel_for(B, nthreads, tron_env, &a, 0, stats);
}

/* ------------------------------------------------------------------ TTT */
typedef struct {
    int n; uint64_t seed; int64_t env0; uint32_t t0; int K, fresh;
    int8_t *board; int32_t *winner, *mover; uint8_t *terminal; int32_t *ep_len;
} ttt_args;

static void ttt_env(int64_t e, int64_t *loc, void *argp, void *scratch) {
    ttt_args *a = (ttt_args *)argp;
    int n = a->n, cells = orc_ttt_cells(n);
    int8_t *bd = a->board + e * cells;
    for (int s = 0; s < a->K; s++) {
        if ((a->fresh && s == 0) || a->terminal[e]) {
            memset(bd, -1, cells); a->winner[e] = -1; a->mover[e] = 0; a->terminal[e] = 0; a->ep_len[e] = 0;
        }
        uint32_t r[4]; env_words(a->seed, (uint64_t)(a->env0 + e), a->t0 + s, 3, r);
        int32_t va[32];
        int nv = orc_ttt_valid_actions(n, bd, va);
        int action = nv ? va[r[0] % (uint32_t)nv] : -1;
        int rew, term, wout, w = a->winner[e], m = a->mover[e];
        int nxt = orc_ttt_next_state(n, bd, &w, m, action, &rew, &term, &wout);
        loc[18] += (int64_t)(m + 1) * rew;
        a->winner[e] = w; a->mover[e] = nxt;
        a->ep_len[e]++; loc[0]++; loc[13] += nv;
        if (term) {
            loc[1]++; loc[2] += a->ep_len[e];
            if (wout < 0) loc[7]++; else loc[3 + wout]++;
            for (int p = 0; p < n; p++) loc[14 + p] += (p == wout) ? 0 : 1; /* default ranking */
            a->terminal[e] = 1;
        }
    }
}

void orc_ttt_rollout(int n, uint64_t seed, int64_t env0, int64_t B, uint32_t t0, int K, int fresh,
                     int8_t *board, int32_t *winner, int32_t *mover, uint8_t *terminal, int32_t *ep_len,
                     int64_t *stats, int nthreads) {
    ttt_args a = {n, seed, env0, t0, K, fresh, board, winner, mover, terminal, ep_len};
    parallel_for(B, nthreads, ttt_env, &a, 0, stats);
}

/* ------------------------------------------------------------------ Blokus */
typedef struct {
    uint64_t seed; int64_t env0; uint32_t t0; int K, fresh;
    int64_t *board; int32_t *round_count; uint8_t *inventory; int64_t *scores; int32_t *mover;
    uint8_t *terminal; int32_t *ep_len;
} blokus_args;

#define BLOKUS_CAP 16384
static void blokus_env(int64_t e, int64_t *loc, void *argp, void *scratch) {
    blokus_args *a = (blokus_args *)argp;
    int32_t *ids = (int32_t *)scratch;
    int64_t *bd = a->board + e * 400, *sc = a->scores + e * 4;
    uint8_t *inv = a->inventory + e * 84;
    for (int s = 0; s < a->K; s++) {
        if ((a->fresh && s == 0) || a->terminal[e]) {
            memset(bd, 0, sizeof(int64_t) * 400); memset(inv, 1, 84); memset(sc, 0, sizeof(int64_t) * 4);
            a->round_count[e] = 0; a->mover[e] = 0; a->terminal[e] = 0; a->ep_len[e] = 0;
        }
        uint32_t r[4]; env_words(a->seed, (uint64_t)(a->env0 + e), a->t0 + s, 2, r);
        int m = a->mover[e];
        int nv = orc_blokus_valid_moves(bd, a->round_count[e], m + 1, inv + m * 21, ids, BLOKUS_CAP);
        int action = nv ? ids[r[0] % (uint32_t)nv] : -1;
        int rew, term, winners, rc = a->round_count[e];
        int nxt = orc_blokus_next_state(bd, &rc, inv, sc, m, action, &rew, &term, &winners);
        loc[18] += (int64_t)(m + 1) * rew;
        a->round_count[e] = rc; a->mover[e] = nxt;
        a->ep_len[e]++; loc[0]++; loc[13] += nv;
        if (term) {
            loc[1]++; loc[2] += a->ep_len[e];
            if (!winners) loc[7]++;
            for (int p = 0; p < 4; p++) {
                if (winners >> p & 1) loc[3 + p]++;
                loc[8 + p] += sc[p];
                loc[14 + p] += (winners >> p & 1) ? 0 : 1;
            }
            a->terminal[e] = 1;
        }
    }
}

void orc_blokus_rollout(uint64_t seed, int64_t env0, int64_t B, uint32_t t0, int K, int fresh,
                        int64_t *board, int32_t *round_count, uint8_t *inventory, int64_t *scores,
                        int32_t *mover, uint8_t *terminal, int32_t *ep_len, int64_t *stats, int nthreads) {
    blokus_args a = {seed, env0, t0, K, fresh, board, round_count, inventory, scores, mover, terminal, ep_len};
    parallel_for(B, nthreads, blokus_env, &a, sizeof(int32_t) * BLOKUS_CAP, stats);
}
