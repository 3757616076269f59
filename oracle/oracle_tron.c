/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's Tron dynamics.
 *
 * Nothing in the product path (colosseumrl_b200/, the CUDA library) may call this.  It is
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
 *
 * Layout follows the reference exactly: board int64[N*N] (0 empty, p+1 owner), heads int64[P]
 * (= y*N + x), directions int64[P] (0 N, 1 E, 2 S, 3 W), deaths int64[P] (0 alive, else 1-based
 * killer id).  Parity is pinned by tests/golden/tron_*.npz, generated from the real reference by
 * oracle/make_golden.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ---- start positions: restates TronGridEnvironment.generate_start_positions
 *      (envs/tron/TronGridEnvironment.py:183-226) with the defaults used by new_state
 *      (ring_offset=1, spawn_offset=2 => randint(2,3) == 2, deterministic; :228,222-224). ---- */

/* python slice semantics on a list of length len with positive step */
static int py_slice(const int64_t *src, int len, int start, int stop, int step, int64_t *dst) {
    int n = 0;
    if (start > len) start = len;
    if (stop > len) stop = len;
    for (int i = start; i < stop; i += step) dst[n++] = src[i];
    return n;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* np.array_split section bounds: first (len % P) sections get len/P + 1 */
static void split_bounds(int len, int P, int sec, int *begin, int *size) {
    int q = len / P, r = len % P;
    if (sec < r) { *begin = sec * (q + 1); *size = q + 1; }
    else { *begin = r * (q + 1) + (sec - r) * q; *size = q; }
}

int orc_tron_start_positions(int N, int P, int ring_offset, int spawn_offset,
                             int64_t *heads, int64_t *directions) {
    int size = N / 2, offset = N % 2;
    double center = -0.5 * (offset - 1);
    int r1 = size - ring_offset - 1, r2 = size - ring_offset;
    int side = 2 * (r1 + 1);
    if (side <= 0 || P <= 0) return -1;
    int64_t *indices = (int64_t *)malloc(sizeof(int64_t) * N * N);
    int L = 0;
    /* np.ogrid[-size+center : size+offset+center] has exactly N points, step 1 (:194) */
    for (int iy = 0; iy < N; iy++) {
        double y = -size + center + iy;
        for (int ix = 0; ix < N; ix++) {
            double x = -size + center + ix;
            int m1 = (fabs(x) <= r1) && (fabs(y) <= r1);
            int m2 = (fabs(x) <= r2) && (fabs(y) <= r2);
            if (m1 ^ m2) indices[L++] = (int64_t)iy * N + ix; /* :197-200 */
        }
    }
    int64_t *top = (int64_t *)malloc(sizeof(int64_t) * 4 * (L + 4));
    int64_t *right = top + (L + 4), *bottom = right + (L + 4), *left = bottom + (L + 4);
    int nt = py_slice(indices, L, 0, side, 1, top);              /* :203 */
    int nr = py_slice(indices, L, side, 3 * side, 2, right);     /* :204 */
    int nb = py_slice(indices, L, 3 * side, L, 1, bottom);       /* :205 */
    int nl = py_slice(indices, L, side + 1, 3 * side + 1, 2, left); /* :206 */
    int M = nt + nr + nb + nl;
    int64_t *ring = (int64_t *)malloc(sizeof(int64_t) * (M + 1));
    int k = 0;
    for (int i = 0; i < nt; i++) ring[k++] = top[i];
    for (int i = 0; i < nr; i++) ring[k++] = right[i];
    for (int i = nb - 1; i >= 0; i--) ring[k++] = bottom[i];     /* bottom[::-1] :213 */
    for (int i = nl - 1; i >= 0; i--) ring[k++] = left[i];       /* left[::-1]   :213 */
    int D = side * 4; /* directions = (arange(4*side)//side + 2) % 4  :209-210 */
    int rc = 0;
    for (int p = 0; p < P; p++) {
        int b, s;
        split_bounds(M, P, p, &b, &s);
        if (s <= 0) { rc = -1; break; }
        heads[p] = ring[b + clampi(s / 2 + spawn_offset, 0, s - 1)];   /* get_centers :216-220 */
        split_bounds(D, P, p, &b, &s);
        if (s <= 0) { rc = -1; break; }
        int di = b + clampi(s / 2 + spawn_offset, 0, s - 1);
        directions[p] = ((di / side) + 2) % 4;
    }
    free(ring); free(top); free(indices);
    return rc;
}

/* new_state (TronGridEnvironment.py:228-263) */
int orc_tron_new_state(int N, int P, int64_t *board, int64_t *heads, int64_t *directions, int64_t *deaths) {
    memset(board, 0, sizeof(int64_t) * N * N);
    int rc = orc_tron_start_positions(N, P, 1, 2, heads, directions);
    if (rc) return rc;
    for (int p = 0; p < P; p++) { deaths[p] = 0; board[heads[p]] = p + 1; } /* :261 */
    return 0;
}

/* next_state_inplace (envs/tron/CyTronGrid.pyx:3-62).  Mutates all four arrays. */
void orc_tron_next_state_inplace(int N, int P, int64_t *board, int64_t *heads, int64_t *directions,
                                 int64_t *deaths, const int64_t *actions) {
    for (int i = 0; i < P; i++) {
        if (deaths[i] > 0) continue;                     /* :16 */
        int64_t x = heads[i] % N, y = heads[i] / N;      /* :21-22 */
        int64_t direction = (directions[i] + actions[i] + 4) % 4; /* :31 */
        if (direction == 0) y -= 1;                      /* :34-41 */
        else if (direction == 1) x += 1;
        else if (direction == 2) y += 1;
        else if (direction == 3) x -= 1;
        directions[i] = direction;                       /* :44 */
        if (x < 0 || x >= N || y < 0 || y >= N) {
            deaths[i] = i + 1;                           /* :47-48 */
        } else if (board[y * N + x] > 0) {
            int64_t enemy = board[y * N + x];            /* :51-53 */
            deaths[i] = enemy;
            if (heads[enemy - 1] == N * y + x) deaths[enemy - 1] = i + 1; /* :56-57 */
        } else {
            board[y * N + x] = i + 1;                    /* :60-62 */
            heads[i] = N * y + x;
        }
    }
}

/* next_state post-processing (TronGridEnvironment.py:309-323).
 * rewards[P]; returns terminal; *alive_mask bit p = p in new_players; *winners_mask = winners
 * (only meaningful when terminal; 0 => empty winners array). */
int orc_tron_next_state(int N, int P, int64_t *board, int64_t *heads, int64_t *directions,
                        int64_t *deaths, const int64_t *actions, int64_t *rewards,
                        int *alive_mask, int *winners_mask) {
    orc_tron_next_state_inplace(N, P, board, heads, directions, deaths, actions);
    int alive = 0, n_alive = 0;
    for (int p = 0; p < P; p++) {
        if (deaths[p] == 0) { alive |= 1 << p; n_alive++; }  /* :310 */
        rewards[p] = -2 * (deaths[p] > 0) + 1;                /* :313 */
    }
    int terminal = n_alive <= 1;                              /* :316 */
    *winners_mask = 0;
    if (terminal) {
        *winners_mask = alive;                                /* :319 */
        for (int p = 0; p < P; p++) if (alive >> p & 1) rewards[p] += 9; /* :320-321 */
    }
    *alive_mask = alive;
    return terminal;
}

/* compute_ranking (TronGridEnvironment.py:483-508), including the deaths[deaths-1] wrap-around for
 * alive players (numpy index -1 == last player) and the Counter semantics (missing key reads 0). */
void orc_tron_compute_ranking(int N, int P, const int64_t *board, const int64_t *deaths, int64_t *rank) {
    int64_t score[64];
    int present[64];
    int order[64], n_order = 0; /* Counter insertion order = first appearance in board.ravel() */
    for (int p = 0; p < P; p++) { score[p] = 0; present[p] = 0; }
    for (int c = 0; c < N * N; c++) {
        int v = (int)board[c] - 1;                        /* :488 */
        if (v >= 0 && v < P) {
            if (!present[v]) { present[v] = 1; order[n_order++] = v; }
            score[v]++;
        }
    }
    /* tie_locations = where(deaths[deaths-1] - arange(P) - 1 == 0) computed up front (:492) */
    int tie[64], n_tie = 0;
    for (int p = 0; p < P; p++) {
        int64_t k = deaths[p] - 1;
        if (k < 0) k += P;                                /* numpy negative index wraps */
        if (deaths[k] - p - 1 == 0) tie[n_tie++] = p;
    }
    for (int t = 0; t < n_tie; t++) {                     /* :493-495, in place, ascending */
        int p = tie[t];
        int64_t killer = deaths[p] - 1;                   /* -1 for alive players: Counter[-1] -> 0 */
        int64_t ks = (killer >= 0) ? (present[killer] ? score[killer] : 0) : 0;
        int64_t ps = present[p] ? score[p] : 0;
        int64_t m = ps < ks ? ps : ks;
        if (!present[p]) { present[p] = 1; order[n_order++] = p; }
        score[p] = m;
    }
    /* most_common(): stable sort by count descending over insertion order (:501) */
    for (int i = 1; i < n_order; i++) {
        int v = order[i], j = i - 1;
        while (j >= 0 && score[order[j]] < score[v]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = v;
    }
    for (int p = 0; p < P; p++) rank[p] = -1;             /* players absent from the Counter get no rank */
    int64_t prev = INT64_MAX; int cur = 0;
    for (int i = 0; i < n_order; i++) {                   /* :499-506 */
        int p = order[i];
        if (score[p] < prev) cur = i;
        rank[p] = cur;
        prev = score[p];
    }
}

/* relative_player_inplace + rolls: state_to_observation (TronGridEnvironment.py:363-405,
 * CyTronGrid.pyx:65-71), fully observable branch. */
void orc_tron_observation(int N, int P, int player, const int64_t *board, const int64_t *heads,
                          const int64_t *directions, const int64_t *deaths, int64_t *oboard,
                          int64_t *oheads, int64_t *odirections, int64_t *odeaths) {
    for (int c = 0; c < N * N; c++) {
        int64_t v = board[c];
        oboard[c] = v > 0 ? ((v - (player + 1) + P) % P) + 1 : v;
    }
    for (int i = 0; i < P; i++) {
        int r = (i + player) % P;                         /* :392 */
        oheads[i] = heads[r]; odeaths[i] = deaths[r]; odirections[i] = directions[r];
    }
}
