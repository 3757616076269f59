/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's generalised Tic Tac Toe.
 *
 * Nothing in the product path may call this (see oracle/README.md).
 *
 * Variants (n players): 2 -> board 3x3, 3 -> board 3x5, 4 -> board 3x3x3
 * (envs/tictactoe/tictactoe_2p_env.py:165, tictactoe_3p_env.py:166, tictactoe_4p_env.py:196).
 * State = (board int8[cells] in C order, -1 empty else player; winner int, -1 == None).
 * The win test restates scipy.signal.correlate(player_mask, pattern, 'valid') containing a 3
 * (2p :297-300, 4p :328-331) by sliding every WINNING_SHAPES pattern over the board.
 */
#include <stdint.h>
#include <string.h>

typedef struct { int d[3]; int cells[3][3]; } pattern_t; /* pattern dims + its three non-zero cells */

/* 2p/3p patterns (tictactoe_2p_env.py:12-17): column (3,1), row (1,3), identity, rot90(identity). Stored
 * as 3-d with a leading unit axis. */
static const pattern_t PAT2D[4] = {
    {{1, 3, 1}, {{0, 0, 0}, {0, 1, 0}, {0, 2, 0}}},
    {{1, 1, 3}, {{0, 0, 0}, {0, 0, 1}, {0, 0, 2}}},
    {{1, 3, 3}, {{0, 0, 0}, {0, 1, 1}, {0, 2, 2}}},
    {{1, 3, 3}, {{0, 0, 2}, {0, 1, 1}, {0, 2, 0}}},
};

/* 4p patterns (tictactoe_4p_env.py:19-38), in list order. */
static const pattern_t PAT3D[13] = {
    {{3, 1, 1}, {{0, 0, 0}, {1, 0, 0}, {2, 0, 0}}},           /* np.full((3,1,1)) */
    {{1, 3, 1}, {{0, 0, 0}, {0, 1, 0}, {0, 2, 0}}},           /* np.full((1,3,1)) */
    {{1, 1, 3}, {{0, 0, 0}, {0, 0, 1}, {0, 0, 2}}},           /* np.full((1,1,3)) */
    {{3, 3, 1}, {{0, 0, 0}, {1, 1, 0}, {2, 2, 0}}},           /* identity, axis=-1 */
    {{3, 3, 1}, {{0, 2, 0}, {1, 1, 0}, {2, 0, 0}}},           /* rot90(identity), axis=-1 */
    {{1, 3, 3}, {{0, 0, 0}, {0, 1, 1}, {0, 2, 2}}},           /* identity, axis=0 */
    {{1, 3, 3}, {{0, 0, 2}, {0, 1, 1}, {0, 2, 0}}},           /* rot90(identity), axis=0 */
    {{3, 1, 3}, {{0, 0, 0}, {1, 0, 1}, {2, 0, 2}}},           /* identity, axis=1 */
    {{3, 1, 3}, {{0, 0, 2}, {1, 0, 1}, {2, 0, 0}}},           /* rot90(identity), axis=1 */
    {{3, 3, 3}, {{0, 0, 0}, {1, 1, 1}, {2, 2, 2}}},           /* _diagonal3d() */
    {{3, 3, 3}, {{2, 0, 0}, {1, 1, 1}, {0, 2, 2}}},           /* rot90 axes=(0,1) */
    {{3, 3, 3}, {{0, 2, 0}, {1, 1, 1}, {2, 0, 2}}},           /* rot90 axes=(1,2) */
    {{3, 3, 3}, {{2, 2, 0}, {1, 1, 1}, {0, 0, 2}}},           /* rot90 (0,1) then (1,2) */
};

static void dims_of(int n, int d[3]) {
    if (n == 2) { d[0] = 1; d[1] = 3; d[2] = 3; }
    else if (n == 3) { d[0] = 1; d[1] = 3; d[2] = 5; }
    else { d[0] = 3; d[1] = 3; d[2] = 3; }
}

int orc_ttt_cells(int n) { int d[3]; dims_of(n, d); return d[0] * d[1] * d[2]; }

static int has_three(int n, const int8_t *board, int player) {
    int d[3]; dims_of(n, d);
    const pattern_t *pats = (n == 4) ? PAT3D : PAT2D;
    int np = (n == 4) ? 13 : 4;
    for (int q = 0; q < np; q++) {
        const pattern_t *pt = &pats[q];
        /* 'valid' correlation: every offset where the pattern fits inside the board */
        for (int a = 0; a + pt->d[0] <= d[0]; a++)
            for (int b = 0; b + pt->d[1] <= d[1]; b++)
                for (int c = 0; c + pt->d[2] <= d[2]; c++) {
                    int s = 0;
                    for (int t = 0; t < 3; t++) {
                        int ia = a + pt->cells[t][0], ib = b + pt->cells[t][1], ic = c + pt->cells[t][2];
                        s += board[(ia * d[1] + ib) * d[2] + ic] == player;
                    }
                    if (s == 3) return 1;
                }
    }
    return 0;
}

/* enumerate all distinct winning lines as sorted cell triples (for table cross-checks) */
int orc_ttt_lines(int n, int32_t *out /* [max 64][3] */) {
    int d[3]; dims_of(n, d);
    const pattern_t *pats = (n == 4) ? PAT3D : PAT2D;
    int np = (n == 4) ? 13 : 4, cnt = 0;
    for (int q = 0; q < np; q++) {
        const pattern_t *pt = &pats[q];
        for (int a = 0; a + pt->d[0] <= d[0]; a++)
            for (int b = 0; b + pt->d[1] <= d[1]; b++)
                for (int c = 0; c + pt->d[2] <= d[2]; c++) {
                    for (int t = 0; t < 3; t++)
                        out[cnt * 3 + t] = ((a + pt->cells[t][0]) * d[1] + b + pt->cells[t][1]) * d[2] + c + pt->cells[t][2];
                    cnt++;
                }
    }
    return cnt;
}

/* next_state (tictactoe_2p_env.py:240-315 / 3p :241-316 / 4p :271-346).
 * action = flat C-order cell index, or -1 for the empty string ''.  Mutates board/winner in place.
 * Outputs: reward, terminal, winner_out (-1 => winners is None); returns next player. */
int orc_ttt_next_state(int n, int8_t *board, int *winner, int player, int action,
                       int *reward, int *terminal, int *winner_out) {
    int cells = orc_ttt_cells(n);
    *reward = 0; *terminal = 0; *winner_out = -1;
    /* len(action) > 0 and is_valid_action (cell == -1) and winner is None  (2p :293) */
    if (action >= 0 && action < cells && board[action] == -1 && *winner < 0) {
        board[action] = (int8_t)player;
        if (has_three(n, board, player)) *winner = player;
    }
    if (*winner >= 0) {                                   /* :302-308 */
        *reward = (*winner == player) ? 1 : -1;
        *winner_out = *winner;
        *terminal = 1;
    }
    int any_empty = 0;                                    /* valid_actions(...) == [''] (:310-311) */
    for (int c = 0; c < cells; c++) any_empty |= board[c] == -1;
    if (!any_empty) *terminal = 1;
    return (player + 1) % n;                              /* :313 */
}

/* valid_actions (2p :317-348): empty cells in C order; returns count (0 => ['']) */
int orc_ttt_valid_actions(int n, const int8_t *board, int32_t *out) {
    int cells = orc_ttt_cells(n), k = 0;
    for (int c = 0; c < cells; c++) if (board[c] == -1) out[k++] = c;
    return k;
}

/* state_to_observation (2p :382-407, relabel :26-27). NOTE 4p uses "% 3" (tictactoe_4p_env.py:50). */
void orc_ttt_observation(int n, const int8_t *board, int player, int8_t *out) {
    int cells = orc_ttt_cells(n);
    int mod = (n == 2) ? 2 : 3;
    for (int c = 0; c < cells; c++) {
        int v = board[c];
        out[c] = (int8_t)(v < 0 ? v : (((v - player) % mod) + mod) % mod);
    }
}
