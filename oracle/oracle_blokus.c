/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's Blokus dynamics.
 *
 * Nothing in the product path may call this (see oracle/README.md).
 *
 * It deliberately follows the reference's *structure* (cell arrays, per-cell legality tests, the
 * piece -> anchor -> orientation -> shift loop nest) and shares no code, tables or bitboard tricks
 * with the CUDA engine, so that a bug cannot be common to both.
 *
 * Layout: board int64[20*20] row-major board[y*20+x] (0 empty, 1..4 colour) as Board.board_contents
 * (envs/blokus/board.py:82-85); inventory uint8[4][21] (1 = piece still held) in PIECE_TYPES order
 * (board.py:24-44 == ai.py:12-22); scores int64[4]; round_count.
 * Action id = ((piece*400 + y*20 + x)*8 + orientation)*5 + shift, -1 = '' (pass).
 */
#include <stdint.h>
#include <string.h>

#define NP 21
#define BS 20

/* PIECE_TYPES offsets (x, y) (board.py:24-44) */
static const int PIECE_SIZE[NP] = {1, 2, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5};
static const int PIECE_OFF[NP][5][2] = {
    {{0, 0}},
    {{0, 0}, {1, 0}},
    {{0, 0}, {1, 0}, {1, 1}},
    {{0, 0}, {1, 0}, {2, 0}},
    {{0, 0}, {1, 0}, {0, 1}, {1, 1}},
    {{0, 0}, {1, -1}, {1, 0}, {2, 0}},
    {{0, 0}, {1, 0}, {2, 0}, {3, 0}},
    {{0, 0}, {1, 0}, {2, 0}, {2, -1}},
    {{0, 0}, {1, 0}, {1, -1}, {2, -1}},
    {{0, 0}, {0, -1}, {1, 0}, {2, 0}, {3, 0}},
    {{0, 0}, {0, -1}, {0, 1}, {1, 0}, {2, 0}},
    {{0, 0}, {0, -1}, {0, -2}, {1, -2}, {2, -2}},
    {{0, 0}, {1, 0}, {1, -1}, {2, -1}, {3, -1}},
    {{0, 0}, {0, 1}, {1, 0}, {2, 0}, {2, -1}},
    {{0, 0}, {1, 0}, {2, 0}, {3, 0}, {4, 0}},
    {{0, 0}, {1, 0}, {2, 0}, {1, -1}, {2, -1}},
    {{0, 0}, {0, 1}, {1, 0}, {1, -1}, {2, -1}},
    {{0, 0}, {1, 0}, {0, 1}, {0, 2}, {1, 2}},
    {{0, 0}, {1, 0}, {1, -1}, {1, 1}, {2, -1}},
    {{0, 0}, {-1, 0}, {1, 0}, {0, -1}, {0, 1}},
    {{0, 0}, {1, 0}, {1, -1}, {2, 0}, {3, 0}},
};
/* GAME_PIECE_VALUES (ai.py:12-22) == piece size */

/* rotate_piece about the origin for an offset (dx, dy) (computation.py:53-86 with rotate_by_deg
 * :24-34 and flip_piece_x/y :37-50).  ORIENTATIONS order (board.py:47):
 * 0 north 1 northeast 2 east 3 southeast 4 south 5 southwest 6 west 7 northwest. */
static void rotate_offset(int o, int dx, int dy, int *rx, int *ry) {
    switch (o) {
    case 0: *rx = dy;  *ry = -dx; break;  /* north: 270 deg */
    case 1: *rx = dx;  *ry = -dy; break;  /* northeast: 0 deg then flip_y */
    case 2: *rx = dx;  *ry = dy;  break;  /* east: identity */
    case 3: *rx = dy;  *ry = dx;  break;  /* southeast: 90 deg then flip_x */
    case 4: *rx = -dy; *ry = dx;  break;  /* south: 90 deg */
    case 5: *rx = -dx; *ry = dy;  break;  /* southwest: 180 deg then flip_y */
    case 6: *rx = -dx; *ry = -dy; break;  /* west: 180 deg */
    default: *rx = -dy; *ry = -dx; break; /* northwest: 270 deg then flip_x */
    }
}

/* is_valid_adjacents (computation.py:89-119) */
static int valid_adjacents(const int64_t *b, int y, int x, int color) {
    int ok = 1;
    if (y != 0 && b[(y - 1) * BS + x] == color) ok = 0;
    if (x != 0 && b[y * BS + x - 1] == color) ok = 0;
    if (y != 19 && b[(y + 1) * BS + x] == color) ok = 0;
    if (x != 19 && b[y * BS + x + 1] == color) ok = 0;
    return ok;
}

/* is_valid_cell (computation.py:122-142) */
static int valid_cell(const int64_t *b, int x, int y, int color) {
    if (x < 0 || x >= 20 || y < 0 || y >= 20) return 0;
    return b[y * BS + x] == 0 && valid_adjacents(b, y, x, color);
}

/* check_valid_corner (board.py:127-154) */
static int valid_corner(const int64_t *b, int color, int row, int col) {
    if (!valid_adjacents(b, row, col, color)) return 0;
    if (row != 0 && col != 19 && b[(row - 1) * BS + col + 1] == color) return 1;
    if (row != 0 && col != 0 && b[(row - 1) * BS + col - 1] == color) return 1;
    if (row != 19 && col != 19 && b[(row + 1) * BS + col + 1] == color) return 1;
    if (row != 19 && col != 0 && b[(row + 1) * BS + col - 1] == color) return 1;
    return 0;
}

/* anchors: round 0 -> PLAYER_DEFAULT_CORNERS (board.py:50,177-179), else gather_empty_corner_indexes
 * (board.py:114-125) in row-major order.  out = (x, y) pairs; returns count. */
int orc_blokus_anchors(const int64_t *board, int round_count, int color, int32_t *out) {
    static const int CORNERS[4][2] = {{0, 0}, {19, 0}, {0, 19}, {19, 19}};
    if (round_count == 0) {
        out[0] = CORNERS[color - 1][0]; out[1] = CORNERS[color - 1][1];
        return 1;
    }
    int n = 0;
    for (int row = 0; row < BS; row++)
        for (int col = 0; col < BS; col++)
            if (board[row * BS + col] == 0 && valid_corner(board, color, row, col)) {
                out[2 * n] = col; out[2 * n + 1] = row; n++;
            }
    return n;
}

/* get_all_valid_moves (board.py:170-193) flattened in the order BlokusEnvironment.valid_actions
 * (:453-500) emits it.  Returns the full count; writes at most cap ids. */
int orc_blokus_valid_moves(const int64_t *board, int round_count, int color, const uint8_t *inventory,
                           int32_t *out_ids, int cap) {
    int32_t anchors[2 * BS * BS];
    int na = orc_blokus_anchors(board, round_count, color, anchors);
    int count = 0;
    for (int piece = 0; piece < NP; piece++) {
        if (!inventory[piece]) continue;
        int sz = PIECE_SIZE[piece];
        for (int a = 0; a < na; a++) {
            int ax = anchors[2 * a], ay = anchors[2 * a + 1];
            for (int o = 0; o < 8; o++) {
                /* get_all_shifted_offsets (computation.py:228-246): rotate the default piece, then
                 * re-origin on cell k; check_shifted (:145-180) then tests every cell. */
                int rx[5], ry[5];
                for (int i = 0; i < sz; i++) rotate_offset(o, PIECE_OFF[piece][i][0], PIECE_OFF[piece][i][1], &rx[i], &ry[i]);
                for (int k = 0; k < sz; k++) {
                    int ok = 1;
                    for (int i = 0; i < sz; i++)
                        if (!valid_cell(board, ax + rx[i] - rx[k], ay + ry[i] - ry[k], color)) ok = 0;
                    if (ok) {
                        if (count < cap) out_ids[count] = ((piece * 400 + ay * 20 + ax) * 8 + o) * 5 + k;
                        count++;
                    }
                }
            }
        }
    }
    return count;
}

static int has_any_move(const int64_t *board, int round_count, int color, const uint8_t *inventory) {
    int32_t tmp[1];
    return orc_blokus_valid_moves(board, round_count, color, inventory, tmp, 0) > 0;
}

/* next_state (BlokusEnvironment.py:357-451).  board / inventory / scores / round_count are updated in
 * place (callers copy first).  NOTE the terminal test (:424) uses the *old* board and old round with
 * the *new* inventories.  Returns next mover. */
int orc_blokus_next_state(int64_t *board, int *round_count, uint8_t *inventory /*[4][21]*/, int64_t *scores,
                          int mover, int action_id, int *reward, int *terminal, int *winners_mask) {
    int64_t old_board[BS * BS];
    memcpy(old_board, board, sizeof(old_board));
    int color = mover + 1;
    if (action_id >= 0) {
        int k = action_id % 5, o = (action_id / 5) % 8, cell = (action_id / 40) % 400, piece = action_id / 16000;
        int ax = cell % 20, ay = cell / 20, sz = PIECE_SIZE[piece];
        /* Board.update_board (board.py:87-98): shift_offsets on the *unrotated* piece, then rotate */
        int kx = PIECE_OFF[piece][k][0], ky = PIECE_OFF[piece][k][1];
        for (int i = 0; i < sz; i++) {
            int rx, ry;
            rotate_offset(o, PIECE_OFF[piece][i][0] - kx, PIECE_OFF[piece][i][1] - ky, &rx, &ry);
            int x = ax + rx, y = ay + ry;
            if (x >= 0 && x < 20 && y >= 0 && y < 20) board[y * BS + x] = color; /* oracle is only fed legal ids */
        }
        /* AI.update_player (ai.py:44-54) */
        uint8_t *inv = inventory + mover * NP;
        inv[piece] = 0;
        int left = 0;
        for (int q = 0; q < NP; q++) left += inv[q];
        if (left == 0 && piece == 0) scores[mover] += 20;
        else if (left == 0) scores[mover] += 15;
        scores[mover] += sz;
    }
    int any = 0;
    for (int p = 0; p < 4 && !any; p++) any = has_any_move(old_board, *round_count, p + 1, inventory + p * NP);
    *winners_mask = 0; *reward = 0; *terminal = 0;
    if (!any) {                                            /* :425-440 */
        *terminal = 1;
        int64_t max_score = 0;
        for (int p = 0; p < 4; p++) if (scores[p] > max_score) max_score = scores[p];
        for (int p = 0; p < 4; p++) if (scores[p] == max_score) *winners_mask |= 1 << p;
        /* index of the mover in a stable ascending sort by score */
        int pos = 0;
        for (int p = 0; p < 4; p++)
            if (scores[p] < scores[mover] || (scores[p] == scores[mover] && p < mover)) pos++;
        *reward = pos;
    }
    if (mover == 3) *round_count += 1;                     /* :446-447 */
    return (mover + 1) % 4;                                /* :449 */
}

/* state_to_observation (BlokusEnvironment.py:721-768): board of relative player ids (-1 empty) rotated by
 * np.rot90(k=-player); pieces uint8[4][21] by relative id; score rolled by -player. */
void orc_blokus_observation(const int64_t *board, const uint8_t *inventory, const int64_t *scores, int player,
                            int64_t *oboard, uint8_t *opieces, int64_t *oscore) {
    int64_t rel[BS * BS];
    for (int c = 0; c < BS * BS; c++) {
        int v = (int)board[c] - 1;                         /* COLOR_TO_PLAYER */
        rel[c] = v < 0 ? -1 : ((v - player) % 4 + 4) % 4;
    }
    /* np.rot90(m, k=-player): apply a clockwise quarter turn `player` times:
     * clockwise: out[i][j] = in[n-1-j][i] */
    int64_t tmp[BS * BS];
    memcpy(oboard, rel, sizeof(rel));
    for (int t = 0; t < player % 4; t++) {
        for (int i = 0; i < BS; i++)
            for (int j = 0; j < BS; j++) tmp[i * BS + j] = oboard[(BS - 1 - j) * BS + i];
        memcpy(oboard, tmp, sizeof(tmp));
    }
    memset(opieces, 0, 4 * NP);
    for (int p = 0; p < 4; p++) {
        int r = ((p - player) % 4 + 4) % 4;
        for (int q = 0; q < NP; q++) if (inventory[p * NP + q]) opieces[r * NP + q] = 1;
    }
    for (int i = 0; i < 4; i++) oscore[i] = scores[(i + player) % 4]; /* np.roll(x, -player)[i] = x[i+player] */
}
