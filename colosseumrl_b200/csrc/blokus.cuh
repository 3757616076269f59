// Blokus 4-player 20x20: legal-move generation and move application on row bitboards (sm_100a).
//
// Replaces (reference file:line):
//   Board.get_all_valid_moves        envs/blokus/board.py:170-193  (anchors :114-154, orientation/shift loop :156-168)
//   check_shifted & friends          envs/blokus/computation.py:145-180, 184-246, 53-142   (numba)
//   BlokusEnvironment.valid_actions  envs/blokus/BlokusEnvironment.py:453-500 (canonical order of the flattened dict)
//   BlokusEnvironment.next_state     :357-451  (Board.update_board board.py:87-98, AI.update_player ai.py:44-54,
//                                    the terminal test `not any(p.check_moves(board, round_count))` :424 which runs on
//                                    the OLD board / OLD round with the NEW inventories, winners/reward :425-440)
//   new_state :248-289, state_to_observation :721-768
//
// HBM layout: 352 bytes per game, AoS (a warp owns a game): 22 x uint4 = 88 words
//   w[20*c + y]  row y of colour c's plane, bit x set iff colour c occupies (x, y)           (c = 0..3, 80 words)
//   w[80 + c]    inventory of player c, bit p set iff piece p (PIECE_TYPES order) is still held
//   w[84]        scores, one byte per player        w[85] round | mover << 8 | terminal << 16
//   w[86]        steps taken in this episode        w[87] unused
// Lanes 0..21 of the warp move the game with one 128-bit access each (352 contiguous bytes).
//
// Legality in bitboard form: with
//   A   = empty & ~N4(own)                                   "allowed" cells (computation.py:122-142, 89-119)
//   ANC = {corner}            if round == 0 (board.py:177-179)
//         A & D4(own)         otherwise      (board.py:114-154)
//   FIT_s[q] = AND_i A[q + cell_i(s)]                       shape s fits with its bounding-box corner at q
// action (piece, anchor a, orientation o, shift k) is legal  <=>  a in ANC  and  FIT_s[a - cell_k(s)],
// s = shape(piece, o).  The list is emitted in the reference's order piece -> anchor (row-major) -> o -> k, with
// lanes = the (o, k) ids of one piece and __ballot_sync / __popc prefix sums for the compaction.
//
// FIT boards: the 91 oriented shapes are exactly the fixed polyominoes of 1..5 cells, so each is a smaller one plus
// a cell and FIT_s[q] = FIT_parent[q + off] & A[q + c]: all boards a player needs cost <= 90 AND steps (lanes = rows),
// generated as straight-line code from blokus_tables.h (every offset an immediate).  Shapes whose pieces are not held
// or whose parent fits nowhere are skipped.  The any-move test of the terminal check needs no anchor loop at all:
// piece p has a move  <=>  OR_s OR_k (ANC & shift(FIT_s, cell_k)) != 0, evaluated level by level with an early exit.
#pragma once
#include "crl_common.cuh"
#include "philox.cuh"
#include "blokus_tables.h"

#define BLK_WORDS 88
#define BLK_VEC 22
#define BLK_WARPS 4            // warps (= games) per CTA
#define BLK_ROWMASK 0xFFFFFu
#define BLK_MAX_ANCHORS 400

#define BLK_FROWS 24           // rows of a FIT board: y = -4..19 at index y + 4 (rows -4..-1 are zero)
#define BLK_FSLOTS (BLK_NSHAPE_LE4 + 8)
#define BLK_SLOT(s, local) ((s) < BLK_NSHAPE_LE4 ? (s) : BLK_NSHAPE_LE4 + (local))

// per-warp shared scratch
struct BlkSmem {
    uint32_t st[BLK_WORDS];          // the game state (old board during a step)
    uint32_t A[24];                  // allowed rows; rows 20..23 are zero (shapes are at most 5 rows tall)
    uint32_t anc[20];                // anchor rows
    uint32_t F[BLK_FSLOTS * BLK_FROWS];   // FIT boards: bit (x + 4) of word [slot * 24 + y + 4]; the padding makes
                                     // FIT_s[a - cell] a plain load + shift for every anchor a and shape cell.
                                     // Slots 0..27: the shapes of <= 4 cells; slots 28..35: the shapes of the ONE
                                     // pentomino piece being processed (5-cell shapes have no children)
    uint16_t ay[BLK_MAX_ANCHORS + 4];    // anchors, row-major: row y ...
    uint16_t ax[BLK_MAX_ANCHORS + 4];    // ... and column x; then four padding entries (column 24: outside every FIT board)
};

__device__ __forceinline__ void blk_load(BlkSmem &sm, const uint4 *__restrict__ st, long long g, int lane) {
    if (lane < BLK_VEC) {
        uint4 v = ld_stream(st + g * BLK_VEC + lane);
        sm.st[4 * lane + 0] = v.x; sm.st[4 * lane + 1] = v.y; sm.st[4 * lane + 2] = v.z; sm.st[4 * lane + 3] = v.w;
    }
    __syncwarp();
}
__device__ __forceinline__ void blk_store(const BlkSmem &sm, uint4 *__restrict__ st, long long g, int lane) {
    __syncwarp();
    if (lane < BLK_VEC)
        st_stream(st + g * BLK_VEC + lane,
                  make_uint4(sm.st[4 * lane + 0], sm.st[4 * lane + 1], sm.st[4 * lane + 2], sm.st[4 * lane + 3]));
}
// new_state (BlokusEnvironment.py:248-289): empty board, round 0, four full inventories, player 0 to move
__device__ __forceinline__ void blk_new_state(BlkSmem &sm, int lane) {
    __syncwarp();
    for (int i = lane; i < BLK_WORDS; i += 32) sm.st[i] = (i >= 80 && i < 84) ? ((1u << BLK_NPIECE) - 1u) : 0u;
    __syncwarp();
}

// A and ANC rows of player c (0-based) for the board in sm.st; fills sm.A, sm.anc, sm.ay / sm.ax; returns #anchors.
__device__ __forceinline__ int blk_allowed_and_anchors(BlkSmem &sm, int c, int round, int lane) {
    __syncwarp();
    uint32_t a = 0, an = 0;
    if (lane < 20) {
        const int y = lane;
        uint32_t own = sm.st[20 * c + y];
        uint32_t occ = sm.st[y] | sm.st[20 + y] | sm.st[40 + y] | sm.st[60 + y];
        uint32_t up = y > 0 ? sm.st[20 * c + y - 1] : 0u, dn = y < 19 ? sm.st[20 * c + y + 1] : 0u;
        uint32_t n4 = up | dn | (own << 1) | (own >> 1);                       // is_valid_adjacents
        a = ~occ & ~n4 & BLK_ROWMASK;                                          // is_valid_cell
        if (round == 0) {                                                      // PLAYER_DEFAULT_CORNERS (board.py:50)
            const int cx = (c & 1) ? 19 : 0, cy = (c & 2) ? 19 : 0;
            an = (y == cy) ? (1u << cx) : 0u;
        } else {
            uint32_t d4 = (up << 1) | (up >> 1) | (dn << 1) | (dn >> 1);       // check_valid_corner (board.py:127-154)
            an = a & d4;
        }
    }
    if (lane < 24) sm.A[lane] = a;
    if (lane < 20) sm.anc[lane] = an;
    // row-major anchor list: exclusive prefix of the per-row counts
    int cnt = __popc(an), pre = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, pre, d);
        if (lane >= d) pre += t;
    }
    int total = __shfl_sync(0xffffffffu, pre, 31);
    int pos = pre - cnt;
    while (an) {
        int x = __ffs((int)an) - 1;
        an &= an - 1;
        sm.ay[pos] = (uint16_t)lane;
        sm.ax[pos++] = (uint16_t)x;
    }
    if (lane < 4) { sm.ay[total + lane] = 0; sm.ax[total + lane] = 24; }   // padding: never fits
    __syncwarp();
    return total;
}

// ---- FIT boards through the polyomino tree -----------------------------------------------------------------
// ne = bit s set iff FIT_s (s < 28) has been built and is not empty; ne5 = the same for the 8 slots of the
// current pentomino piece (both warp-uniform)

// shape 0 (the monomino): FIT = A
__device__ __forceinline__ uint32_t blk_tree_root(BlkSmem &sm, int lane, int frow) {
    const uint32_t f = lane < 20 ? sm.A[lane] << 4 : 0u;
    if (lane < BLK_FROWS) sm.F[frow] = f;
    return __any_sync(0xffffffffu, f != 0u) ? 1u : 0u;
}

// one tree step; every argument except sm / ne / lane / frow is a literal
#define BLK_TREE_STEP(s, local, par, px, py, cx, cy, nevar)                                                   \
    if (ne >> (par) & 1u) {                                                                                   \
        uint32_t f_ = 0u;                                                                                     \
        if (lane < 20 - (py)) f_ = (sm.F[(par) * BLK_FROWS + 4 + (py) + lane] >> (px)) & ((sm.A[lane + (cy)] >> (cx)) << 4); \
        if (lane < BLK_FROWS) sm.F[BLK_SLOT(s, local) * BLK_FROWS + frow] = f_;                                 \
        if (__any_sync(0xffffffffu, f_ != 0u)) nevar |= 1u << ((s) < BLK_NSHAPE_LE4 ? (s) : (local));          \
    }

// the shapes with LEVEL (2..4) cells that some held piece needs
template <int LEVEL>
__device__ __forceinline__ void blk_tree_level(BlkSmem &sm, uint32_t &ne, uint32_t inv, int lane, int frow) {
#define BLK_X(level, s, piece, local, par, px, py, cx, cy, need) \
    if ((level) == LEVEL && (inv & (need))) BLK_TREE_STEP(s, local, par, px, py, cx, cy, ne)
    BLK_TREE_LIST(BLK_X)
#undef BLK_X
    __syncwarp();
}

// the (<= 8) shapes of pentomino piece PIECE into slots 28..35; returns their non-empty flags
template <int PIECE>
__device__ __forceinline__ uint32_t blk_tree_pentomino(BlkSmem &sm, uint32_t ne, int lane, int frow) {
    uint32_t ne5 = 0u;
#define BLK_X(level, s, piece, local, par, px, py, cx, cy, need) \
    if ((level) == 5 && (piece) == PIECE) BLK_TREE_STEP(s, local, par, px, py, cx, cy, ne5)
    BLK_TREE_LIST(BLK_X)
#undef BLK_X
    __syncwarp();
    return ne5;
}

__device__ __forceinline__ uint32_t blk_tree_pentomino_dyn(BlkSmem &sm, int p, uint32_t ne, int lane, int frow) {
    __syncwarp();                // the previous piece's readers are done with slots 28..35
    switch (p) {
    case 9: return blk_tree_pentomino<9>(sm, ne, lane, frow);
    case 10: return blk_tree_pentomino<10>(sm, ne, lane, frow);
    case 11: return blk_tree_pentomino<11>(sm, ne, lane, frow);
    case 12: return blk_tree_pentomino<12>(sm, ne, lane, frow);
    case 13: return blk_tree_pentomino<13>(sm, ne, lane, frow);
    case 14: return blk_tree_pentomino<14>(sm, ne, lane, frow);
    case 15: return blk_tree_pentomino<15>(sm, ne, lane, frow);
    case 16: return blk_tree_pentomino<16>(sm, ne, lane, frow);
    case 17: return blk_tree_pentomino<17>(sm, ne, lane, frow);
    case 18: return blk_tree_pentomino<18>(sm, ne, lane, frow);
    case 19: return blk_tree_pentomino<19>(sm, ne, lane, frow);
    default: return blk_tree_pentomino<20>(sm, ne, lane, frow);
    }
}

// any-move test: OR_k (ANC & shift(FIT_s, cell_k)) over the shapes selected by COND
#define BLK_ANY_CELL(slot, c) (sm.F[(slot) * BLK_FROWS + 4 + lane - ((c) >> 3)] >> (4 - ((c) & 7)))
#define BLK_ANY_STEP(s, local, n, c0, c1, c2, c3, c4, nevar)                                                  \
    if (nevar >> ((s) < BLK_NSHAPE_LE4 ? (s) : (local)) & 1u) {                                               \
        if (lane < 20) {                                                                                      \
            uint32_t d_ = BLK_ANY_CELL(BLK_SLOT(s, local), c0);                                               \
            if ((n) > 1) d_ |= BLK_ANY_CELL(BLK_SLOT(s, local), c1);                                          \
            if ((n) > 2) d_ |= BLK_ANY_CELL(BLK_SLOT(s, local), c2);                                          \
            if ((n) > 3) d_ |= BLK_ANY_CELL(BLK_SLOT(s, local), c3);                                          \
            if ((n) > 4) d_ |= BLK_ANY_CELL(BLK_SLOT(s, local), c4);                                          \
            acc |= d_ & anc;                                                                                  \
        }                                                                                                     \
    }

// held pieces with LEVEL (1..4) cells
template <int LEVEL>
__device__ __forceinline__ bool blk_any_level(BlkSmem &sm, uint32_t ne, uint32_t inv, uint32_t anc, int lane) {
    uint32_t acc = 0u;
#define BLK_Y(s, piece, local, n, c0, c1, c2, c3, c4) \
    if ((n) == LEVEL && (inv >> (piece) & 1u)) BLK_ANY_STEP(s, local, n, c0, c1, c2, c3, c4, ne)
    BLK_SHAPE_LIST(BLK_Y)
#undef BLK_Y
    return __any_sync(0xffffffffu, acc != 0u);
}

template <int PIECE>
__device__ __forceinline__ bool blk_any_pentomino(BlkSmem &sm, uint32_t ne, uint32_t anc, int lane, int frow) {
    const uint32_t ne5 = blk_tree_pentomino<PIECE>(sm, ne, lane, frow);
    uint32_t acc = 0u;
#define BLK_Y(s, piece, local, n, c0, c1, c2, c3, c4) \
    if ((n) == 5 && (piece) == PIECE) BLK_ANY_STEP(s, local, n, c0, c1, c2, c3, c4, ne5)
    BLK_SHAPE_LIST(BLK_Y)
#undef BLK_Y
    return __any_sync(0xffffffffu, acc != 0u);
}

// AI.check_moves (ai.py:36-42): does player c holding `inv` have any move on the board in sm.st?
__device__ __forceinline__ int blk_any_move(BlkSmem &sm, int c, int round, uint32_t inv, int lane) {
    if (inv == 0u || blk_allowed_and_anchors(sm, c, round, lane) == 0) return 0;
    const uint32_t anc = lane < 20 ? sm.anc[lane] : 0u;
    const int frow = lane < 20 ? lane + 4 : lane - 20;
    uint32_t ne = blk_tree_root(sm, lane, frow);
    __syncwarp();
    if (blk_any_level<1>(sm, ne, inv, anc, lane)) return 1;
    blk_tree_level<2>(sm, ne, inv, lane, frow);
    if (blk_any_level<2>(sm, ne, inv, anc, lane)) return 1;
    blk_tree_level<3>(sm, ne, inv, lane, frow);
    if (blk_any_level<3>(sm, ne, inv, anc, lane)) return 1;
    blk_tree_level<4>(sm, ne, inv, lane, frow);
    if (blk_any_level<4>(sm, ne, inv, anc, lane)) return 1;
#define BLK_P5(P)                                                                            \
    if (inv >> (P) & 1u) {                                                                   \
        __syncwarp();                                                                        \
        if (blk_any_pentomino<P>(sm, ne, anc, lane, frow)) return 1;                         \
    }
    BLK_P5(9) BLK_P5(10) BLK_P5(11) BLK_P5(12) BLK_P5(13) BLK_P5(14)
    BLK_P5(15) BLK_P5(16) BLK_P5(17) BLK_P5(18) BLK_P5(19) BLK_P5(20)
#undef BLK_P5
    return 0;
}

// Enumerate the legal moves of player c holding `inv` on the board in sm.st: action ids in the reference's order
// into out[0..cap), returns the full count.
__device__ __forceinline__ int blk_enumerate(BlkSmem &sm, int c, int round, uint32_t inv, int lane,
                                             int32_t *__restrict__ out, int cap) {
    const int na = blk_allowed_and_anchors(sm, c, round, lane);
    if (na == 0 || inv == 0u) return 0;
    const int frow = lane < 20 ? lane + 4 : lane - 20;
    uint32_t ne = blk_tree_root(sm, lane, frow);
    __syncwarp();
    blk_tree_level<2>(sm, ne, inv, lane, frow);
    blk_tree_level<3>(sm, ne, inv, lane, frow);
    blk_tree_level<4>(sm, ne, inv, lane, frow);
    int base = 0;
    const uint32_t lt = (1u << lane) - 1u;
    for (int p = 0; p < BLK_NPIECE; p++) {
        if (!(inv >> p & 1u)) continue;
        const int id0 = BLK_PIECE_ID0[p], nid = BLK_PIECE_ID0[p + 1] - id0, sh0 = BLK_PIECE_SHAPE0[p];
        const bool pent = nid > 32;                                      // 40 ids: the 12 pentominoes
        uint32_t nep = ne;                                               // non-empty flags of shape s: bit s - foff
        int soff = 0, foff = 0;
        if (pent) {
            nep = blk_tree_pentomino_dyn(sm, p, ne, lane, frow);
            if (nep == 0u) continue;                                     // none of the piece's shapes fits anywhere
            soff = BLK_NSHAPE_LE4 - sh0;                                 // shape s of this piece lives in slot s + soff
            foff = sh0;
        }
        // lane = (orientation, shift) id of the piece: its shape's FIT board, the cell that sits on the anchor.
        // A pentomino's ids 32..39 are tested for FOUR anchors per pass: lane = (anchor j = lane >> 3, id 32 + (lane & 7)).
        const uint32_t e0 = lane < nid ? BLK_ID_TAB_G[id0 + lane] : 0u;
        const uint32_t e1 = pent ? BLK_ID_TAB_G[id0 + 32 + (lane & 7)] : 0u;
        const int s0 = lane < nid ? (int)(e0 & 127u) : foff, s1 = pent ? (int)(e1 & 127u) : foff;
        const int sl0 = s0 + soff, sl1 = s1 + soff;
        const bool v0 = lane < nid && (nep >> (s0 - foff) & 1u), v1 = pent && (nep >> (s1 - foff) & 1u);
        if (__ballot_sync(0xffffffffu, v0 || v1) == 0u) continue;
        // FIT_s[anchor - (dx, dy)] lives at bit (x - dx + 4) of row y - dy + 4: with the row shifted right by the
        // (warp-uniform) anchor column the lane's test is one AND against its own constant mask 1 << (4 - dx)
        const uint32_t *f0 = sm.F + sl0 * BLK_FROWS + 4 - (int)((e0 >> 10) & 7u);
        const uint32_t *f1 = sm.F + sl1 * BLK_FROWS + 4 - (int)((e1 >> 10) & 7u);
        const uint32_t k0 = v0 ? 1u << (4 - (int)((e0 >> 7) & 7u)) : 0u, k1 = v1 ? 1u << (4 - (int)((e1 >> 7) & 7u)) : 0u;
        const int ok0 = (int)(e0 >> 13) + p * 16000, ok1 = (int)(e1 >> 13) + p * 16000;
        // ids 32..39 of a pentomino ("pass B") are tested for the next four anchors at once every fourth iteration
        const int j1 = lane >> 3;
        const uint32_t lt8 = (1u << (lane & 7)) - 1u;
        uint32_t mB = 0u;
        bool hit1 = false;
        for (int ai = 0; ai < na; ai++) {
            const int sub = ai & 3;
            if (pent && sub == 0) {                                      // (the list is padded with never-fitting anchors)
                hit1 = ((f1[sm.ay[ai + j1]] >> sm.ax[ai + j1]) & k1) != 0u;
                mB = __ballot_sync(0xffffffffu, hit1);
            }
            const int ay = sm.ay[ai], ax = sm.ax[ai];
            const bool hit = ((f0[ay] >> ax) & k0) != 0u;
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            const uint32_t mh = (mB >> (8 * sub)) & 255u;
            if ((m | mh) == 0u) continue;
            const int code = (ay * 20 + ax) * 40;
            int pos = base + __popc(m & lt);
            if (hit && pos < cap) out[pos] = code + ok0;                 // anchor ai: ids 0..31 ...
            base += __popc(m);
            pos = base + __popc(mh & lt8);
            if (hit1 && j1 == sub && pos < cap) out[pos] = code + ok1;   // ... then ids 32..39
            base += __popc(mh);
        }
    }
    return base;
}

// ---- valid_actions: one warp per game.  player < 0: the game's current mover.
__global__ void __launch_bounds__(32 * BLK_WARPS)
blokus_legal_kernel(const uint4 *__restrict__ st, int32_t *__restrict__ counts, int32_t *__restrict__ ids, int cap,
                    crl_u64 *stats, long long B, int player, int flags) {
    __shared__ BlkSmem smem[BLK_WARPS];
    __shared__ int sm_stat[CRL_NSTAT];
    BlockStats bs{sm_stat};
    if (stats) bs.init();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long g = (long long)blockIdx.x * BLK_WARPS + wid;
    if (g < B) {
        BlkSmem &sm = smem[wid];
        blk_load(sm, st, g, lane);
        if ((flags & CRL_FLAG_AUTO_RESET) && (sm.st[85] >> 16 & 1u)) blk_new_state(sm, lane);
        const uint32_t meta = sm.st[85];
        const int c = player >= 0 ? player : (int)(meta >> 8 & 3u);
        const int n = blk_enumerate(sm, c, (int)(meta & 0xffu), sm.st[80 + c], lane, ids + g * cap, cap);
        if (lane == 0) {
            counts[g] = n;
            if (stats) atomicAdd(&sm_stat[ST_NVALID], n);
        }
    }
    if (stats) bs.flush(stats);
}

// ---- next_state: one warp per game.
// result record, 8 bytes: int8 reward | u8 flags (1 terminal, 2 illegal action, 4 placed) | u8 winners mask |
//                         u8 ranking bits (bit p = rank of p: winners 0, others 1) | u8 next mover | 3 unused
__global__ void __launch_bounds__(32 * BLK_WARPS)
blokus_step_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ outst, const int32_t *__restrict__ actions,
                   uint2 *__restrict__ result, crl_u64 *stats, long long B, int flags) {
    __shared__ BlkSmem smem[BLK_WARPS];
    __shared__ int sm_stat[CRL_NSTAT];
    BlockStats bs{sm_stat};
    if (stats) bs.init();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long g = (long long)blockIdx.x * BLK_WARPS + wid;
    if (g < B) {
        BlkSmem &sm = smem[wid];
        blk_load(sm, in, g, lane);
        if ((flags & CRL_FLAG_AUTO_RESET) && (sm.st[85] >> 16 & 1u)) blk_new_state(sm, lane);
        const uint32_t meta = sm.st[85];
        const int round = (int)(meta & 0xffu), mover = (int)(meta >> 8 & 3u);
        const int aid = actions[g];
        uint32_t inv[4] = {sm.st[80], sm.st[81], sm.st[82], sm.st[83]};
        uint32_t scores = sm.st[84];
        uint32_t new_row = lane < 20 ? sm.st[20 * mover + lane] : 0u;
        int error = 0, placed = 0;
        if (aid >= 0) {
            // decode ((piece*400 + y*20 + x)*8 + o)*5 + k   (string form: BlokusEnvironment.py:55-106)
            const int piece = aid / 16000, rem = aid - piece * 16000, cell = rem / 40, ok = rem - cell * 40;
            const int o = ok / 5, k = ok - o * 5;
            bool legal = piece < BLK_NPIECE;
            const int pc = legal ? piece : 0, size = BLK_PIECE_SIZE[pc];
            uint32_t minv = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) minv |= (q == mover) ? inv[q] : 0u;
            legal = legal && k < size && (minv >> pc & 1u);
            const uint32_t e = BLK_ID_TAB[BLK_PIECE_ID0[pc] + (legal ? o * size + k : 0)];
            const uint32_t cells = BLK_SHAPE_CELLS[e & 127u];
            const int ay = cell / 20, ax = cell - ay * 20;
            const int qx = ax - (int)((e >> 7) & 7u), qy = ay - (int)((e >> 10) & 7u);
            blk_allowed_and_anchors(sm, mover, round, lane);      // validation against A / ANC of the mover
            legal = legal && qx >= 0 && qy >= 0 && (sm.anc[ay] >> ax & 1u);
            uint32_t add = 0;
#pragma unroll
            for (int i = 0; i < 5; i++) {
                const uint32_t cd = cells >> (6 * i);
                const int x = qx + (int)(cd & 7u), y = qy + (int)((cd >> 3) & 7u);
                // y <= 23 by construction (rows 20..23 of A are zero), x may exceed 19 -> bit not in A
                legal = legal && x >= 0 && y >= 0 && x < 20 && (sm.A[min(max(y, 0), 23)] >> max(x, 0) & 1u);
                add |= (y == lane && x >= 0 && x < 20) ? (1u << x) : 0u;
            }
            if (legal) {
                placed = 1;
                new_row |= add;                                                     // Board.update_board
                const uint32_t left = minv & ~(1u << pc);
                int gain = size + (left == 0 ? (pc == 0 ? 20 : 15) : 0);            // AI.update_player (ai.py:44-54)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (q == mover) {
                        inv[q] = left;
                        scores = (scores & ~(0xffu << (8 * q))) | ((((scores >> (8 * q)) & 0xffu) + gain) & 0xffu) << (8 * q);
                    }
                }
            } else {
                error = 1;                                                          // illegal id: flagged, applied as a pass
            }
        }
        // terminal test on the OLD board / OLD round with the NEW inventories (BlokusEnvironment.py:424)
        // `any` is order-independent; start with the players that have not moved yet this round (in round 0 the
        // ones that already moved have an occupied corner and would need a full, fruitless scan) and end with the mover.
        int any = 0;
#pragma unroll 1
        for (int j = 1; j <= 4 && !any; j++) {
            const int q = (mover + j) & 3;
            uint32_t iq = 0;
#pragma unroll
            for (int r = 0; r < 4; r++) iq |= (r == q) ? inv[r] : 0u;
            any = blk_any_move(sm, q, round, iq, lane);
        }
        const int terminal = !any;
        int reward = 0, winners = 0;
        int sc[4] = {(int)(scores & 0xff), (int)(scores >> 8 & 0xff), (int)(scores >> 16 & 0xff), (int)(scores >> 24)};
        if (terminal) {                                                             // :425-440
            int mx = max(max(sc[0], sc[1]), max(sc[2], sc[3]));                     // max_score starts at 0, scores >= 0
            int ms = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) ms |= (q == mover) ? sc[q] : 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                winners |= (sc[q] == mx) ? (1 << q) : 0;
                reward += (sc[q] < ms || (sc[q] == ms && q < mover)) ? 1 : 0;       // index in the stable ascending sort
            }
        }
        const int nround = round + (mover == 3 ? 1 : 0), nmover = (mover + 1) & 3;  // :446-449
        const uint32_t ep_len = sm.st[86] + 1u;
        // commit into the shared copy, then store with 22 x 128-bit lanes
        __syncwarp();
        if (lane < 20) sm.st[20 * mover + lane] = new_row;
        if (lane < 4) sm.st[80 + lane] = lane == 0 ? inv[0] : lane == 1 ? inv[1] : lane == 2 ? inv[2] : inv[3];
        if (lane == 4) sm.st[84] = scores;
        if (lane == 5) sm.st[85] = (uint32_t)(nround & 0xff) | (uint32_t)nmover << 8 | (uint32_t)terminal << 16;
        if (lane == 6) sm.st[86] = ep_len;
        blk_store(sm, outst, g, lane);
        if (lane == 0) {
            const uint32_t rank = 0xfu & ~(uint32_t)winners;
            result[g] = make_uint2(((uint32_t)reward & 0xffu) | (uint32_t)(terminal | error << 1 | placed << 2) << 8 |
                                       (uint32_t)winners << 16 | rank << 24,
                                   (uint32_t)nmover);
            if (stats) {
                atomicAdd(&sm_stat[ST_STEPS], 1);
                if (error) atomicAdd(&sm_stat[ST_ERRORS], 1);
                if (reward) atomicAdd(&sm_stat[ST_REWARD], (mover + 1) * reward);
                if (terminal) {
                    atomicAdd(&sm_stat[ST_EPISODES], 1);
                    atomicAdd(&sm_stat[ST_EPLEN], (int)ep_len);
                    if (!winners) atomicAdd(&sm_stat[ST_NOWIN], 1);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (winners >> q & 1) atomicAdd(&sm_stat[ST_WINS + q], 1);
                        else atomicAdd(&sm_stat[ST_RANK + q], 1);
                        atomicAdd(&sm_stat[ST_SCORE + q], sc[q]);
                    }
                }
            }
        }
    }
    if (stats) bs.flush(stats);
}

// uniform random policy over the generated list: the (r0 % n)-th entry, pass (-1) if n == 0
__global__ void blokus_policy_random_kernel(const int32_t *__restrict__ counts, const int32_t *__restrict__ ids, int cap,
                                            int32_t *__restrict__ actions, long long B, crl_u64 seed, crl_u64 first_env,
                                            uint32_t step) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B) return;
    int n = min(counts[g], cap);
    uint4 r = env_words(seed, first_env + (crl_u64)g, step, CRL_TAG_BLOKUS);
    actions[g] = n > 0 ? ids[g * cap + (long long)(r.x % (uint32_t)n)] : -1;
}

__global__ void blokus_reset_kernel(uint4 *__restrict__ st, const uint8_t *__restrict__ mask, long long B) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * BLK_VEC) return;
    long long g = idx / BLK_VEC;
    int v = (int)(idx - g * BLK_VEC);
    if (mask && !mask[g]) return;
    const uint32_t full = (1u << BLK_NPIECE) - 1u;
    st[idx] = (v == 20) ? make_uint4(full, full, full, full) : make_uint4(0, 0, 0, 0);
}

// state_to_observation (BlokusEnvironment.py:721-768): one warp per game.  The 352-byte state is loaded with 22
// coalesced 128-bit accesses into shared memory; lane l then produces output cells l, l + 32, ... so that every
// store instruction of the warp writes 32 consecutive bytes.
//  player >= 0: board int8[B][20][20] of relative player ids (-1 empty) rotated by np.rot90(k=-player);
//               pieces u8[B][4][21] rows by relative id; score int32[B][4] rolled by -player.
//  player == -1: absolute unpack: board = Board.board_contents (0 empty, 1..4 colour), pieces / score in seat order.
//  player == -2: like player >= 0 with every game seen by its own current mover (CRL_PLAYER_MOVER).
//  meta (optional) int32[B][4] = round, mover, terminal, episode steps.
__global__ void __launch_bounds__(32 * BLK_WARPS)
blokus_observe_kernel(const uint4 *__restrict__ st4, long long B, int player, int8_t *__restrict__ board,
                      uint8_t *__restrict__ pieces, int32_t *__restrict__ score, int32_t *__restrict__ meta) {
    __shared__ uint32_t sst[BLK_WARPS][BLK_WORDS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long g = (long long)blockIdx.x * BLK_WARPS + wid;
    if (g >= B) return;
    uint32_t *s = sst[wid];
    if (lane < BLK_VEC) {
        const uint4 v = ld_stream(st4 + g * BLK_VEC + lane);
        s[4 * lane + 0] = v.x; s[4 * lane + 1] = v.y; s[4 * lane + 2] = v.z; s[4 * lane + 3] = v.w;
    }
    __syncwarp();
    if (player == -2) player = (int)(s[85] >> 8 & 3u);
    for (int idx = lane; idx < 400; idx += 32) {
        const int i = idx / 20, j = idx - i * 20;
        int si = i, sj = j;                                 // source cell of np.rot90(k=-player)
        if (player == 1) { si = 19 - j; sj = i; }
        else if (player == 2) { si = 19 - i; sj = 19 - j; }
        else if (player == 3) { si = j; sj = 19 - i; }
        int v = -1;
#pragma unroll
        for (int c = 0; c < 4; c++) v = (s[20 * c + si] >> sj & 1u) ? c : v;
        if (player >= 0) v = v < 0 ? -1 : ((v - player) & 3);   // _relative_player_id (:46-50)
        else v = v + 1;
        board[g * 400 + idx] = (int8_t)v;
    }
    for (int idx = lane; idx < 84; idx += 32) {             // pieces[rel][piece]
        const int r = idx / 21, p = idx - r * 21;
        const int src = player >= 0 ? ((r + player) & 3) : r;
        pieces[g * 84 + idx] = (uint8_t)(s[80 + src] >> p & 1u);
    }
    if (lane < 4) {
        const int src = player >= 0 ? ((lane + player) & 3) : lane;
        score[g * 4 + lane] = (int)(s[84] >> (8 * src) & 0xffu);
        if (meta) {
            const uint32_t m = s[85];
            meta[g * 4 + lane] = lane == 0 ? (int)(m & 0xffu) : lane == 1 ? (int)(m >> 8 & 3u)
                                 : lane == 2 ? (int)(m >> 16 & 1u) : (int)s[86];
        }
    }
}

// import a reference-layout state: board int8[B][20][20] (0..4), inventory u8[B][4][21], scores int32[B][4],
// meta int32[B][4] = round, mover, terminal, episode steps
__global__ void blokus_pack_kernel(uint4 *__restrict__ st4, long long B, const int8_t *__restrict__ board,
                                   const uint8_t *__restrict__ pieces, const int32_t *__restrict__ score,
                                   const int32_t *__restrict__ meta) {
    uint32_t *st = (uint32_t *)st4;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * BLK_WORDS) return;
    long long g = idx / BLK_WORDS;
    int w = (int)(idx - g * BLK_WORDS);
    uint32_t v = 0;
    if (w < 80) {
        int c = w / 20, y = w - c * 20;
        for (int x = 0; x < 20; x++) v |= (board[g * 400 + y * 20 + x] == c + 1) ? (1u << x) : 0u;
    } else if (w < 84) {
        for (int p = 0; p < BLK_NPIECE; p++) v |= pieces[g * 84 + (w - 80) * 21 + p] ? (1u << p) : 0u;
    } else if (w == 84) {
        for (int q = 0; q < 4; q++) v |= ((uint32_t)score[g * 4 + q] & 0xffu) << (8 * q);
    } else if (w == 85) {
        v = ((uint32_t)meta[g * 4] & 0xffu) | ((uint32_t)meta[g * 4 + 1] & 3u) << 8 | ((uint32_t)meta[g * 4 + 2] & 1u) << 16;
    } else if (w == 86) {
        v = (uint32_t)meta[g * 4 + 3];
    }
    st[idx] = v;
}
