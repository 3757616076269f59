"""Generate csrc/blokus_tables.h: the 21 Blokus pieces' oriented shapes as bit-friendly tables.

Geometry (closed form of the reference): the cells of action (piece, anchor a,
orientation o, shift k) are  a + R_o(c_i) - R_o(c_k)  over the piece's default offsets c_i
(colosseumrl/envs/blokus/board.py:24-44), with R_o from computation.py:53-86.  For every (piece, o) the rotated
cell set is normalised to its bounding box's top-left corner; the 168 (piece, o) pairs collapse onto 91 distinct
normalised shapes.  An action is then  (shape s, origin q = a - cell_k(s))  and is legal iff the anchor is an
anchor and FIT_s[q] is set, where FIT_s = AND_i shift(allowed, cell_i(s)).

Run:  python colosseumrl_b200/csrc/gen_blokus_tables.py   (output is committed; deterministic)
"""
import os

PIECES = [
    ("monomino1", [(0, 0)]),
    ("domino1", [(0, 0), (1, 0)]),
    ("trominoe1", [(0, 0), (1, 0), (1, 1)]),
    ("trominoe2", [(0, 0), (1, 0), (2, 0)]),
    ("tetrominoes1", [(0, 0), (1, 0), (0, 1), (1, 1)]),
    ("tetrominoes2", [(0, 0), (1, -1), (1, 0), (2, 0)]),
    ("tetrominoes3", [(0, 0), (1, 0), (2, 0), (3, 0)]),
    ("tetrominoes4", [(0, 0), (1, 0), (2, 0), (2, -1)]),
    ("tetrominoes5", [(0, 0), (1, 0), (1, -1), (2, -1)]),
    ("pentominoe1", [(0, 0), (0, -1), (1, 0), (2, 0), (3, 0)]),
    ("pentominoe2", [(0, 0), (0, -1), (0, 1), (1, 0), (2, 0)]),
    ("pentominoe3", [(0, 0), (0, -1), (0, -2), (1, -2), (2, -2)]),
    ("pentominoe4", [(0, 0), (1, 0), (1, -1), (2, -1), (3, -1)]),
    ("pentominoe5", [(0, 0), (0, 1), (1, 0), (2, 0), (2, -1)]),
    ("pentominoe6", [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0)]),
    ("pentominoe7", [(0, 0), (1, 0), (2, 0), (1, -1), (2, -1)]),
    ("pentominoe8", [(0, 0), (0, 1), (1, 0), (1, -1), (2, -1)]),
    ("pentominoe9", [(0, 0), (1, 0), (0, 1), (0, 2), (1, 2)]),
    ("pentominoe10", [(0, 0), (1, 0), (1, -1), (1, 1), (2, -1)]),
    ("pentominoe11", [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)]),
    ("pentominoe12", [(0, 0), (1, 0), (1, -1), (2, 0), (3, 0)]),
]

# ORIENTATIONS order: north northeast east southeast south southwest west northwest (board.py:47)
ROT = [
    lambda x, y: (y, -x),    # north      (270 deg)
    lambda x, y: (x, -y),    # northeast  (0 deg, flip y)
    lambda x, y: (x, y),     # east       (identity)
    lambda x, y: (y, x),     # southeast  (90 deg, flip x)
    lambda x, y: (-y, x),    # south      (90 deg)
    lambda x, y: (-x, y),    # southwest  (180 deg, flip y)
    lambda x, y: (-x, -y),   # west       (180 deg)
    lambda x, y: (-y, -x),   # northwest  (270 deg, flip x)
]


ORIENTS = []             # (piece, orientation, shape, cells in shift order k)


def build():
    shapes = []          # list of (piece, tuple(sorted cells))
    shape_index = {}
    piece_shape0 = []
    piece_id0 = []
    ids = []             # (shape, dx, dy, o*5+k)
    ORIENTS.clear()
    for p, (_, offs) in enumerate(PIECES):
        piece_shape0.append(len(shapes))
        piece_id0.append(len(ids))
        for o in range(8):
            rot = [ROT[o](x, y) for x, y in offs]
            mx, my = min(c[0] for c in rot), min(c[1] for c in rot)
            norm = [(x - mx, y - my) for x, y in rot]
            key = (p, tuple(sorted(norm)))
            if key not in shape_index:
                shape_index[key] = len(shapes)
                shapes.append(key)
            s = shape_index[key]
            ORIENTS.append((p, o, s, list(norm)))
            for k, (dx, dy) in enumerate(norm):
                ids.append((s, dx, dy, o * 5 + k))
    piece_shape0.append(len(shapes))
    piece_id0.append(len(ids))
    return shapes, piece_shape0, piece_id0, ids


def connected(cells):
    cells = set(cells)
    seen, todo = set(), [next(iter(cells))]
    while todo:
        c = todo.pop()
        if c in seen:
            continue
        seen.add(c)
        for d in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            n = (c[0] + d[0], c[1] + d[1])
            if n in cells and n not in seen:
                todo.append(n)
    return len(seen) == len(cells)


def build_tree(shapes):
    """The 91 shapes are exactly the fixed polyominoes of 1..5 cells, so every shape s of n >= 2 cells is a shape
    of n - 1 cells ("parent") plus one cell c:   FIT_s[q] = FIT_parent[q + off] & A[q + c]   with off = the parent's
    bounding-box corner inside s.  All 91 FIT boards then cost 90 AND steps instead of 414 - 91.
    need[s] = pieces whose own shapes are s or descend from s (a shape is only built if one of them is held)."""
    index = {cells: i for i, (_, cells) in enumerate(shapes)}
    tree = [None] * len(shapes)
    for i, (_, cells) in enumerate(shapes):
        if len(cells) == 1:
            continue
        best = None
        for c in cells:
            rest = [x for x in cells if x != c]
            if not connected(rest):
                continue
            mx, my = min(x for x, _ in rest), min(y for _, y in rest)
            par = index[tuple(sorted((x - mx, y - my) for x, y in rest))]
            cand = (par, mx, my, c[0], c[1])
            if best is None or cand < best:
                best = cand
        assert best is not None and best[0] < i
        tree[i] = best
    need = [1 << p for p, _ in shapes]
    for i in range(len(shapes) - 1, 0, -1):
        need[tree[i][0]] |= need[i]
    return tree, need


def emit(path):
    shapes, piece_shape0, piece_id0, ids = build()
    assert len(shapes) == 91 and len(ids) == 712
    assert sum(len(s[1]) for s in shapes) == 414
    assert [len(c) for _, c in shapes] == sorted(len(c) for _, c in shapes)       # shapes are ordered by size
    tree, need = build_tree(shapes)
    L = []
    L.append("// GENERATED by gen_blokus_tables.py -- do not edit.  See that file for the derivation.")
    L.append("#pragma once")
    L.append("#define BLK_NPIECE 21")
    L.append("#define BLK_NSHAPE %d" % len(shapes))
    L.append("#define BLK_NID %d" % len(ids))
    L.append("// cells of each normalised shape: 5 x 6 bits (dx | dy << 3), short pieces repeat cell 0 (AND/OR idempotent)")
    L.append("__constant__ uint32_t BLK_SHAPE_CELLS[BLK_NSHAPE] = {")
    row = []
    for _, cells in shapes:
        cl = list(cells) + [cells[0]] * (5 - len(cells))
        v = 0
        for i, (dx, dy) in enumerate(cl):
            assert 0 <= dx <= 4 and 0 <= dy <= 4
            v |= (dx | dy << 3) << (6 * i)
        row.append("0x%08xu" % v)
    for i in range(0, len(row), 8):
        L.append("    " + ", ".join(row[i:i + 8]) + ",")
    L.append("};")
    L.append("// piece sizes (== GAME_PIECE_VALUES, ai.py:12-22)")
    L.append("__constant__ uint8_t BLK_PIECE_SIZE[BLK_NPIECE] = {" + ", ".join(str(len(o)) for _, o in PIECES) + "};")
    L.append("// shapes of piece p are [BLK_PIECE_SHAPE0[p], BLK_PIECE_SHAPE0[p+1])")
    L.append("__constant__ uint8_t BLK_PIECE_SHAPE0[BLK_NPIECE + 1] = {" + ", ".join(map(str, piece_shape0)) + "};")
    L.append("// (orientation, shift) ids of piece p are [BLK_PIECE_ID0[p], BLK_PIECE_ID0[p+1]), ordered o*size + k")
    L.append("__constant__ uint16_t BLK_PIECE_ID0[BLK_NPIECE + 1] = {" + ", ".join(map(str, piece_id0)) + "};")
    L.append("// per id: shape (7 bits) | dx << 7 | dy << 10 | (o*5 + k) << 13 ; (dx, dy) = the shape cell that sits on the anchor")
    L.append("__constant__ uint32_t BLK_ID_TAB[BLK_NID] = {")
    row = ["0x%05xu" % (s | dx << 7 | dy << 10 | ok << 13) for s, dx, dy, ok in ids]
    for i in range(0, len(row), 8):
        L.append("    " + ", ".join(row[i:i + 8]) + ",")
    L.append("};")
    L.append("// the same table in global memory for lane-indexed reads (a divergent __constant__ read is serialised)")
    L.append("__device__ const uint32_t BLK_ID_TAB_G[BLK_NID] = {")
    for i in range(0, len(row), 8):
        L.append("    " + ", ".join(row[i:i + 8]) + ",")
    L.append("};")
    L.append("// Polyomino tree (see build_tree in the generator): X(level, s, piece, local, parent, px, py, cx, cy, need) for")
    L.append("// s = 1..90 in shape order (= by size); FIT_s[q] = FIT_parent[q + (px, py)] & A[q + (cx, cy)]; local = s - first")
    L.append("// shape of its piece; `need` = pieces that need s (own piece + descendants).  Expanded into straight-line code,")
    L.append("// so every field is an immediate.  Shapes of 5 cells are leaves: need == 1 << piece.")
    L.append("#define BLK_NSHAPE_LE4 %d" % sum(1 for _, c in shapes if len(c) <= 4))
    L.append("#define BLK_TREE_LIST(X) \\")
    for i in range(1, len(shapes)):
        par, px, py, cx, cy = tree[i]
        pc = shapes[i][0]
        if len(shapes[i][1]) == 5:
            assert need[i] == 1 << pc and len(shapes[par][1]) == 4
        L.append("    X(%d, %d, %d, %d, %d, %d, %d, %d, %d, 0x%06xu) \\" % (len(shapes[i][1]), i, pc, i - piece_shape0[pc], par, px, py,
                                                                       cx, cy, need[i]))
    L.append("")
    L.append("// Shapes with their cells: Y(s, piece, local, ncells, c0..c4) with c = dx | dy << 3 (unused cells repeat cell 0)")
    L.append("#define BLK_SHAPE_LIST(Y) \\")
    for i, (pc, cells) in enumerate(shapes):
        cl = list(cells) + [cells[0]] * (5 - len(cells))
        L.append("    Y(%d, %d, %d, %d, %s) \\" % (i, pc, i - piece_shape0[pc], len(cells), ", ".join(str(dx | dy << 3) for dx, dy in cl)))
    L.append("")
    L.append("// Per piece: its 8 orientations with the cells in shift order: Z(o, s, local, n, c0..c4), c = dx | dy << 3 = the")
    L.append("// shape cell that sits on the anchor for shift k (id = o * n + k).  Expanded into straight-line tests.")
    for pc in range(len(PIECES)):
        L.append("#define BLK_ORIENTS_P%d(Z) \\" % pc)
        for (p2, o, s_, cells) in ORIENTS:
            if p2 != pc:
                continue
            cl = list(cells) + [cells[0]] * (5 - len(cells))
            L.append("    Z(%d, %d, %d, %d, %s) \\" % (o, s_, s_ - piece_shape0[pc], len(cells), ", ".join(str(dx | dy << 3) for dx, dy in cl)))
        L.append("")
    open(path, "w").write("\n".join(L) + "\n")


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "blokus_tables.h")
    emit(out)
    print("wrote", out)
