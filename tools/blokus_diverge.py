#!/usr/bin/env python
"""Dev tool: first divergence between the GPU path and the oracle in a Blokus rollout (step, game) + diagnosis."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import backends
import cases_blokus as cb
from oracle import oracle as orc

B, K, seed, env0, cap = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 70, 2, 0, 2048
be = backends.Cuda()
ob = orc.BlokusBatch(B)
st, st2 = be.zeros((B, 22, 4), np.int32), be.zeros((B, 22, 4), np.int32)
counts, ids = be.zeros((B,), np.int32), be.zeros((B, cap), np.int32)
act, res = be.zeros((B,), np.int32), be.zeros((B, 8), np.uint8)
be.check(be.lib.crl_blokus_reset(be.ptr(st), None, B, be.stream))
cur, nxt = st, st2
for t in range(K):
    prev_board, prev_inv, prev_round, prev_mover, prev_scores, prev_term = (ob.board.copy(), ob.inventory.copy(), ob.round_count.copy(),
                                                                ob.mover.copy(), ob.scores.copy(), ob.terminal.copy())
    ob.rollout(seed, env0, t, 1, fresh=(t == 0))
    be.check(be.lib.crl_blokus_legal(be.ptr(cur), -1, be.ptr(counts), be.ptr(ids), cap, None, B, 1, be.stream))
    be.check(be.lib.crl_blokus_policy_random(be.ptr(counts), be.ptr(ids), cap, be.ptr(act), seed, env0, t, B, be.stream))
    be.check(be.lib.crl_blokus_step(be.ptr(cur), be.ptr(nxt), be.ptr(act), be.ptr(res), None, B, 1, be.stream))
    cur, nxt = nxt, cur
    board, pieces, score, meta = cb.blk_unpack(be, cur)
    bad = np.where((board != ob.board).any(axis=(1, 2)) | (pieces != ob.inventory).any(axis=(1, 2)))[0]
    c = be.download(counts)
    print("step", t, "max count", int(c.max()), "mismatching games", len(bad))
    if len(bad):
        g = int(bad[0])
        print("game", g, "mover", prev_mover[g], "round", prev_round[g], "terminal before", prev_term[g], "count", c[g], "action", be.download(act)[g])
        if prev_term[g]:
            pst = orc.blokus_new_state()
            exp = orc.blokus_valid_moves(pst, 0, cap=65536)
        else:
            pst = (prev_board[g], int(prev_round[g]), prev_inv[g], prev_scores[g])
            exp = orc.blokus_valid_moves(pst, int(prev_mover[g]), cap=65536)
        got = be.download(ids)[g][:min(c[g], cap)]
        print("oracle list len", len(exp), "gpu", c[g], "equal prefix", (got[:len(exp)] == exp[:len(got)]).all() if len(got) and len(exp) else None)
        if len(exp) == c[g] and not (got == exp).all():
            d = np.where(got != exp)[0]
            print("first diff at", d[:5], got[d[:5]], exp[d[:5]])
        print("anchors", len(orc.blokus_anchors(pst[0], pst[1], (0 if prev_term[g] else int(prev_mover[g])) + 1)))
        break
