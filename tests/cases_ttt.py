"""Tic Tac Toe parity cases against the C ABI (both backends)."""
import ctypes as C
import os

import numpy as np

from oracle import oracle as orc
from colosseumrl_b200._lib import STAT_ROWS
from colosseumrl_b200 import philox

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unpack_result(res):
    res = np.asarray(res).view(np.uint8).reshape(-1, 4)
    return dict(reward=res[:, 0].view(np.int8).astype(np.int64), terminal=(res[:, 1] & 1).astype(bool),
                error=(res[:, 1] >> 1 & 1).astype(bool), placed=(res[:, 1] >> 2 & 1).astype(bool),
                winners=res[:, 2].astype(np.int64), ranking=res[:, 3].astype(np.int64))


def ttt_pack(be, n, board, winner, mover):
    B = board.shape[0]
    st = be.zeros((B, 4), np.int32)
    b = be.upload(np.ascontiguousarray(board.reshape(B, -1), np.int8))
    w, m = be.upload(np.ascontiguousarray(winner, np.int8)), be.upload(np.ascontiguousarray(mover, np.int8))
    be.check(be.lib.crl_ttt_pack(be.ptr(st), be.ptr(b), be.ptr(w), be.ptr(m), B, n, be.stream))
    return st


def ttt_unpack(be, st, n, player=-1):
    B = st.shape[0]
    cells = orc.ttt_cells(n)
    board, winner, mover = be.zeros((B, cells), np.int8), be.zeros((B,), np.int8), be.zeros((B,), np.int8)
    be.check(be.lib.crl_ttt_observe(be.ptr(st), player, be.ptr(board), be.ptr(winner), be.ptr(mover), B, n, be.stream))
    return be.download(board), be.download(winner), be.download(mover)


def case_line_tables(be):
    for n, cnt in ((2, 8), (3, 20), (4, 49)):
        out = np.zeros(64, np.uint32)
        k = be.lib.crl_ttt_lines(n, out.ctypes.data_as(C.POINTER(C.c_uint32)), 64)
        assert k == cnt and be.lib.crl_ttt_cells(n) == orc.ttt_cells(n)
        mine = sorted(int(x) for x in out[:k])
        ref = sorted(sum(1 << int(c) for c in line) for line in orc.ttt_lines(n))
        assert mine == ref and len(set(mine)) == cnt


def case_golden_steps(be, n):
    g = np.load(os.path.join(GOLDEN, "ttt_%dp.npz" % n))
    T = len(g["t"])
    cells = orc.ttt_cells(n)
    first = g["t"] == 0
    prev_board = np.where(first[:, None], np.int8(-1), np.concatenate([np.full((1, cells), -1, np.int8), g["board"][:-1]], 0))
    prev_winner = np.where(first, -1, np.concatenate([[-1], g["winner"][:-1]]))
    st = ttt_pack(be, n, prev_board, prev_winner, g["player"])
    # valid_actions of the state before the move
    vm = be.zeros((T,), np.uint32)
    be.check(be.lib.crl_ttt_valid_actions(be.ptr(st), be.ptr(vm), T, n, be.stream))
    vmask = be.download(vm).view(np.uint32)
    exp_mask = (g["valid_before"].astype(np.uint64) << np.arange(cells, dtype=np.uint64)[None]).sum(1)
    assert (vmask == exp_mask).all()
    act = be.upload(g["action"].astype(np.int8))
    res, out, va = be.zeros((T, 4), np.uint8), be.zeros((T, 4), np.int32), be.zeros((T,), np.uint32)
    be.check(be.lib.crl_ttt_step(be.ptr(st), be.ptr(out), be.ptr(act), be.ptr(res), be.ptr(va), None, T, n, 0, be.stream))
    r = unpack_result(be.download(res))
    board, winner, mover = ttt_unpack(be, out, n)
    assert (board == g["board"]).all() and (winner == g["winner"]).all() and (mover == g["next_player"]).all()
    assert (r["reward"] == g["reward"]).all() and (r["terminal"] == g["terminal"]).all()
    assert (r["winners"] == np.where(g["winners"] >= 0, 1 << np.maximum(g["winners"], 0), 0)).all()
    # invalid == a non-pass action on an occupied cell
    occupied = (g["action"] >= 0) & (prev_board[np.arange(T), np.maximum(g["action"], 0)] != -1)
    assert (r["error"] == occupied).all()
    assert (r["ranking"] == (((1 << n) - 1) & ~r["winners"])).all()
    assert (be.download(va).view(np.uint32) == ((g["board"] == -1).astype(np.uint64) << np.arange(cells, dtype=np.uint64)[None]).sum(1)).all()
    for p in range(n):
        ob, _, _ = ttt_unpack(be, out, n, player=p)
        assert (ob == g["obs"][:, p]).all()


def case_rollout_vs_oracle(be, n, B=300, K=40, seed=4, env0=77):
    ob = orc.TTTBatch(B, n)
    ob.rollout(seed, env0, 0, K, fresh=True)
    st, st2 = be.zeros((B, 4), np.int32), be.zeros((B, 4), np.int32)
    stats = be.zeros((STAT_ROWS, 32), np.int64)
    act, res = be.zeros((B,), np.int8), be.zeros((B, 4), np.uint8)
    be.check(be.lib.crl_ttt_reset(be.ptr(st), None, B, n, be.stream))
    cur, nxt = st, st2
    for t in range(K):
        be.check(be.lib.crl_ttt_policy_random(be.ptr(cur), be.ptr(act), seed, env0, t, B, n, 1, be.stream))
        be.check(be.lib.crl_ttt_step(be.ptr(cur), be.ptr(nxt), be.ptr(act), be.ptr(res), None, be.ptr(stats), B, n, 1, be.stream))
        cur, nxt = nxt, cur
    board, winner, mover = ttt_unpack(be, cur, n)
    assert (board.reshape(ob.board.shape) == ob.board).all() and (winner == ob.winner).all() and (mover == ob.mover).all()
    s = be.download(stats).sum(0)
    assert (s == ob.stats).all(), (s, ob.stats)
    term = unpack_result(be.download(res))["terminal"]
    assert (term == ob.terminal.astype(bool)).all()
    st3, stats3 = be.zeros((B, 4), np.int32), be.zeros((STAT_ROWS, 32), np.int64)
    be.check(be.lib.crl_ttt_reset(be.ptr(st3), None, B, n, be.stream))
    be.check(be.lib.crl_ttt_rollout(be.ptr(st3), None, be.ptr(stats3), seed, env0, 0, 7, B, n, be.stream))
    be.check(be.lib.crl_ttt_rollout(be.ptr(st3), be.ptr(res), be.ptr(stats3), seed, env0, 7, K - 7, B, n, be.stream))
    assert (be.download(st3) == be.download(cur)).all() and (be.download(stats3).sum(0) == ob.stats).all()


def case_masked_reset_and_errors(be):
    B, n = 50, 4
    st = be.upload(np.random.RandomState(0).randint(0, 2 ** 20, size=(B, 4)).astype(np.int32))
    before = be.download(st)
    mask = (np.arange(B) % 3 == 0).astype(np.uint8)
    m = be.upload(mask)
    be.check(be.lib.crl_ttt_reset(be.ptr(st), be.ptr(m), B, n, be.stream))
    after = be.download(st)
    assert (after[mask == 1] == 0).all() and (after[mask == 0] == before[mask == 0]).all()
    assert be.lib.crl_ttt_reset(be.ptr(st), None, B, 5, be.stream) == 3
    assert be.lib.crl_ttt_cells(7) == -1
    # out-of-range action index: flagged, no-op, turn passes
    be.check(be.lib.crl_ttt_reset(be.ptr(st), None, B, 2, be.stream))
    act = be.upload(np.full(B, 9, np.int8))
    res, out = be.zeros((B, 4), np.uint8), be.zeros((B, 4), np.int32)
    be.check(be.lib.crl_ttt_step(be.ptr(st), be.ptr(out), be.ptr(act), be.ptr(res), None, None, B, 2, 0, be.stream))
    r = unpack_result(be.download(res))
    board, winner, mover = ttt_unpack(be, out, 2)
    assert r["error"].all() and not r["placed"].any() and (board == -1).all() and (mover == 1).all()


def case_arbitrary_states(be, n):
    """tests/golden/ttt_states.npz (oracle/make_golden_ttt_states.py): arbitrary states stepped by the REFERENCE --
    carried winners that contradict the board, full boards, occupied-cell actions -- valid masks, next_state outputs
    and one observation each."""
    g = np.load(os.path.join(GOLDEN, "ttt_states.npz"))
    k = lambda name: g["p%d_%s" % (n, name)]
    T, cells = len(k("mover")), orc.ttt_cells(n)
    st = ttt_pack(be, n, k("board"), k("winner"), k("mover"))
    vm = be.zeros((T,), np.uint32)
    be.check(be.lib.crl_ttt_valid_actions(be.ptr(st), be.ptr(vm), T, n, be.stream))
    vmask = be.download(vm).view(np.uint32)
    assert (np.array([bin(int(m)).count("1") for m in vmask]) == k("n_valid")).all()
    assert (vmask == ((k("board") == -1).astype(np.uint64) << np.arange(cells, dtype=np.uint64)[None]).sum(1)).all()
    act = be.upload(k("action").astype(np.int8))
    res, out = be.zeros((T, 4), np.uint8), be.zeros((T, 4), np.int32)
    be.check(be.lib.crl_ttt_step(be.ptr(st), be.ptr(out), be.ptr(act), be.ptr(res), None, None, T, n, 0, be.stream))
    r = unpack_result(be.download(res))
    board, winner, mover = ttt_unpack(be, out, n)
    assert (board == k("o_board")).all() and (winner == k("o_winner")).all() and (mover == k("next_player")).all()
    assert (r["reward"] == k("reward")).all() and (r["terminal"] == k("terminal").astype(bool)).all()
    assert (r["winners"] == np.where(k("winners") >= 0, 1 << np.maximum(k("winners"), 0), 0)).all()
    for p in range(n):
        sel = np.flatnonzero(k("viewer") == p)
        sub = ttt_pack(be, n, k("o_board")[sel], k("o_winner")[sel], k("next_player")[sel])
        ob, _, _ = ttt_unpack(be, sub, n, player=p)
        assert (ob == k("obs")[sel]).all(), p
