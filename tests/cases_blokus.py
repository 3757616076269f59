"""Blokus parity cases against the C ABI (both backends)."""
import os

import numpy as np

from oracle import oracle as orc
from colosseumrl_b200._lib import STAT_ROWS
from colosseumrl_b200 import philox

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unpack_result(res):
    res = np.asarray(res).view(np.uint8).reshape(-1, 8)
    return dict(reward=res[:, 0].view(np.int8).astype(np.int64), terminal=(res[:, 1] & 1).astype(bool),
                error=(res[:, 1] >> 1 & 1).astype(bool), placed=(res[:, 1] >> 2 & 1).astype(bool),
                winners=res[:, 2].astype(np.int64), ranking=res[:, 3].astype(np.int64), next_mover=res[:, 4].astype(np.int64),
                players_mask=res[:, 5].astype(np.int64), terminal_byte=res[:, 6].astype(np.int64))


def blk_pack(be, board, inv, scores, rounds, movers, terminal=None, ep_len=None):
    B = board.shape[0]
    meta = np.zeros((B, 4), np.int32)
    meta[:, 0], meta[:, 1] = rounds, movers
    if terminal is not None:
        meta[:, 2] = terminal
    if ep_len is not None:
        meta[:, 3] = ep_len
    st = be.zeros((B, 22, 4), np.int32)
    args = [be.upload(np.ascontiguousarray(board, np.int8)), be.upload(np.ascontiguousarray(inv, np.uint8)),
            be.upload(np.ascontiguousarray(scores, np.int32)), be.upload(meta)]
    be.check(be.lib.crl_blokus_pack(be.ptr(st), *[be.ptr(a) for a in args], B, be.stream))
    return st


def blk_unpack(be, st, player=-1):
    B = st.shape[0]
    board, pieces = be.zeros((B, 20, 20), np.int8), be.zeros((B, 4, 21), np.uint8)
    score, meta = be.zeros((B, 4), np.int32), be.zeros((B, 4), np.int32)
    be.check(be.lib.crl_blokus_observe(be.ptr(st), player, be.ptr(board), be.ptr(pieces), be.ptr(score), be.ptr(meta), B, be.stream))
    return tuple(be.download(x) for x in (board, pieces, score, meta))


def blk_legal(be, st, player=-1, cap=2048, flags=0, stats=None):
    B = st.shape[0]
    counts, ids = be.zeros((B,), np.int32), be.zeros((B, cap), np.int32)
    be.check(be.lib.crl_blokus_legal(be.ptr(st), player, be.ptr(counts), be.ptr(ids), cap, be.ptr(stats), B, flags, be.stream))
    return be.download(counts), be.download(ids)


def blk_step(be, st, actions, flags=0, stats=None):
    B = st.shape[0]
    a = be.upload(np.ascontiguousarray(actions, np.int32))
    out, res = be.zeros((B, 22, 4), np.int32), be.zeros((B, 8), np.uint8)
    be.check(be.lib.crl_blokus_step(be.ptr(st), be.ptr(out), be.ptr(a), be.ptr(res), be.ptr(stats), B, flags, be.stream))
    return out, unpack_result(be.download(res))


def _golden_prev(g):
    """State before each recorded step."""
    T = len(g["t"])
    first = g["t"] == 0
    board = np.where(first[:, None, None], np.int8(0), np.concatenate([np.zeros((1, 20, 20), np.int8), g["board"][:-1]], 0))
    inv = np.where(first[:, None, None], np.uint8(1), np.concatenate([np.ones((1, 4, 21), np.uint8), g["inventory"][:-1]], 0))
    scores = np.where(first[:, None], 0, np.concatenate([np.zeros((1, 4), np.int64), g["scores"][:-1]], 0))
    rounds = np.where(first, 0, np.concatenate([[0], g["round"][:-1]]))
    return board, inv, scores, rounds


def case_golden_games(be):
    """Every recorded transition of the reference's random games: valid_actions list (ordered), next_state outputs,
    observations."""
    g = np.load(os.path.join(GOLDEN, "blokus_games.npz"))
    T = len(g["t"])
    board, inv, scores, rounds = _golden_prev(g)
    st = blk_pack(be, board, inv, scores, rounds, g["mover"])
    counts, ids = blk_legal(be, st, cap=2048)
    assert (counts == g["n_valid"]).all()
    for i in range(T):
        exp = g["valid_flat"][g["valid_off"][i]:g["valid_off"][i + 1]]
        assert (ids[i, :len(exp)] == exp).all(), i
    # explicit player argument == mover
    for p in range(4):
        sel = np.flatnonzero(g["mover"] == p)[:40]
        c2, i2 = blk_legal(be, blk_pack(be, board[sel], inv[sel], scores[sel], rounds[sel], (g["mover"][sel] + 1) % 4), player=p)
        assert (c2 == counts[sel]).all() and (i2 == ids[sel]).all()
    out, r = blk_step(be, st, g["action"])
    b2, p2, s2, m2 = blk_unpack(be, out)
    assert (b2 == g["board"]).all() and (p2 == g["inventory"]).all() and (s2 == g["scores"]).all()
    assert (m2[:, 0] == g["round"]).all() and (m2[:, 1] == g["next_mover"]).all() and (m2[:, 2] == g["terminal"]).all()
    assert (r["reward"] == g["reward"]).all() and (r["terminal"] == g["terminal"]).all()
    assert (r["winners"] == g["winners"]).all() and (r["next_mover"] == g["next_mover"]).all()
    assert not r["error"].any() and (r["placed"] == (g["action"] >= 0)).all()
    assert (r["ranking"][r["terminal"]] == (0xf & ~r["winners"][r["terminal"]])).all()
    assert (r["players_mask"] == 1 << g["next_mover"]).all() and (r["terminal_byte"] == g["terminal"]).all()
    # observations
    idx = g["obs_idx"]
    sub = blk_pack(be, g["board"][idx], g["inventory"][idx], g["scores"][idx], g["round"][idx], g["next_mover"][idx])
    for p in range(4):
        m = g["obs_player"] == p
        sp = blk_pack(be, g["board"][idx[m]], g["inventory"][idx[m]], g["scores"][idx[m]], g["round"][idx[m]], g["next_mover"][idx[m]])
        ob, op, osc, _ = blk_unpack(be, sp, player=p)
        assert (ob == g["obs_board"][m]).all(), p
        assert (op == g["obs_pieces"][m]).all() and (osc == g["obs_score"][m]).all()


def _list_hashes(counts, ids):
    """oracle/make_golden_wide.py::list_hash of every row's first counts[i] ids."""
    out = np.zeros(len(counts), np.uint64)
    with np.errstate(over="ignore"):
        pw = np.cumprod(np.full(ids.shape[1], np.uint64(0x9E3779B97F4A7C15), np.uint64))
        for i, n in enumerate(counts):
            out[i] = ((ids[i, :n].astype(np.int64) + 1).astype(np.uint64) * pw[:n]).sum(dtype=np.uint64)
    return out


def _check_wide(be, g, pre, board, inv, scores, rounds):
    st = blk_pack(be, board, inv, scores, rounds, g[pre + "mover"])
    counts, ids = blk_legal(be, st, cap=2048)
    assert (counts == g[pre + "n_valid"]).all()
    assert (_list_hashes(counts, ids) == g[pre + "valid_hash"]).all()
    out, r = blk_step(be, st, g[pre + "action"])
    b2, p2, s2, m2 = blk_unpack(be, out)
    assert (b2 == g[pre + "board"]).all() and (p2 == g[pre + "inventory"]).all() and (s2 == g[pre + "scores"]).all()
    assert (m2[:, 0] == g[pre + "round"]).all() and (m2[:, 1] == g[pre + "next_mover"]).all()
    assert (m2[:, 2] == g[pre + "terminal"]).all()
    assert (r["reward"] == g[pre + "reward"]).all() and (r["terminal"] == g[pre + "terminal"]).all()
    assert (r["winners"] == g[pre + "winners"]).all() and (r["next_mover"] == g[pre + "next_mover"]).all()
    assert not r["error"].any() and (r["placed"] == (g[pre + "action"] >= 0)).all()
    assert (r["ranking"][r["terminal"]] == (0xf & ~r["winners"][r["terminal"]])).all()
    # is_valid_action of the recorded (legal) action
    placed = g[pre + "action"] >= 0
    assert blk_is_valid(be, st, g[pre + "action"])[placed].all()


def case_wide_golden(be, stride=1):
    """tests/golden/blokus_wide.npz (oracle/make_golden_wide.py): 64 more reference games (valid lists through length
    + order-sensitive hash) and the hand-built end-game positions (last-piece bonuses, tied winners)."""
    g = np.load(os.path.join(GOLDEN, "blokus_wide.npz"))
    gg = {k: g[k][::stride] if k.startswith("g_") else g[k] for k in g.files}
    gfull = {"t": g["g_t"], "board": g["g_board"], "inventory": g["g_inventory"], "scores": g["g_scores"], "round": g["g_round"]}
    board, inv, scores, rounds = (x[::stride] for x in _golden_prev(gfull))
    _check_wide(be, gg, "g_", board, inv, scores, rounds)
    e = {k: g[k][::stride] if k.startswith("e_") else g[k] for k in g.files}
    _check_wide(be, e, "e_", e["e_i_board"], e["e_i_inventory"], e["e_i_scores"], e["e_i_round"])


def case_arbitrary_positions(be):
    """tests/golden/blokus_boards.npz (oracle/make_golden_blokus_boards.py): arbitrary hand-built positions stepped by the
    REFERENCE -- every seat's ordered valid list (length + order-sensitive hash) and one next_state of the mover."""
    g = np.load(os.path.join(GOLDEN, "blokus_boards.npz"))
    n = len(g["mover"])
    for q in range(4):
        st = blk_pack(be, g["i_board"], g["i_inventory"], g["i_scores"], g["i_round"], np.full(n, (q + 1) % 4))
        counts, ids = blk_legal(be, st, player=q, cap=4096)
        assert (counts == g["n_valid"][:, q]).all(), q
        assert (_list_hashes(counts, ids) == g["valid_hash"][:, q]).all(), q
    st = blk_pack(be, g["i_board"], g["i_inventory"], g["i_scores"], g["i_round"], g["mover"])
    out, r = blk_step(be, st, g["action"])
    b2, p2, s2, m2 = blk_unpack(be, out)
    assert (b2 == g["board"]).all() and (p2 == g["inventory"]).all() and (s2 == g["scores"]).all()
    assert (m2[:, 0] == g["round"]).all() and (m2[:, 1] == g["next_mover"]).all() and (m2[:, 2] == g["terminal"]).all()
    assert (r["reward"] == g["reward"]).all() and (r["terminal"] == g["terminal"]).all() and (r["winners"] == g["winners"]).all()
    assert not r["error"].any()


def case_illegal_actions(be):
    """Ids outside the mover's valid list are flagged and applied as a pass (engine contract)."""
    g = np.load(os.path.join(GOLDEN, "blokus_games.npz"))
    board, inv, scores, rounds = _golden_prev(g)
    sel = np.arange(0, len(g["t"]), 3)
    st = blk_pack(be, board[sel], inv[sel], scores[sel], rounds[sel], g["mover"][sel])
    counts, ids = blk_legal(be, st)
    rng = np.random.RandomState(1)
    acts = np.zeros(len(sel), np.int32)
    legal = np.zeros(len(sel), bool)
    for i in range(len(sel)):
        valid = set(ids[i, :counts[i]].tolist())
        kind = i % 4
        if kind == 0:
            a = int(rng.randint(0, 21 * 16000))                      # arbitrary id
        elif kind == 1 and counts[i]:
            a = int(ids[i, rng.randint(counts[i])]) ^ 1              # neighbour of a legal id
        elif kind == 2:
            a = int(rng.randint(21 * 16000, 2 ** 30))                # piece index out of range
        else:
            a = int(ids[i, rng.randint(counts[i])]) if counts[i] else -1
        acts[i] = a
        legal[i] = a in valid or a < 0
    out, r = blk_step(be, st, acts)
    assert (r["error"] == ~legal).all()
    # an illegal action must behave exactly like a pass
    passes = np.where(legal, acts, -1)
    out2, r2 = blk_step(be, st, passes)
    assert (be.download(out) == be.download(out2)).all()
    for k in ("reward", "terminal", "winners", "next_mover"):
        assert (r[k] == r2[k]).all()
    # and legal ones match the oracle
    for i in np.flatnonzero(legal)[:60]:
        j = sel[i]
        ost = (board[j].astype(np.int64), int(rounds[j]), inv[j], scores[j])
        nst, nxt, rew, term, win = orc.blokus_next_state(ost, int(g["mover"][j]), int(acts[i]))
        assert rew == r["reward"][i] and term == r["terminal"][i] and (win if term else 0) == r["winners"][i]


def case_rollout_vs_oracle(be, B=24, K=80, seed=3, env0=500, cap=2048):
    """legal + policy + step with auto-reset and fused statistics == the oracle's rollout."""
    ob = orc.BlokusBatch(B)
    ob.rollout(seed, env0, 0, K, fresh=True)
    st, st2 = be.zeros((B, 22, 4), np.int32), be.zeros((B, 22, 4), np.int32)
    stats = be.zeros((STAT_ROWS, 32), np.int64)
    counts, ids = be.zeros((B,), np.int32), be.zeros((B, cap), np.int32)
    act, res = be.zeros((B,), np.int32), be.zeros((B, 8), np.uint8)
    be.check(be.lib.crl_blokus_reset(be.ptr(st), None, B, be.stream))
    cur, nxt = st, st2
    for t in range(K):
        be.check(be.lib.crl_blokus_legal(be.ptr(cur), -1, be.ptr(counts), be.ptr(ids), cap, be.ptr(stats), B, 1, be.stream))
        be.check(be.lib.crl_blokus_policy_random(be.ptr(counts), be.ptr(ids), cap, be.ptr(act), seed, env0, t, B, be.stream))
        be.check(be.lib.crl_blokus_step(be.ptr(cur), be.ptr(nxt), be.ptr(act), be.ptr(res), be.ptr(stats), B, 1, be.stream))
        cur, nxt = nxt, cur
    board, pieces, score, meta = blk_unpack(be, cur)
    assert (board == ob.board).all() and (pieces == ob.inventory).all() and (score == ob.scores).all()
    assert (meta[:, 0] == ob.round_count).all() and (meta[:, 1] == ob.mover).all()
    assert (meta[:, 2] == ob.terminal).all() and (meta[:, 3] == ob.ep_len).all()
    s = be.download(stats).sum(0)
    assert (s == ob.stats).all(), (s, ob.stats)
    assert int(be.download(counts).max()) <= cap


def case_reset_and_capacity(be):
    B = 9
    st = be.zeros((B, 22, 4), np.int32)
    be.check(be.lib.crl_blokus_reset(be.ptr(st), None, B, be.stream))
    board, pieces, score, meta = blk_unpack(be, st)
    assert (board == 0).all() and (pieces == 1).all() and (score == 0).all() and (meta == 0).all()
    for p in range(4):
        counts, ids = blk_legal(be, st, player=p)
        assert (counts == 116).all()                                   # 116 first moves for every seat
        exp = orc.blokus_valid_moves(orc.blokus_new_state(), p)
        assert (ids[0, :116] == exp).all()
    # capacity smaller than the list: count is still the full length, the prefix is written
    counts, ids = blk_legal(be, st, player=0, cap=50)
    assert (counts == 116).all() and (ids[3] == orc.blokus_valid_moves(orc.blokus_new_state(), 0)[:50]).all()
    # masked reset
    out, _ = blk_step(be, st, np.full(B, orc.blokus_encode_action(20, 0, 0, 2, 0), np.int32))
    mask = (np.arange(B) % 2).astype(np.uint8)
    m = be.upload(mask)
    before = be.download(out)
    be.check(be.lib.crl_blokus_reset(be.ptr(out), be.ptr(m), B, be.stream))
    after = be.download(out)
    fresh = be.download(st)
    assert (after[mask == 1] == fresh[mask == 1]).all() and (after[mask == 0] == before[mask == 0]).all()
    assert be.lib.crl_blokus_legal(be.ptr(st), 4, be.ptr(st), be.ptr(st), 10, None, B, 0, be.stream) == 1


def blk_is_valid(be, st, acts, player=-1, flags=0):
    B = st.shape[0]
    a, v = be.upload(np.ascontiguousarray(acts, np.int32)), be.zeros((B,), np.uint8)
    be.check(be.lib.crl_blokus_is_valid(be.ptr(st), player, be.ptr(a), be.ptr(v), B, flags, be.stream))
    return be.download(v).astype(bool)


def case_is_valid(be, stride=2, seed=5):
    """is_valid_action (BlokusEnvironment.py:667-719) == membership in the ordered valid list, for every seat of the
    reference's recorded positions: listed ids, their neighbours in id space, arbitrary and out-of-range ids, ''."""
    g = np.load(os.path.join(GOLDEN, "blokus_games.npz"))
    board, inv, scores, rounds = _golden_prev(g)
    sel = np.arange(0, len(g["t"]), stride)
    st = blk_pack(be, board[sel], inv[sel], scores[sel], rounds[sel], g["mover"][sel])
    rng = np.random.RandomState(seed)
    n_true = n_false = 0
    for p in (-1, 0, 1, 2, 3):
        counts, ids = blk_legal(be, st, player=p, cap=4096)
        for rep in range(5):
            acts = np.zeros(len(sel), np.int32)
            exp = np.zeros(len(sel), bool)
            for i in range(len(sel)):
                valid = set(ids[i, :counts[i]].tolist())
                kind = (i + rep) % 5
                if kind == 0 and counts[i]:
                    a = int(ids[i, rng.randint(counts[i])])
                elif kind == 1 and counts[i]:
                    a = int(ids[i, rng.randint(counts[i])]) + int(rng.choice([-40, -5, -1, 1, 5, 40, 800, -800, 16000]))
                elif kind == 2:
                    a = int(rng.randint(0, 21 * 16000))
                elif kind == 3:
                    a = int(rng.choice([-1, -2, 21 * 16000, 21 * 16000 + 7, 2 ** 31 - 1, -2 ** 31, 336000 - 1]))
                else:
                    # a held piece on a real anchor with a random orientation / shift: the most likely near miss
                    a = (int(ids[i, rng.randint(counts[i])]) // 40) * 40 + int(rng.randint(40)) if counts[i] else 0
                acts[i] = a
                exp[i] = a in valid
            got = blk_is_valid(be, st, acts, player=p)
            assert (got == exp).all(), (p, rep, np.flatnonzero(got != exp)[:5], acts[got != exp][:5])
            n_true += int(exp.sum())
            n_false += int((~exp).sum())
    assert n_true > 100 and n_false > 100


def case_random_boards(be, n=96, seed=11):
    """Hand-built (not necessarily reachable) positions: random colour blobs, random inventories, random round and
    mover.  valid_actions for every seat and next_state of a random legal / pass action vs the oracle."""
    rng = np.random.RandomState(seed)
    boards = np.zeros((n, 20, 20), np.int8)
    inv = (rng.rand(n, 4, 21) < rng.uniform(0.1, 0.9, size=(n, 1, 1))).astype(np.uint8)
    scores = rng.randint(0, 90, size=(n, 4))
    rounds = rng.randint(0, 4, size=n)
    movers = rng.randint(0, 4, size=n)
    for i in range(n):
        density = rng.uniform(0.05, 0.7)
        for _ in range(int(density * 60)):
            c = rng.randint(1, 5)
            y, x = rng.randint(0, 20, size=2)
            h, w = rng.randint(1, 4, size=2)
            if rng.rand() < 0.5:
                boards[i, y:y + h, x:x + w] = np.where(boards[i, y:y + h, x:x + w] == 0, c, boards[i, y:y + h, x:x + w])
            else:
                boards[i, y, x] = c
        if i % 7 == 0:
            boards[i] = 0          # empty boards with odd inventories / rounds
    st = blk_pack(be, boards, inv, scores, rounds, movers)
    lists = {}
    for p in range(4):
        counts, ids = blk_legal(be, st, player=p, cap=4096)
        for i in range(n):
            exp = orc.blokus_valid_moves((boards[i].astype(np.int64), int(rounds[i]), inv[i], scores[i]), p, cap=16384)
            assert counts[i] == len(exp), (i, p, counts[i], len(exp))
            assert (ids[i, :len(exp)] == exp).all(), (i, p)
            lists[(i, p)] = exp
    acts = np.full(n, -1, np.int32)
    for i in range(n):
        v = lists[(i, int(movers[i]))]
        if len(v) and i % 5 != 0:
            acts[i] = v[rng.randint(len(v))]
    out, r = blk_step(be, st, acts)
    b2, p2, s2, m2 = blk_unpack(be, out)
    for i in range(n):
        ost = (boards[i].astype(np.int64), int(rounds[i]), inv[i], scores[i].astype(np.int64))
        nst, nxt, rew, term, win = orc.blokus_next_state(ost, int(movers[i]), int(acts[i]))
        assert (b2[i] == nst[0]).all() and (p2[i] == nst[2]).all() and (s2[i] == nst[3]).all(), i
        assert m2[i, 0] == nst[1] and m2[i, 1] == nxt and bool(m2[i, 2]) == term
        assert r["reward"][i] == rew and r["terminal"][i] == term and r["winners"][i] == (win if term else 0), i
        assert not r["error"][i]


def case_many_anchors(be):
    """Lattice positions with far more anchors than random play ever has (up to ~170: several 16-anchor chunks of
    the emission, anchors on every row / column, pieces hanging over all four edges): every seat's full ordered
    list vs the oracle, with short and full inventories."""
    boards, invs = [], []
    for step, off, colour in ((3, 1, 1), (3, 0, 2), (4, 1, 3), (3, 2, 4), (5, 2, 1), (2, 0, 2)):
        b = np.zeros((20, 20), np.int8)
        b[off::step, off::step] = colour
        if step == 2:                                   # checkerboard of the even lattice: anchors on all (odd, odd) cells
            b[:] = 0
            ys, xs = np.mgrid[0:20:2, 0:20:2]
            sel = ((ys + xs) % 4 == 0)
            b[ys[sel], xs[sel]] = colour
        boards.append(b)
        invs.append(np.ones((4, 21), np.uint8))
        boards.append(b.copy())
        short = np.zeros((4, 21), np.uint8)
        short[:, [0, 3, 8, 9, 14, 19, 20]] = 1
        invs.append(short)
    boards, invs = np.stack(boards), np.stack(invs)
    n = len(boards)
    scores = np.zeros((n, 4), np.int64)
    rounds = np.full(n, 3)
    movers = np.zeros(n, np.int64)
    st = blk_pack(be, boards, invs, scores, rounds, movers)
    most = 0
    for p in range(4):
        counts, ids = blk_legal(be, st, player=p, cap=16384)
        for i in range(n):
            most = max(most, len(orc.blokus_anchors(boards[i].astype(np.int64), 3, p + 1)))
            exp = orc.blokus_valid_moves((boards[i].astype(np.int64), 3, invs[i], scores[i]), p, cap=65536)
            assert counts[i] == len(exp), (i, p, counts[i], len(exp))
            assert (ids[i, :min(len(exp), 16384)] == exp[:16384]).all(), (i, p)
    assert most > 128, most
