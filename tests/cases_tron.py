"""Tron parity cases, written once against the C ABI and run on both backends (tests/backends.py)."""
import glob
import os

import numpy as np

from oracle import oracle as orc
from colosseumrl_b200._lib import STAT_ROWS
from colosseumrl_b200 import philox

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def tron_geom(be, N, P):
    """(int8 actions per environment, bytes per result record): 4 / 8 on the tuned path (N <= 19, P <= 4), 8 / 16 on the
    wide path (csrc/tron_wide.cuh)."""
    return be.lib.crl_tron_action_stride(N, P), be.lib.crl_tron_result_bytes(N, P)


def tron_state(be, N, P, B):
    """Zeroed packed state: int32 [13, B, 4] (tuned path) or int32 [W, B] (wide path); B is shape[1] in both."""
    if tron_geom(be, N, P)[0] == 4:
        return be.zeros((13, B, 4), np.int32)
    return be.zeros((be.lib.crl_tron_state_bytes(N, P, 1) // 4, B), np.int32)


def unpack_result(res):
    """uint8 [B, 8] (tuned path) or [B, 16] (wide path) -> dict"""
    res = np.asarray(res).view(np.uint8)
    if res.shape[-1] == 16:
        rk = res[:, 12:16].copy().view(np.uint32)[:, 0].astype(np.int64)
        return dict(rewards=res[:, :8].view(np.int8).astype(np.int64), terminal=res[:, 8].astype(bool),
                    alive=res[:, 9].astype(np.int64), winners=res[:, 10].astype(np.int64),
                    ranking=np.stack([(rk >> (3 * p)) & 7 for p in range(8)], axis=1))
    res = res.reshape(-1, 8)
    rk = res[:, 7]
    return dict(rewards=res[:, :4].view(np.int8).astype(np.int64), terminal=res[:, 4].astype(bool),
                alive=res[:, 5].astype(np.int64), winners=res[:, 6].astype(np.int64),
                ranking=np.stack([(rk >> (2 * p)) & 3 for p in range(4)], axis=1).astype(np.int64))


def tron_pack(be, N, P, board, heads, dirs, deaths):
    B = board.shape[0]
    st = tron_state(be, N, P, B)
    b, h, d, de = (be.upload(np.ascontiguousarray(board, np.int8)), be.upload(np.ascontiguousarray(heads, np.int32)),
                   be.upload(np.ascontiguousarray(dirs, np.int32)), be.upload(np.ascontiguousarray(deaths, np.int32)))
    be.check(be.lib.crl_tron_pack(be.ptr(st), be.ptr(b), be.ptr(h), be.ptr(d), be.ptr(de), B, N, P, be.stream))
    return st


def tron_unpack(be, st, N, P, player=-1):
    B = st.shape[1]
    board = be.zeros((B, N, N), np.int8)
    heads, dirs, deaths = (be.zeros((B, P), np.int32) for _ in range(3))
    term = be.zeros((B,), np.uint8)
    be.check(be.lib.crl_tron_observe(be.ptr(st), player, be.ptr(board), be.ptr(heads), be.ptr(dirs), be.ptr(deaths),
                                     be.ptr(term), B, N, P, be.stream))
    return tuple(be.download(x) for x in (board, heads, dirs, deaths, term))


def tron_step(be, st, actions, N, P, flags=0, stats=None, out=None):
    B = st.shape[1]
    aw, rb = tron_geom(be, N, P)
    act = np.zeros((B, aw), np.int8)
    act[:, :actions.shape[1]] = actions
    a = be.upload(act)
    res = be.zeros((B, rb), np.uint8)
    out = tron_state(be, N, P, B) if out is None else out
    be.check(be.lib.crl_tron_step(be.ptr(st), be.ptr(out), be.ptr(a), be.ptr(res), be.ptr(stats), B, N, P, flags, be.stream))
    return out, unpack_result(be.download(res))


def tron_ranking(be, st, N, P):
    """crl_tron_ranking -> int64 [B, P] (uint8 / 2 bits per player on the tuned path, uint32 / 3 bits on the wide one)."""
    B = st.shape[1]
    wide = tron_geom(be, N, P)[0] != 4
    rk = be.zeros((B,), np.int32 if wide else np.uint8)
    be.check(be.lib.crl_tron_ranking(be.ptr(st), be.ptr(rk), B, N, P, be.stream))
    rk = be.download(rk).astype(np.int64)
    bits = 3 if wide else 2
    return np.stack([(rk >> (bits * p)) & ((1 << bits) - 1) for p in range(P)], axis=1)


def case_start_positions(be):
    for N in list(range(5, 26)) + [31, 40, 64]:
        for P in (2, 3, 4, 5, 6, 8):
            if P > 4 and N < 7:
                continue
            h = (np.zeros(8, np.int32), np.zeros(8, np.int32))
            import ctypes as C
            be.check(be.lib.crl_tron_start_positions(N, P, h[0].ctypes.data_as(C.POINTER(C.c_int32)),
                                                     h[1].ctypes.data_as(C.POINTER(C.c_int32))))
            oh, od = orc.tron_start_positions(N, P)
            assert (h[0][:P] == oh).all() and (h[1][:P] == od).all(), (N, P)


def case_start_positions_golden(be):
    """generate_start_positions for every (N, P, ring_offset, spawn_offset) recorded from the real reference
    (tests/golden/tron_starts.npz, oracle/make_golden_starts.py), and new_state at those spawns."""
    import ctypes as C
    tab = np.load(os.path.join(GOLDEN, "tron_starts.npz"))["table"]
    checked = resets = 0
    for row in tab:
        N, P, ring, spawn, ok = (int(v) for v in row[:5])
        if N > 64 or P > 8 or not ok:
            continue
        h, d = np.zeros(8, np.int32), np.zeros(8, np.int32)
        rc = be.lib.crl_tron_start_positions_at(N, P, ring, spawn, h.ctypes.data_as(C.POINTER(C.c_int32)),
                                                d.ctypes.data_as(C.POINTER(C.c_int32)))
        distinct = len(set(row[5:5 + P].tolist())) == P
        if not distinct:
            assert rc != 0, (N, P, ring, spawn)      # two players on one cell: refused
            continue
        assert rc == 0, (N, P, ring, spawn, be.lib.crl_last_error())
        assert h[:P].tolist() == row[5:5 + P].tolist() and d[:P].tolist() == row[9:9 + P].tolist(), (N, P, ring, spawn)
        oh, od = orc.tron_start_positions(N, P, ring, spawn)
        assert oh.tolist() == h[:P].tolist() and od.tolist() == d[:P].tolist()
        checked += 1
        if (N, P) in ((19, 4), (15, 4), (9, 3), (12, 2), (21, 4), (11, 6), (25, 8)) and spawn in (-2, 0, 3):
            B = 5
            st = tron_state(be, N, P, B)
            be.check(be.lib.crl_tron_reset_at(be.ptr(st), None, B, N, P, ring, spawn, be.stream))
            board, heads, dirs, deaths, term = tron_unpack(be, st, N, P)
            exp = np.zeros(N * N, np.int64)
            exp[row[5:5 + P]] = np.arange(1, P + 1)
            assert (board.reshape(B, -1) == exp[None]).all() and (heads == row[None, 5:5 + P]).all()
            assert (dirs == row[None, 9:9 + P]).all() and (deaths == 0).all() and (term == 0).all()
            resets += 1
    assert checked > 1000 and resets > 20, (checked, resets)


def case_reset(be):
    for N, P in [(19, 4), (9, 4), (7, 3), (8, 2), (15, 4), (19, 2), (6, 4), (5, 2), (21, 4), (11, 6), (25, 8), (20, 2)]:
        B = 70
        st = tron_state(be, N, P, B)
        be.check(be.lib.crl_tron_reset(be.ptr(st), None, B, N, P, be.stream))
        board, heads, dirs, deaths, term = tron_unpack(be, st, N, P)
        ob, oh, od, ode = orc.tron_new_state(N, P)
        assert (board == ob[None]).all() and (heads == oh[None]).all() and (dirs == od[None]).all()
        assert (deaths == 0).all() and (term == 0).all()


def _golden_files():
    # every recorded shape, incl. tron_N21_P4 and tron_N11_P6 (the wide path, csrc/tron_wide.cuh)
    return sorted(glob.glob(os.path.join(GOLDEN, "tron_N*_P*.npz")))


def case_golden_steps(be, path):
    """Every recorded (state, actions) -> (next state, outputs) transition of the real reference, as one batch."""
    g = np.load(path)
    N, P = int(g["N"]), int(g["P"])
    T = len(g["t"])
    s0 = orc.tron_new_state(N, P)
    first = g["t"] == 0
    prev = lambda name, init: np.where(first.reshape((-1,) + (1,) * (g[name].ndim - 1)),
                                       init[None], np.concatenate([init[None], g[name][:-1]], 0))
    st = tron_pack(be, N, P, prev("board", s0[0]), prev("heads", s0[1]), prev("directions", s0[2]), prev("deaths", s0[3]))
    out, res = tron_step(be, st, g["actions"], N, P)
    board, heads, dirs, deaths, term = tron_unpack(be, out, N, P)
    assert (board == g["board"]).all() and (heads == g["heads"]).all()
    assert (dirs == g["directions"]).all() and (deaths == g["deaths"]).all()
    assert (res["rewards"][:, :P] == g["rewards"]).all() and (res["terminal"] == g["terminal"]).all()
    assert (term.astype(bool) == g["terminal"]).all()
    assert (res["alive"] == g["alive"]).all() and (res["winners"] == g["winners"]).all()
    assert (res["ranking"][:, :P] == g["ranking"]).all()
    # observations (state_to_observation) for every player on the sampled states
    idx = g["obs_idx"]
    sub = tron_pack(be, N, P, g["board"][idx], g["heads"][idx], g["directions"][idx], g["deaths"][idx])
    for p in range(P):
        ob, oh, od, ode, _ = tron_unpack(be, sub, N, P, player=p)
        sel = np.arange(len(idx)) * P + p
        assert (ob == g["obs_board"][sel]).all() and (oh == g["obs_heads"][sel]).all()
        assert (od == g["obs_directions"][sel]).all() and (ode == g["obs_deaths"][sel]).all()
    # all views in one pass (CRL_PLAYER_ALL)
    n = len(idx)
    ab = be.zeros((n, P, N, N), np.int8)
    ah, ad, ade = (be.zeros((n, P, P), np.int32) for _ in range(3))
    be.check(be.lib.crl_tron_observe(be.ptr(sub), -3, be.ptr(ab), be.ptr(ah), be.ptr(ad), be.ptr(ade), None, n, N, P, be.stream))
    assert (be.download(ab).reshape(n * P, N, N) == g["obs_board"]).all()
    assert (be.download(ah).reshape(n * P, P) == g["obs_heads"]).all()
    assert (be.download(ad).reshape(n * P, P) == g["obs_directions"]).all()
    assert (be.download(ade).reshape(n * P, P) == g["obs_deaths"]).all()


def case_adversarial(be, name="tron_adversarial.npz"):
    """Hand-built states stepped by the REFERENCE (tests/golden/tron_adversarial.npz; `tron_adversarial_wide.npz`,
    oracle/make_golden_tron_wide.py: 5..8 players and / or boards beyond 19x19 -- the wide path)."""
    g = np.load(os.path.join(GOLDEN, name))
    for N in sorted(set(g["N"].tolist())):
        for P in sorted(set(g["P"].tolist())):
            m = (g["N"] == N) & (g["P"] == P)
            if not m.any():
                continue
            st = tron_pack(be, N, P, g["board"][m][:, :N, :N], g["heads"][m][:, :P], g["directions"][m][:, :P], g["deaths"][m][:, :P])
            out, res = tron_step(be, st, g["actions"][m][:, :P], N, P)
            board, heads, dirs, deaths, term = tron_unpack(be, out, N, P)
            assert (board == g["o_board"][m][:, :N, :N]).all() and (heads == g["o_heads"][m][:, :P]).all()
            assert (dirs == g["o_directions"][m][:, :P]).all() and (deaths == g["o_deaths"][m][:, :P]).all()
            assert (res["rewards"][:, :P] == g["o_rewards"][m][:, :P]).all()
            assert (res["terminal"] == g["o_terminal"][m]).all() and (res["alive"] == g["o_alive"][m]).all()
            assert (res["winners"] == g["o_winners"][m]).all()
            exp = g["o_ranking"][m][:, :P]
            ok = exp >= 0   # players owning no cell are absent from the reference's Counter (hand-built states only)
            assert (res["ranking"][:, :P][ok] == exp[ok]).all()
            # compute_ranking of the INPUT states through the standalone entry point
            got = tron_ranking(be, st, N, P)
            exp = g["i_ranking"][m][:, :P]
            ok = exp >= 0
            assert (got[ok] == exp[ok]).all()


def case_rollout_vs_oracle(be, N=19, P=4, B=200, K=48, seed=5, env0=1000):
    """policy + step (auto-reset, stats) K times == oracle rollout; and the fused rollout kernel == both."""
    ob = orc.TronBatch(B, N, P)
    ob.rollout(seed, env0, 0, K, fresh=True)
    aw, rb = tron_geom(be, N, P)
    st = tron_state(be, N, P, B)
    st2 = tron_state(be, N, P, B)
    stats = be.zeros((STAT_ROWS, 32), np.int64)
    act = be.zeros((B, aw), np.int8)
    res = be.zeros((B, rb), np.uint8)
    be.check(be.lib.crl_tron_reset(be.ptr(st), None, B, N, P, be.stream))
    cur, nxt = st, st2
    policy = be.lib.crl_tron_policy_random if aw == 4 else be.lib.crl_tron_policy_random_wide
    for t in range(K):
        be.check(policy(be.ptr(act), seed, env0, t, B, be.stream))
        if t == 3:
            r = philox.env_step_words(seed, env0 + np.arange(B), t, philox.TAG_TRON)
            exp = np.array([0, 1, -1], np.int8)[(r % 3).astype(np.int64)]
            if aw == 8:                              # players 4..7: (r[p - 4] // 3) % 3
                exp = np.concatenate([exp, np.array([0, 1, -1], np.int8)[((r // 3) % 3).astype(np.int64)]], axis=1)
            assert (be.download(act) == exp).all()
        be.check(be.lib.crl_tron_step(be.ptr(cur), be.ptr(nxt), be.ptr(act), be.ptr(res), be.ptr(stats), B, N, P, 1, be.stream))
        cur, nxt = nxt, cur
    board, heads, dirs, deaths, term = tron_unpack(be, cur, N, P)
    assert (board == ob.board).all() and (heads == ob.heads).all() and (dirs == ob.directions).all()
    assert (deaths == ob.deaths).all() and (term == ob.terminal).all()
    s = be.download(stats).sum(0)
    assert (s == ob.stats).all(), (s, ob.stats)
    # fused K-step kernel, split in two launches
    st3 = tron_state(be, N, P, B)
    stats3 = be.zeros((STAT_ROWS, 32), np.int64)
    be.check(be.lib.crl_tron_reset(be.ptr(st3), None, B, N, P, be.stream))
    be.check(be.lib.crl_tron_rollout(be.ptr(st3), None, be.ptr(stats3), seed, env0, 0, K // 3, B, N, P, be.stream))
    be.check(be.lib.crl_tron_rollout(be.ptr(st3), be.ptr(res), be.ptr(stats3), seed, env0, K // 3, K - K // 3, B, N, P, be.stream))
    assert (be.download(st3) == be.download(cur)).all()
    assert (be.download(stats3).sum(0) == ob.stats).all()


def case_in_place_and_masked_reset(be, N=9, P=4, B=67):
    st = tron_state(be, N, P, B)
    be.check(be.lib.crl_tron_reset(be.ptr(st), None, B, N, P, be.stream))
    rng = np.random.RandomState(0)
    ref = tron_state(be, N, P, B)
    be.check(be.lib.crl_tron_reset(be.ptr(ref), None, B, N, P, be.stream))
    for t in range(6):
        a = rng.randint(-1, 2, size=(B, P)).astype(np.int8)
        ref, _ = tron_step(be, ref, a, N, P)
        st, _ = tron_step(be, st, a, N, P, out=st)       # in place
    assert (be.download(st) == be.download(ref)).all()
    mask = (rng.rand(B) < 0.5).astype(np.uint8)
    m = be.upload(mask)
    be.check(be.lib.crl_tron_reset(be.ptr(st), be.ptr(m), B, N, P, be.stream))
    board, heads, dirs, deaths, term = tron_unpack(be, st, N, P)
    rb, rh, rd, rde, rt = tron_unpack(be, ref, N, P)
    nb = orc.tron_new_state(N, P)
    for e in range(B):
        if mask[e]:
            assert (board[e] == nb[0]).all() and (heads[e] == nb[1]).all() and (deaths[e] == 0).all()
        else:
            assert (board[e] == rb[e]).all() and (heads[e] == rh[e]).all() and (deaths[e] == rde[e]).all()


def case_errors(be):
    st = be.zeros((13, 4, 4), np.int32)
    assert be.lib.crl_tron_reset(be.ptr(st), None, 4, 65, 4, be.stream) == 3      # N too large
    assert be.lib.crl_tron_reset(be.ptr(st), None, 4, 19, 9, be.stream) == 3      # P too large
    assert b"player count" in be.lib.crl_last_error()
    assert be.lib.crl_tron_state_bytes(21, 4, 1) == (4 * 14 + 9) * 4 and be.lib.crl_tron_state_bytes(19, 4, 1) == 208
    assert be.lib.crl_tron_action_stride(19, 4) == 4 and be.lib.crl_tron_action_stride(11, 6) == 8
    assert be.lib.crl_tron_result_bytes(20, 2) == 16 and be.lib.crl_tron_result_bytes(65, 2) == -1
    # compact records / packed actions exist on the tuned path only
    wst, wact, wres = tron_state(be, 21, 4, 4), be.zeros((4, 8), np.int8), be.zeros((4, 16), np.uint8)
    assert be.lib.crl_tron_step(be.ptr(wst), be.ptr(wst), be.ptr(wact), be.ptr(wres), None, 4, 21, 4, 2, be.stream) == 3
    assert be.lib.crl_tron_reset(None, None, 4, 19, 4, be.stream) == 1
    assert be.lib.crl_tron_reset(be.ptr(st), None, 0, 19, 4, be.stream) == 0      # empty batch is fine


def case_compact_result(be, N=19, P=4, B=300, K=40, seed=9):
    """CRL_FLAG_COMPACT_RESULT: the 4-byte record is the second half of the full one, the states are identical, and
    the reference's rewards (TronGridEnvironment.py:313-320) follow from it: alive ? 1 : -1, +9 for the winners of a
    terminal step."""
    rng = np.random.RandomState(seed)
    full = be.zeros((13, B, 4), np.int32)
    be.check(be.lib.crl_tron_reset(be.ptr(full), None, B, N, P, be.stream))
    comp = be.zeros((13, B, 4), np.int32)
    be.check(be.lib.crl_tron_reset(be.ptr(comp), None, B, N, P, be.stream))
    seen_terminal = 0
    for t in range(K):
        act = rng.randint(-1, 2, size=(B, 4)).astype(np.int8)
        a = be.upload(act)
        r8, r4 = be.zeros((B, 8), np.uint8), be.zeros((B, 4), np.uint8)
        be.check(be.lib.crl_tron_step(be.ptr(full), be.ptr(full), be.ptr(a), be.ptr(r8), None, B, N, P, 1, be.stream))
        be.check(be.lib.crl_tron_step(be.ptr(comp), be.ptr(comp), be.ptr(a), be.ptr(r4), None, B, N, P, 1 | 2, be.stream))
        r8, r4 = be.download(r8), be.download(r4)
        assert (r8[:, 4:] == r4).all(), t
        assert (be.download(full) == be.download(comp)).all(), t
        p = np.arange(P)[None, :]
        alive = (r4[:, 1][:, None] >> p) & 1
        win = ((r4[:, 2][:, None] >> p) & 1) * (r4[:, 0][:, None] & 1)
        rewards = 2 * alive.astype(np.int64) - 1 + 9 * win
        assert (rewards == r8[:, :P].view(np.int8)).all(), t
        seen_terminal += int(r4[:, 0].sum())
    assert seen_terminal > 0


def case_packed_actions(be, N=19, P=4, B=300, K=30, seed=11):
    """CRL_FLAG_PACKED_ACTIONS: uint8 [B] with 2 bits per player gives the same states and records as int8 [B, 4]."""
    rng = np.random.RandomState(seed)
    full = be.zeros((13, B, 4), np.int32)
    be.check(be.lib.crl_tron_reset(be.ptr(full), None, B, N, P, be.stream))
    pk = be.zeros((13, B, 4), np.int32)
    be.check(be.lib.crl_tron_reset(be.ptr(pk), None, B, N, P, be.stream))
    for t in range(K):
        act = rng.randint(-1, 2, size=(B, 4)).astype(np.int8)
        a8 = act.astype(np.uint8) & 3
        packed = (a8[:, 0] | a8[:, 1] << 2 | a8[:, 2] << 4 | a8[:, 3] << 6).astype(np.uint8)
        a, ap = be.upload(act), be.upload(packed)
        r1, r2 = be.zeros((B, 8), np.uint8), be.zeros((B, 8), np.uint8)
        prev = be.upload(be.download(full))
        be.check(be.lib.crl_tron_step(be.ptr(full), be.ptr(full), be.ptr(a), be.ptr(r1), None, B, N, P, 1, be.stream))
        be.check(be.lib.crl_tron_step(be.ptr(pk), be.ptr(pk), be.ptr(ap), be.ptr(r2), None, B, N, P, 1 | 4, be.stream))
        assert (be.download(r1) == be.download(r2)).all(), t
        assert (be.download(full) == be.download(pk)).all(), t
        # both flags together, out of place
        r3, out3 = be.zeros((B, 4), np.uint8), be.zeros((13, B, 4), np.int32)
        be.check(be.lib.crl_tron_step(be.ptr(prev), be.ptr(out3), be.ptr(ap), be.ptr(r3), None, B, N, P, 1 | 2 | 4, be.stream))
        assert (be.download(r3) == be.download(r1)[:, 4:]).all() and (be.download(out3) == be.download(full)).all(), t


def case_wide_adversarial(be, seed=3, n=400):
    """Hand-built states on shapes beyond N <= 19, P <= 4 (the wide path) against the oracle: collision-dense boards
    with up to 8 players, every death code, head-on and head-into-head moves, compute_ranking of input and output."""
    rng = np.random.RandomState(seed)
    for (N, P) in ((20, 2), (21, 4), (7, 5), (11, 6), (9, 8), (25, 8)):
        boards, heads, dirs, deaths, acts = [], [], [], [], []
        for _ in range(n // 6 + 1):
            board = np.zeros((N, N), np.int64)
            fill = rng.rand(N, N) < rng.uniform(0.05, 0.5)
            board[fill] = rng.randint(1, P + 1, size=int(fill.sum()))
            h = rng.choice(N * N, size=P, replace=False).astype(np.int64)
            if rng.rand() < 0.5:                      # crowd the heads around one cell
                c = int(rng.randint(1, N - 1)) * N + int(rng.randint(1, N - 1))
                neigh = [c - 1, c + 1, c - N, c + N, c, c - N - 1, c - N + 1, c + N - 1, c + N + 1]
                rng.shuffle(neigh)
                h = np.asarray(neigh[:P], np.int64)
            board.ravel()[h] = np.arange(1, P + 1)
            boards.append(board); heads.append(h); dirs.append(rng.randint(0, 4, size=P))
            deaths.append(np.where(rng.rand(P) < 0.3, rng.randint(1, P + 1, size=P), 0))
            acts.append(rng.randint(-1, 2, size=P))
        boards, heads, dirs, deaths, acts = (np.asarray(x, np.int64) for x in (boards, heads, dirs, deaths, acts))
        st = tron_pack(be, N, P, boards, heads, dirs, deaths)
        i_rank = tron_ranking(be, st, N, P)
        out, res = tron_step(be, st, acts, N, P)
        b2, h2, d2, de2, term = tron_unpack(be, out, N, P)
        for i in range(len(boards)):
            state = (boards[i], heads[i], dirs[i], deaths[i])
            exp_i = orc.tron_compute_ranking(state)
            ok = exp_i >= 0                           # players owning no cell are absent from the reference's Counter
            assert (i_rank[i][ok] == exp_i[ok]).all(), (N, P, i)
            nst, alive, rewards, terminal, winners = orc.tron_next_state(state, acts[i])
            assert (b2[i] == nst[0]).all() and (h2[i] == nst[1]).all() and (d2[i] == nst[2]).all(), (N, P, i)
            assert (de2[i] == nst[3]).all() and bool(term[i]) == terminal, (N, P, i)
            assert (res["rewards"][i, :P] == rewards).all() and res["alive"][i] == alive and res["winners"][i] == winners
            assert res["terminal"][i] == terminal
            exp_o = orc.tron_compute_ranking(nst)
            ok = exp_o >= 0
            assert (res["ranking"][i, :P][ok] == exp_o[ok]).all(), (N, P, i)
