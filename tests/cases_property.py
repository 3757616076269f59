"""Property tests (hypothesis): the engine against the oracle on GENERATED states, through the C ABI on both backends.

hypothesis draws the *shape* of a case -- game configuration, densities, probabilities, a seed -- and a numpy generator
seeded with it fills a small batch of states; a failing case shrinks to small parameters.  States are arbitrary, not
only reachable ones (carried winners on full boards, dead players on live trails, tied scores ...): the reference's
functions are total on them and so is the oracle (oracle/*.c restates them line by line, pinned by tests/golden/).
"""
import numpy as np
from hypothesis import given, settings, strategies as st, HealthCheck

from oracle import oracle as orc
import cases_tron as ct
import cases_ttt as cx
import cases_blokus as cb


def _settings(n):
    return settings(max_examples=n, deadline=None, derandomize=True,
                    suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large, HealthCheck.function_scoped_fixture])


# ----------------------------------------------------------------------------------------------- Tron
def tron_property(be, examples):
    @_settings(examples)
    @given(N=st.integers(5, 26), P=st.integers(2, 8), seed=st.integers(0, 2 ** 31 - 1),
           fill=st.floats(0.0, 0.7), dead=st.floats(0.0, 0.6), crowd=st.booleans())
    def run(N, P, seed, fill, dead, crowd):
        import ctypes as C
        h8, d8 = (C.c_int32 * 8)(), (C.c_int32 * 8)()
        if be.lib.crl_tron_start_positions(N, P, h8, d8) != 0:     # the spawn ring cannot hold the players: the reference's
            return                                                 # new_state fails for this shape too, nothing to compare
        rng = np.random.RandomState(seed)
        B = 12
        boards = np.zeros((B, N, N), np.int64)
        heads = np.zeros((B, P), np.int64)
        for i in range(B):
            m = rng.rand(N, N) < fill
            boards[i][m] = rng.randint(1, P + 1, size=int(m.sum()))
            h = rng.choice(N * N, size=P, replace=False)
            if crowd and N >= 5:                      # heads next to each other: head-on and head-into-head moves
                c = int(rng.randint(1, N - 1)) * N + int(rng.randint(1, N - 1))
                neigh = [c - 1, c + 1, c - N, c + N, c, c - N - 1, c - N + 1, c + N - 1, c + N + 1]
                rng.shuffle(neigh)
                h = np.asarray((neigh + list(h))[:P])
                if len(set(h.tolist())) < P:
                    h = rng.choice(N * N, size=P, replace=False)
            heads[i] = h
            boards[i].ravel()[h] = np.arange(1, P + 1)
        dirs = rng.randint(0, 4, size=(B, P))
        deaths = np.where(rng.rand(B, P) < dead, rng.randint(1, P + 1, size=(B, P)), 0)
        acts = rng.randint(-1, 2, size=(B, P))
        sst = ct.tron_pack(be, N, P, boards, heads, dirs, deaths)
        i_rank = ct.tron_ranking(be, sst, N, P)
        out, res = ct.tron_step(be, sst, acts, N, P)
        b2, h2, d2, de2, term = ct.tron_unpack(be, out, N, P)
        viewer = int(rng.randint(P))
        ob, oh, od, ode, _ = ct.tron_unpack(be, out, N, P, player=viewer)
        for i in range(B):
            state = (boards[i], heads[i], dirs[i], deaths[i])
            exp = orc.tron_compute_ranking(state)
            ok = exp >= 0
            assert (i_rank[i][ok] == exp[ok]).all()
            nst, alive, rewards, terminal, winners = orc.tron_next_state(state, acts[i])
            assert (b2[i] == nst[0]).all() and (h2[i] == nst[1]).all() and (d2[i] == nst[2]).all() and (de2[i] == nst[3]).all()
            assert bool(term[i]) == terminal == bool(res["terminal"][i])
            assert (res["rewards"][i, :P] == rewards).all() and res["alive"][i] == alive and res["winners"][i] == winners
            exp = orc.tron_compute_ranking(nst)
            ok = exp >= 0
            assert (res["ranking"][i, :P][ok] == exp[ok]).all()
            oo = orc.tron_observation(nst, viewer)
            assert (ob[i] == oo["board"]).all() and (oh[i] == oo["heads"]).all()
            assert (od[i] == oo["directions"]).all() and (ode[i] == oo["deaths"]).all()
    run()


# ----------------------------------------------------------------------------------------------- Tic Tac Toe
def ttt_property(be, examples):
    @_settings(examples)
    @given(n=st.sampled_from([2, 3, 4]), seed=st.integers(0, 2 ** 31 - 1), fill=st.floats(0.0, 1.0),
           carried=st.floats(0.0, 0.5), bad=st.floats(0.0, 0.5))
    def run(n, seed, fill, carried, bad):
        rng = np.random.RandomState(seed)
        B, cells = 40, orc.ttt_cells(n)
        board = np.where(rng.rand(B, cells) < fill, rng.randint(0, n, size=(B, cells)), -1).astype(np.int8)
        winner = np.where(rng.rand(B) < carried, rng.randint(0, n, size=B), -1)
        mover = rng.randint(0, n, size=B)
        action = np.where(rng.rand(B) < bad, rng.randint(-2, cells + 3, size=B), -1)
        for i in range(B):                            # the rest: a free cell when there is one
            free = np.flatnonzero(board[i] == -1)
            if action[i] == -1 and len(free) and rng.rand() < 0.9:
                action[i] = free[rng.randint(len(free))]
        action = np.clip(action, -128, 127)
        sst = cx.ttt_pack(be, n, board, winner, mover)
        vm = be.zeros((B,), np.uint32)
        be.check(be.lib.crl_ttt_valid_actions(be.ptr(sst), be.ptr(vm), B, n, be.stream))
        vmask = be.download(vm).view(np.uint32)
        act = be.upload(action.astype(np.int8))
        res, out = be.zeros((B, 4), np.uint8), be.zeros((B, 4), np.int32)
        be.check(be.lib.crl_ttt_step(be.ptr(sst), be.ptr(out), be.ptr(act), be.ptr(res), None, None, B, n, 0, be.stream))
        r = cx.unpack_result(be.download(res))
        b2, w2, m2 = cx.ttt_unpack(be, out, n)
        viewer = int(rng.randint(n))
        ob, _, _ = cx.ttt_unpack(be, out, n, player=viewer)
        shape = orc.TTT_SHAPE[n]
        for i in range(B):
            state = (board[i].reshape(shape), int(winner[i]))
            va = orc.ttt_valid_actions(n, state)
            assert int(vmask[i]) == sum(1 << int(c) for c in va)
            a = int(action[i]) if 0 <= action[i] < cells else -1       # out-of-range indices: a pass with the error bit
            nst, nxt, reward, terminal, wout = orc.ttt_next_state(n, state, int(mover[i]), a)
            assert (b2[i] == nst[0].ravel()).all() and w2[i] == nst[1] and m2[i] == nxt
            assert r["reward"][i] == reward and r["terminal"][i] == terminal
            assert r["winners"][i] == ((1 << wout) if wout >= 0 else 0)
            occupied = 0 <= action[i] < cells and board[i, action[i]] != -1
            assert r["error"][i] == (occupied or action[i] >= cells)
            assert (ob[i] == orc.ttt_observation(n, nst, viewer).ravel()).all()
    run()


# ----------------------------------------------------------------------------------------------- Blokus
def blokus_property(be, examples):
    @_settings(examples)
    @given(seed=st.integers(0, 2 ** 31 - 1), density=st.floats(0.0, 0.8), held=st.floats(0.05, 1.0),
           rnd=st.integers(0, 3), blobs=st.booleans())
    def run(seed, density, held, rnd, blobs):
        rng = np.random.RandomState(seed)
        n = 5
        boards = np.zeros((n, 20, 20), np.int8)
        for i in range(n):
            for _ in range(int(density * 70)):
                c = rng.randint(1, 5)
                y, x = rng.randint(0, 20, size=2)
                h, w = (rng.randint(1, 4, size=2) if blobs else (1, 1))
                blk = boards[i, y:y + h, x:x + w]
                boards[i, y:y + h, x:x + w] = np.where(blk == 0, c, blk)
        inv = (rng.rand(n, 4, 21) < held).astype(np.uint8)
        scores = rng.randint(0, 90, size=(n, 4))
        rounds = np.full(n, rnd)
        movers = rng.randint(0, 4, size=n)
        sst = cb.blk_pack(be, boards, inv, scores, rounds, movers)
        lists = {}
        for p in range(4):
            counts, ids = cb.blk_legal(be, sst, player=p, cap=16384)     # (generated boards reach > 4 096 moves)
            for i in range(n):
                exp = orc.blokus_valid_moves((boards[i].astype(np.int64), int(rounds[i]), inv[i], scores[i]), p, cap=16384)
                assert counts[i] == len(exp) and (ids[i, :len(exp)] == exp).all()
                lists[(i, p)] = exp
        acts = np.full(n, -1, np.int32)
        for i in range(n):
            v = lists[(i, int(movers[i]))]
            if len(v) and rng.rand() < 0.85:
                acts[i] = v[rng.randint(len(v))]
        legal = cb.blk_is_valid(be, sst, acts)
        assert (legal == (acts >= 0)).all()
        out, r = cb.blk_step(be, sst, acts)
        b2, p2, s2, m2 = cb.blk_unpack(be, out)
        for i in range(n):
            ost = (boards[i].astype(np.int64), int(rounds[i]), inv[i], scores[i].astype(np.int64))
            nst, nxt, rew, term, win = orc.blokus_next_state(ost, int(movers[i]), int(acts[i]))
            assert (b2[i] == nst[0]).all() and (p2[i] == nst[2]).all() and (s2[i] == nst[3]).all()
            assert m2[i, 0] == nst[1] and m2[i, 1] == nxt and bool(m2[i, 2]) == term
            assert r["reward"][i] == rew and r["terminal"][i] == term and r["winners"][i] == (win if term else 0)
    run()
