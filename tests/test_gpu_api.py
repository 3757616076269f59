"""GPU tests of the drop-in Python surface (Batched*Environment: the BaseEnvironment method names) vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from colosseumrl_b200 import philox

pytestmark = pytest.mark.gpu


def test_registry_and_surface():
    from colosseumrl_b200.config import get_environment, available_environments
    assert set(available_environments()) == {"blokus", "tron", "tictactoe", "tictactoe_3p", "tictactoe_4p"}
    env = get_environment("tron")("9;4", batch=8)
    assert env.min_players == env.max_players == 4 and env.observation_shape["board"] == (9, 9)
    assert env.observation_names() == ["board", "heads", "directions", "deaths"]
    for name in ("new_state", "next_state", "valid_actions", "is_valid_action", "is_terminal", "compute_ranking",
                 "state_to_observation", "serialize_state", "deserialize_state"):
        assert hasattr(env, name)
    wide = get_environment("tron")("25;4", batch=8)       # beyond N <= 19: the wide layout (csrc/tron_wide.cuh)
    assert wide.wide and wide.observation_shape["board"] == (25, 25)
    with pytest.raises(Exception):
        get_environment("tron")("65;4", batch=8)          # unsupported board size fails loudly
    with pytest.raises(Exception):
        get_environment("tron")("19;9", batch=8)


def test_tron_api_episode():
    from colosseumrl_b200.tron import BatchedTronGridEnvironment
    B, N, P, seed = 257, 9, 4, 11
    env = BatchedTronGridEnvironment("9;4", batch=B, seed=seed)
    state, players = env.new_state()
    assert (players.cpu().numpy() == 15).all()
    ost = [orc.tron_new_state(N, P) for _ in range(B)]
    for t in range(12):
        r = philox.env_step_words(seed, np.arange(B), t, philox.TAG_TRON)
        acts = np.array([0, 1, -1], np.int8)[(r % 3).astype(np.int64)]
        assert (env.random_actions(t).cpu().numpy() == acts).all()
        state, players, rewards, terminal, winners = env.next_state(state, players, torch.from_numpy(acts))   # host tensor in
        rk = env.compute_ranking(state, None, winners).cpu().numpy()
        obs = {p: env.state_to_observation(state, p) for p in (0, 3)}
        for e in range(0, B, 16):
            ost[e], alive, orew, oterm, owin = orc.tron_next_state(ost[e], acts[e])
            assert alive == int(players[e]) and (orew == rewards[e].cpu().numpy()).all()
            assert oterm == bool(terminal[e]) and owin == int(winners[e]) and bool(env.is_terminal(state)[e]) == oterm
            assert (orc.tron_compute_ranking(ost[e]) == rk[e]).all()
            for p in (0, 3):
                oo = orc.tron_observation(ost[e], p)
                for k in ("board", "heads", "directions", "deaths"):
                    assert (oo[k] == obs[p][k][e].cpu().numpy()).all()
    blob = env.serialize_state(state)
    back = env.deserialize_state(blob)
    assert (back.packed == state.packed).all()
    st2 = env.state_from_arrays(*(env.state_to_observation(state, -1)[k] for k in ("board", "heads", "directions", "deaths")))
    assert (st2.packed[:12] == state.packed[:12]).all()


def test_tron_api_wide_shapes():
    """Shapes beyond N <= 19, P <= 4 through the Python surface (csrc/tron_wide.cuh): next_state / is_terminal /
    compute_ranking / state_to_observation / rollout vs the oracle."""
    from colosseumrl_b200.tron import BatchedTronGridEnvironment
    for config, N, P, B, seed in (("21;4", 21, 4, 65, 3), ("11;6", 11, 6, 130, 4)):
        env = BatchedTronGridEnvironment(config, batch=B, seed=seed)
        assert env.wide and env.action_stride == 8 and env.result_bytes == 16
        state, players = env.new_state()
        assert (players.cpu().numpy() == (1 << P) - 1).all()
        ost = [orc.tron_new_state(N, P) for _ in range(B)]
        for t in range(14):
            acts = env.random_actions(t)
            a = acts.cpu().numpy()
            state, players, rewards, terminal, winners = env.next_state(state, players, acts)
            rk = env.compute_ranking(state, None, winners).cpu().numpy()
            obs = {p: env.state_to_observation(state, p) for p in (0, P - 1)}
            for e in range(0, B, 8):
                ost[e], alive, orew, oterm, owin = orc.tron_next_state(ost[e], a[e, :P])
                assert alive == int(players[e]) and (orew == rewards[e].cpu().numpy()).all()
                assert oterm == bool(terminal[e]) and owin == int(winners[e]) and bool(env.is_terminal(state)[e]) == oterm
                assert (orc.tron_compute_ranking(ost[e]) == rk[e]).all()
                for p in (0, P - 1):
                    oo = orc.tron_observation(ost[e], p)
                    for k in ("board", "heads", "directions", "deaths"):
                        assert (oo[k] == obs[p][k][e].cpu().numpy()).all()
        # ranking of a state that no step produced (crl_tron_ranking) == the fused one
        fresh = env.state_from_arrays(*(env.state_to_observation(state, -1)[k] for k in ("board", "heads", "directions", "deaths")))
        assert (env.compute_ranking(fresh).cpu().numpy() == rk).all() and (env.is_terminal(fresh) == env.is_terminal(state)).all()
        # fused rollout with auto-reset and statistics == the oracle's rollout
        env2 = BatchedTronGridEnvironment(config, batch=B, seed=seed, auto_reset=True)
        st2, _ = env2.new_state()
        env2.rollout(st2, 0, 50)
        ob = orc.TronBatch(B, N, P)
        ob.rollout(seed, 0, 0, 50, fresh=True)
        o2 = env2.state_to_observation(st2, -1)
        assert (o2["board"].cpu().numpy() == ob.board).all() and (o2["deaths"].cpu().numpy() == ob.deaths).all()
        assert (env2.stats.cpu().numpy() == ob.stats).all()
        with pytest.raises(ValueError):
            env.host_stepper(state, compact=2)


def test_ttt_api():
    from colosseumrl_b200.tictactoe import BatchedTicTacToe4PlayerEnv, BatchedTicTacToe2PlayerEnv
    B, seed = 300, 2
    env = BatchedTicTacToe4PlayerEnv(batch=B, seed=seed, auto_reset=True)
    state, players = env.new_state()
    ob = orc.TTTBatch(B, 4)
    for t in range(40):
        acts = env.random_actions(state, t)
        va = env.valid_actions(state)
        assert env.is_valid_action(state, None, acts).cpu().numpy().astype(bool).tolist() == (acts.cpu().numpy() >= 0).tolist() or t > 0
        state, players, reward, terminal, winners = env.next_state(state, players, acts)
    ob.rollout(seed, 0, 0, 40, fresh=True)
    board, winner, mover = env.state_arrays(state)
    assert (board.cpu().numpy() == ob.board).all() and (winner.cpu().numpy() == ob.winner).all()
    assert (mover.cpu().numpy() == ob.mover).all() and (env.stats.cpu().numpy() == ob.stats).all()
    assert (env.is_terminal(state).cpu().numpy() == ob.terminal).all()
    assert (players.cpu().numpy() == (1 << ob.mover)).all()
    rk = env.compute_ranking(state, None, winners).cpu().numpy()
    w = winners.cpu().numpy()
    assert (rk == 1 - ((w[:, None] >> np.arange(4)[None]) & 1)).all()
    e2 = BatchedTicTacToe2PlayerEnv(batch=4)
    s2, _ = e2.new_state()
    s2, pl, rew, term, win = e2.next_state(s2, None, torch.tensor([4, 4, -1, 8], dtype=torch.int8))
    assert e2.state_to_observation(s2, 1)["board"].shape == (4, 3, 3)
    assert (pl.cpu().numpy() == 2).all() and (term.cpu().numpy() == 0).all()


def test_blokus_api():
    from colosseumrl_b200.blokus import BatchedBlokusEnvironment, action_to_string, string_to_action
    B, seed = 48, 6
    env = BatchedBlokusEnvironment(batch=B, seed=seed, auto_reset=True)
    state, players = env.new_state()
    ob = orc.BlokusBatch(B)
    for t in range(90):
        valid = env.valid_actions(state)
        acts = env.random_actions(valid, t)
        if t == 5:
            assert env.is_valid_action(state, -1, acts).cpu().numpy().all()
            s = action_to_string(int(acts[0]))
            assert string_to_action(s) == int(acts[0]) and s.count(";") == 2
        state, players, reward, terminal, winners = env.next_state(state, players, acts)
    ob.rollout(seed, 0, 0, 90, fresh=True)
    obs = env.state_to_observation(state, -1)
    assert (obs["board"].cpu().numpy() == ob.board).all() and (obs["pieces"].cpu().numpy() == ob.inventory).all()
    assert (obs["score"].cpu().numpy() == ob.scores).all() and (obs["mover"].cpu().numpy() == ob.mover).all()
    assert (env.is_terminal(state).cpu().numpy() == ob.terminal).all()
    assert (env.stats.cpu().numpy() == ob.stats).all()
    assert (env.current_rewards(state).cpu().numpy() == ob.scores).all()
    for p in range(4):
        o = env.state_to_observation(state, p)
        e = 7
        oo = orc.blokus_observation((ob.board[e], ob.round_count[e], ob.inventory[e], ob.scores[e]), p)
        assert (o["board"][e].cpu().numpy() == oo["board"]).all() and (o["pieces"][e].cpu().numpy() == oo["pieces"]).all()
        assert (o["score"][e].cpu().numpy() == oo["score"]).all()


def test_host_stepper_matches_next_state():
    """host_stepper (graph-fused H2D + step + D2H) == next_state with the same actions."""
    from colosseumrl_b200.tron import BatchedTronGridEnvironment
    B = 777
    a_env = BatchedTronGridEnvironment("", batch=B, seed=3)
    b_env = BatchedTronGridEnvironment("", batch=B, seed=3)
    sa, _ = a_env.new_state()
    sb, _ = b_env.new_state()
    stepper = b_env.host_stepper(sb)                      # warm-up applies one all-forward step
    sa, *_ = a_env.next_state(sa, None, torch.zeros((B, 4), dtype=torch.int8))
    rng = np.random.RandomState(0)
    for t in range(12):
        a = torch.from_numpy(rng.randint(-1, 2, size=(B, 4)).astype(np.int8))
        sa, pa, ra, ta, wa = a_env.next_state(sa, None, a)
        stepper.actions.copy_(a)
        res = stepper()
        assert stepper.result.is_pinned() and (res == sa.result.cpu().numpy()).all()
    assert (sa.packed == sb.packed).all()
    # a second stepper on its own stream, pipelined with the first (two independent batches in flight)
    c_env = BatchedTronGridEnvironment("", batch=B, seed=3)
    sc, _ = c_env.new_state()
    side = torch.cuda.Stream()
    stepper2 = c_env.host_stepper(sc, stream=side)
    ref = BatchedTronGridEnvironment("", batch=B, seed=3)
    sr, _ = ref.new_state()
    sr, *_ = ref.next_state(sr, None, torch.zeros((B, 4), dtype=torch.int8))
    for t in range(6):
        a = rng.randint(-1, 2, size=(B, 4)).astype(np.int8)
        stepper.actions_np[...] = a
        stepper2.actions_np[...] = a
        stepper.launch(); stepper2.launch()
        sr, *_ = ref.next_state(sr, None, torch.from_numpy(a))
        r1, r2 = stepper.wait().copy(), stepper2.wait().copy()
        assert (r2 == sr.result.cpu().numpy()).all() and r1.shape == r2.shape
    torch.cuda.synchronize()
    assert (sc.packed == sr.packed).all()


@pytest.mark.parametrize("compact", [True, 2, "zero_copy"])
def test_compact_host_stepper(compact):
    """host_stepper(compact=True / 2): 4- / 2-byte records whose decode equals next_state's return values; and the
    zero-copy transport (the kernel reads / writes the pinned host buffers itself): same records, same states."""
    from colosseumrl_b200.tron import BatchedTronGridEnvironment
    B = 513
    zero_copy = compact == "zero_copy"
    compact = 2 if zero_copy else compact
    a_env = BatchedTronGridEnvironment("", batch=B, seed=4, auto_reset=True)
    b_env = BatchedTronGridEnvironment("", batch=B, seed=4, auto_reset=True)
    sa, _ = a_env.new_state()
    sb, _ = b_env.new_state()
    stepper = b_env.host_stepper(sb, compact=compact, packed_actions=True, zero_copy=zero_copy)   # warm-up applies one all-forward step
    sa, *_ = a_env.next_state(sa, None, torch.zeros((B, 4), dtype=torch.int8))
    rng = np.random.RandomState(1)
    terminals = 0
    for t in range(40):
        a = rng.randint(-1, 2, size=(B, 4)).astype(np.int8)
        sa, pa, ra, ta, wa = a_env.next_state(sa, None, torch.from_numpy(a))
        stepper.actions_np[...] = b_env.pack_actions(a)
        rec = stepper() if t % 2 == 0 else stepper.launch_then_wait(stepper)      # (one foreign call: launch + wait)
        assert rec.shape == (B, 2 if compact == 2 else 4) and stepper.actions_np.shape == (B,)
        alive, rewards, terminal, winners, ranking = b_env.decode_compact(rec)
        assert (alive == pa.cpu().numpy()).all() and (rewards == ra.cpu().numpy()).all()
        assert (terminal == ta.cpu().numpy()).all() and (winners == wa.cpu().numpy()).all()
        assert (ranking == a_env.compute_ranking(sa).cpu().numpy()).all()
        terminals += int(terminal.sum())
    assert terminals > 0 and (sa.packed == sb.packed).all()


def test_tron_new_state_spawn_arguments():
    """new_state(ring_offset, spawn_offset) (TronGridEnvironment.py:228) and generate_start_positions through the Python
    surface, batched and single-environment, against the reference-recorded table."""
    import os
    from colosseumrl_b200.tron import BatchedTronGridEnvironment
    from colosseumrl_b200.single import TronGridEnvironment
    tab = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tron_starts.npz"))["table"]
    want = {tuple(int(v) for v in r[:4]): r for r in tab if r[4]}
    env = BatchedTronGridEnvironment("15;4", batch=33)
    single = TronGridEnvironment("15;4")
    for ring, spawn in ((1, 2), (2, 0), (3, -2), (0, 3)):
        row = want[(15, 4, ring, spawn)]
        h, d = env.generate_start_positions(ring, spawn)
        assert h.tolist() == row[5:9].tolist() and d.tolist() == row[9:13].tolist()
        state, _ = env.new_state(ring_offset=ring, spawn_offset=spawn)
        obs = env.state_to_observation(state, -1)
        assert (obs["heads"].cpu().numpy() == row[None, 5:9]).all() and (obs["directions"].cpu().numpy() == row[None, 9:13]).all()
        (board, heads, dirs, deaths), players = single.new_state(ring_offset=ring, spawn_offset=spawn)
        assert heads.tolist() == row[5:9].tolist() and dirs.tolist() == row[9:13].tolist() and not deaths.any()
        assert board.ravel()[heads].tolist() == [1, 2, 3, 4] and int((board != 0).sum()) == 4
    # the tuple form draws one offset in [lo, hi) PER PLAYER (:222-224)
    for _ in range(8):
        (board, heads, dirs, deaths), _ = single.new_state(spawn_offset=(-1, 2))
        for p in range(4):
            assert any(int(heads[p]) == int(want[(15, 4, 1, s)][5 + p]) and int(dirs[p]) == int(want[(15, 4, 1, s)][9 + p])
                       for s in (-1, 0, 1))
    # explicit per-player offsets: player p takes the spawn of the table row of ITS offset; auto-reset keeps them
    offs = [2, -1, 0, 3]
    h, d = env.generate_start_positions(1, offs)
    assert h.tolist() == [int(want[(15, 4, 1, o)][5 + p]) for p, o in enumerate(offs)]
    assert d.tolist() == [int(want[(15, 4, 1, o)][9 + p]) for p, o in enumerate(offs)]
    renv = BatchedTronGridEnvironment("15;4", batch=64, auto_reset=True, seed=9)
    state, _ = renv.new_state(ring_offset=1, spawn_offset=offs)
    start = renv.state_to_observation(state, -1)
    assert (start["heads"].cpu().numpy() == h[None]).all()
    ended = np.zeros(64, bool)
    for t in range(60):
        state, _, _, terminal, _ = renv.next_state(state, None, renv.random_actions(t), out=state)
        ended |= terminal.cpu().numpy().astype(bool)
    assert ended.all()
    # step once more with everything finished at some point: finished games restarted from THESE spawns
    term = renv.is_terminal(state).cpu().numpy().astype(bool)
    assert term.any()
    state, *_ = renv.next_state(state, None, torch.zeros((64, 4), dtype=torch.int8), out=state)
    obs = renv.state_to_observation(state, -1)
    N = 15
    step = np.array([-N, 1, N, -1])                                  # one forward move from the spawn, per direction
    assert (obs["heads"].cpu().numpy()[term] == (h + step[d])[None]).all()
    with pytest.raises(Exception):
        env.new_state(ring_offset=9)                     # no such ring on a 15 x 15 board


def test_ttt_compact_host_stepper():
    """host_stepper(compact=True): 1-byte records whose decode equals next_state's return values, same states."""
    from colosseumrl_b200.tictactoe import BatchedTicTacToe4PlayerEnv, BatchedTicTacToe2PlayerEnv
    for cls, cells in ((BatchedTicTacToe4PlayerEnv, 27), (BatchedTicTacToe2PlayerEnv, 9)):
        B = 700
        a_env, b_env = cls("", batch=B, seed=2, auto_reset=True), cls("", batch=B, seed=2, auto_reset=True)
        sa, _ = a_env.new_state()
        sb, _ = b_env.new_state()
        stepper = b_env.host_stepper(sb, compact=True)                      # warm-up applies one cell-0 step
        sa, *_ = a_env.next_state(sa, None, torch.zeros((B,), dtype=torch.int8))
        rng = np.random.RandomState(3)
        terminals = wins = 0
        for t in range(60):
            a = rng.randint(-1, cells + 1, size=B).astype(np.int8)           # incl. passes, occupied and out-of-range cells
            sa, pa, ra, ta, wa = a_env.next_state(sa, None, torch.from_numpy(a))
            stepper.actions_np[...] = a
            rec = stepper()
            assert rec.shape == (B, 1)
            players, reward, terminal, winners = b_env.decode_compact(rec)
            assert (players == pa.cpu().numpy()).all() and (reward == ra.cpu().numpy()).all()
            assert (terminal == ta.cpu().numpy()).all() and (winners == wa.cpu().numpy()).all()
            terminals += int(terminal.sum()); wins += int((winners != 0).sum())
        assert terminals > 0 and wins > 0 and (sa.packed == sb.packed).all()


def test_blokus_host_stepper():
    """env.host_stepper (BlokusHostStepper: legal -> counts on the host -> the policy's index -> pick + step -> records)
    plays the same games as valid_actions + next_state with the same choices; indices past the list are passes."""
    from colosseumrl_b200.blokus import BatchedBlokusEnvironment
    B = 96
    a_env = BatchedBlokusEnvironment("", batch=B, seed=6, auto_reset=True)
    b_env = BatchedBlokusEnvironment("", batch=B, seed=6, auto_reset=True)
    sa, pa = a_env.new_state()
    sb, _ = b_env.new_state()
    stepper = b_env.host_stepper(sb)
    rng = np.random.RandomState(8)
    terminals = 0
    for t in range(90):
        counts_a, ids_a = a_env.valid_actions(sa)
        counts = stepper.legal().copy()
        assert (counts == counts_a.cpu().numpy()).all(), t
        choice = np.where(counts > 0, rng.randint(0, 1 << 30, size=B) % np.maximum(counts, 1), -1).astype(np.int32)
        choice[t % B] = counts[t % B] + 3                      # past the end of the list: a pass
        ids = ids_a.cpu().numpy()
        acts = np.where((choice >= 0) & (choice < counts), ids[np.arange(B), np.clip(choice, 0, ids.shape[1] - 1)], -1).astype(np.int32)
        sa, pa, ra, ta, wa = a_env.next_state(sa, pa, torch.from_numpy(acts))
        stepper.choice_np[...] = choice
        rec = stepper.step()
        assert (rec == sa.result.cpu().numpy()).all(), t
        terminals += int(ta.sum())
    assert terminals > 0 and (sa.packed == sb.packed).all()
