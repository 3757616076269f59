"""Single-environment string-action adapters (colosseumrl_b200/single.py) replaying the reference's recorded games
through the reference's own call shapes: lists of players, action STRINGS, numpy states."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRON_ACTION = {0: "forward", 1: "right", -1: "left"}


def test_tron_episodes_through_strings():
    from colosseumrl_b200.single import TronGridEnvironment
    for name, config in (("tron_N19_P4", ""), ("tron_N7_P3", "7;3"), ("tron_N8_P2", "8;2"),
                         ("tron_N21_P4", "21;4"), ("tron_N11_P6", "11;6")):      # the last two: the wide path
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        P = int(g["P"])
        env = TronGridEnvironment(config)
        assert env.min_players == env.max_players == P and env.valid_actions(None, 0) == ["forward", "right", "left"]
        rows = np.flatnonzero(g["env"] < 3)                      # the first three recorded episodes
        state = None
        for i in rows:
            if g["t"][i] == 0:
                state, players = env.new_state()
                assert (state[0] == g["start_board"]).all() and (state[1] == g["start_heads"]).all()
                assert (state[2] == g["start_directions"]).all() and list(players) == list(range(P))
            acts = [TRON_ACTION[int(a)] for a in g["actions"][i][players]]
            new, players, rewards, terminal, winners = env.next_state(state, list(players), acts)
            assert state is not new and (new[0] == g["board"][i]).all() and (new[1] == g["heads"][i]).all()
            assert (new[2] == g["directions"][i]).all() and (new[3] == g["deaths"][i]).all()
            assert (rewards == g["rewards"][i]).all() and terminal == bool(g["terminal"][i])
            assert sum(1 << int(p) for p in players) == g["alive"][i]
            assert (sum(1 << int(p) for p in winners) if winners is not None else 0) == g["winners"][i]
            rk = env.compute_ranking(new, list(range(P)), winners)
            assert [rk[p] for p in range(P)] == list(g["ranking"][i])
            state = new
        with pytest.raises(KeyError):
            env.next_state(state, [0], ["sideways"])
        # observations of sampled states
        for j, i in enumerate(g["obs_idx"][:6]):
            st = (g["board"][i].astype(np.int64), g["heads"][i], g["directions"][i], g["deaths"][i])
            for p in range(P):
                ob = env.state_to_observation(st, p)
                assert (ob["board"] == g["obs_board"][j * P + p]).all() and (ob["heads"] == g["obs_heads"][j * P + p]).all()
                assert (ob["directions"] == g["obs_directions"][j * P + p]).all()
                assert (ob["deaths"] == g["obs_deaths"][j * P + p]).all()
        blob = env.serialize_state(state)
        back = env.deserialize_state(blob)
        assert all((a == b).all() for a, b in zip(state, back))


def test_blokus_game_through_strings():
    from colosseumrl_b200.single import BlokusEnvironment
    from colosseumrl_b200.blokus import action_to_string
    g = np.load(os.path.join(GOLDEN, "blokus_games.npz"))
    env = BlokusEnvironment()
    rows = np.flatnonzero(g["game"] == 0)
    state, players = env.new_state()
    assert players == [0] and env.valid_actions(state, 0)[0] == "monomino1;(0, 0);north0"
    for i in rows:
        mover = int(g["mover"][i])
        assert players == [mover]
        if i % 4 == 0 or i >= rows[-1] - 8:                      # ordered valid-action STRINGS (subset of steps: ~ms each)
            exp = [action_to_string(int(a)) for a in g["valid_flat"][g["valid_off"][i]:g["valid_off"][i + 1]]] or [""]
            assert env.valid_actions(state, mover) == exp
        action = action_to_string(int(g["action"][i]))
        assert env.is_valid_action(state, mover, action) == (action != "")     # '' is never "valid" (reference :702)
        new, players, rewards, terminal, winners = env.next_state(state, [mover], [action])
        assert (new[0].board_contents == g["board"][i]).all() and new[1] == g["round"][i]
        assert [pl.player_score for pl in new[2]] == list(g["scores"][i])
        inv = np.array([[n in pl.current_pieces for n in env.all_piece_types()] for pl in new[2]])
        assert (inv == g["inventory"][i].astype(bool)).all()
        assert rewards == [int(g["reward"][i])] and terminal == bool(g["terminal"][i]) and players == [int(g["next_mover"][i])]
        assert (sum(1 << w for w in winners) if winners is not None else 0) == g["winners"][i]
        state = new
    assert terminal and env.compute_ranking(state, [0, 1, 2, 3], winners) == {p: (0 if p in winners else 1) for p in range(4)}
    assert not env.is_valid_action(env.new_state()[0], 0, "pentominoe6;(10, 10);east0")
    for j in np.flatnonzero(np.isin(g["obs_idx"], rows))[:8]:
        i, p = g["obs_idx"][j], int(g["obs_player"][j])
        st = (type(state[0])(g["board"][i]), int(g["round"][i]),
              [type(state[2][0])(q + 1, [n for k, n in enumerate(env.all_piece_types()) if g["inventory"][i][q, k]], g["scores"][i][q])
               for q in range(4)])
        ob = env.state_to_observation(st, p)
        assert (ob["board"] == g["obs_board"][j]).all() and (ob["pieces"] == g["obs_pieces"][j]).all()
        assert (ob["score"] == g["obs_score"][j]).all() and ob["player"][0] == p


@pytest.mark.parametrize("n", [2, 3, 4])
def test_tictactoe_through_strings(n):
    from colosseumrl_b200 import single
    env = {2: single.TicTacToe2PlayerEnv, 3: single.TicTacToe3PlayerEnv, 4: single.TicTacToe4PlayerEnv}[n]()
    g = np.load(os.path.join(GOLDEN, "ttt_%dp.npz" % n))
    shape = env._shape
    rows = np.flatnonzero(g["env"] < 6)
    state = None
    for i in rows:
        if g["t"][i] == 0:
            state, players = env.new_state()
        p, a = int(g["player"][i]), int(g["action"][i])
        action = "" if a < 0 else str(tuple(int(x) for x in np.unravel_index(a, shape)))
        valid = env.valid_actions(state, p)
        exp = [str(tuple(int(x) for x in np.unravel_index(c, shape))) for c in np.flatnonzero(g["valid_before"][i])] or [""]
        assert valid == exp
        new, players, rewards, terminal, winners = env.next_state(state, [p], [action])
        assert (new[0].reshape(-1) == g["board"][i]).all() and (-1 if new[1] is None else new[1]) == g["winner"][i]
        assert players == [int(g["next_player"][i])] and rewards == [int(g["reward"][i])] and terminal == bool(g["terminal"][i])
        assert (winners[0] if winners else -1) == g["winners"][i]
        for q in range(n):
            assert (env.state_to_observation(new, q)["board"].reshape(-1) == g["obs"][i, q]).all()
        state = new
    assert single.get_environment("tictactoe_4p") is single.TicTacToe4PlayerEnv


def test_blokus_perspective_valid_actions():
    from colosseumrl_b200.single import BlokusEnvironment, Board, AI
    from colosseumrl_b200.blokus import action_to_string
    g = np.load(os.path.join(GOLDEN, "blokus_games.npz"))
    gp = np.load(os.path.join(GOLDEN, "blokus_perspective.npz"))
    env = BlokusEnvironment()
    names = env.all_piece_types()
    for j, i in enumerate(gp["pos_idx"]):
        st = (Board(g["board"][i]), int(g["round"][i]),
              [AI(q + 1, [n for k, n in enumerate(names) if g["inventory"][i][q, k]], g["scores"][i][q]) for q in range(4)])
        p = int(gp["pos_player"][j])
        exp = [action_to_string(int(a)) for a in gp["pos_flat"][gp["pos_off"][j]:gp["pos_off"][j + 1]]]
        got = env.player_perspective_valid_actions(st, p)
        assert got == exp
        assert env.convert_player_perspective_action_to_real_action(got[0], p) == env.valid_actions(st, p)[0]


def test_vector_env():
    import torch
    from colosseumrl_b200 import BatchedTronGridEnvironment, BatchedBlokusEnvironment, BatchedTicTacToe4PlayerEnv
    from colosseumrl_b200.vector import VectorEnv
    # Tron: observation of every seat, rewards [B, P]
    env = BatchedTronGridEnvironment("", batch=64, auto_reset=True, seed=3)
    v = VectorEnv(env)
    obs = v.reset()
    assert obs["board"].shape == (64, 4, 19, 19) and obs["heads"].shape == (64, 4, 4)
    done_seen = 0
    for t in range(40):
        obs, rewards, dones, info = v.step(env.random_actions(t))
        assert rewards.shape == (64, 4) and dones.shape == (64,)
        for p in range(4):
            ref = env.state_to_observation(v.state, p)
            assert (obs["board"][:, p] == ref["board"]).all() and (obs["deaths"][:, p] == ref["deaths"]).all()
        done_seen += int(dones.sum())
    assert done_seen > 0
    # Blokus: every game seen from its own mover's seat
    env = BatchedBlokusEnvironment("", batch=32, auto_reset=True, seed=1)
    v = VectorEnv(env)
    obs = v.reset()
    for t in range(12):
        valid = v.valid_actions()
        obs, rewards, dones, info = v.step(env.random_actions(valid, t))
        mover = info["mover"].to(torch.int64)
        assert (mover == (t + 1) % 4).all() and (obs["player"][:, 0] == mover).all()
        per = [env.state_to_observation(v.state, p) for p in range(4)]
        for k in ("board", "pieces", "score"):
            exp = torch.stack([per[int(mover[g])][k][g] for g in range(32)])
            assert (obs[k] == exp).all(), k
    # Tic Tac Toe 4p
    env = BatchedTicTacToe4PlayerEnv("", batch=128, auto_reset=True, seed=2)
    v = VectorEnv(env)
    obs = v.reset()
    for t in range(30):
        obs, rewards, dones, info = v.step(env.random_actions(v.state, t))
        mover = info["mover"].to(torch.int64)
        per = [env.state_to_observation(v.state, p)["board"] for p in range(4)]
        exp = torch.stack([per[int(mover[g])][g] for g in range(128)])
        assert (obs["board"] == exp).all()


def test_vector_env_episode_boundary():
    """auto_reset: the step that ends an episode already returns the FRESH game's observation, mover and valid actions
    (what the next action is applied to), and the finished board is available as info["final_observation"]."""
    import torch
    from colosseumrl_b200 import BatchedTronGridEnvironment, BatchedBlokusEnvironment, BatchedTicTacToe4PlayerEnv
    from colosseumrl_b200.vector import VectorEnv
    # Tic Tac Toe 4p: a finished game shows the empty board, mover 0 and all 27 cells valid; a full-board draw does not
    # leave a zero mask behind (the agent is never forced to pass on a fresh board)
    env = BatchedTicTacToe4PlayerEnv("", batch=256, auto_reset=True, seed=5)
    v = VectorEnv(env)
    v.reset()
    ended = 0
    for t in range(60):
        act = env.random_actions(v.state, t)
        prev_valid = v.valid_actions().clone()
        assert (prev_valid != 0).all()                       # never a forced pass
        assert (((prev_valid >> act.to(torch.int32)) & 1) == 1).all()      # the policy's cell is free on the board we showed
        obs, rewards, dones, info = v.step(act, final_observation=True)
        d = dones.bool()
        if d.any():
            ended += int(d.sum())
            assert (obs["board"][d] == -1).all() and (info["mover"][d] == 0).all()
            assert (v.valid_actions()[d] == (1 << 27) - 1).all()
            assert (info["final_observation"]["board"][d] != -1).any(dim=(1, 2, 3)).all()
        assert not env.is_terminal(v.state).any()
    assert ended > 100
    # Blokus: after the terminal step the mover's list is the fresh board's 116 openings
    env = BatchedBlokusEnvironment("", batch=16, auto_reset=True, seed=2)
    v = VectorEnv(env)
    v.reset()
    ended = 0
    for t in range(90):
        valid = v.valid_actions()
        obs, rewards, dones, info = v.step(env.random_actions(valid, t))
        d = dones.bool()
        if d.any():
            ended += int(d.sum())
            assert (obs["board"][d] == -1).all() and (info["mover"][d] == 0).all()
            counts, _ = v.valid_actions()
            assert (counts[d] == 116).all()
    assert ended >= 8
    # Tron: every seat sees the start position again
    env = BatchedTronGridEnvironment("", batch=64, auto_reset=True, seed=4)
    v = VectorEnv(env)
    first = v.reset()
    ended = 0
    for t in range(40):
        obs, rewards, dones, info = v.step(env.random_actions(t))
        d = dones.bool()
        if d.any():
            ended += int(d.sum())
            assert (obs["board"][d] == first["board"][d]).all() and (obs["deaths"][d] == 0).all()
            assert (info["players"][d] == 15).all()
    assert ended > 0


def test_blokus_valid_actions_dict_and_current_rewards():
    """valid_actions_dict (BlokusEnvironment.py:630-665) against dictionaries recorded from the real reference
    (oracle/make_golden_dict.py): same keys, same insertion order, same orientation lists; TTT current_rewards."""
    import json
    from colosseumrl_b200.single import BlokusEnvironment, TicTacToe2PlayerEnv, TicTacToe3PlayerEnv, TicTacToe4PlayerEnv
    from colosseumrl_b200.blokus import action_to_string
    gold = json.load(open(os.path.join(GOLDEN, "blokus_valid_dict.json")))
    g = np.load(os.path.join(GOLDEN, "blokus_games.npz"))
    env = BlokusEnvironment()
    want = {}
    for rec in gold["blokus"]:
        want.setdefault(rec["row"], []).append(rec)
    state, players = env.new_state()
    seen = 0
    for i in np.flatnonzero(g["game"] == 0):
        for rec in want.get(int(i), []):
            d = env.valid_actions_dict(state, rec["player"])
            got = [[piece, [[k[0], k[1], v] for k, v in idx.items()]] for piece, idx in d.items()]
            assert got == rec["dict"], (i, rec["player"])
            assert all(isinstance(k, tuple) for idx in d.values() for k in idx)
            seen += 1
        state, players, *_ = env.next_state(state, [int(g["mover"][i])], [action_to_string(int(g["action"][i]))])
    assert seen == len(gold["blokus"]) >= 10
    assert env.current_rewards(state) == [float(s) for s in g["scores"][np.flatnonzero(g["game"] == 0)[-1]]]
    envs = {2: TicTacToe2PlayerEnv(), 3: TicTacToe3PlayerEnv(), 4: TicTacToe4PlayerEnv()}
    for rec in gold["ttt_current_rewards"]:
        e = envs[rec["n"]]
        st, _ = e.new_state()
        assert e.current_rewards((st[0], rec["winner"])) == rec["rewards"]
        b = e._b.current_rewards(e._pack((st[0], rec["winner"]), 0)).cpu().numpy()
        assert b[0].tolist() == rec["rewards"]


def test_tron_create():
    from colosseumrl_b200.single import TronGridEnvironment
    from colosseumrl_b200.tron import BatchedTronGridEnvironment, create_tron_config, parse_tron_config
    assert create_tron_config(15, 3, -1, False) == "15;3;-1;False"
    assert list(parse_tron_config("15;3;-1;False")) == [15, 3, -1, False]
    env = TronGridEnvironment.create(board_size=11, num_players=3)
    assert env.N == 11 and env.num_players == 3 and env.observation_shape["board"] == (11, 11)
    benv = BatchedTronGridEnvironment.create(board_size=9, num_players=2, batch=4)
    assert benv.N == 9 and benv.num_players == 2 and "9x9" in repr(benv)


def test_adapter_continues_a_reference_serialized_game():
    """The bytes an untouched match server pushes (tests/golden/wire_blokus_step20.dill = the reference's own
    dill.dumps after 20 recorded steps, oracle/make_golden_wire.py) -> deserialize_state -> the engine plays the next
    recorded steps: valid_actions lists and next_state outputs equal the reference's recording; and our own
    serialize_state round-trips mid-game."""
    from colosseumrl_b200.single import BlokusEnvironment
    g = np.load(os.path.join(GOLDEN, "blokus_games.npz"))
    env = BlokusEnvironment()
    state = env.deserialize_state(open(os.path.join(GOLDEN, "wire_blokus_step20.dill"), "rb").read())
    players = [int(g["mover"][20])]
    for i in range(20, 32):
        exp = g["valid_flat"][g["valid_off"][i]:g["valid_off"][i + 1]]
        assert env.valid_actions(state, players[0]) == ([orc.blokus_action_to_string(int(a)) for a in exp] or [""])
        state, players, rewards, terminal, winners = env.next_state(state, players, [orc.blokus_action_to_string(int(g["action"][i]))])
        assert (state[0].board_contents == g["board"][i]).all() and players[0] == g["next_mover"][i]
        assert [p.player_score for p in state[2]] == g["scores"][i].tolist() and rewards[0] == g["reward"][i]
        if i == 25:
            state = env.deserialize_state(env.serialize_state(state))
