"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the REAL reference.

Run in the build container (where /root/reference exists):

    make -C oracle ref          # compiles the reference's own CyTronGrid.pyx into oracle/_ref/
    python oracle/make_golden.py

The reference's unmodified Python classes (TronGridEnvironment, BlokusEnvironment,
TicTacToe{2,3,4}PlayerEnv) are driven through ``next_state`` / ``valid_actions`` /
``compute_ranking`` / ``state_to_observation`` with Philox4x32-10 action streams and every
input / output is recorded as plain integer arrays.  These files are the pin for the C oracle
(tests/test_oracle_golden.py) and are also compared directly with the CUDA engine on the GPU
(tests/cases_{tron,ttt,blokus}.py through tests/test_gpu_*.py).  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so this is the only pin available.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.oracle import (PIECE_NAMES, blokus_action_to_string, blokus_string_to_action)  # noqa: E402
from colosseumrl_b200.philox import env_step_words, TAG_TRON, TAG_BLOKUS, TAG_TTT  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TRON_MOVES = ["forward", "right", "left"]     # index r % 3 -> {0, +1, -1}
TRON_MOVE_VAL = [0, 1, -1]


# ----------------------------------------------------------------------------------- Tron
def tron_trajectories(R, N, P, seed, n_env, max_steps, extra_after_terminal):
    """Per environment: play from new_state with Philox actions; keep stepping `extra_after_terminal`
    steps past the terminal flag (reference semantics on finished games: T3 death overwrite etc.)."""
    env = R["TronGridEnvironment"]("{};{}".format(N, P))
    rec = {k: [] for k in ("env", "t", "actions", "board", "heads", "directions", "deaths", "alive",
                           "rewards", "terminal", "winners", "ranking")}
    start = env.new_state()[0]
    for e in range(n_env):
        state, players = env.new_state()
        after = 0
        for t in range(max_steps):
            r = env_step_words(seed, [e], t, TAG_TRON)[0]
            idx = [int(r[p] % 3) if p < 4 else int(r[p % 4] // 3 % 3) for p in range(P)]
            actions = [TRON_MOVES[i] for i in idx]
            # all players are passed every step (dense action vector; dead players' entries are ignored)
            state, players, rewards, terminal, winners = env.next_state(state, list(range(P)), actions)
            ranking = env.compute_ranking(state, list(range(P)), winners)
            rec["env"].append(e); rec["t"].append(t)
            rec["actions"].append([TRON_MOVE_VAL[i] for i in idx])
            rec["board"].append(state[0].astype(np.int8)); rec["heads"].append(state[1].copy())
            rec["directions"].append(state[2].copy()); rec["deaths"].append(state[3].copy())
            rec["alive"].append(sum(1 << int(p) for p in players))
            rec["rewards"].append(np.asarray(rewards, np.int64))
            rec["terminal"].append(bool(terminal))
            rec["winners"].append(sum(1 << int(p) for p in winners) if winners is not None else 0)
            rec["ranking"].append([int(ranking.get(p, -1)) for p in range(P)])
            if terminal:
                after += 1
                if after > extra_after_terminal:
                    break
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["start_board"] = start[0].astype(np.int8)
    out["start_heads"] = start[1]
    out["start_directions"] = start[2]
    out["N"] = np.int64(N); out["P"] = np.int64(P); out["seed"] = np.int64(seed)
    # observations for every player on a sample of states
    obs_idx = np.arange(0, len(rec["t"]), max(1, len(rec["t"]) // 24))
    ob = {k: [] for k in ("board", "heads", "directions", "deaths")}
    for i in obs_idx:
        st = (out["board"][i].astype(np.int64), out["heads"][i], out["directions"][i], out["deaths"][i])
        for p in range(P):
            o = env.state_to_observation(st, p)
            for k in ob:
                ob[k].append(np.asarray(o[k]))
    out["obs_idx"] = obs_idx
    for k in ob:
        out["obs_" + k] = np.asarray(ob[k]).astype(np.int64 if k != "board" else np.int8)
    return out


def tron_adversarial(R, seed=7, n=1500):
    """Hand-built (random, not reachable-only) states: exercises T1-T5 quirks of next_state and the
    compute_ranking wrap-around on arbitrary death vectors."""
    rng = np.random.RandomState(seed)
    rec = {k: [] for k in ("N", "P", "board", "heads", "directions", "deaths", "actions", "o_board", "o_heads",
                           "o_directions", "o_deaths", "o_alive", "o_rewards", "o_terminal", "o_winners", "o_ranking",
                           "i_ranking")}
    for _ in range(n):
        P = int(rng.choice([2, 3, 4]))
        N = int(rng.choice([5, 6, 7]))
        env = R["TronGridEnvironment"]("{};{}".format(N, P))
        board = np.zeros((N, N), np.int64)
        # random trails
        fill = rng.rand(N, N) < rng.uniform(0.1, 0.6)
        board[fill] = rng.randint(1, P + 1, size=int(fill.sum()))
        heads = rng.choice(N * N, size=P, replace=False).astype(np.int64)
        board.ravel()[heads] = np.arange(1, P + 1)
        # bias heads towards each other sometimes
        if rng.rand() < 0.5:
            c = int(rng.randint(1, N - 1)) * N + int(rng.randint(1, N - 1))
            neigh = [c - 1, c + 1, c - N, c + N, c]
            rng.shuffle(neigh)
            for p in range(P):
                board.ravel()[heads[p]] = rng.randint(0, P + 1)
            heads = np.asarray(neigh[:P], np.int64)
            board.ravel()[heads] = np.arange(1, P + 1)
            if rng.rand() < 0.5:
                board.ravel()[c] = 0 if c not in heads else board.ravel()[c]
        directions = rng.randint(0, 4, size=P).astype(np.int64)
        deaths = np.where(rng.rand(P) < 0.3, rng.randint(1, P + 1, size=P), 0).astype(np.int64)
        actions = rng.randint(-1, 2, size=P)
        names = {0: "forward", 1: "right", -1: "left"}
        pad = lambda a, L, v=0: np.concatenate([np.asarray(a, np.int64), np.full(L - len(a), v, np.int64)])
        bpad = np.zeros((7, 7), np.int8); bpad[:N, :N] = board
        i_rank = env.compute_ranking((board, heads, directions, deaths), list(range(P)), None)
        st, players, rewards, terminal, winners = env.next_state(
            (board, heads, directions, deaths), list(range(P)), [names[int(a)] for a in actions])
        o_rank = env.compute_ranking(st, list(range(P)), winners)
        obpad = np.zeros((7, 7), np.int8); obpad[:N, :N] = st[0]
        rec["N"].append(N); rec["P"].append(P); rec["board"].append(bpad); rec["heads"].append(pad(heads, 4))
        rec["directions"].append(pad(directions, 4)); rec["deaths"].append(pad(deaths, 4))
        rec["actions"].append(pad(actions, 4)); rec["o_board"].append(obpad); rec["o_heads"].append(pad(st[1], 4))
        rec["o_directions"].append(pad(st[2], 4)); rec["o_deaths"].append(pad(st[3], 4))
        rec["o_alive"].append(sum(1 << int(p) for p in players)); rec["o_rewards"].append(pad(rewards, 4))
        rec["o_terminal"].append(bool(terminal))
        rec["o_winners"].append(sum(1 << int(p) for p in winners) if winners is not None else 0)
        rec["o_ranking"].append(pad([int(o_rank.get(p, -1)) for p in range(P)], 4, -1))
        rec["i_ranking"].append(pad([int(i_rank.get(p, -1)) for p in range(P)], 4, -1))
    return {k: np.asarray(v) for k, v in rec.items()}


# ----------------------------------------------------------------------------------- TTT
def ttt_trajectories(R, n, seed, n_env, extra_after_terminal=3):
    env = R["TicTacToe%dPlayerEnv" % n]()
    shape = env.observation_shape["board"]
    cells = int(np.prod(shape))
    rec = {k: [] for k in ("env", "t", "player", "action", "board", "winner", "next_player", "reward", "terminal",
                           "winners", "valid_before", "obs")}
    for e in range(n_env):
        state, players = env.new_state()
        after = 0
        for t in range(200):
            r = env_step_words(seed, [e], t, TAG_TTT)[0]
            board = state[0]
            empty = np.flatnonzero(board.ravel() == -1)
            # reference's valid_actions (index set in C order; strings are unparsable under numpy 2 - SURVEY X1)
            va = env.valid_actions(state, players[0])
            assert (va == [""]) == (len(empty) == 0) and (va == [""] or len(va) == len(empty))
            mode = r[1] % 20
            if len(empty) == 0 or mode == 0:
                action = -1                       # '' pass (5 %)
            elif mode < 7:
                action = int(r[2] % cells)        # any cell, often occupied (30 %)
            else:
                action = int(empty[r[0] % len(empty)])
            s = "" if action < 0 else str(tuple(int(v) for v in np.unravel_index(action, shape)))
            vmask = np.zeros(cells, np.int8); vmask[empty] = 1
            player = players[0]
            state, players, rewards, terminal, winners = env.next_state(state, [player], [s])
            rec["env"].append(e); rec["t"].append(t); rec["player"].append(player); rec["action"].append(action)
            rec["board"].append(state[0].ravel().copy()); rec["winner"].append(-1 if state[1] is None else state[1])
            rec["next_player"].append(players[0]); rec["reward"].append(rewards[0]); rec["terminal"].append(bool(terminal))
            rec["winners"].append(-1 if winners is None else winners[0]); rec["valid_before"].append(vmask)
            rec["obs"].append([env.state_to_observation(state, p)["board"].ravel() for p in range(n)])
            if terminal:
                after += 1
                if after > extra_after_terminal:
                    break
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["board"] = out["board"].astype(np.int8); out["obs"] = out["obs"].astype(np.int8)
    out["n"] = np.int64(n); out["seed"] = np.int64(seed)
    return out


# ----------------------------------------------------------------------------------- Blokus
def blokus_state_arrays(state):
    board, round_count, players = state
    inv = np.zeros((4, 21), np.uint8)
    for i, p in enumerate(players):
        for name in p.current_pieces:
            inv[i, PIECE_NAMES.index(name)] = 1
    scores = np.asarray([p.player_score for p in players], np.int64)
    return board.board_contents.astype(np.int8), int(round_count), inv, scores


def blokus_games(R, seed, n_games, extra_after_terminal=2):
    env = R["BlokusEnvironment"]()
    rec = {k: [] for k in ("game", "t", "mover", "action", "n_valid", "board", "round", "inventory", "scores",
                           "reward", "terminal", "winners", "next_mover")}
    valid_flat, valid_off = [], [0]
    obs = {k: [] for k in ("idx", "player", "board", "pieces", "score")}
    for g in range(n_games):
        state, players = env.new_state()
        after = 0
        for t in range(200):
            mover = players[0]
            r = env_step_words(seed, [g], t, TAG_BLOKUS)[0]
            va = env.valid_actions(state, mover)
            ids = [blokus_string_to_action(s) for s in va if s != ""]
            assert all(blokus_action_to_string(i) == s for i, s in zip(ids, va))
            action = ids[int(r[0] % len(ids))] if ids else -1
            state, players, rewards, terminal, winners = env.next_state(state, [mover], [blokus_action_to_string(action)])
            b, rc, inv, sc = blokus_state_arrays(state)
            rec["game"].append(g); rec["t"].append(t); rec["mover"].append(mover); rec["action"].append(action)
            rec["n_valid"].append(len(ids)); rec["board"].append(b); rec["round"].append(rc); rec["inventory"].append(inv)
            rec["scores"].append(sc); rec["reward"].append(rewards[0]); rec["terminal"].append(bool(terminal))
            rec["winners"].append(sum(1 << int(p) for p in winners) if winners is not None else 0)
            rec["next_mover"].append(players[0])
            valid_flat.extend(ids); valid_off.append(len(valid_flat))
            if t % 9 == 4 or terminal:
                for p in range(4):
                    o = env.state_to_observation(state, p)
                    obs["idx"].append(len(rec["t"]) - 1); obs["player"].append(p)
                    obs["board"].append(np.asarray(o["board"]).astype(np.int8))
                    obs["pieces"].append(o["pieces"]); obs["score"].append(np.asarray(o["score"], np.int64))
            if terminal:
                after += 1
                if after > extra_after_terminal:
                    break
        print("  blokus game", g, "steps", t + 1, "scores", sc, flush=True)
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["valid_flat"] = np.asarray(valid_flat, np.int32)
    out["valid_off"] = np.asarray(valid_off, np.int64)
    for k, v in obs.items():
        out["obs_" + k] = np.asarray(v)
    out["seed"] = np.int64(seed)
    return out


def main():
    R = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    # Tron: default config + small boards (collision-dense) + other player counts + an even board + P > 4
    for (N, P, n_env, extra) in [(19, 4, 40, 2), (9, 4, 60, 6), (7, 3, 60, 6), (8, 2, 40, 4), (15, 4, 10, 1),
                                 (19, 2, 6, 1), (21, 4, 4, 1), (11, 6, 30, 4), (6, 4, 60, 8)]:
        d = tron_trajectories(R, N, P, seed=0, n_env=n_env, max_steps=400, extra_after_terminal=extra)
        np.savez_compressed(os.path.join(OUT, "tron_N%d_P%d.npz" % (N, P)), **d)
        print("tron", N, P, "steps", len(d["t"]), "episodes", int(d["terminal"].sum()))
    d = tron_adversarial(R)
    np.savez_compressed(os.path.join(OUT, "tron_adversarial.npz"), **d)
    print("tron adversarial", len(d["N"]))
    for n, n_env in [(2, 150), (3, 120), (4, 60)]:
        d = ttt_trajectories(R, n, seed=0, n_env=n_env)
        np.savez_compressed(os.path.join(OUT, "ttt_%dp.npz" % n), **d)
        print("ttt", n, "steps", len(d["t"]), "wins", int((d["winners"] >= 0).sum()))
    d = blokus_games(R, seed=0, n_games=8)
    np.savez_compressed(os.path.join(OUT, "blokus_games.npz"), **d)
    print("blokus steps", len(d["t"]), "max valid", int(d["n_valid"].max()))


if __name__ == "__main__":
    main()
