"""TEST INFRASTRUCTURE ONLY -- reference-recorded hand-built Tic Tac Toe states (tests/golden/ttt_states.npz).

Run in the build container (where /root/reference exists):   python oracle/make_golden_ttt_states.py

The trajectories of oracle/make_golden.py only visit reachable positions.  Here the unmodified reference
(`TicTacToe{2,3,4}PlayerEnv.next_state`, `valid_actions`, `state_to_observation`) is run on ARBITRARY states: random
boards of any fill (several completed lines, full boards), a carried winner that may contradict the board, any mover,
and actions on free cells, on occupied cells (a silent no-op, tictactoe_2p_env.py:293) and '' (pass).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main(seed=23, per_variant=500):
    R = ref_shim.load()
    rng = np.random.RandomState(seed)
    out = {}
    for n in (2, 3, 4):
        env = R["TicTacToe%dPlayerEnv" % n]()
        shape = env.observation_shape["board"]
        cells = int(np.prod(shape))
        rec = {k: [] for k in ("board", "winner", "mover", "action", "n_valid", "o_board", "o_winner", "next_player", "reward",
                               "terminal", "winners", "viewer", "obs")}
        for _ in range(per_variant):
            fill = rng.choice([0.0, 0.2, 0.5, 0.8, 1.0]) if rng.rand() < 0.5 else rng.rand()
            board = np.where(rng.rand(cells) < fill, rng.randint(0, n, size=cells), -1).astype(np.int8).reshape(shape)
            winner = int(rng.randint(0, n)) if rng.rand() < 0.25 else None
            mover = int(rng.randint(0, n))
            free = np.flatnonzero(board.ravel() == -1)
            kind = rng.rand()
            if kind < 0.15 or (len(free) == 0 and kind < 0.6):
                action = -1
            elif kind < 0.4:
                action = int(rng.randint(cells))                  # any cell, often occupied
            else:
                action = int(free[rng.randint(len(free))]) if len(free) else int(rng.randint(cells))
            s = "" if action < 0 else str(tuple(int(v) for v in np.unravel_index(action, shape)))
            va = env.valid_actions((board, winner), mover)
            nst, players, rewards, terminal, winners = env.next_state((board, winner), [mover], [s])
            viewer = int(rng.randint(0, n))
            rec["board"].append(board.ravel().copy()); rec["winner"].append(-1 if winner is None else winner)
            rec["mover"].append(mover); rec["action"].append(action)
            rec["n_valid"].append(0 if va == [""] else len(va))
            rec["o_board"].append(nst[0].ravel().copy()); rec["o_winner"].append(-1 if nst[1] is None else nst[1])
            rec["next_player"].append(players[0]); rec["reward"].append(rewards[0]); rec["terminal"].append(bool(terminal))
            rec["winners"].append(-1 if winners is None else winners[0]); rec["viewer"].append(viewer)
            rec["obs"].append(env.state_to_observation(nst, viewer)["board"].ravel().copy())
        for k, v in rec.items():
            out["p%d_%s" % (n, k)] = np.asarray(v).astype(np.int8 if k in ("board", "o_board", "obs") else np.int64)
        print("ttt", n, "states", per_variant, "terminal", int(np.sum(rec["terminal"])), "carried winners", int(np.sum(np.asarray(rec["winner"]) >= 0)))
    np.savez_compressed(os.path.join(OUT, "ttt_states.npz"), **out)
    print("file", os.path.getsize(os.path.join(OUT, "ttt_states.npz")), "bytes")


if __name__ == "__main__":
    main()
