"""TEST INFRASTRUCTURE ONLY -- reference-recorded hand-built Tron states beyond N <= 19 / P <= 4
(tests/golden/tron_adversarial_wide.npz).

Run in the build container (where /root/reference exists):   python oracle/make_golden_tron_wide.py

Same recipe as `tron_adversarial` in oracle/make_golden.py (random trails, heads crowded around one cell, random death
codes incl. self-kills and mutual pairs, random actions), on shapes only the wide path of the engine runs
(csrc/tron_wide.cuh): 5..8 players and / or boards larger than 19x19.  The unmodified reference
(`TronGridEnvironment.next_state` -> its own compiled `CyTronGrid.next_state_inplace`, `compute_ranking` of the input and
of the output state) records every case; arrays are padded to 8 players / 21x21 cells.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
NMAX, PMAX = 21, 8
SHAPES = [(7, 5), (9, 6), (9, 8), (11, 7), (20, 2), (21, 4), (21, 8)]


def main(seed=17, per_shape=140):
    R = ref_shim.load()
    rng = np.random.RandomState(seed)
    names = {0: "forward", 1: "right", -1: "left"}
    keys = ("N", "P", "board", "heads", "directions", "deaths", "actions", "o_board", "o_heads", "o_directions", "o_deaths",
            "o_alive", "o_rewards", "o_terminal", "o_winners", "o_ranking", "i_ranking")
    rec = {k: [] for k in keys}
    pad = lambda a, v=0: np.concatenate([np.asarray(a, np.int64), np.full(PMAX - len(a), v, np.int64)])
    for N, P in SHAPES:
        env = R["TronGridEnvironment"]("{};{}".format(N, P))
        for _ in range(per_shape):
            board = np.zeros((N, N), np.int64)
            fill = rng.rand(N, N) < rng.uniform(0.05, 0.6)
            board[fill] = rng.randint(1, P + 1, size=int(fill.sum()))
            heads = rng.choice(N * N, size=P, replace=False).astype(np.int64)
            if rng.rand() < 0.6:
                c = int(rng.randint(1, N - 1)) * N + int(rng.randint(1, N - 1))
                neigh = [c - 1, c + 1, c - N, c + N, c, c - N - 1, c - N + 1, c + N - 1, c + N + 1]
                rng.shuffle(neigh)
                cand = np.asarray((neigh + list(heads))[:P], np.int64)
                if len(set(cand.tolist())) == P:
                    heads = cand
            board.ravel()[heads] = np.arange(1, P + 1)
            directions = rng.randint(0, 4, size=P).astype(np.int64)
            deaths = np.where(rng.rand(P) < 0.35, rng.randint(1, P + 1, size=P), 0).astype(np.int64)
            actions = rng.randint(-1, 2, size=P)
            i_rank = env.compute_ranking((board, heads, directions, deaths), list(range(P)), None)
            st, players, rewards, terminal, winners = env.next_state(
                (board, heads, directions, deaths), list(range(P)), [names[int(a)] for a in actions])
            o_rank = env.compute_ranking(st, list(range(P)), winners)
            bpad = np.zeros((NMAX, NMAX), np.int8); bpad[:N, :N] = board
            opad = np.zeros((NMAX, NMAX), np.int8); opad[:N, :N] = st[0]
            rec["N"].append(N); rec["P"].append(P); rec["board"].append(bpad); rec["heads"].append(pad(heads))
            rec["directions"].append(pad(directions)); rec["deaths"].append(pad(deaths)); rec["actions"].append(pad(actions))
            rec["o_board"].append(opad); rec["o_heads"].append(pad(st[1])); rec["o_directions"].append(pad(st[2]))
            rec["o_deaths"].append(pad(st[3])); rec["o_alive"].append(sum(1 << int(p) for p in players))
            rec["o_rewards"].append(pad(rewards)); rec["o_terminal"].append(bool(terminal))
            rec["o_winners"].append(sum(1 << int(p) for p in winners) if winners is not None else 0)
            rec["o_ranking"].append(pad([int(o_rank.get(p, -1)) for p in range(P)], -1))
            rec["i_ranking"].append(pad([int(i_rank.get(p, -1)) for p in range(P)], -1))
    out = {k: np.asarray(v) for k, v in rec.items()}
    np.savez_compressed(os.path.join(OUT, "tron_adversarial_wide.npz"), **out)
    print("cases", len(out["N"]), "terminal", int(out["o_terminal"].sum()), "file",
          os.path.getsize(os.path.join(OUT, "tron_adversarial_wide.npz")), "bytes")


if __name__ == "__main__":
    main()
