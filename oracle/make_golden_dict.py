"""TEST INFRASTRUCTURE ONLY -- tests/golden/blokus_valid_dict.json from the REAL reference: BlokusEnvironment.
valid_actions_dict (:630-665) at a few positions of recorded game 0 (tests/golden/blokus_games.npz is replayed through
the reference's own next_state), for the mover and one other seat, plus TicTacToe current_rewards (2p :219-238).

    python oracle/make_golden_dict.py        (build container only: needs /root/reference)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.oracle import blokus_action_to_string  # noqa: E402


def main():
    R = ref_shim.load()
    g = np.load(os.path.join(ROOT, "tests", "golden", "blokus_games.npz"))
    env = R["BlokusEnvironment"]()
    rows = np.flatnonzero(g["game"] == 0)
    state, players = env.new_state()
    out = []
    for i in rows:
        mover = int(g["mover"][i])
        if int(g["t"][i]) in (0, 1, 5, 14, 27, 41, 55):
            for player in (mover, (mover + 2) % 4):
                d = env.valid_actions_dict(state, player)
                out.append({"row": int(i), "player": player,
                            "dict": [[piece, [[int(k[0]), int(k[1]), list(v)] for k, v in idx.items()]] for piece, idx in d.items()]})
        state, players, *_ = env.next_state(state, [mover], [blokus_action_to_string(int(g["action"][i]))])
    ttt = []
    for name, n in (("TicTacToe2PlayerEnv", 2), ("TicTacToe3PlayerEnv", 3), ("TicTacToe4PlayerEnv", 4)):
        e = R[name]()
        st, _ = e.new_state()
        for winner in [None] + list(range(n)):
            ttt.append({"n": n, "winner": winner, "rewards": [int(x) for x in e.current_rewards((st[0], winner))]})
    path = os.path.join(ROOT, "tests", "golden", "blokus_valid_dict.json")
    json.dump({"blokus": out, "ttt_current_rewards": ttt}, open(path, "w"))
    print("wrote", path, len(out), "dicts,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
