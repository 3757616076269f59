"""TEST INFRASTRUCTURE ONLY -- reference-recorded ARBITRARY Blokus positions (tests/golden/blokus_boards.npz).

Run in the build container (where /root/reference exists):   python oracle/make_golden_blokus_boards.py [n]

Hand-built, not necessarily reachable positions -- random colour blobs and single cells on the 20x20 board, random
inventories, rounds 0..3 (round 0 with free AND occupied start corners), any mover -- on which the unmodified reference
computes `valid_actions` for all four seats (recorded as length + order-sensitive hash, oracle/make_golden_wide.py's
`list_hash`) and one `next_state` of the mover (a random entry of its list, or a pass).  Pins the oracle's legality and
ordering on geometries random self-play never produces: anchors on every edge, blocked corners, crowded boards.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.oracle import PIECE_NAMES, blokus_action_to_string, blokus_string_to_action  # noqa: E402
from oracle.make_golden_wide import list_hash, state_arrays, winners_mask  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    R = ref_shim.load()
    env = R["BlokusEnvironment"]()
    Board, AI = R["blokus_board"].Board, R["blokus_ai"].AI
    rng = np.random.RandomState(29)
    rec = {k: [] for k in ("i_board", "i_round", "i_inventory", "i_scores", "mover", "n_valid", "valid_hash", "action",
                           "board", "round", "inventory", "scores", "reward", "terminal", "winners", "next_mover")}
    for i in range(n):
        cells = np.zeros((20, 20), np.int64)
        density = rng.uniform(0.0, 0.7)
        for _ in range(int(density * 60)):
            c = rng.randint(1, 5)
            y, x = rng.randint(0, 20, size=2)
            h, w = rng.randint(1, 4, size=2) if rng.rand() < 0.5 else (1, 1)
            blk = cells[y:y + h, x:x + w]
            cells[y:y + h, x:x + w] = np.where(blk == 0, c, blk)
        if i % 9 == 0:
            cells[:] = 0
        board = Board()
        board.board_contents = cells
        players = []
        held = rng.uniform(0.1, 1.0)
        for q in range(4):
            p = AI(board, q + 1)
            p.current_pieces = [name for name in PIECE_NAMES if rng.rand() < held]
            p.player_score = int(rng.randint(0, 90))
            players.append(p)
        round_count = int(rng.randint(0, 4))
        mover = int(rng.randint(0, 4))
        state = (board, round_count, players)
        ib, irc, iinv, isc = state_arrays(state)
        nv, hs, lists = [], [], []
        for q in range(4):
            va = env.valid_actions(state, q)
            ids = [blokus_string_to_action(s) for s in va if s != ""]
            assert all(blokus_action_to_string(a) == s for a, s in zip(ids, va))
            nv.append(len(ids)); hs.append(list_hash(ids)); lists.append(ids)
        ids = lists[mover]
        action = ids[int(rng.randint(len(ids)))] if ids and rng.rand() < 0.85 else -1
        nst, nplayers, rewards, terminal, winners = env.next_state(state, [mover], [blokus_action_to_string(action)])
        b, rc, inv, sc = state_arrays(nst)
        rec["i_board"].append(ib); rec["i_round"].append(irc); rec["i_inventory"].append(iinv); rec["i_scores"].append(isc)
        rec["mover"].append(mover); rec["n_valid"].append(nv); rec["valid_hash"].append(hs); rec["action"].append(action)
        rec["board"].append(b); rec["round"].append(rc); rec["inventory"].append(inv); rec["scores"].append(sc)
        rec["reward"].append(rewards[0]); rec["terminal"].append(bool(terminal)); rec["winners"].append(winners_mask(winners))
        rec["next_mover"].append(nplayers[0])
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["valid_hash"] = out["valid_hash"].astype(np.uint64)
    np.savez_compressed(os.path.join(OUT, "blokus_boards.npz"), **out)
    print("positions", n, "max list", int(out["n_valid"].max()), "terminal", int(out["terminal"].sum()),
          "empty lists", int((out["n_valid"] == 0).sum()), "file", os.path.getsize(os.path.join(OUT, "blokus_boards.npz")))


if __name__ == "__main__":
    main()
