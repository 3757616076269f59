"""TEST INFRASTRUCTURE ONLY -- tests/golden/tron_starts.npz from the REAL reference: generate_start_positions
(TronGridEnvironment.py:183-226) for a grid of (N, P, ring_offset, spawn_offset), integer spawn offsets only (an int
offset makes np.random.randint(o, o + 1) deterministic, :222-224).  Combinations the reference itself cannot place
(IndexError / empty arcs) are recorded with ok = 0.

    python oracle/make_golden_starts.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402


def main():
    ref = ref_shim.load()
    Tron = ref["TronGridEnvironment"]
    rows = []
    for N in range(5, 22):
        for P in (2, 3, 4):
            env = Tron.create(board_size=N, num_players=P)
            for ring in (0, 1, 2, 3):
                for spawn in (-4, -2, -1, 0, 1, 2, 3, 5):
                    try:
                        heads, dirs = env.generate_start_positions(ring, spawn)
                        ok = int(len(heads) == P and len(set(int(h) for h in heads)) == P)
                        h = [int(x) for x in heads] + [0] * (4 - P)
                        d = [int(x) for x in dirs] + [0] * (4 - P)
                    except Exception:
                        ok, h, d = 0, [0] * 4, [0] * 4
                    rows.append([N, P, ring, spawn, ok] + h + d)
    a = np.array(rows, np.int32)
    out = os.path.join(ROOT, "tests", "golden", "tron_starts.npz")
    np.savez_compressed(out, table=a)
    print("wrote", out, a.shape, "placeable:", int(a[:, 4].sum()))


if __name__ == "__main__":
    main()
