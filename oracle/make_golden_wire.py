"""TEST INFRASTRUCTURE ONLY -- record the reference's OWN serialized states (tests/golden/wire_*.dill).

Run in the build container (where /root/reference exists):   python oracle/make_golden_wire.py

The unmodified reference (through oracle/ref_shim.py) replays the first recorded Blokus game of
tests/golden/blokus_games.npz up to step 20 and writes `BlokusEnvironment.serialize_state(state)` (= dill.dumps,
BlokusEnvironment.py:305-320) to tests/golden/wire_blokus_step20.dill; the same for a Tic Tac Toe 2-player state after
four moves (tictactoe_2p_env.py:185-200) and a Tron state after five steps (TronGridEnvironment.py:434-447).  These
are the byte streams an untouched match server sends (match_server.py:206-207); tests/test_wire_format.py and
tests/test_gpu_single.py check that colosseumrl_b200 reads them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.oracle import blokus_action_to_string  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
BLOKUS_STEPS = 20


def main():
    R = ref_shim.load()
    g = np.load(os.path.join(OUT, "blokus_games.npz"))
    env = R["BlokusEnvironment"]()
    state, players = env.new_state()
    for i in range(BLOKUS_STEPS):
        assert g["game"][i] == 0 and g["t"][i] == i and players[0] == g["mover"][i]
        state, players, _, terminal, _ = env.next_state(state, players, [blokus_action_to_string(int(g["action"][i]))])
        assert not terminal and (state[0].board_contents == g["board"][i]).all()
    open(os.path.join(OUT, "wire_blokus_step%d.dill" % BLOKUS_STEPS), "wb").write(bytes(env.serialize_state(state)))

    t2 = R["TicTacToe2PlayerEnv"]()
    st, pl = t2.new_state()
    for a in ("(1, 1)", "(0, 0)", "(2, 1)", "(0, 2)"):
        st, pl, _, _, _ = t2.next_state(st, pl, [a])
    open(os.path.join(OUT, "wire_ttt2_step4.dill"), "wb").write(bytes(t2.serialize_state(st)))

    tr = R["TronGridEnvironment"]("")
    st, pl = tr.new_state()
    for a in (["forward"] * 4, ["left", "right", "forward", "forward"], ["forward"] * 4, ["right"] * 4, ["forward"] * 4):
        st, pl, _, _, _ = tr.next_state(st, list(range(4)), a)
    open(os.path.join(OUT, "wire_tron_step5.dill"), "wb").write(bytes(tr.serialize_state(st)))
    for f in ("wire_blokus_step20.dill", "wire_ttt2_step4.dill", "wire_tron_step5.dill"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
