#!/usr/bin/env python
"""Small-batch invocation of every hot kernel for compute-sanitizer:

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_cases.py

Each case is checked against the CPU oracle as well, so a sanitizer run is also a parity run."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from colosseumrl_b200 import BatchedTronGridEnvironment, BatchedBlokusEnvironment, BatchedTicTacToe4PlayerEnv  # noqa: E402
from colosseumrl_b200 import BatchedTicTacToe2PlayerEnv, BatchedTicTacToe3PlayerEnv  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    seed = 7
    # Tron: step (ragged last tile: 1000 = 15 x 64 + 40), packed actions + compact record, rollout, observe (3 forms), ranking
    B, K = 1000, 10
    env = BatchedTronGridEnvironment("", batch=B, seed=seed, auto_reset=True)
    st, _ = env.new_state()
    for t in range(K):
        st = env.step_(st, env.random_actions(t), out=st)
    ob = orc.TronBatch(B, 19, 4); ob.rollout(seed, 0, 0, K, fresh=True)
    o = env.state_to_observation(st, -1)
    assert (o["board"].cpu().numpy() == ob.board).all()
    env.state_to_observation(st, 2); env.state_to_observation(st, -3); env.compute_ranking(st)
    sp = env.host_stepper(st, compact=2, packed_actions=True)
    sp(); sp()
    env.rollout(st, K + 3, 5)
    env2 = BatchedTronGridEnvironment("9;3", batch=77, seed=seed, auto_reset=True)
    s2, _ = env2.new_state(ring_offset=1, spawn_offset=[1, 0, -1])
    for t in range(12):
        s2 = env2.step_(s2, env2.random_actions(t), out=s2)
    env2.state_to_observation(s2, -3)
    print("tron ok")
    # Blokus: legal + pick + step over a whole game, is_valid, observe
    Bb, Kb = 48, 72
    benv = BatchedBlokusEnvironment(batch=Bb, seed=seed, auto_reset=True, capacity=2304)
    bs, _ = benv.new_state()
    for t in range(Kb):
        valid = benv.valid_actions(bs)
        act = benv.random_actions(valid, t)
        if t % 9 == 0:
            assert benv.is_valid_action(bs, -1, act).cpu().numpy()[act.cpu().numpy() >= 0].all()
        bs = benv.step_(bs, act, out=bs)
    bob = orc.BlokusBatch(Bb); bob.rollout(seed, 0, 0, Kb, fresh=True)
    o = benv.state_to_observation(bs, -1)
    assert (o["board"].cpu().numpy() == bob.board).all()
    benv.state_to_observation(bs, -2); benv.state_to_observation(bs, 3)
    print("blokus ok")
    # Tic Tac Toe 2/3/4p: policy + step, fused rollout, observe
    for cls, n in ((BatchedTicTacToe2PlayerEnv, 2), (BatchedTicTacToe3PlayerEnv, 3), (BatchedTicTacToe4PlayerEnv, 4)):
        Bt, Kt = 3001, 12
        tenv = cls(batch=Bt, seed=seed, auto_reset=True)
        ts, _ = tenv.new_state()
        for t in range(Kt):
            ts = tenv.step_(ts, tenv.random_actions(ts, t), out=ts)
        tob = orc.TTTBatch(Bt, n); tob.rollout(seed, 0, 0, Kt, fresh=True)
        board, winner, mover = tenv.state_arrays(ts)
        assert (board.cpu().numpy() == tob.board).all()
        tenv.rollout(ts, Kt, 7)
        tenv.state_to_observation(ts, -2); tenv.state_to_observation(ts, 1)
        assert int(tenv.stats[0]) == Bt * (Kt + 7)
    torch.cuda.synchronize()
    print("ttt ok")


if __name__ == "__main__":
    main()
