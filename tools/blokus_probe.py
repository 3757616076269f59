#!/usr/bin/env python
"""Dev probe: Blokus kernel times by game phase (CUDA events, 16,384 games in lock-step from the start position).

usage: blokus_probe.py [steps=110] [batch=16384]
For ncu: `ncu -k regex:blokus_legal --launch-skip <3*t> --launch-count 1 ... python tools/blokus_probe.py` captures the
legal launch of step t (the probe launches legal 3x per step: 1 for the game, 2 timed repeats).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseumrl_b200.blokus import BatchedBlokusEnvironment  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 110
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    env = BatchedBlokusEnvironment("", batch=B, device="cuda:0", seed=0, auto_reset=True)
    st, _ = env.new_state()
    out = (torch.empty((B,), dtype=torch.int32, device="cuda:0"), torch.empty((B, env.capacity), dtype=torch.int32, device="cuda:0"))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    rows = []
    for t in range(steps):
        valid = env.valid_actions(st, out=out)
        ev[0].record()
        env.valid_actions(st, out=out, count_stats=False)
        env.valid_actions(st, out=out, count_stats=False)
        ev[1].record()
        act = env.random_actions(valid, t)
        ev[2].record()
        env.step_(st, act, out=st)
        ev[3].record()
        torch.cuda.synchronize()
        rows.append((t, ev[0].elapsed_time(ev[1]) * 500.0, ev[2].elapsed_time(ev[3]) * 1000.0, float(valid[0].float().mean())))
    print("step  legal_us  step_us  mean_valid")
    for r in rows:
        if r[0] % 4 == 0:
            print("%4d  %8.1f  %7.1f  %9.1f" % r)
    n = len(rows)
    print("mean legal %.1f us, step %.1f us, valid %.1f" % (sum(r[1] for r in rows) / n, sum(r[2] for r in rows) / n,
                                                           sum(r[3] for r in rows) / n))


if __name__ == "__main__":
    main()
