#!/usr/bin/env python
"""Development probe: time of the state_to_observation kernels at the BASELINE.json batch sizes (CUDA events, graph of
20 launches, best of 5) and their share of the measured HBM bandwidth (bytes = packed state read + dense tensors written)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from colosseumrl_b200 import BatchedTronGridEnvironment, BatchedBlokusEnvironment, BatchedTicTacToe4PlayerEnv  # noqa: E402


def timeit(fn, n=20, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


def main():
    peak = 6551.4
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    for name, env, player in (("tron 65,536", BatchedTronGridEnvironment("", batch=65536, seed=1), 1),
                              ("tron 65,536 (all 4 views)", BatchedTronGridEnvironment("", batch=65536, seed=1), -3),
                              ("blokus 16,384", BatchedBlokusEnvironment("", batch=16384, seed=1), 1),
                              ("blokus 16,384 (mover)", BatchedBlokusEnvironment("", batch=16384, seed=1), -2),
                              ("ttt4 1,048,576", BatchedTicTacToe4PlayerEnv("", batch=1 << 20, seed=1), 1)):
        st, _ = env.new_state()
        obs = env.state_to_observation(st, player)
        out_bytes = sum(v.numel() * v.element_size() for v in obs.values())
        in_bytes = st.packed.numel() * 4
        us = timeit(lambda: env.state_to_observation(st, player))
        print("%-24s %8.2f us   %6.1f MB in + %6.1f MB out   %6.0f GB/s  (%.2f of measured HBM; includes torch.empty)" %
              (name, us, in_bytes / 1e6, out_bytes / 1e6, (in_bytes + out_bytes) / us / 1e3, (in_bytes + out_bytes) / us / 1e3 / peak))


if __name__ == "__main__":
    main()
