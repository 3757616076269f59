#!/usr/bin/env python
"""Development probe (not the benchmark): Tron step kernel with phases switched off by the diagnostic flag bits
(csrc/tron.cuh), several graph replays per variant, min / median us per step.  Results are meaningless for parity --
this only attributes time to the phases.

    python tools/tron_probe.py [--steps 1000] [--reps 7]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from colosseumrl_b200.tron import BatchedTronGridEnvironment  # noqa: E402

VARIANTS = [("full", 0), ("full, loads staggered 100 ns / SM rank", 0x8000), ("full, loads staggered 200 ns / SM rank", 0x10000),
            ("full, loads staggered 300 ns / SM rank", 0x18000),
            ("data movement only, staggered 200 ns", 0x100 | 0x800 | 0x10000), ("full, PDL trigger at the very top", 0x4000), ("full, late PDL trigger", 0x1000), ("full, no L2 prefetch", 0x2000),
            ("full, round-1 (late trigger, no prefetch)", 0x3000),
            ("no stats", 0x800), ("no phase3, no stats", 0xC00), ("no phase2/3, no stats", 0xE00),
            ("data movement only", 0x100 | 0x800), ("data movement only, round-1", 0x100 | 0x800 | 0x3000),
            ("data movement only, trigger at top", 0x100 | 0x800 | 0x4000)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--replicas", type=int, default=39)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B, G, K = args.batch, args.replicas, args.steps
    envs = [BatchedTronGridEnvironment("", batch=B, device=dev, seed=0, auto_reset=True, first_env_id=g * B) for g in range(G)]
    states = [e.new_state()[0] for e in envs]
    for e in envs[1:]:
        e.stats_rows = envs[0].stats_rows
    acts = [e.random_actions(i) for i, e in enumerate(envs)]
    for g in range(G):                       # a few real steps so that the boards are not all fresh
        for t in range(6):
            envs[g].step_(states[g], envs[g].random_actions(t), out=states[g])
    torch.cuda.synchronize()
    print("B = %d, %d replicas, %d steps per graph, tile %s" % (B, G, K, os.environ.get("CRL_TRON_TILE", "64")))
    for name, dbg in VARIANTS:
        for e in envs:
            e._debug_flags = dbg
        graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(dev)
        with torch.cuda.graph(graph, stream=s):
            for k in range(K):
                g = k % G
                envs[g].step_(states[g], acts[g], out=states[g])
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / K * 1e3)
        ts.sort()
        print("%-44s min %.2f  median %.2f us/step   (%.0f GB/s algorithmic at min)" % (name, ts[0], ts[len(ts) // 2], 424 * B / ts[0] / 1e3))


if __name__ == "__main__":
    main()
