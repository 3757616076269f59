#!/usr/bin/env python
"""Dev probe: TTT-4p kernels one by one (1 048 576 envs, 32 replicas round-robin so the state comes from HBM, 320
launches in one CUDA graph, 1 and 4 parallel chains): next_state alone (resident actions), the random policy alone,
and the fused policy + step kernel that bench.py times."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseumrl_b200.tictactoe import BatchedTicTacToe4PlayerEnv

B, G, K = 1 << 20, 32, 320
envs = [BatchedTicTacToe4PlayerEnv("", batch=B, seed=0, auto_reset=True, first_env_id=g * B) for g in range(G)]
states = [e.new_state()[0] for e in envs]
for g in range(G):
    envs[g].rollout(states[g], 0, 1 + g % 7)              # desynchronise the replicas a little
acts = [e.random_actions(s, 9) for e, s in zip(envs, states)]
torch.cuda.synchronize()
PEAK = 6551.4


def timed(name, fn, bytes_per_env, chains):
    cap = torch.cuda.Stream()
    side = [torch.cuda.Stream() for _ in range(chains - 1)]
    lanes = [cap] + side
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=cap):
        ev = torch.cuda.Event(); ev.record(cap)
        for s in side: s.wait_event(ev)
        for k in range(K):
            with torch.cuda.stream(lanes[(k % G) % chains]):
                fn(k % G, k)
        for s in side:
            j = torch.cuda.Event(); j.record(s); cap.wait_event(j)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / K * 1e3
    print("%-44s chains=%d  %7.2f us  %6.1f G env-steps/s  %5.0f GB/s  %.2f of HBM" %
          (name, chains, us, B / us / 1e3, bytes_per_env * B / us / 1e3, bytes_per_env * B / us / 1e3 / PEAK))


for stats in (True, False):
  for e in envs:
      e.collect_stats = stats
  print("statistics", "on" if stats else "off")
  for ch in (1, 4):
      timed("next_state (crl_ttt_step, resident actions)", lambda g, k: envs[g].step_(states[g], acts[g], out=states[g]), 41, ch)
      timed("random policy (crl_ttt_policy_random)", lambda g, k: envs[g].random_actions(states[g], k, out=acts[g]), 17, ch)
      timed("fused policy + step (crl_ttt_rollout, K=1)", lambda g, k: envs[g].rollout(states[g], k, 1), 36, ch)
