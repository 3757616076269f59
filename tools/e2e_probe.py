#!/usr/bin/env python
"""Dev probe: where does the Tron e2e step go?  Same 8-deep HostStepper pipeline as bench.py's e2e leg with the graph
reduced to (a) H2D + step + D2H (compact), (b) step only, (c) copies only; plus the bare Python loop cost."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseumrl_b200.tron import BatchedTronGridEnvironment
from colosseumrl_b200.base import HostStepper

B, G, D, K = 65536, 16, 8, 400
envs = [BatchedTronGridEnvironment("", batch=B, seed=0, auto_reset=True, first_env_id=g * B) for g in range(G)]
states = [e.new_state()[0] for e in envs]
streams = [torch.cuda.Stream() for _ in range(D)]


def run(name, make):
    sps = [make(e, s, streams[i % D]) for i, (e, s) in enumerate(zip(envs, states))]
    for sp in sps: sp()
    torch.cuda.synchronize()
    infl = []
    t0 = time.perf_counter()
    for k in range(K):
        sp = sps[k % G]; sp.launch(); infl.append(sp)
        if len(infl) >= D: int(infl.pop(0).wait()[0, 0])
    while infl: infl.pop(0).wait()
    dt = time.perf_counter() - t0
    print("%-34s %.2f us/step  %.2f G env-steps/s" % (name, dt / K * 1e6, B * K / dt / 1e9))


run("packed actions + 2-byte record", lambda e, s, st: e.host_stepper(s, stream=st, compact=2, packed_actions=True))
run("packed actions + 4-byte record", lambda e, s, st: e.host_stepper(s, stream=st, compact=True, packed_actions=True))
run("H2D + step + D2H (compact)", lambda e, s, st: e.host_stepper(s, stream=st, compact=True))
run("H2D + step + D2H (full record)", lambda e, s, st: e.host_stepper(s, stream=st))


def step_only(e, s, st):
    rec = torch.empty((B, 4), dtype=torch.uint8, device=e.device)
    dev_a = torch.zeros((B, 4), dtype=torch.int8, device=e.device)
    hs = HostStepper(e, s, (1, 4), torch.int8, stream=st, step=lambda a: (e.step_(s, dev_a, out=s), rec[:1])[1])
    return hs


run("step only (tiny copies)", step_only)


def copies_only(e, s, st):
    rec = torch.empty((B, 4), dtype=torch.uint8, device=e.device)
    return HostStepper(e, s, (B, 4), torch.int8, stream=st, step=lambda a: rec)


run("copies only (256 KB each way)", copies_only)
