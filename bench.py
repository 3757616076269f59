#!/usr/bin/env python
"""Benchmark of the batched game-dynamics hot path (BASELINE.json metric: batched env-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tron|ttt4|blokus] [--impl b200|reference]

One "step" = one next_state pass over one batch of environments (the workload's configured batch per GPU).
N > 1 is launched by torchrun (one rank per GPU, contiguous slices of global environment ids, no data-path
collective; one NCCL all-reduce of the episode statistics per measurement window).

What is timed
  value  : device-resident.  Per step: [untimed: random-policy action kernel, L2 flush]  ->  CUDA event ->
           step kernel -> CUDA event.  ms_per_step = mean device time of the step kernel; max over ranks.
  e2e    : through the public Python API (Batched*Environment.next_state) with HOST buffers: per step the actions
           are copied from pinned host memory, the step runs, the result record is copied back to pinned host
           memory (all inside the timed event pair, L2 flushed before it).
  roofline: algorithmic bytes per env-step x envs / mean step-kernel time vs MEASURED_PEAKS.json.
  cpu_baseline / --impl reference: the CPU oracle port (oracle/liboracle.so, plain C restatement of the
           reference's Python; the reference itself is Python and cannot travel to the GPU box) on all host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (description, per-GPU batch, algorithmic bytes per env-step)
    "tron": ("Tron 4-player 19x19, 65,536 batched envs, random actions (BASELINE.json configs[1])", 65536, 424),
}
FLUSH_BYTES = 512 << 20     # > 126 MB L2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / throttle reasons with NVML while the timed regions run."""
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._run, daemon=True)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in list(self.BAD.items()) + list(self.NOTE.items()):
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self.t.start()

    def stop(self):
        if self.nv:
            self._stop.set()
            self.t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_tron(B, K, nthreads, seed=0):
    from oracle import oracle as orc
    ob = orc.TronBatch(B, 19, 4)
    ob.rollout(seed, 0, 0, 1, fresh=True, nthreads=nthreads)
    t0 = time.perf_counter()
    ob.rollout(seed, 0, 1, K, fresh=False, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return B * K / dt, dt


def cpu_baseline(workload, target_s=10.0):
    from oracle import oracle as orc
    cores = orc.num_threads()
    B = WORKLOADS[workload][1]
    fn = {"tron": cpu_tron}[workload]
    rate, _ = fn(B, 4, cores)
    K = max(4, int(rate * target_s / B))
    rate, dt = fn(B, K, cores)
    return {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": "%d envs x %d steps (%.1f s) of the C oracle port, %d pthreads" % (B, K, dt, cores)}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port, all host threads), same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    desc, B, _ = WORKLOADS[args.workload]
    cores = orc.num_threads()
    ob = orc.TronBatch(B, 19, 4)
    ob.rollout(0, 0, 0, 1, fresh=True, nthreads=cores)
    t = 1
    for _ in range(args.warmup):
        ob.rollout(0, 0, t, 1, nthreads=cores); t += 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ob.rollout(0, 0, t, 1, nthreads=cores); t += 1
    dt = time.perf_counter() - t0
    value = B * args.steps / dt
    line = {"impl": "reference", "metric": "batched env-steps/sec", "value": value, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": desc, "batch_per_step": B, "policy": "philox4x32-10 uniform random",
                       "note": "reference is Python (cannot travel to the GPU box); this is its plain-C restatement "
                               "oracle/liboracle.so, which is ~100x faster than the Python original"},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": "%d envs x %d steps" % (B, args.steps)},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from colosseumrl_b200.tron import BatchedTronGridEnvironment

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    desc, B, bytes_per_step = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup

    env = BatchedTronGridEnvironment("", batch=B, device=dev, seed=0, auto_reset=True, first_env_id=rank * B)
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    state, players = env.new_state()
    spare, _ = env.new_state()
    actions = torch.empty((B, 4), dtype=torch.int8, device=dev)
    lib, stream = env._lib, torch.cuda.current_stream(dev)

    def one_step(t, timed):
        nonlocal state, spare
        env.random_actions(t, out=actions)
        flush.fill_(t & 0xff)                                   # evict the state from L2 (untimed)
        ev0, ev1 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if timed else (None, None)
        if timed:
            ev0.record(stream)
        new, _, _, _, _ = env.next_state(state, None, actions, out=spare)   # the C-ABI call crl_tron_step
        if timed:
            ev1.record(stream)
        state, spare = new, state
        return ev0, ev1

    clocks = ClockSampler(local)
    t = 0
    for _ in range(W):
        one_step(t, False); t += 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.start()
    wall0 = time.perf_counter()
    evs = []
    for _ in range(K):
        evs.append(one_step(t, True)); t += 1
    ar0, ar1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ar0.record(stream)
    total_stats = env.all_reduce_stats()                          # the only collective: <= 256 B, once per window
    ar1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = sum(step_ms) + (ar0.elapsed_time(ar1) if world > 1 else 0.0)
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tmax.item())
    ms_per_step = dev_ms_max / K
    value = world * B * K / (dev_ms_max * 1e-3)
    kernel_ms = sum(step_ms) / K

    # ---- end to end through the public API with host buffers (pinned): H2D actions, step, D2H result
    h_actions = [torch.empty((B, 4), dtype=torch.int8).pin_memory() for _ in range(2)]
    h_result = torch.empty((B, 8), dtype=torch.uint8).pin_memory()
    rng = np.random.RandomState(rank)
    for h in h_actions:
        h.copy_(torch.from_numpy(rng.randint(-1, 2, size=(B, 4)).astype(np.int8)))
    e2e_evs = []
    Ke = min(K, 200)
    for k in range(W + Ke):
        flush.fill_(k & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        new, _, rewards, terminal, winners = env.next_state(state, None, h_actions[k & 1], out=spare)   # H2D inside
        h_result.copy_(new.result, non_blocking=True)                                                   # D2H
        e1.record(stream)
        e1.synchronize()                    # the host consumes the result before issuing the next step
        state, spare = new, state
        if k >= W:
            e2e_evs.append(e0.elapsed_time(e1))
    e2e_ms = torch.tensor([sum(e2e_evs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * Ke / (float(e2e_ms.item()) * 1e-3)
    clk = clocks.stop()

    if rank == 0:
        peak, peak_src = peaks()
        achieved = bytes_per_step * B / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        line = {
            "metric": "batched env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": B, "policy": "philox4x32-10 uniform random, auto-reset",
                       "l2": "flushed between timed steps (%d MiB write)" % (FLUSH_BYTES >> 20),
                       "state_bytes_per_env": 208, "parallelism": "env-sharded x%d, no data-path collective" % world},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "tron_step_kernel",
                         "algorithmic_bytes_per_env_step": bytes_per_step, "kernel_ms": kernel_ms},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": B * 4, "d2h_bytes_per_step": B * 8,
                    "steps": Ke},
            "gpu_launches": K * world,
            "clocks": clk,
            "wall_s_timed_region": wall,
            "episodes": int(total_stats[1].item()), "env_steps_counted": int(total_stats[0].item()),
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="tron", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
