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


class ZeroCopyStepper:
    """Experiment: the step kernel reads the packed actions from and writes the 2-byte records to PINNED HOST memory
    directly (UVA: a cudaHostAlloc'ed buffer has the same address on the device), so the graph is ONE kernel node -- no
    copy-engine transfers, no memcpy nodes."""

    def __init__(self, env, state, stream, graph=True, device_buffers=False):
        from colosseumrl_b200 import _cudart, _lib
        self.env, self.state = env, state
        if device_buffers:
            self.actions = torch.zeros((B,), dtype=torch.uint8, device=env.device)
            self.result = torch.zeros((B, 2), dtype=torch.uint8, device=env.device)
        else:
            self.actions = torch.zeros((B,), dtype=torch.uint8).pin_memory()
            self.result = torch.zeros((B, 2), dtype=torch.uint8).pin_memory()
        self.flags = env.flags | _lib.FLAG_COMPACT2_RESULT | _lib.FLAG_PACKED_ACTIONS
        self._rt = _cudart.rt()
        self._done = _cudart.new_event()
        self._stream_handle = stream.cuda_stream
        self.stream = stream
        self.graph = None
        with torch.cuda.stream(stream):
            self._call()
        torch.cuda.synchronize()
        if graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=stream):
                self._call()
            self._exec = self.graph.raw_cuda_graph_exec()
            _cudart.check(self._rt.cudaGraphUpload(self._exec, self._stream_handle), "upload")
        self.result_np = self.result.numpy() if not device_buffers else None

    def _call(self):
        e, p = self.env, self.state.packed.data_ptr()
        e._check(e._step_call(p, p, self.actions.data_ptr(), self.result.data_ptr(), self.flags))

    def launch(self):
        if self.graph is not None:
            rc = self._rt.cudaGraphLaunch(self._exec, self._stream_handle)
        else:
            with torch.cuda.stream(self.stream):
                self._call()
            rc = 0
        rc = rc or self._rt.cudaEventRecord(self._done, self._stream_handle)
        assert rc == 0, rc

    def wait(self):
        assert self._rt.cudaEventSynchronize(self._done) == 0
        return self.result_np if self.result_np is not None else np.zeros((1, 1), np.uint8)

    def __call__(self):
        self.launch()
        return self.wait()


run("zero-copy, 1-node graph", lambda e, s, st: ZeroCopyStepper(e, s, st))
run("device buffers, 1-node graph (floor)", lambda e, s, st: ZeroCopyStepper(e, s, st, device_buffers=True))
run("zero-copy, direct launch", lambda e, s, st: ZeroCopyStepper(e, s, st, graph=False))
# parity of the zero-copy path: same states as the copy path after the same actions
a = ZeroCopyStepper(envs[0], states[0], streams[0])
b = envs[1].host_stepper(states[1], stream=streams[1], compact=2, packed_actions=True)
states[1].packed.copy_(states[0].packed)
torch.cuda.synchronize()
rng = np.random.RandomState(0)
for t in range(20):
    acts = rng.randint(0, 256, size=B).astype(np.uint8)
    a.actions.numpy()[:] = acts
    b.actions_np[:] = acts
    ra, rb = a().copy(), b().copy()
    assert (ra == rb).all(), t
torch.cuda.synchronize()
assert (states[0].packed == states[1].packed).all()
print("zero-copy parity OK")


def run_threads(name, nthreads, depth):
    """`nthreads` host threads, each driving its own share of the zero-copy steppers (ctypes releases the GIL during the
    runtime calls): does the host-bound leg scale with actor threads?"""
    import threading
    sps = [envs[i].host_stepper(states[i], stream=torch.cuda.Stream(), compact=2, packed_actions=True, zero_copy=True) for i in range(G)]
    for sp in sps: sp()
    torch.cuda.synchronize()
    per = K // nthreads

    def worker(mine):
        launch = [sp.launch for sp in mine]; wait = [sp.wait for sp in mine]
        n = len(mine); r = 0
        for k in range(per):
            launch[k % n]()
            j = k - depth + 1
            if j >= 0: r += int(wait[j % n]()[0, 0])
        for j in range(max(0, per - depth + 1), per): r += int(wait[j % n]()[0, 0])

    ths = [threading.Thread(target=worker, args=(sps[t::nthreads],)) for t in range(nthreads)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    print("%-34s %.2f us/step  %.2f G env-steps/s" % (name, dt / (per * nthreads) * 1e6, B * per * nthreads / dt / 1e9))


K = 4000
run_threads("zero-copy, 1 thread, depth 16", 1, 16)
run_threads("zero-copy, 2 threads, depth 8 each", 2, 8)
run_threads("zero-copy, 4 threads, depth 4 each", 4, 4)
