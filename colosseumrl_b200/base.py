"""BatchedBaseEnvironment: the reference's plugin surface (colosseumrl/BaseEnvironment.py:10-283), stepping a
whole batch of independent games per call.

Method names, argument order and the order of return values are the reference's; Python objects become
tensors on the environment's CUDA device:

=====================  ==============================  ===============================================
reference               type there                       here (B = batch, P = players)
=====================  ==============================  ===============================================
state                   opaque object                    opaque ``*BatchState`` (bit-packed int32 tensor)
players / new_players   List[int]                        uint8 [B] bit mask (bit p = player p acts next)
actions                 List[str]                        integer tensor (per game; see each class)
rewards                 List[float] / ndarray            int8 [B, P] (Tron) or int8 [B] (mover's reward)
terminal                bool                             uint8 [B] (0 / 1)
winners                 List[int] | None                 uint8 [B] bit mask (0 = None / nobody)
ranking                 Dict[int, int]                   uint8 [B, P]
=====================  ==============================  ===============================================

``is_terminal`` does not exist in the reference (terminal is next_state's 4th return value); here it reads the
flag the step kernel fused into the state.  All compute happens in libcolosseum_b200.so (hand-written sm_100a
kernels); there is no CPU or PyTorch fallback path.
"""
from abc import ABC, abstractmethod
from typing import Dict, List, Tuple

import torch

from . import _lib


class HostStepper:
    """One env-step for an actor whose policy runs on the HOST, as a single CUDA-graph launch:
    H2D copy of `actions` (pinned) -> step kernel (in place on `state`) -> D2H copy of the result record (pinned).

        stepper = env.host_stepper(state)
        stepper.actions[...] = my_policy(...)      # write into the pinned action buffer
        result = stepper()                         # replay + wait; `result` is a numpy view of the pinned uint8 records
    An actor that serves two (or more) environment batches can pipeline them: `a.launch(); b.wait(); ...`; give each
    stepper its own `stream` and the copy engines of one batch overlap the step kernel of another.

    Two transports:
      * copy engines (default): the graph is  memcpy H2D -> step kernel -> memcpy D2H.  Right for wide records: with
        the full 8-byte Tron record zero-copy was measured ~4x slower (posted PCIe writes of scattered 8-byte stores).
      * zero copy (`zero_copy_step`): the step kernel itself reads the actions from and writes the records to the
        PINNED HOST buffers (a cudaHostAlloc'ed buffer has the same address on the device under UVA), so the graph is
        ONE kernel node.  With Tron's packed actions (1 B / env, one coalesced 64-byte read per CTA) and 2-byte records
        (one 128-byte segment per CTA) the PCIe traffic is the same 64 KB up + 128 KB down per step, but the two
        memcpy nodes -- ~3 us each in stream order, the cost that bounded this leg -- are gone: 8.3 -> 5.7 us per step,
        7.9 -> 11.6 G env-steps/s (tools/e2e_probe.py; the floor with device-resident buffers is 4.7 us).
    """

    def __init__(self, env, state, action_shape, action_dtype, stream=None, step=None, zero_copy_step=None,
                 result_shape=None):
        """step(dev_actions) -> device tensor holding the step's result record; default: env.step_ in place on
        `state`, the full record.  (Tron passes a compact-record step: half the bytes to read back.)
        zero_copy_step(host_actions, host_result): launches the step with the pinned buffers as its action / result
        arguments (result_shape = shape of the uint8 record buffer)."""
        self.env, self.state, self.stream = env, state, stream
        with torch.cuda.device(env.device):
            if zero_copy_step is not None:
                self._build_zero_copy(env, action_shape, action_dtype, result_shape, stream, zero_copy_step)
            else:
                self._build(env, state, action_shape, action_dtype, stream, step)

    def _build_zero_copy(self, env, action_shape, action_dtype, result_shape, stream, zstep):
        self.actions = torch.zeros(action_shape, dtype=action_dtype).pin_memory()
        self.result = torch.zeros(result_shape, dtype=torch.uint8).pin_memory()
        self.result_np, self.actions_np = self.result.numpy(), self.actions.numpy()
        self._step = zstep
        s = stream if stream is not None else torch.cuda.current_stream(env.device)
        with torch.cuda.stream(s):
            zstep(self.actions, self.result)                    # warms the launch path (applies one step)
        torch.cuda.synchronize(env.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            zstep(self.actions, self.result)
        self._finish(env, s)

    def _finish(self, env, s):
        # replay / completion through the CUDA runtime directly: three ctypes calls per step instead of torch's
        # stream context + event objects (the host loop is CPU-bound)
        from . import _cudart
        self._rt = _cudart.rt()
        self._exec = self.graph.raw_cuda_graph_exec()
        self._done = _cudart.new_event()
        self._stream_handle = s.cuda_stream
        _cudart.check(self._rt.cudaGraphUpload(self._exec, self._stream_handle), "cudaGraphUpload")
        # ONE foreign call per launch (graph launch + event record) and one per wait: libcolosseum_b200's host helpers
        raw = _lib.load()
        self._launch_fn, self._wait_fn = raw.crl_host_graph_launch, raw.crl_host_event_wait
        self._launch_wait_fn = raw.crl_host_graph_launch_wait

    def _build(self, env, state, action_shape, action_dtype, stream, step):
        self.actions = torch.zeros(action_shape, dtype=action_dtype).pin_memory()
        self._dev_actions = torch.zeros(action_shape, dtype=action_dtype, device=env.device)
        if step is None:
            def step(dev_actions):
                return env.step_(state, dev_actions, out=state).result
        # the graph below bakes in the addresses of everything `step` touches: keep the closure (and with it any
        # buffer it owns) and the record tensor alive for as long as the graph can be replayed
        self._step = step
        dev_result = self._dev_result = step(self._dev_actions)   # allocates the record; warms the launch path
        self.result = torch.empty(dev_result.shape, dtype=torch.uint8).pin_memory()
        self.result_np = self.result.numpy()                    # zero-copy numpy view of the pinned record (cheap reads)
        self.actions_np = self.actions.numpy()
        torch.cuda.synchronize(env.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._dev_actions.copy_(self.actions, non_blocking=True)
            self.result.copy_(step(self._dev_actions), non_blocking=True)

        self._finish(env, stream if stream is not None else torch.cuda.current_stream(env.device))

    def launch(self):
        """Enqueue the step (one graph launch) on the stepper's stream; returns immediately."""
        rc = self._launch_fn(self._exec, self._stream_handle, self._done)
        if rc:
            if rc == _lib.ERR_ARG:
                # the calling thread's current device is not the stepper's.  The hot loop of a one-process-per-GPU actor
                # never pays for a device guard; a multi-device process lands here and retries under one (the failed
                # call launched nothing).
                with torch.cuda.device(self.env.device):
                    rc = self._launch_fn(self._exec, self._stream_handle, self._done)
            if rc:
                raise RuntimeError("HostStepper launch failed: %s" % _lib.load().crl_last_error().decode())

    def launch_then_wait(self, oldest: "HostStepper"):
        """A pipelined actor's whole step in ONE foreign call: enqueue this batch's step, then block until the step of
        `oldest` (the batch launched longest ago) has delivered its records; returns `oldest.result_np`."""
        rc = self._launch_wait_fn(self._exec, self._stream_handle, self._done, oldest._done)
        if rc:
            if rc == _lib.ERR_ARG:                 # wrong current device: retry under a guard (nothing was launched)
                with torch.cuda.device(self.env.device):
                    rc = self._launch_wait_fn(self._exec, self._stream_handle, self._done, oldest._done)
            if rc:
                raise RuntimeError("HostStepper launch failed: %s" % _lib.load().crl_last_error().decode())
        return oldest.result_np

    def wait(self):
        """Block until the launched step's result record is in `self.result` (pinned host memory)."""
        if self._wait_fn(self._done):
            raise RuntimeError("HostStepper wait failed: %s" % _lib.load().crl_last_error().decode())
        return self.result_np

    def __del__(self):
        try:
            self._rt.cudaEventDestroy(self._done)
        except Exception:
            pass

    def __call__(self):
        self.launch()
        return self.wait()


class _DeviceBoundLib:
    """The C ABI takes no device argument: a launch goes to the CUDA context that is current on the calling thread.
    Every call made through this proxy runs with the environment's device current (and restores the caller's), so
    environments on different GPUs can live in one process and the user's `torch.cuda.current_device()` never moves."""

    def __init__(self, lib, index):
        self._lib, self._index = lib, index

    def __getattr__(self, name):
        fn, index = getattr(self._lib, name), self._index

        def call(*args):
            if torch.cuda.current_device() == index:
                return fn(*args)
            with torch.cuda.device(index):
                return fn(*args)
        self.__dict__[name] = call            # resolved once per entry point
        return call


class BatchedBaseEnvironment(ABC):
    def __init__(self, config: str = "", batch: int = 1, device="cuda:0", seed: int = 0, auto_reset: bool = False,
                 first_env_id: int = 0):
        self._config = config
        self.batch = int(batch)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CrlError("colosseumrl_b200 runs on CUDA devices only (no CPU fallback); got %r" % (device,))
        if not torch.cuda.is_available():
            raise _lib.CrlError("no CUDA device available (colosseumrl_b200 has no CPU fallback)")
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        self._lib = _DeviceBoundLib(_lib.init(index), index)
        self.seed = int(seed)
        self.auto_reset = bool(auto_reset)
        self.first_env_id = int(first_env_id)     # global id of env 0 of this shard (Philox counter)
        # int64 [STAT_ROWS, NSTAT]; partial sums per row, see include/colosseum_b200.h.  `stats` sums the rows.
        self.stats_rows = torch.zeros((_lib.STAT_ROWS, _lib.NSTAT), dtype=torch.int64, device=self.device)
        self.collect_stats = True                 # False: the step kernels skip the fused episode statistics

    # -- helpers ------------------------------------------------------------------------------------
    @property
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _check(self, rc):
        _lib.check(rc)

    def _dev(self, t, dtype):
        """Move an action tensor to the device (non-blocking from pinned host memory)."""
        if not torch.is_tensor(t):
            t = torch.as_tensor(t)
        if t.dtype != dtype:
            t = t.to(dtype)
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        return t.contiguous()

    def _new_result(self, shape):
        """Result-record tensor (device memory)."""
        return torch.empty(shape, dtype=torch.uint8, device=self.device)

    @property
    def _stats_ptr(self):
        return self.stats_rows.data_ptr() if self.collect_stats else None

    @property
    def stats(self) -> torch.Tensor:
        """Episode statistics int64 [NSTAT] (slots: include/colosseum_b200.h CRL_ST_*)."""
        out = torch.empty((_lib.NSTAT,), dtype=torch.int64, device=self.device)
        self._check(self._lib.crl_stats_reduce(self.stats_rows.data_ptr(), out.data_ptr(), 0, self._stream))
        return out

    @property
    def flags(self):
        return (_lib.FLAG_AUTO_RESET if self.auto_reset else 0) | getattr(self, "_debug_flags", 0)

    # -- reference surface --------------------------------------------------------------------------
    @property
    @abstractmethod
    def min_players(self) -> int: ...

    @property
    @abstractmethod
    def max_players(self) -> int: ...

    @staticmethod
    @abstractmethod
    def observation_names() -> List[str]: ...

    @property
    @abstractmethod
    def observation_shape(self) -> Dict[str, tuple]: ...

    @abstractmethod
    def new_state(self, num_players: int = None) -> Tuple[object, torch.Tensor]: ...

    @abstractmethod
    def next_state(self, state, players, actions): ...

    @abstractmethod
    def valid_actions(self, state, player): ...

    @abstractmethod
    def is_valid_action(self, state, player, action): ...

    @abstractmethod
    def state_to_observation(self, state, player: int) -> Dict[str, torch.Tensor]: ...

    @abstractmethod
    def is_terminal(self, state) -> torch.Tensor: ...

    def compute_ranking(self, state, players, winners) -> torch.Tensor:
        """Default (BaseEnvironment.py:173-195): winners rank 0, everybody else 1. uint8 [B, P]."""
        P = self.max_players
        bits = (winners.to(torch.int32)[:, None] >> torch.arange(P, device=winners.device)[None]) & 1
        return (1 - bits).to(torch.uint8)

    @staticmethod
    def serializable() -> bool:
        return True

    @staticmethod
    def serialize_state(state) -> bytes:
        import io
        buf = io.BytesIO()
        torch.save(state, buf)
        return buf.getvalue()

    @staticmethod
    def deserialize_state(serialized_state: bytes):
        import io
        return torch.load(io.BytesIO(serialized_state), weights_only=False)

    # -- statistics (fused into the step kernels) ----------------------------------------------------
    def reset_stats(self):
        self.stats_rows.zero_()

    def all_reduce_stats(self, async_op: bool = False):
        """Sum the episode statistics over all ranks (the only collective of the engine): one crl_stats_reduce launch
        (256 rows -> the 32-slot vector) + one all-reduce of 256 bytes, both on the current stream.  async_op=True
        returns (tensor, work handle) without waiting on the host."""
        from .sharding import all_reduce_stats
        return all_reduce_stats(self.stats, async_op=async_op)
