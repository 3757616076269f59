"""Test helper: two ways to call the SAME C ABI (include/colosseum_b200.h).

* ``Cuda``    -- the product library libcolosseum_b200.so on a real GPU (torch CUDA tensors own the memory).
* ``HostSim`` -- the kernel *source* compiled with g++ against tests/hostsim/cuda_shim.h (a SIMT emulator),
                 numpy arrays own the memory.  CPU unit tests only; the product never loads it.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from colosseumrl_b200 import _lib  # noqa: E402


class HostSim:
    name = "hostsim"
    stream = None

    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
        import build_hostsim
        self.lib = _lib.declare(C.CDLL(build_hostsim.build()))

    def zeros(self, shape, dtype):
        return np.zeros(shape, dtype)

    def upload(self, arr):
        return np.array(arr, copy=True, order="C")

    def download(self, buf):
        return np.array(buf, copy=True)

    def ptr(self, buf):
        return None if buf is None else buf.ctypes.data

    def sync(self):
        pass

    def check(self, rc):
        assert rc == 0, self.lib.crl_last_error().decode()


class Cuda:
    name = "cuda"

    def __init__(self):
        import torch
        self.torch = torch
        assert torch.cuda.is_available()
        self.dev = torch.device("cuda:0")
        self.lib = _lib.init(0)
        self._np2t = {np.dtype(np.int8): torch.int8, np.dtype(np.uint8): torch.uint8, np.dtype(np.int32): torch.int32,
                      np.dtype(np.int64): torch.int64, np.dtype(np.uint32): torch.int32}

    @property
    def stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def zeros(self, shape, dtype):
        return self.torch.zeros(shape, dtype=self._np2t[np.dtype(dtype)], device=self.dev)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.uint32:
            arr = arr.view(np.int32)
        return self.torch.from_numpy(arr.copy()).to(self.dev)

    def download(self, buf):
        self.torch.cuda.synchronize()
        return buf.cpu().numpy()

    def ptr(self, buf):
        return None if buf is None else buf.data_ptr()

    def sync(self):
        self.torch.cuda.synchronize()

    def check(self, rc):
        _lib.check(rc, self.lib)
