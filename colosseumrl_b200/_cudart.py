"""Thin ctypes view of the CUDA runtime torch has already loaded: graph launch / upload and events without the
Python-side stream bookkeeping (the host-policy step loop is CPU-bound on exactly that)."""
import ctypes as C

_rt = None


def rt():
    global _rt
    if _rt is None:
        import torch  # noqa: F401  (loads libcudart.so.12)
        lib = C.CDLL("libcudart.so.12")
        lib.cudaGraphLaunch.argtypes = [C.c_void_p, C.c_void_p]
        lib.cudaGraphUpload.argtypes = [C.c_void_p, C.c_void_p]
        lib.cudaEventCreateWithFlags.argtypes = [C.POINTER(C.c_void_p), C.c_uint]
        lib.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
        lib.cudaEventSynchronize.argtypes = [C.c_void_p]
        lib.cudaEventDestroy.argtypes = [C.c_void_p]
        lib.cudaGetLastError.argtypes = []
        lib.cudaGetLastError.restype = C.c_int
        for f in (lib.cudaGraphLaunch, lib.cudaGraphUpload, lib.cudaEventCreateWithFlags, lib.cudaEventRecord,
                  lib.cudaEventSynchronize, lib.cudaEventDestroy):
            f.restype = C.c_int
        _rt = lib
    return _rt


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: cudaError %d" % (what, rc))


def new_event():
    ev = C.c_void_p()
    check(rt().cudaEventCreateWithFlags(C.byref(ev), 0x2), "cudaEventCreateWithFlags")   # cudaEventDisableTiming
    return ev
