"""TEST INFRASTRUCTURE ONLY -- compile the kernel sources with g++ against the SIMT emulator shim.

Produces tests/hostsim/libcrl_hostsim.so exporting the same C ABI as libcolosseum_b200.so but taking HOST
pointers.  Used by the CPU unit tests to exercise the kernel source before GPU time is spent; never
loaded by the product package.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "colosseumrl_b200", "csrc")
LIB = os.path.join(HERE, "libcrl_hostsim.so")


def build(force=False, sanitize=None):
    """sanitize = "address" | "thread" (or the CRL_HOSTSIM_SANITIZE environment variable): the emulator library built
    with -fsanitize=address,undefined / -fsanitize=thread.  One host thread per CUDA thread with real barriers, so
    AddressSanitizer checks every shared / global access of the kernel source for bounds (compute-sanitizer memcheck's
    job) and ThreadSanitizer flags accesses that are not ordered by __syncthreads / __syncwarp / atomics (racecheck's
    job).  Run through tools/hostsim_sanitize.sh (the sanitizer runtime must be LD_PRELOADed into python)."""
    global LIB
    sanitize = sanitize or os.environ.get("CRL_HOSTSIM_SANITIZE")
    if sanitize:
        LIB = os.path.join(HERE, "libcrl_hostsim_%s.so" % sanitize)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps += [os.path.join(HERE, f) for f in ("cuda_shim.h", "hostsim.cpp")]
    deps += [os.path.join(ROOT, "include", "colosseum_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    san = {"address": ["-fsanitize=address,undefined", "-fno-omit-frame-pointer"],
           "thread": ["-fsanitize=thread", "-fno-omit-frame-pointer"], None: []}[sanitize]
    cmd = [cxx, "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-DCRL_HOSTSIM", "-I", HERE, "-I", CSRC] + san + [
"-Wno-unknown-pragmas", "-x", "c++", os.path.join(CSRC, "crl_api.cu"), os.path.join(HERE, "hostsim.cpp"),
           "-o", LIB]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
