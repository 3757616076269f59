"""Build libcolosseum_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcolosseum_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--shared"]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(HERE, "..", "include", "colosseum_b200.h")]
    return any(os.path.getmtime(s) > t for s in deps)


def build(force=False, verbose=False):
    """Compile the CUDA library in-tree. Returns the .so path."""
    if not force and not is_stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("CRL_NVCC_EXTRA", "").split()          # e.g. -DTTT_DBG=7 for a timing-attribution build
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB, os.path.join(CSRC, "crl_api.cu")]
    env = dict(os.environ)
    # nvcc's host compiler must be the system g++ (the image's CC/CXX wrappers lack some spec files)
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (exit %d)" % proc.returncode)
    with open(os.path.join(HERE, "csrc", "ptxas_info.txt"), "w") as f:     # registers / shared memory per kernel
        f.write("".join(l for l in proc.stdout.splitlines(True) if "Compile time" not in l))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
