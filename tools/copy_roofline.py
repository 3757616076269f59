"""Speed-of-light probe for a given batch: how fast can this GPU move the same bytes as one env-step?

For a workload with S state bytes per env and B envs, one step reads S*B and writes S*B.  This script times a
plain device-to-device copy of S*B bytes (torch copy_ kernel and cudaMemcpyAsync), cycling over enough replicas to
defeat L2 and replaying from a CUDA graph, i.e. exactly the harness bench.py uses for the step kernels.
"""
import argparse
import json

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=208 * 65536)
    ap.add_argument("--steps", type=int, default=400)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n = args.bytes
    G = max(2, -(-4 * (126 << 20) // n))
    src = [torch.empty(n, dtype=torch.uint8, device=dev).random_(0, 255) for _ in range(G)]
    dst = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(G)]
    out = {}
    for name in ("copy_kernel_int4", "inplace_add"):
        def step(k):
            g = k % G
            if name == "copy_kernel_int4":
                dst[g].view(torch.int32).copy_(src[g].view(torch.int32))
            else:
                src[g].view(torch.int32).add_(1)          # read + write the same bytes in place
        for k in range(G + 5):
            step(k)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for k in range(args.steps):
                step(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / args.steps * 1e3
        out[name] = {"us_per_step": us, "GBps": 2 * n / (us * 1e-6) / 1e9}
    print(json.dumps({"bytes": n, "replicas": G, **out}))


if __name__ == "__main__":
    main()
