#!/bin/bash
# round 2, GPU call 12 (4 GPUs): multi-rank parity tests, weak-scaling bench at N = 1, 2, 4 at the driver's flags, reference arm
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharding.py -m gpu -q > gpurun_out/r2_gpu_sharding12.log 2>&1; echo "sharding tests rc=$?"
tail -3 gpurun_out/r2_gpu_sharding12.log
for n in 1 2 4; do
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_scale_g1.json 2> gpurun_out/r2_scale_g1.err; echo "bench g1 rc=$?"
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_scale_g$n.json 2> gpurun_out/r2_scale_g$n.err; echo "bench g$n rc=$?"
  fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 --workload tron > gpurun_out/r2_ref_g2.json 2> gpurun_out/r2_ref_g2.err; echo "ref g2 rc=$?"
python tools/show_bench.py gpurun_out/r2_scale_g1.json gpurun_out/r2_scale_g2.json gpurun_out/r2_scale_g4.json gpurun_out/r2_ref_g2.json
python - <<'PY'
import json
base = json.load(open("gpurun_out/r2_scale_g1.json"))
for n in (2, 4):
    d = json.load(open("gpurun_out/r2_scale_g%d.json" % n))
    print("N=%d tron eff %.3f per-rank ms %s" % (n, d["value"] / base["value"] / n, [round(x * 1e3, 3) for x in d["method"]["per_rank_ms_per_step"]]))
    for w in ("blokus", "ttt4"):
        print("     %s eff %.3f" % (w, d["workloads"][w]["value"] / base["workloads"][w]["value"] / n))
PY
