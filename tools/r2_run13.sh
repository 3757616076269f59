#!/bin/bash
# round 2, GPU call 13 (8 GPUs): the driver's scaling sweep rehearsed -- N = 1, 2, 4, 8 at the driver's flags, both arms at N = 8
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_s8_g1.json 2> gpurun_out/r2_s8_g1.err; echo "bench g1 rc=$?"
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_s8_g$n.json 2> gpurun_out/r2_s8_g$n.err; echo "bench g$n rc=$?"
  fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 bench.py --impl reference --gpus 8 --steps 3 --warmup 3 --workload tron > gpurun_out/r2_s8_ref.json 2> gpurun_out/r2_s8_ref.err; echo "ref g8 rc=$?"
python tools/show_bench.py gpurun_out/r2_s8_g1.json gpurun_out/r2_s8_g2.json gpurun_out/r2_s8_g4.json gpurun_out/r2_s8_g8.json
python - <<'PY'
import json
base = json.load(open("gpurun_out/r2_s8_g1.json"))
for n in (2, 4, 8):
    d = json.load(open("gpurun_out/r2_s8_g%d.json" % n))
    pr = [round(x * 1e3, 3) for x in d["method"]["per_rank_ms_per_step"]]
    print("N=%d tron eff %.3f e2e %.4g per-rank us %s" % (n, d["value"] / base["value"] / n, d["e2e"]["value"], pr))
    for w in ("blokus", "ttt4"):
        print("     %s eff %.3f" % (w, d["workloads"][w]["value"] / base["workloads"][w]["value"] / n))
PY
