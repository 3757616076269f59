#!/bin/bash
# round 2, GPU call 8: full GPU parity suite, TTT rollout register-budget A/B, bench at the driver's flags
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_gputests8.log 2>&1; echo "gpu tests rc=$?"
tail -8 gpurun_out/r2_gputests8.log
for mb in 8 6; do
  echo "== CRL_TTT_MINB=$mb"
  CRL_TTT_MINB=$mb python tools/ttt_probe.py 2>&1 | grep -E "fused|statistics"
done
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2_bench8.json
CRL_TTT_MINB=6 python bench.py --workload ttt4 --no-cpu --steps 20 --warmup 5 > gpurun_out/r2_bench8_ttt6.json 2> gpurun_out/r2_bench8_ttt6.err
python tools/show_bench.py gpurun_out/r2_bench8_ttt6.json
