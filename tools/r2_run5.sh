#!/bin/bash
# round 2, GPU call 5 (re-entry): parity, bench at the driver's flags (both arms), then the profile refresh
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_gputests.log 2>&1; echo "gpu tests rc=$?"
tail -4 gpurun_out/r2_gputests.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench.err
python bench.py --workload tron --no-cpu --steps 20 --warmup 5 --streams 1 > gpurun_out/r2_bench_tron_s1.json 2> gpurun_out/r2_bench_tron_s1.err
python tools/tron_probe.py --steps 1000 --reps 5 > gpurun_out/r2_tron_probe.log 2>&1
cat gpurun_out/r2_tron_probe.log
python tools/ttt_probe.py > gpurun_out/r2_ttt_probe.log 2>&1; tail -12 gpurun_out/r2_ttt_probe.log
python tools/show_bench.py gpurun_out/r2_bench.json gpurun_out/r2_bench_tron_s1.json gpurun_out/r2_bench_ref.json
bash tools/r2_profile.sh r02 2>&1 | tail -40
