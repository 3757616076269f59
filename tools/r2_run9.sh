#!/bin/bash
# round 2, GPU call 9: parity (zero-copy stepper, TTT compact record), bench at the driver's flags, compute-sanitizer attempt
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_gputests9.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r2_gputests9.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench9.err
python tools/show_bench.py gpurun_out/r2_bench9.json
for tool in memcheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool --log-file gpurun_out/r2_sanitizer_$tool.log python tools/sanitize_cases.py > gpurun_out/r2_sanitizer_$tool.out 2>&1; echo "$tool rc=$?"
  tail -5 gpurun_out/r2_sanitizer_$tool.log; tail -5 gpurun_out/r2_sanitizer_$tool.out
done
