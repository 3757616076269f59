#!/bin/bash
# round 2, GPU call 10: host helpers (one foreign call per launch / wait) + tight e2e loop
mkdir -p gpurun_out
python -m pytest tests/test_gpu_api.py tests/test_gpu_single.py -m gpu -q > gpurun_out/r2_gputests10.log 2>&1; echo "gpu api tests rc=$?"
tail -3 gpurun_out/r2_gputests10.log
python tools/e2e_probe.py 2>&1 | grep -E "2-byte|zero-copy|floor"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench10.err
python tools/show_bench.py gpurun_out/r2_bench10.json
python bench.py --workload tron --no-cpu --steps 20 --warmup 5 --e2e-depth 4 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('depth 4 e2e %.4g'%d['e2e']['value'])"
python bench.py --workload tron --no-cpu --steps 20 --warmup 5 --e2e-depth 16 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('depth 16 e2e %.4g'%d['e2e']['value'])"
