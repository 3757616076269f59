#!/bin/bash
# round 2, GPU call 6 (2 GPUs): multi-device parity tests, 2-GPU weak-scaling bench at the driver's flags, 1-GPU bench with the Blokus phase desync
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharding.py tests/test_gpu_blokus.py -m gpu -q > gpurun_out/r2_gpu_sharding.log 2>&1; echo "sharding + blokus tests rc=$?"
tail -4 gpurun_out/r2_gpu_sharding.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_g2.json 2> gpurun_out/r2_bench_g2.err; echo "bench g2 rc=$?"
tail -c 300 gpurun_out/r2_bench_g2.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_g1.json 2> gpurun_out/r2_bench_g1.err; echo "bench g1 rc=$?"
python tools/show_bench.py gpurun_out/r2_bench_g1.json gpurun_out/r2_bench_g2.json
