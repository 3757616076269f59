#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharding.py tests/test_gpu_single.py tests/test_gpu_api.py -m gpu -q > gpurun_out/r2_gpu_sharding.log 2>&1; echo "sharding+api tests rc=$?"
tail -4 gpurun_out/r2_gpu_sharding.log
