#!/bin/bash
# parity suite, then a short bench (each under its own timeout)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu > gpurun_out/r2_t1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t1.log
tail -15 gpurun_out/r2_t1.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-exact > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.log
echo "bench rc=$?"
tail -12 gpurun_out/r2_b1.log
