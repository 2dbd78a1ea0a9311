#!/bin/bash
# quick: a few parity tests + short bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "small_shapes or config1 or unit_gaussian or skewed or duplicates or mid_size or wide_rows or config5 or scaled" > gpurun_out/r2_t2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t2.log
tail -5 gpurun_out/r2_t2.log
VI_B200_TRACE=1 timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu --no-exact --no-search --no-e2e > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.log
echo "bench rc=$?"
grep -E "cycles per sub-tree|fast build" gpurun_out/r2_b2.log | tail -2
