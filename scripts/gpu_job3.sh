#!/bin/bash
mkdir -p gpurun_out
VI_B200_TRACE=1 timeout 120 python bench.py --steps 2 --warmup 1 --no-cpu --no-exact --no-search --no-e2e > gpurun_out/r2_b3.json 2> gpurun_out/r2_b3.log
echo "bench rc=$?"
grep -E "sub-tree|fast build" gpurun_out/r2_b3.log | tail -6
