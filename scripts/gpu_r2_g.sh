#!/bin/bash
# round 2, job G: wide rows (1M x 768): T_BIG x WIDE_CH sweep; ncu source profile of the sub-tree kernel at 10M x 96
mkdir -p gpurun_out
SWEEP_T_BIG=512,128,64 SWEEP_WIDE_CH=6,3,2,1 timeout 600 python scripts/sweep_tbig.py 1000000 768 > gpurun_out/r2_sweep_768.log 2>&1
cat gpurun_out/r2_sweep_768.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_subtree_fast -c 1 -o gpurun_out/r2_subtree_final -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-exact --no-e2e --no-search > gpurun_out/r2_subtree_final.log 2>&1
echo "ncu rc=$?"
