#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-search --no-cpu --no-exact --no-e2e"
timeout 300 $B > gpurun_out/r2_plain_sub.json 2> gpurun_out/r2_plain_sub.log || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_subtree_cta -s 1 -c 1 -o gpurun_out/r2_subtree -f $B > gpurun_out/r2_ncu_sub.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/r2_ncu_sub.log
