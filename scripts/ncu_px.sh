#!/bin/bash
# one ncu capture of the exact-mode pipeline kernel at 2M x 96 (level 0 launch)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --rows 2000000 --steps 2 --warmup 1 --no-cpu --no-search --no-e2e"
timeout 300 $B 2>&1 | grep -E "^exact build" | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stats_big_exact_px -c 1 -o gpurun_out/r1_px_$1 -f $B > gpurun_out/ncu_px_$1.log 2>&1
echo "ncu rc=$?"
