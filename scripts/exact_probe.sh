#!/bin/bash
# exact-mode level-0 probes: ring depth, copy shape; then one ncu capture of the big-range kernel (2M rows)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --rows 2000000 --steps 2 --warmup 1 --no-cpu --no-search --no-e2e"
for cfg in "VI_B200_EX_NG=6 VI_B200_EX_VEC=1" "VI_B200_EX_NG=10 VI_B200_EX_VEC=1" "VI_B200_EX_NG=4 VI_B200_EX_VEC=1" "VI_B200_EX_NG=6 VI_B200_EX_VEC=0" "VI_B200_EX_NG=10 VI_B200_EX_VEC=0"; do
  echo "== $cfg"
  env $cfg $B 2>&1 | grep -E "^exact build" | cut -c1-260
done
ncu --set full --clock-control none --import-source on -k regex:k_stats_big_exact -c 1 -o gpurun_out/r1_bigexact -f $B > gpurun_out/ncu_bigexact.log 2>&1
echo "ncu rc=$?"
