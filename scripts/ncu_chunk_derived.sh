# ncu --set full of one k_stats_big_fast launch at a level with sibling derivation (level 2 of the second build)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-search --no-cpu --no-e2e --no-exact"
timeout 200 $B > /dev/null 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none -k regex:k_stats_big_fast -s 17 -c 1 -o gpurun_out/r1_chunk_derived -f $B > gpurun_out/ncu_chunk_derived.log 2>&1
echo "ncu rc=$?"
