# one ncu --set full capture of k_flags and k_scatter at a top level of 10M x 96 (after the plain command passed)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-search --no-cpu --no-e2e --no-exact"
timeout 200 $B > /dev/null 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none -k regex:"k_flags|k_scatter" -s 8 -c 2 -o gpurun_out/r1_partition -f $B > gpurun_out/ncu_partition.log 2>&1
echo "ncu rc=$?"
