# round-1 final launch list (B200_PROFILING.md): plain run first, then the same command under ncu (durations only)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-search --no-cpu --no-e2e"
timeout 300 $B > gpurun_out/plain_final.json 2> gpurun_out/plain_final.log || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1_launches_final2.csv $B > gpurun_out/ncu_final.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r1_launches_final2.csv
