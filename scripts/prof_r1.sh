# round-1 profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one full capture
set -x
B="python bench.py --steps 2 --warmup 1 --no-search --no-cpu --no-exact --no-e2e"
timeout 300 $B > gpurun_out/plain_r1.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1_final.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 300 $B > gpurun_out/plain_r1b.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stats_big_fast -s 3 -c 2 -o gpurun_out/prof_big_final $B > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stats_small_fast -s 2 -c 1 -o gpurun_out/prof_warp_final $B > gpurun_out/ncu_w.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_flags|k_scatter" -s 4 -c 2 -o gpurun_out/prof_part_final $B > gpurun_out/ncu_p.log 2>&1
ls -la gpurun_out | tail -8
