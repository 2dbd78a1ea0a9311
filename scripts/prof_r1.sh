set -x
timeout 600 python bench.py --steps 3 --warmup 3 --no-search --no-cpu > gpurun_out/bench2.json 2> gpurun_out/bench2.log; echo rc=$?
B="python bench.py --steps 1 --warmup 0 --no-search --no-cpu --no-exact --no-e2e"
timeout 300 $B > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stats_big_q30 -c 2 -o gpurun_out/prof_big $B > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stats_small_q30 -s 6 -c 1 -o gpurun_out/prof_small_l16 $B > gpurun_out/ncu_s1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stats_small_q30 -s 12 -c 1 -o gpurun_out/prof_small_l22 $B > gpurun_out/ncu_s2.log 2>&1
ls -la gpurun_out
