#!/bin/bash
# round 2, job D: the whole -m gpu suite, the full default bench line, the ncu launch list of the same command
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_tD.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tD.log
tail -8 gpurun_out/r2_tD.log
( time timeout 900 python bench.py ) > gpurun_out/r2_bD.json 2> gpurun_out/r2_bD.log
echo "bench rc=$?"
grep -E "build:|search|e2e:|cpu baseline|divergence|sql-mode" gpurun_out/r2_bD.log | cut -c1-400 | tail -20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-exact --no-e2e --no-search > gpurun_out/r2_ncu_list.log 2>&1
echo "ncu list rc=$?"
wc -l gpurun_out/r2_launches.csv
