#!/bin/bash
# round 2, job F: SQL search test, BASELINE configs[4] (1M x 768) bench + launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sql_mode.py -x -q -m gpu > gpurun_out/r2_tF.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tF.log
tail -4 gpurun_out/r2_tF.log
( time timeout 900 python bench.py --rows 1000000 --dims 768 ) > gpurun_out/r2_b768.json 2> gpurun_out/r2_b768.log
echo "bench768 rc=$?"
grep -E "build:|search|e2e:|cpu baseline|sql-mode" gpurun_out/r2_b768.log | cut -c1-600 | tail -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_768.csv \
  python bench.py --rows 1000000 --dims 768 --steps 1 --warmup 3 --no-cpu --no-exact --no-e2e --no-search > gpurun_out/r2_ncu_list768.log 2>&1
echo "ncu list rc=$?"
