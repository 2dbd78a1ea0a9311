#!/bin/bash
# round 2, job B: SQL mode parity, search probe (thread vs warp), ncu of the traversal kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sql_mode.py tests/test_golden.py -x -q -m gpu > gpurun_out/r2_tB.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tB.log
tail -15 gpurun_out/r2_tB.log
timeout 300 python scripts/search_probe.py > gpurun_out/r2_search_probe.log 2>&1; echo "probe rc=$?"
VI_B200_SEARCH_PATH=0 timeout 300 python scripts/search_probe.py >> gpurun_out/r2_search_probe.log 2>&1
VI_B200_SEARCH_PATH=1 VI_B200_SEARCH_POOL=0 timeout 300 python scripts/search_probe.py >> gpurun_out/r2_search_probe.log 2>&1
cat gpurun_out/r2_search_probe.log
# ncu: the traversal kernels of one probe run (reps = 0: only the two warm-up searches per p)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_search' -c 12 -o gpurun_out/r2_search_ncu -f \
  python scripts/search_probe.py 10000000 96 1000000 0 > gpurun_out/r2_search_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_search_ncu.log
ls -la gpurun_out/r2_search_ncu.ncu-rep
