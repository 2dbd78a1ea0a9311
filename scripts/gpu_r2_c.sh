#!/bin/bash
# round 2, job C: SQL mode + HDF5 ingest + search paths on the GPU, search probe after the kernel trims
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sql_mode.py tests/test_hdf5.py tests/test_golden.py -x -q -m gpu > gpurun_out/r2_tC.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tC.log
tail -15 gpurun_out/r2_tC.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "search or topk or csv or find_entry or textindex" > gpurun_out/r2_tC2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tC2.log
tail -5 gpurun_out/r2_tC2.log
timeout 300 python scripts/search_probe.py > gpurun_out/r2_search_probe2.log 2>&1; echo "probe rc=$?"
VI_B200_SEARCH_PATH=1 VI_B200_SEARCH_POOL=0 timeout 300 python scripts/search_probe.py >> gpurun_out/r2_search_probe2.log 2>&1
cat gpurun_out/r2_search_probe2.log
