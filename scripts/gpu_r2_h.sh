#!/bin/bash
# round 2, job H: build_copy + reference fixtures + wide-row defaults on the GPU; e2e with the overlapped D2H
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_reference_fixtures.py tests/test_gpu_parity.py -x -q -m gpu -k "reference_fixture or build_copy or wide_rows or config5 or small_shapes or mid_size" > gpurun_out/r2_tH.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tH.log
tail -6 gpurun_out/r2_tH.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-exact --no-search > gpurun_out/r2_bH.json 2> gpurun_out/r2_bH.log
echo "bench rc=$?"; grep -E "e2e|fast build" gpurun_out/r2_bH.log | tail -8
timeout 600 python bench.py --rows 1000000 --dims 768 --steps 5 --warmup 3 --no-cpu --no-exact --no-search > gpurun_out/r2_bH768.json 2> gpurun_out/r2_bH768.log
echo "bench768 rc=$?"; grep -E "e2e|fast build" gpurun_out/r2_bH768.log | tail -4
