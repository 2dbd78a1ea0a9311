#!/bin/bash
# round 2, job J: bulk-async ring of the warp-per-range kernel: parity + timing, and the small all-paths parity script
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_small.py > gpurun_out/r2_small_ring0.log 2>&1; echo "small ring0 rc=$?"; tail -2 gpurun_out/r2_small_ring0.log
VI_B200_RING=1 timeout 300 python scripts/sanitize_small.py > gpurun_out/r2_small_ring1.log 2>&1; echo "small ring1 rc=$?"; tail -3 gpurun_out/r2_small_ring1.log
VI_B200_RING=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "small_shapes or config1 or unit_gaussian or mid_size or duplicates or skewed or one_million" > gpurun_out/r2_tJ.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tJ.log; tail -4 gpurun_out/r2_tJ.log
for r in 0 1; do
  VI_B200_RING=$r SWEEP_T_BIG=512 timeout 200 python scripts/sweep_tbig.py 2>&1 | sed "s/^/RING=$r /" | tee -a gpurun_out/r2_sweep_ring.log
done
