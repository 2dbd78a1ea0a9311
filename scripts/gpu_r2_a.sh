#!/bin/bash
# round 2, job A: search paths + mid-range derivation parity, T_SLOT sweep, search bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "search or topk or mid_size or config1 or unit_gaussian or wide_rows or skewed or duplicates or find_entry or csv" > gpurun_out/r2_tA.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tA.log
tail -15 gpurun_out/r2_tA.log
for ts in 512 256 128 64; do
  VI_B200_T_SLOT=$ts SWEEP_T_BIG=512 timeout 120 python scripts/sweep_tbig.py 2>&1 | sed "s/^/T_SLOT=$ts /" | tee -a gpurun_out/r2_sweep_tslot.log
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-exact > gpurun_out/r2_bA.json 2> gpurun_out/r2_bA.log
echo "bench rc=$?"
grep -E "search|e2e|fast build" gpurun_out/r2_bA.log | tail -12
