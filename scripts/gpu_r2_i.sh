#!/bin/bash
# round 2, job I: the whole -m gpu suite on one GPU, smoke(), the default bench line
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu ) > gpurun_out/r2_tI.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tI.log
tail -8 gpurun_out/r2_tI.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r2_bI.json 2> gpurun_out/r2_bI.log
echo "bench rc=$?"
grep -E "build:|search|e2e:|cpu baseline|sql-mode" gpurun_out/r2_bI.log | cut -c1-300 | tail -12
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2_bI_ref.json 2> gpurun_out/r2_bI_ref.log
echo "ref rc=$?"; tail -2 gpurun_out/r2_bI_ref.log | cut -c1-300; cut -c1-400 gpurun_out/r2_bI_ref.json
