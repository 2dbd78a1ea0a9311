#!/bin/bash
# round 2, job K (2 GPUs): multi-rank parity tests at world 2 + the bench line under torchrun
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu ) > gpurun_out/r2_tK.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tK.log; tail -6 gpurun_out/r2_tK.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bK2.json 2> gpurun_out/r2_bK2.log
echo "bench 2 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bK2.json"))
print("N=2 ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],1), "checksum", d["config"]["table_checksum"], "replicate", round(d["replicate_ms"],1), "search", {k:(round(v["queries_per_sec"]/1e6,1) if isinstance(v,dict) else v) for k,v in (d["search"] or {}).items()}, "launches", d["gpu_launches"])
PY
