#!/bin/bash
# round 2, job L (8 GPUs): the bench line under torchrun at N = 8 (and its reference arm: rank 0 only)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bL8.json 2> gpurun_out/r2_bL8.log
echo "bench 8 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bL8.json"))
print("N=8 ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],1), "checksum", d["config"]["table_checksum"], "replicate", round(d["replicate_ms"],1), "search", {k:(round(v["queries_per_sec"]/1e6,1) if isinstance(v,dict) else v) for k,v in (d["search"] or {}).items()}, "launches", d["gpu_launches"])
PY
tail -5 gpurun_out/r2_bL8.log | cut -c1-300
