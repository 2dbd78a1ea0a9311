#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_multi_t8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_multi_t8.log
tail -4 gpurun_out/r2_multi_t8.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-exact --no-e2e --no-search > gpurun_out/r2_multi_b1.json 2> gpurun_out/r2_multi_b1.log
python -c "
import json; d=json.load(open('gpurun_out/r2_multi_b1.json')); print('N=1 ms/step', round(d['ms_per_step'],2), 'checksum', d['config']['table_checksum'])"
for G in 2 4 8; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/r2_multi_b${G}.json 2> gpurun_out/r2_multi_b${G}.log
    echo "bench $G rc=$?"
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_multi_b${G}.json"))
    print("N=${G}", "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],1), "checksum", d["config"]["table_checksum"], "replicate", round(d["replicate_ms"],1), "search", {k:(round(v["queries_per_sec"]/1e6,1) if isinstance(v,dict) else v) for k,v in (d["search"] or {}).items()}, "launches", d["gpu_launches"])
except Exception as e: print("parse failed", e)
PY
done
