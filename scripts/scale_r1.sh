set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -8
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2970$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.log; echo rc=$?
  grep "GPUs\]" gpurun_out/bench_g$N.log | cut -c1-200
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.log; echo rc=$?
grep -E "fast build|exact build|search|e2e:" gpurun_out/bench_g1.log | cut -c1-250
