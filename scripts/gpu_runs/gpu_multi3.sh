#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/pytest_multi.log 2>&1; tail -15 gpurun_out/pytest_multi.log | cut -c1-700
for m in peer allreduce; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-e2e --multi $m > gpurun_out/bench_c4_n2_$m.json 2> gpurun_out/bench_c4_n2_$m.err; cut -c1-160 gpurun_out/bench_c4_n2_$m.json; grep -iE "error|Traceback" -A5 gpurun_out/bench_c4_n2_$m.err | head -20
done
