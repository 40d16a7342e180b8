#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/pytest_multi.log 2>&1; tail -8 gpurun_out/pytest_multi.log | cut -c1-600
for m in hybrid; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 27 --warmup 3 --no-e2e --multi $m > gpurun_out/bench_c4_n${N}_$m.json 2> gpurun_out/bench_c4_n${N}_$m.err; cut -c1-160 gpurun_out/bench_c4_n${N}_$m.json; grep -iE "error|Traceback" -A5 gpurun_out/bench_c4_n${N}_$m.err | head -20
done
