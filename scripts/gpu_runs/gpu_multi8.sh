#!/bin/bash
mkdir -p gpurun_out
N=$1
for m in hybrid allreduce; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 27 --warmup 3 --no-e2e --multi $m > gpurun_out/bench_c4_n${N}_$m.json 2> gpurun_out/bench_c4_n${N}_$m.err; cut -c1-160 gpurun_out/bench_c4_n${N}_$m.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_n${N}_$m.json')); print(d.get('quality'))"; grep -iE "error|Traceback" -A5 gpurun_out/bench_c4_n${N}_$m.err | head -20
done
