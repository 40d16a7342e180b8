#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 27 --warmup 3 > gpurun_out/bench_c4_n${N}_hybrid_b1.json 2> gpurun_out/bench_c4_n${N}_hybrid_b1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_n${N}_hybrid_b1.json')); print('hybrid', d['value'], d['ms_per_step'], d['quality']['stress_final'], d['e2e']['value'])"; grep -iE "error|Traceback" -A5 gpurun_out/bench_c4_n${N}_hybrid_b1.err | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 27 --warmup 3 --no-e2e --multi allreduce > gpurun_out/bench_c4_n${N}_allreduce_b1.json 2> gpurun_out/bench_c4_n${N}_allreduce_b1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_n${N}_allreduce_b1.json')); print('allreduce', d['value'], d['ms_per_step'], d['quality']['stress_final'])"
