#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/bench_c4_n2_default.json 2> gpurun_out/bench_c4_n2_default.err; cat gpurun_out/bench_c4_n2_default.json | cut -c1-1800; grep -iE "error|Traceback" -A8 gpurun_out/bench_c4_n2_default.err | head -30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 12 --warmup 3 > gpurun_out/bench_ref_n2.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_n2.json
