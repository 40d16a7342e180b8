#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/pytest_multi.log 2>&1; tail -12 gpurun_out/pytest_multi.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-e2e > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err; cut -c1-250 gpurun_out/bench_c4_n2.json; tail -3 gpurun_out/bench_c4_n2.err
