#!/bin/bash
# Round 2, call 32 (2 GPUs): the multi-GPU tests on the final library and bands; one N=2 bench line (auto).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_host.py tests/test_gpu_more.py -q -k "two_rank or single_process or two_gpus or sharded" 2>&1 | tail -12 > gpurun_out/r02_c32_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_c32_bench_c4_n2.json 2> gpurun_out/r02_c32_bench_c4_n2.err
cat gpurun_out/r02_c32_pytest_2gpu.log; head -c 1500 gpurun_out/r02_c32_bench_c4_n2.json; tail -3 gpurun_out/r02_c32_bench_c4_n2.err
