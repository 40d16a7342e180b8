#!/bin/bash
# Round 2, first 2-GPU call: the LPA 1D peer-mode fix (hub margin 2 in peer phases), the single-process path, path-sharded records.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_runs/r02_first_2gpu.sh'
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_multi.py tests/test_gpu_host.py tests/test_gpu_more.py -q -k "two_rank or single_process or two_gpus" 2>&1 | tail -40 > gpurun_out/r02_pytest_2gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --multi sharded \
    > gpurun_out/r02_bench_c4_n2_sharded.json 2> gpurun_out/r02_bench_c4_n2_sharded.err
tail -12 gpurun_out/r02_pytest_2gpu.log; head -c 700 gpurun_out/r02_bench_c4_n2_sharded.json
