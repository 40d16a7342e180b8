#!/bin/bash
# Round 2, call 14 (2 GPUs): the multi-GPU tests on the final library (single mode, fixed sharded test, in-process path).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_host.py tests/test_gpu_more.py -q -k "two_rank or single_process or two_gpus or sharded" 2>&1 | tail -30 > gpurun_out/r02_c14_pytest_2gpu.log
tail -12 gpurun_out/r02_c14_pytest_2gpu.log
