#!/bin/bash
# Round 2, call 31 (1 GPU): the GPU suite against the widened small-graph bands; the parity file three more times (flakiness).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_c31_pytest_gpu.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py tests/test_gpu_host.py -q -m gpu 2>&1 | tail -3; done > gpurun_out/r02_c31_pytest_repeat.log 2>&1
cat gpurun_out/r02_c31_pytest_gpu.log gpurun_out/r02_c31_pytest_repeat.log; uptime
