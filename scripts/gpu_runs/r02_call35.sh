#!/bin/bash
# Round 2, call 35 (1 GPU): the GPU suite and smoke on the final commit (final bands).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_c35_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c35_smoke.log 2>&1
cat gpurun_out/r02_c35_pytest_gpu.log; tail -2 gpurun_out/r02_c35_smoke.log
