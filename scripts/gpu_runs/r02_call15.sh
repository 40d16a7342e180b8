#!/bin/bash
# Round 2, call 15 (1 GPU): compute-sanitizer over every round-2 device path, then the whole GPU suite and smoke on the final library.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/gpu_sanitize.py > gpurun_out/r02_c15_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_c15_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/gpu_sanitize.py > gpurun_out/r02_c15_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_c15_racecheck.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_c15_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c15_smoke.log 2>&1
tail -8 gpurun_out/r02_c15_memcheck.log; tail -8 gpurun_out/r02_c15_racecheck.log; tail -6 gpurun_out/r02_c15_pytest_gpu.log; tail -2 gpurun_out/r02_c15_smoke.log
