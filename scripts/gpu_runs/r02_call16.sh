#!/bin/bash
# Round 2, call 16 (1 GPU): memcheck again after padding the 1D coordinate buffer (call 15 found a 16-byte read one double past an odd N).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/gpu_sanitize.py > gpurun_out/r02_c16_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_c16_memcheck.log
timeout 600 python -m pytest tests -q -m gpu -k "1d or sort or order or goodness" 2>&1 | tail -8 > gpurun_out/r02_c16_pytest_1d.log
tail -12 gpurun_out/r02_c16_memcheck.log; tail -4 gpurun_out/r02_c16_pytest_1d.log
