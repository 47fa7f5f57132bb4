#!/bin/bash
# 1 GPU: full parity suite twice (stability of the shared-GPU multi-rank tests), ncu captures for profiles/, compute-sanitizer
mkdir -p gpurun_out
for k in 1 2; do python -m pytest tests -m gpu -q --no-header -rf 2>&1 | tail -4; done > gpurun_out/r2_tests20.log 2>&1
cat gpurun_out/r2_tests20.log
bash tools/run14.sh
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/r2_memcheck.log 2>&1; tail -3 gpurun_out/r2_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/r2_racecheck.log 2>&1; tail -3 gpurun_out/r2_racecheck.log
