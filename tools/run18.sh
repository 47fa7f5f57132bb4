#!/bin/bash
cd /root/repo
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests18.log 2>&1; tail -5 gpurun_out/r2_tests18.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench18.json 2> gpurun_out/r2_bench18.err; tail -c 500 gpurun_out/r2_bench18.json
