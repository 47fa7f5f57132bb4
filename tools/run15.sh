#!/bin/bash
cd /root/repo
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests15.log 2>&1; tail -5 gpurun_out/r2_tests15.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err; tail -c 600 gpurun_out/r2_bench15.json
