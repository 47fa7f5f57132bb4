#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf -x 2>&1 | tail -30 > gpurun_out/r2_tests6.log
tail -12 gpurun_out/r2_tests6.log
for wl in cfg2 cfg4; do MV_NOSTATS=1 python tools/pass_times.py $wl 60; done
