#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf -x 2>&1 | tail -30 > gpurun_out/r2_tests15.log
tail -6 gpurun_out/r2_tests15.log
for wl in cfg3 cfg4; do python tools/pass_times.py $wl 40; MV_NOSTATS=1 python tools/pass_times.py $wl 60; done
bash tools/run14.sh
