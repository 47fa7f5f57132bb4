#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf -x 2>&1 | tail -30 > gpurun_out/r2_tests3.log
tail -3 gpurun_out/r2_tests3.log
for b in 0 16 32 64; do for wl in cfg2 cfg3 cfg4; do echo "bricks=$b $wl"; MV_OCC_BRICKS=$b python tools/pass_times.py $wl 40; MV_OCC_BRICKS=$b MV_NOSTATS=1 python tools/pass_times.py $wl 40; done; done > gpurun_out/r2_occ_ab.log 2>&1
cat gpurun_out/r2_occ_ab.log
