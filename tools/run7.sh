#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf 2>&1 | tail -30 > gpurun_out/r2_tests7.log
tail -12 gpurun_out/r2_tests7.log
for wl in cfg2 cfg4; do MV_NOSTATS=1 python tools/pass_times.py $wl 60; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_resolve_oit|k_postprocess" -s 40 -c 2 -f -o gpurun_out/r2_cfg4_v2 python tools/pass_times.py cfg4 2 > gpurun_out/r2_ncu7.log 2>&1
