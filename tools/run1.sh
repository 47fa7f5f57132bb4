#!/bin/bash
# round 2, GPU call 1: both builds through the parity suite, strict-vs-fast per-pass times, ncu captures of the hot kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf 2>&1 | tail -40 > gpurun_out/r2_tests.log
for s in 1 0; do for wl in cfg2 cfg3 cfg4; do echo "strict=$s $wl"; MV_STRICT_FP=$s MV_NOSTATS=1 python tools/pass_times.py $wl 60; done; done > gpurun_out/r2_ab_strict_fast.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_resolve_oit|k_postprocess|k_ray_march_v|k_ray_cast_direct" -s 80 -c 4 -f -o gpurun_out/r2_cfg4_fast python tools/pass_times.py cfg4 2 > gpurun_out/r2_ncu1.log 2>&1
tail -5 gpurun_out/r2_tests.log; cat gpurun_out/r2_ab_strict_fast.log
