#!/bin/bash
# ncu evidence of a round: (1) per-launch device times of the bench command, (2) one full-metric capture of a cfg4 frame.
# Run on the GPU box; then tools/ncu_summarise.py gpurun_out/r02_cfg4_frame.ncu-rep cfg4 writes profiles/ncu_summary.json.
mkdir -p gpurun_out
# (1) launch list of the bench command (per-launch device time, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_launch_bench.log 2>&1
tail -2 gpurun_out/r2_launch_bench.log | cut -c1-300
# (2) one frame's kernels of cfg4 with the full metric set
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ray_march_v|k_ray_cast_direct|k_resolve_oit|k_postprocess|k_ray_march_l|k_light_ao|k_light_classify|k_light_emit|k_light_finalize|k_cull|k_environment" -s 330 -c 11 -f -o gpurun_out/r02_cfg4_frame python tools/pass_times.py cfg4 2 > gpurun_out/r2_ncu14.log 2>&1
tail -3 gpurun_out/r2_ncu14.log | cut -c1-300
ls -la gpurun_out/r02_cfg4_frame.ncu-rep
