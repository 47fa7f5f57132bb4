#!/bin/bash
run() { python tools/pass_times.py ${WL:-cfg2} | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ray_march_light'], d['ray_march_view'], d['resolve_oit'], d['postprocess'], d['total'])"; }
MV_NOSTATS=1 run "mb4 nostats"
for v in 5 6; do MV_NOSTATS=1 MV_B200_LIB=$PWD/multivolumes_b200/libmv_var_mb$v.so run "mb$v nostats"; MV_B200_LIB=$PWD/multivolumes_b200/libmv_var_mb$v.so run "mb$v stats"; done
