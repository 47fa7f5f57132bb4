#!/bin/bash
run() { python tools/pass_times.py ${WL:-cfg2} | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ray_march_light'], d['ray_march_view'], d['resolve_oit'], d['postprocess'], d['total'])"; }
MV_NOSTATS=1 run "cfg2"
WL=cfg4 MV_NOSTATS=1 run "cfg4"
WL=cfg3 MV_NOSTATS=1 run "cfg3"
