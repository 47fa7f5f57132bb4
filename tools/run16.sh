#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf -x 2>&1 | tail -30 > gpurun_out/r2_tests16.log
tail -4 gpurun_out/r2_tests16.log
for ov in 0 1; do for wl in cfg3 cfg4; do echo "direct overlap $ov $wl"; MV_DIRECT_OVERLAP=$ov MV_NOSTATS=1 python tools/pass_times.py $wl 60; done; done
for ov in 0 1; do MV_DIRECT_OVERLAP=$ov python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('overlap $ov cfg4 N=1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v,3) for k,v in d['per_pass_ms'].items()}, 'tex', round(d['roofline']['frac'],3))"; done
