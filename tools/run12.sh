#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf 2>&1 | tail -30 > gpurun_out/r2_tests12.log
tail -12 gpurun_out/r2_tests12.log
for wl in cfg2 cfg3 cfg4; do MV_NOSTATS=1 python tools/pass_times.py $wl 60; done
python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; tail -3 gpurun_out/r2_bench12.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench12.json') if l.startswith('{')][0]); print('cfg4 N=1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['checksum'], {k: round(v,3) for k,v in d['per_pass_ms'].items()}, 'tex', round(d['roofline']['frac'],3), 'parity', d.get('parity_checked'), d.get('parity'))"
