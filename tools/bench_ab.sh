#!/bin/bash
# uninstrumented frames/s of bench.py under environment variants: tools/bench_ab.sh cfg2 "MV_OVERLAP=0" "MV_OVERLAP=1 MV_OVERLAP_V_BLOCKS=4"
wl=$1; shift
for envs in "$@"; do
  env $envs python bench.py --workload $wl --steps ${STEPS:-150} --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$wl %-44s fps %.1f  e2e %.1f  ms %.4f' % ('$envs', d['value'], d['e2e']['value'], d['ms_per_step']))"
done
