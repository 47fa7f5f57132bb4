#!/bin/bash
# strong-scaling sweep as the driver launches it: N = 1, 2, 4, 8 on one box
WLS=${WLS:-"cfg4"}
for wl in $WLS; do
for n in ${NS:-1 2 4 8}; do
  if [ $n -eq 1 ]; then cmd="python bench.py --gpus 1"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n"; fi
  $cmd --steps ${STEPS:-100} --warmup 20 --workload $wl --no-cpu-baseline ${EXTRA} 2>gpurun_out/scale_${wl}_$n.err | grep "^{" > gpurun_out/scale_${wl}_$n.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_${wl}_$n.json")); print("$wl N=$n", round(d["value"],1), "fps  e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d["per_pass_ms"].items()})
except Exception as e:
    print("$wl N=$n FAILED", e); print(open("gpurun_out/scale_${wl}_$n.err").read()[-1500:])
PY
done; done
