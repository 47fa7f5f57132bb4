#!/bin/bash
# N = 4, 8 on one box (cfg4 headline + cfg2), fused exchange, pipelined sharded frames
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for wl in cfg4 cfg2; do for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 100 --warmup 10 --workload $wl --no-cpu-baseline 2>gpurun_out/r2_scale_${wl}_$n.err | grep "^{" > gpurun_out/r2_scale_${wl}_$n.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_${wl}_$n.json")); print("$wl N=$n", round(d["value"],1), "fps  e2e", round(d["e2e"]["value"],1), "blocking", round(d["e2e"]["blocking_readback_value"],1), "checksum", d["e2e"]["checksum"], {k: round(v,3) for k,v in d["per_pass_ms"].items()})
except Exception as e:
    print("$wl N=$n FAILED", e); print(open("gpurun_out/r2_scale_${wl}_$n.err").read()[-1500:])
PY
done; done
