#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_baseline.py -m gpu -q --no-header -rf -k "volume_sharded" 2>&1 | tail -8
MV_STEP=3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/multi_check.py cfg4 24 pipelined 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6
for wl in cfg4 cfg5s; do
  steps=100; [ $wl = cfg5s ] && steps=40
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps $steps --warmup 10 --workload $wl --no-cpu-baseline 2>gpurun_out/r2_s11_${wl}_8.err | grep "^{" > gpurun_out/r2_s11_${wl}_8.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_s11_${wl}_8.json")); print("$wl N=8", round(d["value"],1), "fps  e2e", round(d["e2e"]["value"],1), "blocking", round(d["e2e"]["blocking_readback_value"],1), "checksum", d["e2e"]["checksum"], {k: round(v,3) for k,v in d["per_pass_ms"].items()})
except Exception as e:
    print("$wl N=8 FAILED", e); print(open("gpurun_out/r2_s11_${wl}_8.err").read()[-2500:])
PY
done
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -2
