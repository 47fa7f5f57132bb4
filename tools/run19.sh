#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/d2h_probe.py 4 2>&1 | grep "ranks x"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/r2_s19_n8.err | grep "^{" > gpurun_out/r2_s19_n8.json
python -c "
import json; d=json.load(open('gpurun_out/r2_s19_n8.json')); print('n8', round(d['value'],1), 'fps  e2e', round(d['e2e']['value'],1), 'blocking', round(d['e2e']['blocking_readback_value'],1), 'checksum', d['e2e']['checksum'])"
