#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_baseline.py -m gpu -q --no-header -rf -k "volume_sharded" 2>&1 | tail -30 > gpurun_out/r2_tests10.log
tail -25 gpurun_out/r2_tests10.log
MV_STEP=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 tools/multi_check.py cfg4 90 pipelined 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 60 --warmup 10 --workload cfg5s-mini --no-cpu-baseline 2>gpurun_out/r2_cfg5smini.err | grep "^{" > gpurun_out/r2_cfg5smini.json; tail -3 gpurun_out/r2_cfg5smini.err; python -c "
import json; d=json.load(open('gpurun_out/r2_cfg5smini.json')); print('cfg5s-mini N=2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v,3) for k,v in d['per_pass_ms'].items()})"
