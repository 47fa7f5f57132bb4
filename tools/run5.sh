#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf 2>&1 | tail -30 > gpurun_out/r2_tests5.log
tail -12 gpurun_out/r2_tests5.log
for pipe in 1 0; do
MV_SHARD_PIPELINE=$pipe python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2_bench_cfg4_2_p$pipe.json 2> gpurun_out/r2_bench_cfg4_2_p$pipe.err; tail -5 gpurun_out/r2_bench_cfg4_2_p$pipe.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_cfg4_2_p$pipe.json')); print('pipe=$pipe N=2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'blocking', round(d['e2e']['blocking_readback_value'],1), d['e2e']['checksum'], {k: round(v,3) for k,v in d['per_pass_ms'].items()})"
done
