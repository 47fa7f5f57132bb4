#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m pytest tests -m gpu -q --no-header -rf -x 2>&1 | tail -30 > gpurun_out/r2_tests4.log
tail -8 gpurun_out/r2_tests4.log
python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench_cfg4_1.json 2> gpurun_out/r2_bench_cfg4_1.err; tail -3 gpurun_out/r2_bench_cfg4_1.err; cat gpurun_out/r2_bench_cfg4_1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2_bench_cfg4_2.json 2> gpurun_out/r2_bench_cfg4_2.err; tail -5 gpurun_out/r2_bench_cfg4_2.err; cat gpurun_out/r2_bench_cfg4_2.json
