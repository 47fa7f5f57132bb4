#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -rf 2>&1 | tail -30 > gpurun_out/r2_tests9.log
tail -15 gpurun_out/r2_tests9.log
for mode in pipelined serial; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 tools/multi_check.py cfg4 6 $mode 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12
done
