python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for vb in 0 2 3 4; do echo "MV_SHARD_V_BLOCKS=$vb"; MV_SHARD_V_BLOCKS=$vb NS="2" WLS="cfg2 cfg4" STEPS=100 bash tools/scale.sh 2>&1; done | tee gpurun_out/s6_scale2_overlap.log
NS="1" WLS="cfg2" STEPS=150 bash tools/scale.sh 2>&1 | tee -a gpurun_out/s6_scale2_overlap.log
