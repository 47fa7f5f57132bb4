python -m pytest tests -m gpu -x -q -k "light or golden or frame_parity or density or pipelined" 2>&1 | tail -3
MV_LIGHT_TIMING=1 FRAMES=64 bash tools/ab.sh "noagg base" "cfg2 cfg4" 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/s6_emit_agg.log
FRAMES=40 bash tools/ab.sh "noagg base" "cfg5" 2>&1 | tee -a gpurun_out/s6_emit_agg.log
