python -m pytest tests -m gpu -x -q -k "light" 2>&1 | tail -4
FRAMES=60 bash tools/ab.sh "base" "cfg2 cfg4 cfg5" 2>&1 | tee gpurun_out/s6_cluster.log
