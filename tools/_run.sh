for wl in cfg4 cfg3; do echo "== $wl"; MV_NOSTATS=1 bash tools/ncu_times.sh "k_" 400 13 python tools/pass_times.py $wl 10 | grep gpu__time; done 2>&1 | tee gpurun_out/s6_ncu_cfg4_times.log
