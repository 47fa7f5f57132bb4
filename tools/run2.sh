#!/bin/bash
mkdir -p gpurun_out
python tools/fast_diag.py > gpurun_out/r2_fast_diag.log 2>&1
tail -100 gpurun_out/r2_fast_diag.log
