#!/bin/bash
cd multivolumes_b200/csrc && make -j8 > /dev/null 2>&1; cd ../..
for cap in 0 4 3; do echo "MV_VIEW_BLOCKS=$cap"; MV_VIEW_BLOCKS=$cap python tools/overlap_probe.py cfg4 150; done
