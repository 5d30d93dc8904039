#!/bin/bash
mkdir -p gpurun_out
P="python tools/prof_small.py 64"
timeout 300 $P > gpurun_out/prof_small_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rpn_targets|decode_clip|sort_nms' -s 9 -c 4 -o gpurun_out/prof_small $P > gpurun_out/ncu_small.log 2>&1
tail -3 gpurun_out/ncu_small.log
