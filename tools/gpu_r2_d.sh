#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/prof_tgt2.ncu-rep
export RADNET_TARGETS_TWO_LAUNCHES=1
timeout 300 python tools/prof_small.py 64 > gpurun_out/prof_tgt2_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rpn_targets_kernel -s 5 -c 1 -o gpurun_out/prof_tgt2 python tools/prof_small.py 64 > gpurun_out/ncu_tgt2.log 2>&1
tail -3 gpurun_out/ncu_tgt2.log
