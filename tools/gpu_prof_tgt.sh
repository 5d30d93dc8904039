#!/bin/bash
mkdir -p gpurun_out
P="python tools/prof_small.py ${1:-512}"
timeout 300 $P > gpurun_out/prof_small_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rpn_targets_kernel' -s 2 -c 1 -o gpurun_out/prof_tgt $P > gpurun_out/ncu_tgt.log 2>&1
tail -2 gpurun_out/ncu_tgt.log
