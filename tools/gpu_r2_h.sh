#!/bin/bash
mkdir -p gpurun_out
P="python tools/prof_pool_forms.py"
timeout 300 $P > gpurun_out/prof_forms_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'roi_pool' -c 9 -o gpurun_out/prof_forms $P > gpurun_out/ncu_forms.log 2>&1
cat gpurun_out/prof_forms_plain.log; tail -3 gpurun_out/ncu_forms.log
