#!/bin/bash
mkdir -p gpurun_out
PROF="python bench.py --steps 2 --warmup 3 --panels 16 --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/prof_plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sort_nms -s 3 -c 1 -o gpurun_out/prof_nms2 $PROF > gpurun_out/ncu_nms2.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu_dur.log 2>&1
tail -15 gpurun_out/pytest_gpu_dur.log
