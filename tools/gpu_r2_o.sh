#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python tools/bench_kernels.py --out gpurun_out/kernels.json > gpurun_out/kernels.log 2>&1
P="python tools/prof_small.py 512"
timeout 300 $P > gpurun_out/prof_small_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'decode_clip|sort_nms' -s 2 -c 2 -o gpurun_out/prof_small $P > gpurun_out/ncu_small.log 2>&1
tail -1 gpurun_out/ncu_small.log | cut -c1-120
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench exit $?"
