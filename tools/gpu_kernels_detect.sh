#!/bin/bash
# detection-stage GPU tests + per-kernel microbench (no profiler)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_detect.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_detect.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_detect.log
tail -3 gpurun_out/pytest_detect.log
timeout 900 python tools/bench_kernels.py --out gpurun_out/kernels.json > gpurun_out/kernels.log 2>&1; echo "kernels exit $?"; grep -i "classify\|final_nms\|class_nms\|Traceback\|Error" gpurun_out/kernels.log | cut -c1-400
