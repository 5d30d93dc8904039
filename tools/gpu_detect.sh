#!/bin/bash
# GPU check of the detection post-processing kernels (f1/f2): parity tests, single-GPU sharded-path check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_detect.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_detect.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_detect.log
tail -30 gpurun_out/pytest_detect.log
timeout 600 python tools/check_sharded_detect.py --panels 1 2>&1 | tail -3
