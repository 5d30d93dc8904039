#!/bin/bash
# GPU check of the detection post-processing kernels (f1/f2): parity tests under a timeout, then memcheck
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_detect.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_detect.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_detect.log
tail -40 gpurun_out/pytest_detect.log
