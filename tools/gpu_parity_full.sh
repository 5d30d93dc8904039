#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "full_size or sweep_records" > gpurun_out/pytest_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_full.log
tail -30 gpurun_out/pytest_full.log
