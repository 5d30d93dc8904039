#!/bin/bash
# round 2, call A: GPU tests with the single-launch K3 + batched a4, K3 phase stamps, kernel microbench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so timeout 300 python tools/tgt_phase_profile.py 64 > gpurun_out/tgt_phase_64.log 2>&1
RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so timeout 300 python tools/tgt_phase_profile.py 512 > gpurun_out/tgt_phase_512.log 2>&1
timeout 900 python tools/bench_kernels.py --out gpurun_out/kernels.json > gpurun_out/kernels.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -12 gpurun_out/tgt_phase_64.log; tail -3 gpurun_out/kernels.log
