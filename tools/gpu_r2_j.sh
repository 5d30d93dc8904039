#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py tests/test_gpu_random_shapes.py -m gpu -x -q -p no:cacheprovider -k "target or region or rpn or calc" > gpurun_out/pytest_gpu_j.log 2>&1; tail -3 gpurun_out/pytest_gpu_j.log
export RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so
timeout 300 python tools/tgt_phase_profile.py 64 > gpurun_out/tgt_phase_64.log 2>&1; tail -16 gpurun_out/tgt_phase_64.log
for z in -1 8192 16384 0; do echo "fill_bulk=$z"; RADNET_TARGETS_FILL_BULK=$z timeout 300 python tools/tgt_phase_profile.py 64 2>&1 | tail -2; done
for c in 40 56 64; do echo "compute_ctas=$c"; RADNET_TARGETS_COMPUTE_CTAS=$c timeout 300 python tools/tgt_phase_profile.py 64 2>&1 | tail -2; done
echo 512; timeout 300 python tools/tgt_phase_profile.py 512 > gpurun_out/tgt_phase_512.log 2>&1; tail -2 gpurun_out/tgt_phase_512.log
echo 512 plain; RADNET_TARGETS_FILL_BULK=-1 timeout 300 python tools/tgt_phase_profile.py 512 2>&1 | tail -2
