#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_targets.py tests/test_gpu_parity.py tests/test_gpu_random_shapes.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_c.log
RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so timeout 300 python tools/tgt_phase_profile.py 64 > gpurun_out/tgt_phase_64.log 2>&1
RADNET_TARGETS_TWO_LAUNCHES=1 RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so timeout 300 python tools/tgt_phase_profile.py 64 > gpurun_out/tgt_phase_64_two.log 2>&1
RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so timeout 300 python tools/tgt_phase_profile.py 512 > gpurun_out/tgt_phase_512.log 2>&1
tail -4 gpurun_out/pytest_gpu_c.log; tail -16 gpurun_out/tgt_phase_64.log; tail -16 gpurun_out/tgt_phase_64_two.log; tail -3 gpurun_out/tgt_phase_512.log
