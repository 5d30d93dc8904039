#!/bin/bash
# round 2, call B: targets + sampling tests, K3 phase stamps at several compute-SM counts, K3 microbench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_targets.py tests/test_gpu_sampling.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_b.log
for n in 0 24 40; do
RADNET_TARGETS_COMPUTE_CTAS=$n RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so timeout 300 python tools/tgt_phase_profile.py 64 > gpurun_out/tgt_phase_64_c$n.log 2>&1
done
RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so timeout 300 python tools/tgt_phase_profile.py 512 > gpurun_out/tgt_phase_512.log 2>&1
tail -5 gpurun_out/pytest_gpu_b.log; for n in 0 24 40; do tail -14 gpurun_out/tgt_phase_64_c$n.log; done; grep "event time" gpurun_out/tgt_phase_512.log
