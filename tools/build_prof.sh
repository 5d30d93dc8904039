#!/bin/bash
# profiling build of the library (phase timestamps in the NMS kernel); not used by tests or bench
set -e
cd "$(dirname "$0")/../rock_art_radnet_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared \
  -DRADNET_NMS_PROFILE -o _C/libradnet_b200_prof.so csrc/capi.cu csrc/decode.cu csrc/sort_nms.cu csrc/roipool.cu csrc/targets.cu csrc/detect.cu
echo built _C/libradnet_b200_prof.so
