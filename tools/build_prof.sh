#!/bin/bash
# profiling build of the library (phase timestamps in the NMS and target kernels); not used by tests or bench
set -e
cd "$(dirname "$0")/../rock_art_radnet_b200"
SRC=$(PYTHONPATH=.. python - <<'PY'
from rock_art_radnet_b200.build import SOURCES
print(" ".join("csrc/" + s for s in SOURCES))
PY
)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared \
  -DRADNET_NMS_PROFILE -DRADNET_TGT_PROFILE -o _C/libradnet_b200_prof.so $SRC
echo built _C/libradnet_b200_prof.so
