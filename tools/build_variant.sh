#!/bin/bash
# experiment build of the library with extra -D flags: tools/build_variant.sh NAME -DFOO=1 ...  ->  _C/libradnet_b200_NAME.so
set -e
NAME=$1; shift
cd "$(dirname "$0")/../rock_art_radnet_b200"
SRC=$(PYTHONPATH=.. python - <<'PY'
from rock_art_radnet_b200.build import SOURCES
print(" ".join("csrc/" + s for s in SOURCES))
PY
)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC -shared "$@" -o _C/libradnet_b200_$NAME.so $SRC
echo built _C/libradnet_b200_$NAME.so
