#!/bin/bash
mkdir -p gpurun_out
python tools/exp_pool.py base 1:0:0:0 1:0:0:0:0:0:148 1:0:0:0:0:0:128 1:0:0:0:0:0:96 1:0:0:0:0:0:64 1:0:0:0:0:0:144 1:0:0:0:0:0:136 1:0:4:0:0:0:128 1:0:4:0:0:0:148 \
  64,1024,14,38,38 32,1024,14,38,50 64,512,7,38,38 > gpurun_out/exp_pool.log 2>&1
cat gpurun_out/exp_pool.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "roi_pool or pool" 2>&1 | tail -3
