#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "roi_pool or pool" > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_g.log
timeout 600 python tools/bench_kernels.py --out gpurun_out/kernels.json > gpurun_out/kernels.log 2>&1
tail -15 gpurun_out/pytest_gpu_g.log; tail -3 gpurun_out/kernels.log
python -c "
import json; k=json.load(open('gpurun_out/kernels.json'))
for n in k:
    if 'roi_pool' in n or 'decode_clip' in n: print(n, round(k[n]['p50_ms'],4), round(k[n]['frac_of_measured_peak'],4))"
