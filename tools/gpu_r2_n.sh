#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_shapes.py -m gpu -x -q -p no:cacheprovider -k "rpn_to_roi or decode or apply_regr or random or full_size" 2>&1 | tail -2
timeout 600 python tools/bench_kernels.py --out gpurun_out/kernels_n.json > gpurun_out/kernels_n.log 2>&1
python -c "
import json; k=json.load(open('gpurun_out/kernels_n.json'))
for n in k:
    if 'decode' in n or 'rpn_to_roi' in n: print(n, k[n] if not isinstance(k[n],dict) else (round(k[n]['p50_ms'],4), round(k[n].get('frac_of_measured_peak',0),4)))"
