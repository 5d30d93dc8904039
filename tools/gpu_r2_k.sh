#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_kernels.py --out gpurun_out/kernels.json > gpurun_out/kernels.log 2>&1
python -c "
import json; k=json.load(open('gpurun_out/kernels.json'))
for n in k:
    if 'frac_of_measured_peak' in k[n]: print(n, round(k[n]['p50_ms'],4), round(k[n]['frac_of_measured_peak'],4))
    else: print(n, {a:(round(b,4) if isinstance(b,float) else b) for a,b in k[n].items() if a in ('p50_ms','p50_us','ms')})"
