#!/bin/bash
mkdir -p gpurun_out
python tools/exp_pool.py base 1:0:0:0 1:0:0:0:0:0 1:0:0:0:0:1 1:0:0:0:2:2 1:0:0:0:2:4 1:0:0:0:2:3 1:0:0:0 > gpurun_out/exp_pool.log 2>&1
cat gpurun_out/exp_pool.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "roi_pool or pool" 2>&1 | tail -3
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --skip-extra > gpurun_out/bench_i.log 2> gpurun_out/bench_i.err; tail -2 gpurun_out/bench_i.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_i.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['kernels_ms_per_step'], d['roofline']['frac'], d['clocks'], d['e2e']['value'])
PY
