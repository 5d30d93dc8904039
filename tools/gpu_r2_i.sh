#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_shapes.py -m gpu -x -q -p no:cacheprovider -k "roi_pool or pool or full_size or sweep" 2>&1 | tail -3
python tools/exp_pool.py pair 1:0:0:0 1:0:0:0:0:0 1:0:0:0 > gpurun_out/exp_pool.log 2>&1
cat gpurun_out/exp_pool.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_i.log 2> gpurun_out/bench_i.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_i.log').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], d['kernels_ms_per_step'], round(d['roofline']['frac'],4), d['roofline']['launch_form'], d['clocks'])
PY
