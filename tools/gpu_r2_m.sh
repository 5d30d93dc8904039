#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_quick.log') if l.startswith('{')][-1])
print(round(d['value']), d['ms_per_step'], d['kernels_ms_per_step'], round(d['roofline']['frac'],4), d['roofline']['launch_form'], round(d['e2e']['value']), 'sweep', round(d['sweep']['value']), 'tiled', round(d['tiled']['value'],1))
PY
