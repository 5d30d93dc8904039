#!/bin/bash
# final refresh of the round-2 records on ONE GPU: tests, bench (both lengths), kernel microbench, launch list + K4 capture
mkdir -p gpurun_out
rm -f gpurun_out/prof_pool.ncu-rep
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err
timeout 600 python tools/bench_kernels.py --out gpurun_out/kernels.json > gpurun_out/kernels.log 2>&1
FORM=$(python -c "
import json
f=json.loads([l for l in open('gpurun_out/bench.log') if l.startswith('{')][-1])['roofline'].get('launch_form') or {}
print('%d,%d' % (f.get('cluster', 0), f.get('sync_every', 0)))")
echo "K4 launch form (cluster,sync_every): $FORM"
PROF="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-extra --k4-lockstep $FORM"
timeout 300 $PROF > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1
timeout 300 $PROF > gpurun_out/prof_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:roi_pool_slice -s 3 -c 1 -o gpurun_out/prof_pool $PROF > gpurun_out/ncu_pool.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; tail -1 gpurun_out/bench.err; tail -2 gpurun_out/ncu_pool.log | cut -c1-200
