#!/bin/bash
# one gpurun call: parity tests, smoke, bench, ncu launch list + one full capture of the pool kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
PROF="python bench.py --steps 2 --warmup 3 --panels 16 --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1
timeout 300 $PROF > gpurun_out/prof_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:roi_pool_slice -s 3 -c 1 -o gpurun_out/prof_pool $PROF > gpurun_out/ncu_pool.log 2>&1
timeout 300 $PROF > gpurun_out/prof_plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sort_nms -s 3 -c 1 -o gpurun_out/prof_nms $PROF > gpurun_out/ncu_nms.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
