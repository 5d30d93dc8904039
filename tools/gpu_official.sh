#!/bin/bash
# round artefacts: smoke, bench (both arms), ncu launch list + full captures of K4 / K2 on the bench command
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
PROF="python bench.py --steps 2 --warmup 3 --panels 16 --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1
timeout 300 $PROF > gpurun_out/prof_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:roi_pool_slice -s 3 -c 1 -o gpurun_out/prof_pool $PROF > gpurun_out/ncu_pool.log 2>&1
tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_reference.log; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
