#!/bin/bash
# round artefacts: GPU tests, smoke, bench (both arms), kernel microbench, ncu launch lists + full captures
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err
timeout 900 python tools/bench_kernels.py --out gpurun_out/kernels.json > gpurun_out/kernels.log 2>&1
# launch list + full capture of the dominant kernel on the bench command itself (64 panels)
PROF="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1
timeout 300 $PROF > gpurun_out/prof_plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:roi_pool_slice -s 3 -c 1 -o gpurun_out/prof_pool $PROF > gpurun_out/ncu_pool.log 2>&1
# the small kernels (decode, sort+NMS, targets) on a 512-panel workload
P="python tools/prof_small.py 512"
timeout 300 $P > gpurun_out/prof_small_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rpn_targets_kernel|decode_clip|sort_nms' -s 3 -c 3 -o gpurun_out/prof_small $P > gpurun_out/ncu_small.log 2>&1
# latency-bound kernels: cluster sort+NMS on one panel, detection post-processing
P="python tools/prof_detect.py"
timeout 300 $P > gpurun_out/prof_detect_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_detect.csv $P > gpurun_out/ncu_launches_detect.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sort_nms|class_nms_kernel|cluster_kernel' -s 8 -c 5 -o gpurun_out/prof_detect $P > gpurun_out/ncu_detect.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_reference.log; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err; tail -2 gpurun_out/ncu_pool.log; tail -2 gpurun_out/ncu_small.log; tail -2 gpurun_out/ncu_detect.log
