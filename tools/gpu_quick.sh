#!/bin/bash
# quick GPU check: parity tests + short bench (no profiler)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench exit $?" >> gpurun_out/bench_quick.err
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_quick.log; tail -5 gpurun_out/bench_quick.err
