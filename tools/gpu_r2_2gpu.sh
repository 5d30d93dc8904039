#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded_detect.py --panels 5 > gpurun_out/sharded_detect_${N}gpu.log 2>&1; echo "exit $?" >> gpurun_out/sharded_detect_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "exit $?" >> gpurun_out/bench_${N}gpu.err
tail -3 gpurun_out/sharded_detect_${N}gpu.log; tail -3 gpurun_out/bench_${N}gpu.err; python - <<PY
import json
l=json.loads([x for x in open('gpurun_out/bench_${N}gpu.log') if x.startswith('{')][-1])
print({k:l[k] for k in ('value','n_gpus','ms_per_step')}, l['e2e']['value'])
print('sweep', l['sweep']['value'], l['sweep']['ms'], l['sweep']['records_checksum'])
print('tiled', l['tiled']['value'], l['tiled']['ms_per_step'], l['tiled']['detections_per_panel'])
PY
