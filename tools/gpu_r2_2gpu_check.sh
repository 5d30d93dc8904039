#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sweep_tiled.py tests/test_gpu_detect.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
for P in 1 5; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$P tools/check_sharded_detect.py --panels $P > gpurun_out/sharded_detect_2gpu_p$P.log 2>&1; echo "exit $?" >> gpurun_out/sharded_detect_2gpu_p$P.log
tail -2 gpurun_out/sharded_detect_2gpu_p$P.log
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "exit $?"
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/bench_2gpu.log') if x.startswith('{')][-1])
print({k:l[k] for k in ('value','n_gpus','ms_per_step')}, l['e2e']['value'], l['roofline']['launch_form'], round(l['roofline']['frac'],3))
print('sweep', l['sweep']['value'], l['sweep']['records_checksum'], 'tiled', l['tiled']['value'])
PY
