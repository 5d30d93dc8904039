#!/bin/bash
# 2 GPUs: fewer panels than ranks (rank 1 owns no panel), then as many, then the sweep with a ragged tail that used to re-tune K4
mkdir -p gpurun_out
for P in 1 3; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$P tools/check_sharded_detect.py --panels $P > gpurun_out/sharded_detect_2gpu_p$P.log 2>&1; echo "exit $?" >> gpurun_out/sharded_detect_2gpu_p$P.log
tail -2 gpurun_out/sharded_detect_2gpu_p$P.log
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 2 --steps 10 --warmup 3 --sweep-panels 4420 > gpurun_out/bench_2gpu_tail.log 2> gpurun_out/bench_2gpu_tail.err; echo "exit $?"
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/bench_2gpu_tail.log') if x.startswith('{')][-1])
print({k:l[k] for k in ('value','n_gpus','ms_per_step')}, l['e2e']['value'], l['roofline']['launch_form'])
print('sweep', l['sweep']['value'], l['sweep']['ms'], l['sweep']['tail_batch'], l['sweep']['steps_per_rank'])
print('tiled', l['tiled']['value'])
PY
