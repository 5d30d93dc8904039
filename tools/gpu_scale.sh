#!/bin/bash
# scaling evidence on one box: bench.py at N = 1, 2, 4, 8 (as many as the box has), both e2e and device-resident
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
: > gpurun_out/scale.log
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline >> gpurun_out/scale.log 2> gpurun_out/scale_$n.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
        bench.py --gpus $n --steps 30 --warmup 5 2> gpurun_out/scale_$n.err | grep '^{' >> gpurun_out/scale.log
    fi
    echo "N=$n exit $?"
  fi
done
python - <<'PY'
import json
for ln in open('gpurun_out/scale.log'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print(d['n_gpus'], round(d['value']), 'panels/s  e2e', round(d['e2e']['value']), ' ms/step', round(d['ms_per_step'], 3), d['clocks'])
PY
if [ $NG -ge 2 ]; then
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tools/check_sharded_detect.py --panels 4 2>&1 | tail -2
fi
