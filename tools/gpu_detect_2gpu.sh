#!/bin/bash
# 2-GPU: sharded tiled-panel check + the bench at N=2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/check_sharded_detect.py --panels 2 > gpurun_out/sharded_detect_2gpu.log 2>&1; echo "check exit $?" >> gpurun_out/sharded_detect_2gpu.log
tail -5 gpurun_out/sharded_detect_2gpu.log
