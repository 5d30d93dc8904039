#!/bin/bash
# K2 iteration loop: full GPU parity suite (both NMS forms), phase profile (profiling build), single-panel latency
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
RADNET_NMS_CLUSTER=0 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_shapes.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_nocluster.log 2>&1; echo "pytest (no cluster) exit $?" >> gpurun_out/pytest_gpu_nocluster.log
tail -2 gpurun_out/pytest_gpu_nocluster.log
bash tools/build_prof.sh > gpurun_out/build_prof.log 2>&1
for c in 1 0; do
RADNET_NMS_CLUSTER=$c RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so timeout 300 python tools/nms_phase_profile.py > gpurun_out/nms_phase_c$c.log 2>&1
tail -4 gpurun_out/nms_phase_c$c.log
RADNET_NMS_CLUSTER=$c timeout 300 python - <<'PY'
import os, sys, torch, numpy as np
sys.path.insert(0, '.')
from rock_art_radnet_b200 import synthetic as S
from rock_art_radnet_b200.pipeline import ProposalPipeline
C = S.HotPathConfig()
for (H, W, tag) in ((38, 38, "600px"), (100, 100, "90k")):
    res = []
    for seed in range(4):
        cls, regr = S.rpn_maps(seed, H, W, realistic=bool(seed % 2))
        pipe = ProposalPipeline(C, 1, H, W, alloc_pooled=False)
        pipe.decode(torch.from_numpy(cls).cuda(), torch.from_numpy(regr).cuda())
        ts = []
        for i in range(40):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); pipe.sort_nms(); b.record(); b.synchronize()
            if i >= 8: ts.append(a.elapsed_time(b) * 1e3)
        ts.sort(); res.append(round(ts[len(ts)//2], 1))
    print("cluster=%s sort_nms single panel p50 us" % os.environ.get("RADNET_NMS_CLUSTER"), tag, res)
for B in (8, 16, 64):
    cls = torch.from_numpy(np.concatenate([S.rpn_maps(s)[0] for s in range(B)])).cuda()
    regr = torch.from_numpy(np.concatenate([S.rpn_maps(s)[1] for s in range(B)])).cuda()
    pipe = ProposalPipeline(C, B, 38, 38, alloc_pooled=False)
    pipe.decode(cls, regr)
    ts = []
    for i in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pipe.sort_nms(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); print("  sort_nms %d panels p50 us" % B, round(ts[15], 1))
PY
done
