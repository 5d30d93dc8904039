#!/bin/bash
# K2 quick loop: parity tests of the NMS paths, phase profile, latency for cluster sizes 8 and 16
mkdir -p gpurun_out
for cs in 0 8; do
RADNET_NMS_CLUSTER_SIZE=$cs timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_shapes.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_nms_cs$cs.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_nms_cs$cs.log
tail -2 gpurun_out/pytest_gpu_nms_cs$cs.log
done
bash tools/build_prof.sh > gpurun_out/build_prof.log 2>&1
for cs in 0; do
RADNET_NMS_CLUSTER_SIZE=$cs RADNET_NMS_CLUSTER=1 RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so timeout 300 python tools/nms_phase_profile.py > gpurun_out/nms_phase_cs$cs.log 2>&1
tail -3 gpurun_out/nms_phase_cs$cs.log
RADNET_NMS_CLUSTER_SIZE=$cs timeout 300 python - <<'PY'
import os, sys, torch, numpy as np
sys.path.insert(0, '.')
from rock_art_radnet_b200 import synthetic as S
from rock_art_radnet_b200.pipeline import ProposalPipeline
C = S.HotPathConfig()
res, resg = [], []
for seed in range(4):
    cls, regr = S.rpn_maps(seed, 38, 38, realistic=bool(seed % 2))
    pipe = ProposalPipeline(C, 1, 38, 38, alloc_pooled=False)
    pipe.decode(torch.from_numpy(cls).cuda(), torch.from_numpy(regr).cuda())
    ts = []
    for i in range(40):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pipe.sort_nms(); b.record(); b.synchronize()
        if i >= 8: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); res.append(round(ts[len(ts)//2], 1))
    # device latency: the same call captured in a CUDA graph (no Python / ctypes launch path inside the timed region)
    st = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        pipe.sort_nms(); st.synchronize()
        with torch.cuda.graph(g, stream=st):
            pipe.sort_nms()
    ts = []
    for i in range(40):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize()
        if i >= 8: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); resg.append(round(ts[len(ts)//2], 1))
print("cluster size", os.environ.get("RADNET_NMS_CLUSTER_SIZE"), "sort_nms single panel p50 us: python call", res, " graph replay", resg)
for tag, (Hh, Ww, scales) in {"12anchors_38x38": (38, 38, (64, 128, 256, 512)), "9anchors_38x50": (38, 50, (128, 256, 512)), "12anchors_38x50": (38, 50, (64, 128, 256, 512))}.items():
    C2 = S.HotPathConfig(scales)
    cls, regr = S.rpn_maps(1, Hh, Ww, C2.num_anchors)
    pipe = ProposalPipeline(C2, 1, Hh, Ww, alloc_pooled=False)
    pipe.decode(torch.from_numpy(cls).cuda(), torch.from_numpy(regr).cuda())
    ts = []
    for i in range(40):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pipe.sort_nms(); b.record(); b.synchronize()
        if i >= 8: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); print("  ", tag, "single panel python-call p50 us", round(ts[len(ts)//2], 1))
for B in (4, 8, 12, 64):
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
