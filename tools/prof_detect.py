#!/usr/bin/env python
"""Small fixed workload for ncu captures of the latency-bound kernels: the cluster form of sort+NMS on a
single panel, and the detection post-processing (K5+K6 fused, K7 final_nms, K6 on records)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rock_art_radnet_b200 import detect as DT  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402
from rock_art_radnet_b200.pipeline import DetectionPipeline, ProposalPipeline  # noqa: E402

C = S.HotPathConfig()
cls1, regr1 = S.rpn_maps(0)
single = ProposalPipeline(C, 1, 38, 38, alloc_pooled=False)
single.decode(torch.from_numpy(cls1).cuda(), torch.from_numpy(regr1).cuda())

B = 64
base = [S.rpn_maps(s) for s in range(4)]
cls = torch.from_numpy(np.concatenate([base[i % 4][0] for i in range(B)])).cuda()
regr = torch.from_numpy(np.concatenate([base[i % 4][1] for i in range(B)])).cuda()
dp = DetectionPipeline(C, B, 38, 38, alloc_pooled=False)
dp.decode(cls, regr)
dp.sort_nms()
g = torch.Generator(device="cuda").manual_seed(1)
logits = torch.randn((B, 300, 7), device="cuda", generator=g) * 1.5
logits[..., 6] += 1.0
boost = torch.rand((B, 300, 1), device="cuda", generator=g) < 0.25
pick = torch.randint(0, 6, (B, 300, 1), device="cuda", generator=g)
logits.scatter_add_(2, pick, boost.float() * 6.0)
P_cls = torch.softmax(logits, dim=-1).contiguous()
P_regr = (torch.randn((B, 300, 24), device="cuda", generator=g) * 0.8).contiguous()
ratio = torch.ones((B,), dtype=torch.float64, device="cuda")
origin = torch.zeros((B, 2), dtype=torch.int32, device="cuda")
for it in range(4):
    single.sort_nms()
    dp.classify(P_cls, P_regr, ratio=ratio, origin=origin)
    tiles36 = DT.ClassRecords(36, 300, "cuda", raw=dp.class_records.raw[:36])
    merged = DT.final_nms_records(tiles36, 1, 36, 7)
    final = DT.class_nms(merged, 1, 1, 7, 0.4)
    torch.cuda.synchronize()
print("ok", int(single.records.counts.sum()), int(final.header[0, 0]))
