#!/usr/bin/env python
"""Small fixed workload for ncu captures of the non-dominant kernels (decode, sort+NMS, targets)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rock_art_radnet_b200 import synthetic as S  # noqa: E402
from rock_art_radnet_b200.pipeline import ProposalPipeline  # noqa: E402
from rock_art_radnet_b200.utils import rpn_targets_device  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
C = S.HotPathConfig()
base = [S.rpn_maps(s) for s in range(4)]
cls = torch.from_numpy(np.concatenate([base[i % 4][0] for i in range(B)])).cuda()
regr = torch.from_numpy(np.concatenate([base[i % 4][1] for i in range(B)])).cuda()
pipe = ProposalPipeline(C, B, 38, 38, alloc_pooled=False)
G = 20
gt = np.zeros((B, G, 4)); bg = np.zeros((B, G), np.uint8)
for b in range(B):
    for k, bb in enumerate(S.gt_figures(b, G)["bboxes"]):
        gt[b, k] = [bb["x1"], bb["x2"], bb["y1"], bb["y2"]]
gt_d = torch.from_numpy(gt).cuda(); bg_d = torch.from_numpy(bg).cuda()
cnt_d = torch.full((B,), G, dtype=torch.int32, device="cuda")
wh_d = torch.tensor([[600.0, 600.0]] * B, dtype=torch.float64, device="cuda")
for it in range(4):
    pipe.decode(cls, regr)
    pipe.sort_nms()
    out = rpn_targets_device(C, gt_d, bg_d, cnt_d, 38, 38, wh_d)
    torch.cuda.synchronize()
print("ok", int(pipe.records.counts.sum()), float(out[0].sum()))
