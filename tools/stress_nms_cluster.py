#!/usr/bin/env python
"""Stress of the cluster form of sort+NMS: thousands of launches over random panels, batch sizes and NMS
settings; every result must equal the one-CTA form (the nms_cluster option is read per launch)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rock_art_radnet_b200 import synthetic as S  # noqa: E402
from rock_art_radnet_b200 import _lib  # noqa: E402
from rock_art_radnet_b200.pipeline import ProposalPipeline  # noqa: E402

torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(0)
rng = np.random.default_rng(0)
t0 = time.time()
n_launch = 0
for trial in range(60):
    B = int(rng.integers(1, 13))
    scales = (128, 256, 512) if trial % 3 else (64, 128, 256, 512)
    C = S.HotPathConfig(scales)
    A = C.num_anchors
    H, W = [(38, 38), (38, 50), (25, 19), (60, 60)][trial % 4]
    if A * H * W > 18900:
        H, W = 38, 38
    thr = [0.7, 0.5, 0.9, 0.3][trial % 4]
    mb = [300, 300, 100, 600][(trial // 4) % 4]
    N = H * W * A
    pipe = ProposalPipeline(C, B, H, W, max_boxes=mb, overlap_thresh=thr, alloc_pooled=False)
    ref = ProposalPipeline(C, B, H, W, max_boxes=mb, overlap_thresh=thr, alloc_pooled=False)
    for rep in range(40):
        if rep % 3 == 0:                        # many equal scores (crowded buckets -> general path)
            cls = (torch.randint(0, 50, (B, H, W, A), device="cuda", generator=g).float() / 50.0).contiguous()
        else:
            perm = torch.stack([torch.randperm(N, device="cuda", generator=g) for _ in range(B)])
            cls = ((perm.float() + 0.5) / N).reshape(B, H, W, A).contiguous()
        spread = [0.5, 1.5, 0.1][rep % 3]       # 0.1: heavy overlap, the slice may run out (second round)
        regr = (spread * torch.randn((B, H, W, 4 * A), device="cuda", generator=g)).contiguous()
        _lib.set_option("nms_cluster", 1)
        pipe.decode(cls, regr)
        pipe.sort_nms()
        _lib.set_option("nms_cluster", 0)
        ref.decode(cls, regr)
        ref.sort_nms()
        torch.cuda.synchronize()
        a, b = pipe.records.raw, ref.records.raw
        if not torch.equal(a, b):
            ha, hb = pipe.records.header.cpu().numpy(), ref.records.header.cpu().numpy()
            # n_sorted / tie statistics may differ between the forms only if they sorted different slices
            same = torch.equal(pipe.records.boxes, ref.records.boxes) and (ha[:, :2] == hb[:, :2]).all()
            if not same:
                print("MISMATCH trial", trial, "rep", rep, "B", B, "HxWxA", H, W, A, "thr", thr, "mb", mb, ha.tolist(), hb.tolist())
                sys.exit(1)
        n_launch += 2
print("cluster stress ok: %d launches in %.1f s" % (n_launch, time.time() - t0))
