#!/usr/bin/env python
"""Small fixed workload of the training-side kernels for ncu captures and launch lists: K3 rpn_targets (64 panels,
20 figures, NHWC x std_scaling), the 256-region balancing, batched calc_iou from K2's records, get_selected_samples,
the four losses, and the mAP matching + AP kernels."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rock_art_radnet_b200 import evaluation as E  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402
from rock_art_radnet_b200.losses import RpnLossBatch, class_losses_device  # noqa: E402
from rock_art_radnet_b200.pipeline import ProposalPipeline  # noqa: E402
from rock_art_radnet_b200.rpn import RoiTargetBatch, gt_feature_cells  # noqa: E402
from rock_art_radnet_b200.sampling import RpnSubsampler, SampleSelector, seed_states  # noqa: E402
from rock_art_radnet_b200.utils import LAYOUT_NHWC, RpnTargetBatch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
C = S.HotPathConfig()
G, H, W, A = 20, 38, 38, 9
gt = np.zeros((B, G, 4)); gtc = np.zeros((B, G, 4)); gcl = np.zeros((B, G), np.int32)
for b in range(B):
    img = S.gt_figures(b, G, 600, 600, classes=("boat", "human"))
    for k, bb in enumerate(img["bboxes"]):
        gt[b, k] = [bb["x1"], bb["x2"], bb["y1"], bb["y2"]]
    gtc[b], gcl[b] = gt_feature_cells(img, C, C.class_mapping)
gt_d = torch.from_numpy(gt).cuda()
bg_d = torch.zeros((B, G), dtype=torch.uint8, device="cuda")
cnt_d = torch.full((B,), G, dtype=torch.int32, device="cuda")
wh_d = torch.tensor([[600.0, 600.0]] * B, dtype=torch.float64, device="cuda")
base = [S.rpn_maps(s) for s in range(8)]
cls = torch.from_numpy(np.concatenate([base[i % 8][0] for i in range(B)])).cuda()
regr = torch.from_numpy(np.concatenate([base[i % 8][1] for i in range(B)])).cuda()
pipe = ProposalPipeline(C, B, H, W, alloc_pooled=False)
tb = RpnTargetBatch(C, B, G, H, W, layout=LAYOUT_NHWC, regr_scale=C.std_scaling)
sub = RpnSubsampler(B, H, W, A, layout=LAYOUT_NHWC)
rt = RoiTargetBatch(C, C.class_mapping, B, 300, G)
sel = SampleSelector(B, 300, 7, int(C.n_rois))
lb = RpnLossBatch(B, H, W, A)
g = torch.Generator(device="cuda").manual_seed(0)
p_cls = torch.sigmoid(3 * torch.randn((B, H, W, A), device="cuda", generator=g))
p_regr = torch.randn((B, H, W, 4 * A), device="cuda", generator=g)
q_cls = torch.softmax(torch.randn((B, int(C.n_rois), 7), device="cuda", generator=g), dim=-1)
q_regr = torch.randn((B, int(C.n_rois), 24), device="cuda", generator=g)
gtc_d, gcl_d = torch.from_numpy(gtc).cuda(), torch.from_numpy(gcl).cuda()
det, gts = S.eval_set(1, 300, 1500)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(4):
    flush.fill_(it)
    pipe.decode(cls, regr)
    pipe.sort_nms()
    y_cls, y_regr, _, _ = tb.run(gt_d, bg_d, cnt_d, wh_d)
    states = seed_states(np.arange(B) + it)
    sub.run(y_cls, states)
    loss = lb.run(y_cls, y_regr, p_cls, p_regr)
    x_roi, y_class, y2, _, _, count = rt.run(gtc_d, gcl_d, cnt_d, det=pipe.records)
    s_idx, rep = sel.run(y_class, count, states)
    closs = class_losses_device(y_class, y2, q_cls, q_regr, sel=s_idx, n_sel_per_panel=rep[:, 0].contiguous())
    torch.cuda.synchronize()
T, P = E.get_objects(det, gts, 0.5)
aps = [E.calc_class_ap(T[k], P[k])[0] for k in T]
print("ok", float(loss.sum()), float(closs.sum()), float(np.mean(aps)))
