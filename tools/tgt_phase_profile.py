#!/usr/bin/env python
"""Phase timing of the RPN target kernel K3 (needs the profiling build, tools/build_prof.sh):

    RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so python tools/tgt_phase_profile.py [B]

Every CTA stamps %globaltimer (ns): fill CTAs at 0 start and 9 share written; compute CTAs (last panel they
handled) at 0 start, 1 figures + floors + anchor tables, 2 windows, 3 phase 1 (IoU over the windows),
4 phase 2 + positives parked, 5 panel filled (wait over), 6 positives written.  Relative to the earliest start."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rock_art_radnet_b200 import _lib  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402
from rock_art_radnet_b200.utils import RpnTargetBatch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
G, H, W = 20, 38, 38
C = S.HotPathConfig()
lib = _lib.load()
A = 9
stamps = torch.zeros((320 * 16,), dtype=torch.int64, device="cuda")
lib.radnet_debug_set_tgt_stamps.argtypes = [ctypes.c_void_p]
lib.radnet_debug_set_tgt_stamps(ctypes.c_void_p(stamps.data_ptr()))
gt = np.zeros((B, G, 4))
for b in range(B):
    for k, bb in enumerate(S.gt_figures(b, G)["bboxes"]):
        gt[b, k] = [bb["x1"], bb["x2"], bb["y1"], bb["y2"]]
gt_d = torch.from_numpy(gt).cuda()
bg_d = torch.zeros((B, G), dtype=torch.uint8, device="cuda")
cnt_d = torch.full((B,), G, dtype=torch.int32, device="cuda")
wh_d = torch.tensor([[600.0, 600.0]] * B, dtype=torch.float64, device="cuda")
tb = RpnTargetBatch(C, B, G, H, W)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(4):
    flush.fill_(it)
    stamps.zero_()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tb.run(gt_d, bg_d, cnt_d, wh_d)
    e.record()
    torch.cuda.synchronize()
    st = stamps.cpu().numpy().reshape(320, 16).astype(np.float64)
    st = st[st[:, 0] > 0]
    t0 = st[:, 0].min()
    names = {0: "start", 9: "fill: share written", 1: "setup", 2: "windows", 3: "phase1", 4: "phase2+parked", 5: "panel filled", 6: "positives written"}
    print("launch %d: event time %.1f us" % (it, a.elapsed_time(e) * 1e3))
    for k in (0, 9, 1, 2, 3, 4, 5, 6):
        v = st[:, k]
        v = v[v > 0] - t0
        if len(v):
            print("   %-18s n=%4d  min %7.2f  median %7.2f  max %7.2f us" % (names[k], len(v), v.min() / 1e3, np.median(v) / 1e3, v.max() / 1e3))
    comp = st[st[:, 6] > 0]
    for k, nm in ((10, "chunks"), (11, "hits"), (12, "winners"), (13, "forced")):
        print("   per panel %-8s min %5d  median %5d  max %5d" % (nm, comp[:, k].min(), np.median(comp[:, k]), comp[:, k].max()))

# ---- steady state: back-to-back launches over rotating output sets (4 x 66.6 MB at B = 64: larger than L2) ----
n_sets = max(2, int(np.ceil(260e6 / (B * 1040320.0))) + 1)
sets = [RpnTargetBatch(C, B, G, H, W) for _ in range(n_sets)]
for lay, name in ((0, "channel-first"), (1, "nhwc x std_scaling")):
    sets = [RpnTargetBatch(C, B, G, H, W, layout=lay, regr_scale=4.0 if lay else 1.0) for _ in range(n_sets)]
    for tb_ in sets:
        tb_.run(gt_d, bg_d, cnt_d, wh_d)
    torch.cuda.synchronize()
    reps = 40
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        sets[i % n_sets].run(gt_d, bg_d, cnt_d, wh_d)
    e.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(e) * 1e3 / reps
    print("stream of %d launches over %d output sets (%s): %.1f us per launch = %.0f GB/s" % (
        reps, n_sets, name, us, B * 1040320.0 / us / 1e3))
    del sets
