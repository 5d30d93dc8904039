#!/usr/bin/env python
"""K4 timing sweep over launch forms on the three benchmark shapes (library chosen by RADNET_B200_LIB).
usage: exp_pool.py TAG [form:bands:lanes ...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rock_art_radnet_b200 import _lib  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402
from rock_art_radnet_b200.pipeline import ProposalPipeline  # noqa: E402

tag = sys.argv[1]
args = [a for a in sys.argv[2:] if "," not in a]
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:] if "," in a]      # B,C,pool,H,W
forms = [tuple(int(v) for v in a.split(":")) for a in args] or [(1, 0, 0)]
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
pk = float(pk.get("hbm_gbs_burst") or pk.get("hbm_gbs") or 6559.4) if isinstance(pk, dict) else 6559.4
C = S.HotPathConfig()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
default = ((64, 1024, 14, "resnet50", 38, 38), (64, 512, 7, "vgg16", 38, 38), (32, 1024, 14, "r50_38x50", 38, 50))
for (B, Cn, pool, name, H, W) in ([(b, c, pl, "%dx%dx%d_p%d_B%d" % (h, w, c, pl, b), h, w) for (b, c, pl, h, w) in shapes] or default):
    base = [S.rpn_maps(s, H, W, 9) for s in range(4)]
    cls = torch.from_numpy(np.concatenate([base[i % 4][0] for i in range(B)])).cuda()
    regr = torch.from_numpy(np.concatenate([base[i % 4][1] for i in range(B)])).cuda()
    feat = torch.randn((B, H, W, Cn), dtype=torch.float32, device="cuda")
    pipe = ProposalPipeline(C, B, H, W, channels=Cn, pool_size=pool)
    pipe.decode(cls, regr)
    pipe.sort_nms()
    kept = int(pipe.records.counts.sum().item())
    nbytes = B * H * W * Cn * 4 + kept * 16 + kept * pool * pool * Cn * 4
    for fm in forms:
        form, bands, lanes = fm[:3]
        pace = fm[3] if len(fm) > 3 else 0
        _lib.set_option("roipool_cluster", fm[4] if len(fm) > 4 else -1)
        _lib.set_option("roipool_sync_every", fm[5] if len(fm) > 5 else 0)
        _lib.set_option("roipool_ctas", fm[6] if len(fm) > 6 else 0)
        _lib.set_option("roipool_form", form)
        _lib.set_option("roipool_bands", bands)
        _lib.set_option("roipool_lanes", lanes)
        ts = []
        for it in range(8):
            flush.fill_(it)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pipe.pool(feat)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = float(np.median(ts[2:]))
        print("%-10s %-22s form=%d bands=%d lanes=%2d pace=%2d cl=%s  %.4f ms  frac %.4f" % (tag, name, form, bands, lanes, pace, ":".join(str(v) for v in fm[4:]), t, nbytes / t / 1e6 / pk), flush=True)
    del pipe, feat
    torch.cuda.empty_cache()
