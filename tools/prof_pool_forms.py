#!/usr/bin/env python
"""K4 once per form on the headline shape (64 panels, 38x38x1024, pool 14) for ncu captures.
usage: prof_pool_forms.py [B] [H] [W] [C] [pool]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rock_art_radnet_b200 import _lib  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402
from rock_art_radnet_b200.pipeline import ProposalPipeline  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
B, H, W, Cn, pool = (a + [64, 38, 38, 1024, 14][len(a):])[:5]
C = S.HotPathConfig()
base = [S.rpn_maps(s, H, W, 9) for s in range(4)]
cls = torch.from_numpy(np.concatenate([base[i % 4][0] for i in range(B)])).cuda()
regr = torch.from_numpy(np.concatenate([base[i % 4][1] for i in range(B)])).cuda()
feat = torch.randn((B, H, W, Cn), dtype=torch.float32, device="cuda")
pipe = ProposalPipeline(C, B, H, W, channels=Cn, pool_size=pool)
pipe.decode(cls, regr)
pipe.sort_nms()
for form, bands in ((1, 0), (2, 2), (2, 3)):
    _lib.set_option("roipool_form", form)
    _lib.set_option("roipool_bands", bands)
    for it in range(3):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        pipe.pool(feat)
        ev1.record()
        torch.cuda.synchronize()
    print("form", form, "bands", bands, "ms", ev0.elapsed_time(ev1))
