#!/usr/bin/env python
"""Phase timing of the sort+NMS kernel (needs the profiling build):

    nvcc ... -DRADNET_NMS_PROFILE  (tools/build_prof.sh)  ->  _C/libradnet_b200_prof.so
    RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so python tools/nms_phase_profile.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rock_art_radnet_b200 import synthetic as S  # noqa: E402
from rock_art_radnet_b200.pipeline import ProposalPipeline  # noqa: E402

C = S.HotPathConfig()
CLUSTER = os.environ.get("RADNET_NMS_CLUSTER", "1") != "0"
# the cluster form (single panel) has one more stamp: the cluster-wide overlap matrix
names = (["start", "keys+table", "select", "bucket_sort", "ties", "gather", "matrix_rows", "blocks_landed", "chain+tiles", "record"] if CLUSTER
         else ["start", "keys+table", "select", "bucket_sort", "ties", "nms", "record"])
for seed in range(3):
    cls, regr = S.rpn_maps(seed)
    pipe = ProposalPipeline(C, 1, 38, 38, alloc_pooled=False)
    cls_d, regr_d = torch.from_numpy(cls).cuda(), torch.from_numpy(regr).cuda()
    for _ in range(3):
        pipe.decode(cls_d, regr_d)
        pipe.sort_nms()
    torch.cuda.synchronize()
    ws = pipe._ws.cpu().numpy()
    # stamps live 32768 bytes after the kept list: find offset = ws_stride - (32768+256) rounded... scan for plausible int64s
    K = 300
    kept_bytes = ((K * 16 + 64 + 255) // 256) * 256
    off = pipe._ws_bytes - (kept_bytes + 32768 + 256) + 32768
    st = ws[off:off + 16 * 8].view(np.int64)
    d = np.diff(st[:len(names)])
    rows = ws[off + 16 * 8:off + (16 + 32 * 8) * 8].view(np.int64).reshape(32, 8)
    if seed == 0 and not CLUSTER:
        t0 = rows[0, 0]
        for w in range(0, 24):
            r = rows[w] - t0
            print("  row %2d gathered=%6d matrix=%6d wait_pred=%6d turn=%6d retired=%6d  (turn->retired %5d, prev retired->my turn %5d)" % (
                w, r[0], r[1], r[2], r[3], r[4], r[4] - r[3], (r[3] - (rows[w - 1, 4] - t0)) if w else 0))
    if seed == 0 and CLUSTER:
        print("  chain groups (cycles between the publications of consecutive groups):",
              [int(rows[g, 7] - rows[g - 1, 7]) if g else 0 for g in range(24)])
        print("  first publication at", int(rows[0, 7] - st[7]), "cycles after 'blocks landed'; last at", int(rows[:24, 7].max() - st[7]))
    if seed == 0:
        bs = ws[off + 300 * 8:off + 306 * 8].view(np.int64)
        print("  bucket sort (zero, pass 1 + histogram, scan, scatter, insertion + ties):", np.diff(bs).tolist())
    ps = ws[off + 8 * 8:off + 16 * 8].view(np.int64)
    print("seed", seed, " ".join("%s=%d" % (n, v) for n, v in zip(names[1:], d)), "total cycles", st[len(names) - 1] - st[0],
          "n_sorted", pipe.records.to_numpy()[0]["n_sorted"])
