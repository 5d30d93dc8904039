import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch, numpy as np
sys.path.insert(0, %r)
from rock_art_radnet_b200 import synthetic as S
from rock_art_radnet_b200.pipeline import ProposalPipeline
C = S.HotPathConfig()
res = []
for seed in range(4):
    cls, regr = S.rpn_maps(seed, realistic=bool(seed %% 2))
    pipe = ProposalPipeline(C, 1, 38, 38, alloc_pooled=False)
    pipe.decode(torch.from_numpy(cls).cuda(), torch.from_numpy(regr).cuda())
    ts = []
    for i in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pipe.sort_nms(); b.record(); b.synchronize()
        if i >= 5: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); res.append(round(ts[len(ts)//2], 1))
print(res)
''' % ROOT
for la in (1, 2, 4, 6, 10, 16, 32):
    for st in (2048, 1024):
        env = dict(os.environ, RADNET_NMS_LOOKAHEAD=str(la), RADNET_NMS_SEL_TARGET=str(st))
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print("look_ahead", la, "sel_target", st, out.stdout.strip(), out.stderr.strip()[-200:])
