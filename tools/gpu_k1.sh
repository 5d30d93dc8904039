#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_random_shapes.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tools')
from rock_art_radnet_b200 import synthetic as S
from rock_art_radnet_b200.pipeline import ProposalPipeline
from oracle import radnet_oracle as O
import bench_kernels as BK
C = S.HotPathConfig()
for B in (64, 512):
    cls, regr = BK.tile_maps(B)
    pipe = ProposalPipeline(C, B, 38, 38, alloc_pooled=False)
    t = BK.time_ms(lambda: pipe.decode(cls, regr))
    nbytes = 36 * pipe.N * B
    print("decode_clip B=%d p50 %.1f us  %.0f GB/s  frac %.3f" % (B, t["p50_ms"] * 1e3, nbytes / t["p50_ms"] / 1e6, nbytes / t["p50_ms"] / 1e6 / BK.peak()))
# stress: many seeds, large regression spread (more fallbacks), compare boxes bit-exact with the oracle
bad = 0
for seed in range(40):
    rng = np.random.default_rng(seed)
    H, W = int(rng.integers(3, 60)), int(rng.integers(3, 60))
    A = 9
    cls = rng.random((1, H, W, A)).astype(np.float32)
    scale = [0.5, 2.0, 6.0, 20.0][seed % 4]
    regr = (scale * rng.standard_normal((1, H, W, 4 * A))).astype(np.float32)
    if seed % 5 == 0:
        regr.flat[rng.integers(0, regr.size, 20)] = np.float32(np.nan)
        regr.flat[rng.integers(0, regr.size, 20)] = np.float32(np.inf)
    pipe = ProposalPipeline(C, 1, H, W, alloc_pooled=False)
    pipe.decode(torch.from_numpy(cls).cuda(), torch.from_numpy(regr).cuda())
    boxes = pipe.boxes.cpu().numpy()[0]
    keys = pipe.keys.cpu().numpy()[0]
    with np.errstate(all="ignore"):
        all_boxes, probs, keep = O.decode_proposals(cls, regr, C)[:3]
    finite = np.isfinite(all_boxes).all(axis=1)
    ok = keep & finite
    km = (keys != 0) != ok
    both = ok & (keys != 0)
    bm = (boxes[both].astype(np.float64) != all_boxes[both]).any(axis=1)
    if km.any() or bm.any():
        bad += 1
        i = int(np.flatnonzero(km)[0]) if km.any() else int(np.flatnonzero(both)[np.flatnonzero(bm)[0]])
        print("seed", seed, "key mismatches", int(km.sum()), "box mismatches", int(bm.sum()), "example", i, boxes[i], all_boxes[i], bool(keep[i]), int(keys[i]), pipe.stats.cpu().numpy())
print("stress mismatches:", bad)
PY
