"""High-resolution panels processed as overlapping 600-px tiles (BASELINE configs[3]; reference
`RADNet.predict`, faster_rcnn/RADNet.py:502-718, tile grid RADNet.py:511-540), batched and sharded:

    tile g = panel * T + t  on rank g % world:   K1 decode -> K2 sort+NMS -> K4 RoI pool -> [classifier head,
                                                 user code] -> K5+K6 head decode + per-class NMS (0.2)
    owner routing (`sharding.OwnerRoutedTiles`): the labelled records of panel p go to rank p % world
    panel p on its owner:                        K7 final_nms (cluster-and-average) -> K6 per-class NMS (0.4)
    one all-gather of the final records:         every rank holds every panel's detections, global order

`TiledPanelRunner` keeps every buffer resident and launches everything on the caller's stream.  The classifier
head is not part of the path (it is the user's network); its outputs for the kept RoIs are handed in per tile.
"""
import numpy as np
import torch

from . import _device as D
from . import detect as DT
from . import sharding
from .pipeline import DetectionPipeline


class TiledPanelRunner:
    def __init__(self, C, n_panels, tiles, chunk=64, H=38, W=38, channels=1024, pool_size=14, max_boxes=300,
                 overlap_thresh=0.7, bbox_threshold=0.7, rank=None, world=None, group=None, device=None,
                 alloc_pooled=True):
        D.require_cuda()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        dev = self.device
        self.C, self.tiles, self.T = C, list(tiles), len(tiles)
        # width of a merged panel record: final_nms / class_nms keep room for every box of the panel's T tiles
        self.final_stride = DT.record_bytes(self.T * int(max_boxes))
        self.router = sharding.OwnerRoutedTiles(n_panels, self.T, rank=rank, world=world, group=group, device=dev,
                                                final_stride=self.final_stride)
        self.rank, self.world, self.n_panels = self.router.rank, self.router.world, int(n_panels)
        self.local_ids = self.router.local_ids
        self.n_local = len(self.local_ids)
        self.chunk = min(int(chunk), max(self.n_local, 1))
        kw = dict(channels=channels, pool_size=pool_size, max_boxes=max_boxes, overlap_thresh=overlap_thresh,
                  bbox_threshold=bbox_threshold, device=dev)
        self.pipe = DetectionPipeline(C, self.chunk, H, W, alloc_pooled=alloc_pooled, **kw)
        tail = self.n_local % self.chunk
        self.tail_pipe = DetectionPipeline(C, tail, H, W, alloc_pooled=False, **kw) if tail else None
        self.n_cls = self.pipe.n_cls
        self.max_boxes = self.pipe.max_boxes
        # labelled records of all local tiles, written chunk by chunk
        self.tile_records = DT.ClassRecords(max(self.n_local, 1), self.max_boxes, dev)
        origin = np.array([[self.tiles[int(g) % self.T][0], self.tiles[int(g) % self.T][1]] for g in self.local_ids],
                          dtype=np.int32).reshape(-1, 2)
        self.origin = torch.from_numpy(origin).to(dev)
        self.ratio = torch.ones((max(self.n_local, 1),), dtype=torch.float64, device=dev)
        self.launches = 0
        self.has_pooled = alloc_pooled

    def chunks(self):
        """(first local slot, size, pipeline) of every chunk."""
        out = []
        for lo in range(0, self.n_local, self.chunk):
            n = min(self.chunk, self.n_local - lo)
            out.append((lo, n, self.pipe if n == self.chunk else self.tail_pipe))
        return out

    def propose(self, cls, regr, lo, n, pipe):
        pipe.decode(cls[lo:lo + n], regr[lo:lo + n])
        pipe.sort_nms()
        self.launches += 2

    def run(self, cls, regr, feat, P_cls, P_regr):
        """cls (n_local,H,W,A), regr (n_local,H,W,4A), feat (n_local,H,W,C) or None, P_cls (n_local,max_boxes,n_cls),
        P_regr (n_local,max_boxes,4(n_cls-1)): float32 CUDA tensors in local slot order.  Returns the final
        records of all panels, (n_panels, stride) uint8 in global panel order (on every rank)."""
        self.launches = 0
        for lo, n, pipe in self.chunks():
            self.propose(cls, regr, lo, n, pipe)
            if feat is not None and self.has_pooled:
                pipe.pool(feat[lo:lo + n], out=self.pipe.pooled[:n])
                self.launches += 1
            out = DT.ClassRecords(n, self.max_boxes, self.device, raw=self.tile_records.raw[lo:lo + n])
            DT.classify_nms(P_cls[lo:lo + n], P_regr[lo:lo + n], self.C, det=pipe.records,
                            bbox_threshold=pipe.bbox_threshold, nms_thresh=0.2, max_boxes=300,
                            ratio=self.ratio[lo:lo + n], origin=self.origin[lo:lo + n], out=out)
            self.launches += 1
        routed = self.router.exchange(self.tile_records.raw[:self.n_local])
        n_owned = len(self.router.owned)
        if n_owned:
            rec = DT.ClassRecords(n_owned * self.T, self.max_boxes, self.device, raw=routed)
            merged = DT.final_nms_records(rec, n_owned, self.T, self.n_cls)
            final = DT.class_nms(merged, n_owned, 1, self.n_cls, 0.4, max_boxes=300)
            self.launches += 2
            final_raw = final.raw
        else:
            # a rank that owns no panel (fewer panels than ranks) still takes part in the gather, with rows of the
            # SAME width as everybody else's final records - not the (narrower) tile records
            final_raw = None
        return self.router.gather_final(final_raw)


def synthetic_tile_inputs(runner, with_features=False, channels=1024, pool_of=8):
    """Seeded synthetic inputs for the local tiles of `runner` (tests, tools, bench): RPN maps of tile (p, t) =
    `synthetic.rpn_maps(1000*p + t)`; classifier-head outputs from `synthetic.tiled_panel_head_outputs` for the
    RoIs K2 keeps on those maps (computed here by running K1 + K2 once and reading the records back; neighbouring
    tiles see the same scene objects, so the merge has real clusters).  Feature maps, if asked for, are drawn
    from a pool of `pool_of` distinct maps (they only feed K4, whose output the head stand-in does not read).
    Returns (cls, regr, feat or None, P_cls, P_regr) CUDA tensors in local slot order."""
    from . import synthetic as S
    dev, T = runner.device, runner.T
    ids = [int(g) for g in runner.local_ids]
    maps = [S.rpn_maps(1000 * (g // T) + g % T, runner.pipe.H, runner.pipe.W, runner.pipe.A) for g in ids]
    n = len(ids)
    H, W, A = runner.pipe.H, runner.pipe.W, runner.pipe.A
    cls = torch.from_numpy(np.concatenate([m[0] for m in maps]) if n else np.zeros((0, H, W, A), np.float32)).to(dev)
    regr = torch.from_numpy(np.concatenate([m[1] for m in maps]) if n else np.zeros((0, H, W, 4 * A), np.float32)).to(dev)
    K = runner.max_boxes
    P_cls = np.zeros((n, K, runner.n_cls), dtype=np.float32)
    P_regr = np.zeros((n, K, 4 * (runner.n_cls - 1)), dtype=np.float32)
    for lo, m, pipe in runner.chunks():
        runner.propose(cls, regr, lo, m, pipe)
        dets = pipe.records.to_numpy()
        for i in range(m):
            g = ids[lo + i]
            R = dets[i]["boxes"].copy()
            R[:, 2] -= R[:, 0]
            R[:, 3] -= R[:, 1]
            P_cls[lo + i], P_regr[lo + i] = S.tiled_panel_head_outputs(runner.C, g // T, runner.tiles[g % T], R, n_slots=K)
    feat = None
    if with_features:
        pool = torch.from_numpy(np.concatenate([S.feature_map(s, H, W, channels) for s in range(pool_of)])).to(dev)
        idx = torch.tensor([g % pool_of for g in ids], dtype=torch.int64, device=dev)
        feat = pool.index_select(0, idx) if n else pool[:0]
    return cls, regr, feat, torch.from_numpy(P_cls).to(dev), torch.from_numpy(P_regr).to(dev)
