"""Archive-scale sweep (BASELINE configs[4]): `n_panels` synthetic 600-px panels,
panel i on rank i % world, processed in batches through K1 -> K2 -> K4; the detection records of every batch are
all-gathered over NCCL while the RoI pool of the same batch runs, and end up in global panel order on every rank.

The panels are generated ON THE DEVICE from their id (`radnet_synth_panels`, counter-based), batch by batch, right
before they are processed - no host-to-device traffic - so any panel can be regenerated anywhere: the records of a
sweep do not depend on the batch size, the number of ranks or which rank a panel lands on (tests check exactly that).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _device as D
from . import _lib
from . import sharding
from .pipeline import DetectionRecords, ProposalPipeline


def generate_panels(seed, first_panel, panel_stride, cls, regr, feat):
    """Fill cls (B,H,W,A), regr (B,H,W,4A), feat (B,H,W,C) float32 CUDA tensors with the synthetic panels
    first_panel + b*panel_stride, b = 0..B-1 (asynchronous, current stream)."""
    B, H, W, A = (int(v) for v in cls.shape)
    _lib.call("radnet_synth_panels", int(seed), int(first_panel), int(panel_stride), B, H, W, A, int(feat.shape[3]),
              D.ptr(cls), D.ptr(regr), D.ptr(feat), D.stream_ptr(cls.device))


class ArchiveSweep:
    """One sweep = `run()`: every rank processes its panels in batches of `batch` (the last one ragged) and
    returns the (n_panels, record stride) uint8 tensor of ALL detection records in global panel order.

    steps      number of batches per rank (the same on every rank: ranks without panels in the last step still
               take part in its gather with an empty batch)
    launches   kernels of ours per run (generator + K1 + K2 + K4 per non-empty batch)"""

    def __init__(self, C, n_panels, batch=64, seed=0, H=38, W=38, channels=1024, pool_size=14, max_boxes=300,
                 overlap_thresh=0.7, rank=None, world=None, group=None, device=None):
        D.require_cuda()
        self.group = group
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rank, self.world = int(rank), int(world)
        self.n_panels, self.batch, self.seed = int(n_panels), int(batch), int(seed)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        dev = self.device
        self.n_local = len(range(self.rank, self.n_panels, self.world))
        per_rank_max = -(-self.n_panels // self.world)
        self.steps = -(-per_rank_max // self.batch)
        kw = dict(channels=channels, pool_size=pool_size, max_boxes=max_boxes, overlap_thresh=overlap_thresh, device=dev)
        self.pipe = ProposalPipeline(C, self.batch, H, W, **kw)
        tail = self.n_local % self.batch
        self.tail = tail
        self.tail_pipe = ProposalPipeline(C, tail, H, W, alloc_pooled=False, **kw) if tail else None
        A = self.pipe.A
        self.cls = D.empty((self.batch, H, W, A), np.float32, dev)
        self.regr = D.empty((self.batch, H, W, 4 * A), np.float32, dev)
        self.feat = D.empty((self.batch, H, W, channels), np.float32, dev)
        stride = self.pipe.records.stride
        self.stride = stride
        # every step's gathered records: [step][rank][slot]
        self.gathered = torch.zeros((self.steps, self.world, self.batch, stride), dtype=torch.uint8, device=dev)
        self._send_tail = torch.zeros((self.batch, stride), dtype=torch.uint8, device=dev)
        # global panel i lives at step (i // world) // batch, rank i % world, slot (i // world) % batch
        i = np.arange(self.n_panels, dtype=np.int64)
        j = i // self.world
        flat = ((j // self.batch) * self.world + i % self.world) * self.batch + j % self.batch
        self._order = torch.from_numpy(flat).to(dev)
        self.launches = 0

    def run(self, check=True):
        pipe, w, r = self.pipe, self.world, self.rank
        self.launches = 0
        for step in range(self.steps):
            first_local = step * self.batch
            n = max(0, min(self.batch, self.n_local - first_local))
            first_global = r + w * first_local
            if n == self.batch:
                generate_panels(self.seed, first_global, w, self.cls, self.regr, self.feat)
                pipe.decode(self.cls, self.regr)
                pipe.sort_nms()
                send = pipe.records.raw
                self.launches += 3
            elif n > 0:
                tp = self.tail_pipe
                generate_panels(self.seed, first_global, w, self.cls[:n], self.regr[:n], self.feat[:n])
                tp.decode(self.cls[:n], self.regr[:n])
                tp.sort_nms()
                self._send_tail[:n].copy_(tp.records.raw)
                send = self._send_tail
                self.launches += 3
            else:
                send = self._send_tail.zero_()
            work = None
            if w > 1:
                _, work = sharding.gather_detections(send, group=self.group, async_op=True, out=self.gathered[step])
            else:
                self.gathered[step, 0].copy_(send)
            if n == self.batch:
                pipe.pool(self.feat)
                self.launches += 1
            elif n > 0:
                self.tail_pipe.pool(self.feat[:n], out=pipe.pooled[:n])
                self.launches += 1
            if work is not None:
                work.wait()
            if check and n > 0:
                self._check = True
        out = self.gathered.reshape(-1, self.stride).index_select(0, self._order)
        return out

    def records(self, raw):
        """Typed views over the result of `run`."""
        return DetectionRecords(self.n_panels, self.pipe.max_boxes, raw.device, raw=raw)
