"""Batched, device-resident proposal -> NMS -> RoI-pool pipeline.

`ProposalPipeline` keeps every buffer of a batch of panels in HBM (decoded boxes,
sort scratch, detection records, pooled features) and launches the three kernels
back to back on the caller's stream:

    K1 radnet_decode_clip_i32   (reference rpn.py:91-166)
    K2 radnet_sort_nms_i32      (reference rpn.py:380-456, called at rpn.py:170)
    K4 radnet_roi_pool          (reference RoiPoolingConv.py:48-88, fed as RADNet.py:564-568)

It is what `bench.py` times and what the single-panel drop-in functions in
`rpn.py` / `RoiPoolingConv.py` are thin wrappers of (B = 1).
"""
import ctypes

import numpy as np
import torch

from . import _device as D
from . import _lib


def anchor_cells(C):
    """Anchor (w,h) in feature cells, a = size-major, ratio-minor (reference rpn.py:108-113)."""
    out = []
    for scale in C.anchor_box_scales:
        for ratio in C.anchor_box_ratios:
            out.append(((scale * ratio[0]) / C.rpn_stride, (scale * ratio[1]) / C.rpn_stride))
    return D.host_f64(out)


def anchor_pixels(C):
    """Anchor (w,h) in pixels, a = ratio_idx + n_ratios*size_idx (reference utils.py:616-620,725)."""
    out = []
    for scale in C.anchor_box_scales:
        for ratio in C.anchor_box_ratios:
            out.append((scale * ratio[0], scale * ratio[1]))
    return D.host_f64(out)


class DetectionRecords:
    """Typed views over the per-panel detection records written by K2
    (layout: include/radnet_b200.h, `radnet_det_record_bytes`)."""

    def __init__(self, batch, max_boxes, device, raw=None):
        self.batch = batch
        self.max_boxes = max_boxes
        self.stride = _lib.det_record_bytes(max_boxes)
        self.raw = raw if raw is not None else torch.zeros((batch, self.stride), dtype=torch.uint8, device=device)
        words = self.raw.view(torch.int32)
        k = max_boxes
        self.header = words[:, 0:4]                                   # count, n_candidates, ties, n_sorted
        self.boxes = words[:, 4:4 + 4 * k].view(batch, k, 4)          # x1,y1,x2,y2
        self.scores = words[:, 4 + 4 * k:4 + 5 * k].view(torch.float32)
        self.index = words[:, 4 + 5 * k:4 + 6 * k]

    @property
    def counts(self):
        return self.header[:, 0]

    def to_numpy(self):
        """-> list of dicts with int64 boxes (k,4), float32 scores, int64 flat indices."""
        hdr = self.header.cpu().numpy()
        if (hdr[:, 0] < 0).any():
            raise RuntimeError("radnet_sort_nms_i32: a row hand-off did not complete within 2 s; results invalid")
        boxes = self.boxes.cpu().numpy()
        scores = self.scores.cpu().numpy()
        index = self.index.cpu().numpy()
        out = []
        for b in range(self.batch):
            n = int(hdr[b, 0])
            out.append({"boxes": boxes[b, :n].astype(np.int64), "scores": scores[b, :n].copy(),
                        "index": index[b, :n].astype(np.int64), "n_candidates": int(hdr[b, 1]),
                        "score_ties": int(hdr[b, 2]), "n_sorted": int(hdr[b, 3])})
        return out


class ProposalPipeline:
    """Decode + NMS + RoI pool for `batch` panels with H x W x A anchors each."""

    def __init__(self, C, batch, H, W, channels=1024, pool_size=14, max_boxes=300,
                 overlap_thresh=0.7, device=None, alloc_pooled=True):
        D.require_cuda()
        _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.C = C
        self.batch, self.H, self.W = int(batch), int(H), int(W)
        self.A = len(C.anchor_box_scales) * len(C.anchor_box_ratios)
        self.N = self.H * self.W * self.A
        self.channels, self.pool_size = int(channels), int(pool_size)
        self.max_boxes, self.overlap_thresh = int(max_boxes), float(overlap_thresh)
        self.std_scaling = float(C.std_scaling)
        self._anchors = anchor_cells(C)
        dev = self.device
        self.boxes = D.empty((self.batch, self.N, 4), np.int32, dev)
        self.keys = torch.empty((self.batch, self.N), dtype=torch.int32, device=dev)      # uint32 bits
        self.stats = D.zeros((self.batch, 4), np.int32, dev)
        self.records = DetectionRecords(self.batch, self.max_boxes, dev)
        lib = _lib.load()
        self._ws_bytes = int(lib.radnet_sort_nms_i32_workspace_bytes(self.batch, self.N, self.H, self.W, self.max_boxes))
        self._ws = torch.empty((max(self._ws_bytes, 16),), dtype=torch.uint8, device=dev)
        self.pooled = None
        if alloc_pooled:
            self.pooled = D.empty((self.batch, self.max_boxes, self.pool_size, self.pool_size, self.channels),
                                  np.float32, dev)

    # -- individual stages (all asynchronous on the current stream) ------------------
    def decode(self, cls, regr, use_regr=True):
        assert cls.shape == (self.batch, self.H, self.W, self.A), cls.shape
        assert regr.shape == (self.batch, self.H, self.W, 4 * self.A), regr.shape
        _lib.call("radnet_decode_clip_i32", D.ptr(cls), D.ptr(regr), self.batch, self.H, self.W, self.A,
                  D.ptr(self._anchors), ctypes.c_float(self.std_scaling), 1 if use_regr else 0,
                  D.ptr(self.boxes), D.ptr(self.keys), D.ptr(self.stats), D.stream_ptr(self.device))

    def sort_nms(self):
        _lib.call("radnet_sort_nms_i32", D.ptr(self.boxes), D.ptr(self.keys), self.batch, self.N, self.H, self.W,
                  self.overlap_thresh, self.max_boxes, D.ptr(self.records.raw), D.ptr(self._ws),
                  self._ws_bytes, D.stream_ptr(self.device))

    def pool(self, feat, out=None):
        out = self.pooled if out is None else out
        assert feat.shape == (self.batch, self.H, self.W, self.channels), feat.shape
        _lib.call("radnet_roi_pool", D.ptr(feat), self.batch, self.H, self.W, self.channels,
                  D.ptr(self.records.raw), self.max_boxes, None, None, self.max_boxes, self.pool_size,
                  D.ptr(out), D.stream_ptr(self.device))
        return out

    def pool_form(self):
        """The launch form K4 settled on for this shape (include/radnet_b200.h, radnet_roi_pool_form):
        {"lanes", "cluster", "sync_every"}, or None before the first large call."""
        out = (ctypes.c_int * 3)()
        _lib.call("radnet_roi_pool_form", self.batch, self.H, self.W, self.channels, self.pool_size, self.max_boxes, out)
        return None if out[0] < 0 else {"lanes": int(out[0]), "cluster": int(out[1]), "sync_every": int(out[2])}

    def __call__(self, cls, regr, feat):
        """cls (B,H,W,A), regr (B,H,W,4A), feat (B,H,W,C): float32 CUDA tensors.
        Returns (DetectionRecords, pooled (B,max_boxes,pool,pool,C)); nothing is synchronised."""
        self.decode(cls, regr)
        self.sort_nms()
        pooled = self.pool(feat)
        return self.records, pooled

    # -- bookkeeping ----------------------------------------------------------------
    def launches_per_step(self):
        """Kernels of ours launched by one __call__ (decode, sort+nms, pool)."""
        return 3

    def check_stats(self):
        """Synchronises.  Raises like the reference when a panel has no candidate left or a
        non-finite box reaches the NMS assert (reference rpn.py:170, rpn.py:400-401)."""
        st = self.stats.cpu().numpy()
        if (st[:, 1] > 0).any():
            raise AssertionError("non-finite proposal coordinates (np.testing.assert_array_less, rpn.py:400)")
        if (st[:, 0] == 0).any():
            raise ValueError("not enough values to unpack (expected 2, got 0)")
        return st


class PipelinedProposalStream:
    """Software pipeline over a stream of device-resident batches.

    Decode + sort + NMS (tens of microseconds, a few SMs) of batch k+1 run on their own stream while the
    RoI pool of batch k streams its 15 GB to HBM; two sets of the small buffers (boxes, keys, records,
    scratch) alternate and the pooled output is shared, so batch k+1's proposals are ready the moment
    batch k's pool retires.  Results are those of `ProposalPipeline.__call__`, batch by batch.
    Measured on the bench workload: 2.685 ms per 64-panel step instead of 2.725 ms (+1.5 %); `bench.py` keeps
    the plain sequential step so that its per-kernel event times stay kernel durations."""

    def __init__(self, C, batch, H, W, device=None, **kw):
        kw.pop("alloc_pooled", None)
        self.pipes = [ProposalPipeline(C, batch, H, W, device=device, alloc_pooled=(i == 0), **kw) for i in range(2)]
        self.device = self.pipes[0].device
        self.pooled = self.pipes[0].pooled
        self.prop_stream = torch.cuda.Stream(device=self.device)
        self.pool_stream = torch.cuda.Stream(device=self.device)
        self.proposed = [torch.cuda.Event() for _ in range(2)]
        self.pooled_done = [torch.cuda.Event() for _ in range(2)]
        self._n = 0

    def begin(self):
        """Order both internal streams after the work already queued on the caller's stream."""
        cur = torch.cuda.current_stream(self.device)
        self.prop_stream.wait_stream(cur)
        self.pool_stream.wait_stream(cur)

    def submit(self, cls, regr, feat, timing=None, after_nms=None):
        """Queue one batch; returns (DetectionRecords of this batch's slot, pooled).  `timing`, if given, is
        a list of four timing events recorded around decode, NMS and pool on the streams they run on;
        `after_nms(pipe)` is called on the proposal stream right after the NMS (e.g. to start a gather)."""
        s = self._n % 2
        pipe = self.pipes[s]
        with torch.cuda.stream(self.prop_stream):
            if self._n >= 2:
                self.prop_stream.wait_event(self.pooled_done[s])      # slot's records no longer read by a pool
            if timing:
                timing[0].record(self.prop_stream)
            pipe.decode(cls, regr)
            if timing:
                timing[1].record(self.prop_stream)
            pipe.sort_nms()
            if timing:
                timing[2].record(self.prop_stream)
            if after_nms is not None:
                after_nms(pipe)
            self.proposed[s].record(self.prop_stream)
        with torch.cuda.stream(self.pool_stream):
            self.pool_stream.wait_event(self.proposed[s])
            if timing:
                timing[3].record(self.pool_stream)
            pipe.pool(feat, out=self.pooled)
            if timing:
                timing[4].record(self.pool_stream)
            self.pooled_done[s].record(self.pool_stream)
        self._n += 1
        return pipe.records, self.pooled

    def end(self):
        """Make the caller's stream wait for everything queued so far."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.prop_stream)
        cur.wait_stream(self.pool_stream)


class HostPanelStream:
    """Public end-to-end entry for HOST buffers: batches of panels whose RPN maps and feature
    maps live in (pinned) host memory go host -> device -> decode -> sort+NMS -> RoI pool, and
    the detection records come back to the host.  Pooled features stay on the device, where
    the classifier head consumes them (in the reference they never leave the TF graph either,
    RoiPoolingConv.py:75-86).

    Two device input slots and a dedicated copy stream overlap the upload of batch k+1 with
    the kernels of batch k; `submit` returns immediately, `collect` blocks for one batch.
    """

    def __init__(self, pipe, slots=2):
        self.pipe = pipe
        dev = pipe.device
        B, H, W, A, Cn = pipe.batch, pipe.H, pipe.W, pipe.A, pipe.channels
        self.slots = slots
        self.cls = [D.empty((B, H, W, A), np.float32, dev) for _ in range(slots)]
        self.regr = [D.empty((B, H, W, 4 * A), np.float32, dev) for _ in range(slots)]
        self.feat = [D.empty((B, H, W, Cn), np.float32, dev) for _ in range(slots)]
        self.rec_host = [torch.empty((B, pipe.records.stride), dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.compute_stream = torch.cuda.Stream(device=dev)
        self.uploaded = [torch.cuda.Event() for _ in range(slots)]
        self.consumed = [torch.cuda.Event() for _ in range(slots)]
        self.done = [torch.cuda.Event() for _ in range(slots)]
        # `pipe.pooled` is ONE buffer shared by consecutive batches: a consumer on another stream waits on
        # `pooled_ready` before reading it and hands back an event through `release_pooled` when it is done
        self.pooled_ready = torch.cuda.Event()
        self.pooled_consumed = None
        self._n = 0
        self._pending = []
        self.h2d_bytes_per_batch = 4 * (B * H * W * A * 5 + B * H * W * Cn)
        self.d2h_bytes_per_batch = B * pipe.records.stride

    def submit(self, cls_h, regr_h, feat_h):
        """cls_h/regr_h/feat_h: float32 CPU tensors (pinned for asynchronous copies).  At most `slots` batches may
        be in flight: a further submit before `collect` would overwrite the host records of an uncollected batch."""
        if len(self._pending) >= self.slots:
            raise RuntimeError("HostPanelStream: %d batches in flight - collect() before submitting more" % self.slots)
        s = self._n % self.slots
        if self._n >= self.slots:
            self.copy_stream.wait_event(self.consumed[s])      # slot inputs no longer being read
        with torch.cuda.stream(self.copy_stream):
            self.cls[s].copy_(cls_h, non_blocking=True)
            self.regr[s].copy_(regr_h, non_blocking=True)
            self.feat[s].copy_(feat_h, non_blocking=True)
            self.uploaded[s].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.uploaded[s])
            if self.pooled_consumed is not None:
                self.compute_stream.wait_event(self.pooled_consumed)   # the head has read the previous pooled batch
            rec, _ = self.pipe(self.cls[s], self.regr[s], self.feat[s])
            self.pooled_ready.record(self.compute_stream)
            self.consumed[s].record(self.compute_stream)
            self.rec_host[s].copy_(rec.raw, non_blocking=True)
            self.done[s].record(self.compute_stream)
        self._pending.append(s)
        self._n += 1

    def release_pooled(self, event):
        """`event`: recorded by the consumer of `pipe.pooled` (the classifier head) on its own stream once it has
        read the current batch; the next RoI pool waits for it before overwriting the buffer."""
        self.pooled_consumed = event

    def collect(self):
        """Wait for the oldest submitted batch; returns its records as a host uint8 tensor."""
        s = self._pending.pop(0)
        self.done[s].synchronize()
        return self.rec_host[s]


class DetectionPipeline(ProposalPipeline):
    """ProposalPipeline plus the stages that follow the classifier head:

        K1 -> K2 -> K4 -> [classifier head, user code on the same stream] -> K5+K6 fused
        (radnet_classify_nms: head decode, per-class NMS at 0.2, real coordinates, tile offset)

    and, once the records of all tiles of a panel are on one GPU (`sharding.gather_detections`
    works on any fixed-size record), `merge`: K7 radnet_final_nms per image followed by the
    cross-image per-class NMS at 0.4 (reference RADNet.py:566-716).  Every buffer is resident; nothing
    synchronises."""

    def __init__(self, C, batch, H, W, n_cls=None, bbox_threshold=0.7, **kw):
        super().__init__(C, batch, H, W, **kw)
        from . import detect as DT
        self._DT = DT
        self.n_cls = int(n_cls if n_cls is not None else len(C.class_mapping))
        self.bbox_threshold = float(bbox_threshold)
        self.class_records = DT.ClassRecords(self.batch, self.max_boxes, self.device)

    def classify(self, P_cls, P_regr, ratio=None, origin=None, nms_thresh=0.2):
        """P_cls (B,max_boxes,n_cls), P_regr (B,max_boxes,4(n_cls-1)) float32 CUDA tensors for the kept
        boxes of `self.records` (rows >= count are ignored); ratio (B,) / origin (B,2) optional."""
        return self._DT.classify_nms(P_cls, P_regr, self.C, det=self.records, bbox_threshold=self.bbox_threshold,
                                     nms_thresh=nms_thresh, max_boxes=300, ratio=ratio, origin=origin,
                                     out=self.class_records)

    def merge(self, tile_records, n_images, tiles_per_image, final_thresh=0.4):
        """tile_records: ClassRecords of n_images * tiles_per_image tiles in (image, tile) order.
        Returns (per-image merged ClassRecords, final ClassRecords(1)) - RADNet.py:672, 698."""
        DT = self._DT
        merged = DT.final_nms_records(tile_records, n_images, tiles_per_image, self.n_cls)
        final = DT.class_nms(merged, 1, n_images, self.n_cls, final_thresh, max_boxes=300)
        return merged, final


class SharedBatchCounter:
    """A job-wide counter of batches handed out to the ranks of one box (rate-aware host streaming): every rank pulls
    the index of its next batch, so ranks whose host link is faster take more batches - a static equal split runs at
    the pace of the slowest link (on the 8-GPU pool box four links move 23 GB/s and four 36 GB/s,
    profiles/r01_topo_h2d.txt).  Backed by the process group's key-value store (`store.add` is atomic; one round
    trip of tens of microseconds per batch of milliseconds).  Without a process group it is a local counter."""

    def __init__(self, total, name="radnet_batches"):
        import torch.distributed as dist
        self.total, self.name = int(total), name
        self._local = 0
        self._store = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._store = dist.distributed_c10d._get_default_store()

    def next(self):
        """Index of the next batch, or None when all `total` batches have been handed out."""
        if self._store is None:
            i = self._local
            self._local += 1
        else:
            i = int(self._store.add(self.name, 1)) - 1
        return i if i < self.total else None
