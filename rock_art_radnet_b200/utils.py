"""Drop-in for the hot-path functions of the reference `faster_rcnn/utils.py`:
`calc_region_props` (upstream keras-frcnn name `calc_rpn`), `get_new_img_size`, `iou`.

The anchor x GT IoU / label / regression-target computation runs in
libradnet_b200.so (`radnet_rpn_targets`), and so does the 256-region subsampling the
reference draws from the global NumPy generator (`radnet_rpn_subsample` replays the
legacy MT19937 `np.random.choice` stream; the host only hands the generator state over).
"""
import threading

import numpy as np
import torch

from . import _device as D
from . import _lib
from .pipeline import anchor_pixels


def get_new_img_size(width, height, img_min_side=300):
    """(resized_width, resized_height) with the short side at img_min_side (reference utils.py:65-75)."""
    if width <= height:
        f = float(img_min_side) / width
        return img_min_side, int(f * height)
    f = float(img_min_side) / height
    return int(f * width), img_min_side


def iou_pairs(a, b):
    """IoU of n box pairs, a and b (n,4) as (x1,y1,x2,y2) -> (n,) float64 (reference utils.py:77-109,
    evaluated by `radnet_iou_pairs` on the device)."""
    D.require_cuda()
    dev = torch.device("cuda:%d" % torch.cuda.current_device())
    a_d = D.to_device(np.asarray(a, dtype=np.float64).reshape(-1, 4), np.float64, dev)
    b_d = D.to_device(np.asarray(b, dtype=np.float64).reshape(-1, 4), np.float64, dev)
    if a_d.shape != b_d.shape:
        raise ValueError("iou_pairs: a and b must hold the same number of boxes")
    out = D.empty((int(a_d.shape[0]),), np.float64, dev)
    _lib.call("radnet_iou_pairs", D.ptr(a_d), D.ptr(b_d), int(a_d.shape[0]), D.ptr(out), D.stream_ptr(dev))
    return out.cpu().numpy()


def iou(a, b):
    """Scalar IoU of (x1,y1,x2,y2) boxes (reference utils.py:99-109): 0.0 for degenerate boxes, else
    inter / (union + 1e-6).  Same device routine as the target-assignment kernels use."""
    return float(iou_pairs([a[:4]], [b[:4]])[0])


LAYOUT_CHANNEL_FIRST = 0      # RADNET_TARGETS_CHANNEL_FIRST: the reference's return layout (utils.py:815-816)
LAYOUT_NHWC = 1               # RADNET_TARGETS_NHWC: what the training loop feeds the losses (utils.py:477-478)


class RpnTargetBatch:
    """Pre-allocated batched RPN target assignment (K3) for B panels with up to Gmax figures each.

    `run` launches `radnet_rpn_targets` (one kernel) on the current stream and returns the resident
    output tensors (overwritten by the next call) - BEFORE the RNG subsampling of utils.py:777-813:

      layout=LAYOUT_CHANNEL_FIRST  y_rpn_cls (B,2A,H,W), y_rpn_regr (B,8A,H,W) float64
      layout=LAYOUT_NHWC           y_rpn_cls (B,H,W,2A), y_rpn_regr (B,H,W,8A) float64 with the regr half
                                   multiplied by `regr_scale` (C.std_scaling: utils.py:475)
      best_anchor (B,Gmax,4) int32, n_hits (B,Gmax) int32."""

    def __init__(self, C, batch, Gmax, H, W, device=None, layout=LAYOUT_CHANNEL_FIRST, regr_scale=1.0):
        D.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.C, self.B, self.Gmax, self.H, self.W = C, int(batch), int(Gmax), int(H), int(W)
        self.A = len(C.anchor_box_scales) * len(C.anchor_box_ratios)
        self.layout, self.regr_scale = int(layout), float(regr_scale)
        dev = self.device
        A, B = self.A, self.B
        if self.layout == LAYOUT_NHWC:
            self.y_cls = D.empty((B, self.H, self.W, 2 * A), np.float64, dev)
            self.y_regr = D.empty((B, self.H, self.W, 8 * A), np.float64, dev)
        else:
            self.y_cls = D.empty((B, 2 * A, self.H, self.W), np.float64, dev)
            self.y_regr = D.empty((B, 8 * A, self.H, self.W), np.float64, dev)
        self.best = D.empty((B, max(self.Gmax, 1), 4), np.int32, dev)
        self.hits = D.zeros((B, max(self.Gmax, 1)), np.int32, dev)
        self.ws_bytes = int(self.lib.radnet_rpn_targets_workspace_bytes(B, self.Gmax, self.H, self.W, A))
        self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=dev)
        self.reset_workspace()
        self._anchors = anchor_pixels(C)

    def reset_workspace(self):
        """Zero the per-panel state at the head of the workspace (once after allocation; every
        successful launch leaves it zeroed again)."""
        _lib.call("radnet_rpn_targets_workspace_init", D.ptr(self.ws), self.ws_bytes, self.B, self.Gmax,
                  D.stream_ptr(self.device))

    def run(self, gt_boxes, gt_is_bg, gt_count, img_wh):
        """gt_boxes (B,Gmax,4) float64 x1,x2,y1,y2 in resized pixels, gt_is_bg (B,Gmax) uint8,
        gt_count (B,) int32, img_wh (B,2) float64 - CUDA tensors."""
        C = self.C
        _lib.call("radnet_rpn_targets", D.ptr(gt_boxes), D.ptr(gt_is_bg), D.ptr(gt_count), self.B, self.Gmax,
                  self.H, self.W, self.A, len(C.anchor_box_ratios), D.ptr(self._anchors), float(C.rpn_stride),
                  D.ptr(img_wh), float(C.rpn_max_overlap), self.layout, self.regr_scale, D.ptr(self.y_cls),
                  D.ptr(self.y_regr), D.ptr(self.best), D.ptr(self.hits), D.ptr(self.ws), self.ws_bytes,
                  D.stream_ptr(self.device))
        return self.y_cls, self.y_regr, self.best[:, :self.Gmax], self.hits[:, :self.Gmax]


def rpn_targets_device(C, gt_boxes, gt_is_bg, gt_count, H, W, img_wh, device=None,
                       layout=LAYOUT_CHANNEL_FIRST, regr_scale=1.0):
    """One-shot batched call (allocates its outputs).  gt_boxes (B,Gmax,4) float64 x1,x2,y1,y2 in
    resized pixels, gt_is_bg (B,Gmax) uint8, gt_count (B,) int32, img_wh (B,2) float64; NumPy or
    CUDA tensors.  Returns CUDA tensors as RpnTargetBatch.run."""
    D.require_cuda()
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    gt_count = D.to_device(gt_count, np.int32, dev)
    B = int(gt_count.shape[0])
    Gmax = int(gt_boxes.shape[1]) if gt_boxes is not None else 0
    batch = RpnTargetBatch(C, B, Gmax, H, W, device=dev, layout=layout, regr_scale=regr_scale)
    gt_dev = D.to_device(gt_boxes, np.float64, dev) if Gmax else None
    bg_dev = D.to_device(gt_is_bg, np.uint8, dev) if Gmax else None
    wh_dev = D.to_device(img_wh, np.float64, dev)
    return batch.run(gt_dev, bg_dev, gt_count, wh_dev)


class _RegionPropsContext:
    """Everything one `calc_region_props` call needs, allocated once per (config, map size, figure capacity)
    and reused: K3 for one panel, the device sampler, ONE packed device buffer for the small inputs and
    outputs with a pinned host mirror, pinned landing buffers for the two target tensors.

    Packed buffer (bytes): [gt Gcap*32 | img_wh 16 | count 4 (+12) | is_bg Gcap (padded) | rng state 2512 |
    best_anchor Gcap*16 | n_hits Gcap*4 (padded) | sampler report 32].  One H2D copy ships everything up to
    and including the rng state, one D2H copy brings back everything from the rng state on."""

    def __init__(self, C, H, W, Gcap, dev):
        from .sampling import RpnSubsampler
        self.C, self.H, self.W, self.Gcap, self.dev = C, H, W, Gcap, dev
        self.batch = RpnTargetBatch(C, 1, Gcap, H, W, device=dev)
        self.A = self.batch.A
        self.sampler = RpnSubsampler(1, H, W, self.A, device=dev)
        pad16 = lambda v: (v + 15) // 16 * 16
        self.o_gt = 0
        self.o_wh = Gcap * 32
        self.o_cnt = self.o_wh + 16
        self.o_bg = self.o_cnt + 16
        self.o_state = self.o_bg + pad16(Gcap)
        self.o_best = self.o_state + pad16(625 * 4)
        self.o_hits = self.o_best + Gcap * 16
        self.o_rep = self.o_hits + pad16(Gcap * 4)
        self.n_bytes = self.o_rep + 32
        self.dev_buf = torch.zeros((self.n_bytes,), dtype=torch.uint8, device=dev)
        self.host_buf = torch.zeros((self.n_bytes,), dtype=torch.uint8).pin_memory()
        hb = self.host_buf.numpy()
        self.h_gt = hb[self.o_gt:self.o_wh].view(np.float64).reshape(Gcap, 4)
        self.h_wh = hb[self.o_wh:self.o_wh + 16].view(np.float64)
        self.h_cnt = hb[self.o_cnt:self.o_cnt + 4].view(np.int32)
        self.h_bg = hb[self.o_bg:self.o_bg + Gcap]
        self.h_state = hb[self.o_state:self.o_state + 2500].view(np.uint32)
        self.h_best = hb[self.o_best:self.o_hits].view(np.int32).reshape(Gcap, 4)
        self.h_hits = hb[self.o_hits:self.o_hits + 4 * Gcap].view(np.int32)
        self.h_rep = hb[self.o_rep:self.o_rep + 32].view(np.int32)
        db = self.dev_buf
        self.d_gt = db[self.o_gt:self.o_wh]
        self.d_wh = db[self.o_wh:self.o_wh + 16]
        self.d_cnt = db[self.o_cnt:self.o_cnt + 4]
        self.d_bg = db[self.o_bg:self.o_bg + Gcap]
        self.d_state = db[self.o_state:self.o_state + 2500]
        self.d_rep = db[self.o_rep:self.o_rep + 32]
        # K3 writes best_anchor / n_hits straight into the packed buffer
        self.batch.best = db[self.o_best:self.o_hits].view(torch.int32).view(1, Gcap, 4)
        self.batch.hits = db[self.o_hits:self.o_hits + 4 * Gcap].view(torch.int32).view(1, Gcap)
        self.h_cls = torch.empty(tuple(self.batch.y_cls.shape), dtype=torch.float64).pin_memory()
        self.h_regr = torch.empty(tuple(self.batch.y_regr.shape), dtype=torch.float64).pin_memory()


_CONTEXTS = {}
_CONTEXT_LOCK = threading.Lock()


def _region_props_context(C, H, W, G, dev):
    Gcap = max(32, (G + 31) // 32 * 32)
    key = (tuple(C.anchor_box_scales), tuple(map(tuple, C.anchor_box_ratios)), float(C.rpn_stride),
           float(C.rpn_max_overlap), H, W, Gcap, dev.index)
    ctx = _CONTEXTS.get(key)
    if ctx is None:
        if len(_CONTEXTS) > 16:
            _CONTEXTS.clear()
        ctx = _RegionPropsContext(C, H, W, Gcap, dev)
        _CONTEXTS[key] = ctx
    return ctx


def calc_region_props(C, img_data, width, height, width_resized, height_resized, get_feat_map_size,
                      verbose=False):
    """RPN anchor targets for one image (reference utils.py:554-821), including the random balancing to 256
    regions with NumPy's global legacy generator (utils.py:777-813; `np.random`'s state is read, replayed on
    the device and advanced exactly as the reference advances it).

    Returns (y_rpn_cls (1,2A,fh,fw) float64 = [valid | overlap],
             y_rpn_regr (1,8A,fh,fw) float64 = [repeat(overlap,4) | regr],
             best_anchor_for_bbox (G,4) int64, n_pos).
    KeyError, like the reference, when more than 128 anchors are positive and one of them sits in an anchor
    channel without any negative (utils.py:789-795)."""
    from .sampling import numpy_state_words, set_numpy_state
    D.require_cuda()
    dev = torch.device("cuda:%d" % torch.cuda.current_device())
    fw, fh = get_feat_map_size(width_resized, height_resized)                 # utils.py:592
    bboxes = img_data['bboxes']
    G = len(bboxes)
    with _CONTEXT_LOCK:
        ctx = _region_props_context(C, int(fh), int(fw), G, dev)
        sx, sy = width_resized / float(width), height_resized / float(height)
        for k, bb in enumerate(bboxes):                                       # utils.py:608-613
            ctx.h_gt[k, 0] = bb['x1'] * sx
            ctx.h_gt[k, 1] = bb['x2'] * sx
            ctx.h_gt[k, 2] = bb['y1'] * sy
            ctx.h_gt[k, 3] = bb['y2'] * sy
            ctx.h_bg[k] = 1 if bb['class'] == 'bg' else 0
        ctx.h_wh[0], ctx.h_wh[1] = float(width_resized), float(height_resized)
        ctx.h_cnt[0] = G
        ctx.h_state[:625] = numpy_state_words()
        ctx.dev_buf[:ctx.o_best].copy_(ctx.host_buf[:ctx.o_best], non_blocking=True)
        y_cls_d, y_regr_d, _, _ = ctx.batch.run(ctx.d_gt, ctx.d_bg, ctx.d_cnt, ctx.d_wh)
        ctx.sampler.run(y_cls_d, ctx.d_state, out=ctx.d_rep)
        ctx.h_cls.copy_(y_cls_d, non_blocking=True)
        ctx.h_regr.copy_(y_regr_d, non_blocking=True)
        ctx.host_buf[ctx.o_state:].copy_(ctx.dev_buf[ctx.o_state:], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        rep = ctx.h_rep
        if G and (ctx.h_hits[:G] < 0).any():
            raise RuntimeError("radnet_rpn_targets: the fill of the panel never completed; results invalid")
        if rep[3] == 1:
            raise KeyError("anchor channel without negatives (pos_probs[l], reference utils.py:795)")
        if rep[5] > 0:
            set_numpy_state(ctx.h_state[:625])
        y_cls = ctx.h_cls.numpy().copy()
        y_rpn_regr = ctx.h_regr.numpy().copy()
        best_anchor_for_bbox = ctx.h_best[:G].astype(np.int64).reshape(G, 4)
        n_pos = int(rep[0])
    return y_cls, y_rpn_regr, best_anchor_for_bbox, n_pos


calc_rpn = calc_region_props
