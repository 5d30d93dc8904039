"""Drop-in for the hot-path functions of the reference `faster_rcnn/utils.py`:
`calc_region_props` (upstream keras-frcnn name `calc_rpn`), `get_new_img_size`, `iou`.

The anchor x GT IoU / label / regression-target computation runs in
libradnet_b200.so (`radnet_rpn_targets`); the host keeps what the reference also
does on the host with the global NumPy RNG: the 256-region subsampling.
"""
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


def _subsample_regions(y_is_box_valid, y_rpn_overlap, max_n_regions=256):
    """Host half of calc_region_props: random balancing to 256 regions with the legacy global
    NumPy RNG, in place on the (1,A,H,W) arrays (reference utils.py:777-813).  Returns n_pos."""
    where_pos = np.where(np.logical_and(y_rpn_overlap[0] == 1, y_is_box_valid[0] == 1))
    where_neg = np.where(np.logical_and(y_rpn_overlap[0] == 0, y_is_box_valid[0] == 1))
    n_pos, n_neg = len(where_pos[0]), len(where_neg[0])
    half = int(max_n_regions / 2)

    def channel_weights(channels, total):
        # per-candidate probability = (share of its anchor channel) / (size of that channel);
        # the table is keyed by the NEGATIVE channel ids in both branches, as in utils.py:789,804
        ids, counts = np.unique(where_neg[0], return_counts=True)
        share = dict(zip(ids, counts / total))
        size = dict(zip(ids, counts))
        return [share[c] / size[c] for c in channels]

    if n_pos > max_n_regions / 2:
        drop = np.random.choice(n_pos, n_pos - half, replace=False, p=channel_weights(where_pos[0], n_pos))
        y_is_box_valid[0, where_pos[0][drop], where_pos[1][drop], where_pos[2][drop]] = 0
        n_pos = half
    if n_neg + n_pos > max_n_regions:
        drop = np.random.choice(n_neg, n_neg - n_pos, replace=False, p=channel_weights(where_neg[0], n_neg))
        y_is_box_valid[0, where_neg[0][drop], where_neg[1][drop], where_neg[2][drop]] = 0
    return n_pos


def calc_region_props(C, img_data, width, height, width_resized, height_resized, get_feat_map_size,
                      verbose=False):
    """RPN anchor targets for one image (reference utils.py:554-821).

    Returns (y_rpn_cls (1,2A,fh,fw) float64 = [valid | overlap],
             y_rpn_regr (1,8A,fh,fw) float64 = [repeat(overlap,4) | regr],
             best_anchor_for_bbox (G,4) int64, n_pos)."""
    fw, fh = get_feat_map_size(width_resized, height_resized)                 # utils.py:592
    bboxes = img_data['bboxes']
    G = len(bboxes)
    gt = np.zeros((1, max(G, 1), 4))
    is_bg = np.zeros((1, max(G, 1)), dtype=np.uint8)
    for k, bb in enumerate(bboxes):                                           # utils.py:608-613
        gt[0, k, 0] = bb['x1'] * (width_resized / float(width))
        gt[0, k, 1] = bb['x2'] * (width_resized / float(width))
        gt[0, k, 2] = bb['y1'] * (height_resized / float(height))
        gt[0, k, 3] = bb['y2'] * (height_resized / float(height))
        is_bg[0, k] = 1 if bb['class'] == 'bg' else 0
    y_cls_d, y_regr_d, best_d, _ = rpn_targets_device(
        C, gt if G else None, is_bg if G else None, np.array([G], dtype=np.int32), int(fh), int(fw),
        np.array([[float(width_resized), float(height_resized)]]))
    A = y_cls_d.shape[1] // 2
    y_cls = y_cls_d.cpu().numpy()
    y_rpn_regr = y_regr_d.cpu().numpy()
    y_is_box_valid = y_cls[:, :A]
    y_rpn_overlap = y_cls[:, A:]
    n_pos = _subsample_regions(y_is_box_valid, y_rpn_overlap)
    best_anchor_for_bbox = best_d[0].cpu().numpy().astype(np.int64).reshape(G, 4)
    return np.copy(y_cls), np.copy(y_rpn_regr), best_anchor_for_bbox, n_pos


calc_rpn = calc_region_props
