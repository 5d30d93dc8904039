"""Drop-in for the hot-path functions of the reference `faster_rcnn/rpn.py`.

Same names, argument meaning, return arity, dtypes and error behaviour as the
reference; the arithmetic runs in libradnet_b200.so on the GPU.  Host code here
only validates arguments, stages NumPy arrays to the device and shapes results.
"""
import ctypes
import threading

import numpy as np
import torch

from . import _device as D
from . import _lib
from .pipeline import ProposalPipeline, anchor_cells
from .utils import get_new_img_size

_PIPELINES = {}
_PIPELINE_LOCK = threading.Lock()       # the cached single-panel pipelines own device buffers: one caller at a time


def _device():
    D.require_cuda()
    return torch.device("cuda:%d" % torch.cuda.current_device())


def _single_panel_pipeline(C, H, W, max_boxes, overlap_thresh):
    """One cached B=1 pipeline per (anchor set, map size, NMS setting) - buffers are reused."""
    key = (tuple(C.anchor_box_scales), tuple(map(tuple, C.anchor_box_ratios)), float(C.rpn_stride),
           float(C.std_scaling), H, W, int(max_boxes), float(overlap_thresh), torch.cuda.current_device())
    pipe = _PIPELINES.get(key)
    if pipe is None:
        if len(_PIPELINES) > 16:
            _PIPELINES.clear()
        pipe = ProposalPipeline(C, 1, H, W, max_boxes=max_boxes, overlap_thresh=overlap_thresh,
                                alloc_pooled=False)
        _PIPELINES[key] = pipe
    return pipe


def _int_table_ok(H, W):
    return 2 * (H - 1) * (W - 1) + 1 <= 65535


def rpn_to_roi(rpn_layer, regr_layer, C, use_regr=True, max_boxes=300, overlap_thresh=0.9):
    """RPN maps -> kept proposals, (K'<=max_boxes, 4) int64 x1,y1,x2,y2 in feature cells,
    score-descending (reference rpn.py:68-172).

    rpn_layer (1,H,W,A) and regr_layer (1,H,W,4A) may be NumPy arrays or CUDA tensors.
    AssertionError if batch != 1 (rpn.py:96) or a non-finite box reaches the NMS
    (rpn.py:400-401); ValueError when no candidate survives (rpn.py:170)."""
    assert rpn_layer.shape[0] == 1
    dev = _device()
    H, W, A = int(rpn_layer.shape[1]), int(rpn_layer.shape[2]), int(rpn_layer.shape[3])
    cls = D.to_device(rpn_layer, np.float32, dev)
    regr = D.to_device(regr_layer, np.float32, dev)
    if use_regr and _int_table_ok(H, W):
        with _PIPELINE_LOCK:
            pipe = _single_panel_pipeline(C, H, W, max(1, min(int(max_boxes), H * W * A)), overlap_thresh)
            pipe.decode(cls, regr, use_regr=True)
            pipe.sort_nms()
            pipe.check_stats()
            det = pipe.records.to_numpy()[0]
        return det["boxes"]
    # general path: float64 boxes (anchors without regression may be half-integers)
    n = H * W * A
    boxes = D.empty((1, n, 4), np.float64, dev)
    scores = D.empty((1, n), np.float32, dev)
    valid = D.empty((1, n), np.uint8, dev)
    stats = D.zeros((1, 4), np.int32, dev)
    _lib.call("radnet_decode_clip_f64", D.ptr(cls), D.ptr(regr), 1, H, W, A, D.ptr(anchor_cells(C)),
              ctypes.c_float(float(C.std_scaling)), 1 if use_regr else 0, D.ptr(boxes), D.ptr(scores),
              D.ptr(valid), D.ptr(stats), D.stream_ptr(dev))
    st = stats.cpu().numpy()[0]
    if st[1] > 0:
        raise AssertionError("non-finite proposal coordinates (np.testing.assert_array_less, rpn.py:400)")
    if st[0] == 0:
        raise ValueError("not enough values to unpack (expected 2, got 0)")
    pick, _ = _nms_device(boxes[0], scores[0].to(torch.float64), valid[0], overlap_thresh, max_boxes)
    return boxes[0][torch.from_numpy(pick).to(dev)].cpu().numpy().astype("int")


def _nms_device(boxes_dev, probs_dev, valid_dev, overlap_thresh, max_boxes):
    """float64 NMS on device tensors; returns (pick int64 ndarray, n_score_ties)."""
    dev = boxes_dev.device
    M = int(boxes_dev.shape[0])
    lib = _lib.load()
    # the reference picks the top box before it tests max_boxes (rpn.py:425,449): at least 1
    mb = max(1, min(int(max_boxes), M))
    ws_bytes = int(lib.radnet_nms_f64_workspace_bytes(M, mb))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    pick = D.empty((min(mb, M),), np.int32, dev)
    count = D.zeros((3,), np.int32, dev)
    _lib.call("radnet_nms_f64", D.ptr(boxes_dev), D.ptr(probs_dev), D.ptr(valid_dev), M,
              float(overlap_thresh), mb, D.ptr(pick), D.ptr(count), D.ptr(ws), ws_bytes, D.stream_ptr(dev))
    cnt = count.cpu().numpy()
    if cnt[0] < 0:
        raise RuntimeError("radnet_nms_f64: a row hand-off did not complete within 2 s; results invalid")
    return pick[:int(cnt[0])].cpu().numpy().astype(np.int64), int(cnt[1])


def non_max_suppression_fast(boxes, probs, overlap_thresh=0.9, max_boxes=300):
    """Greedy NMS (reference rpn.py:380-456): returns (boxes[pick].astype(int), probs[pick]);
    `[]` for empty input (rpn.py:391-392); AssertionError unless x1<x2 and y1<y2 (rpn.py:400-401).

    IoU arithmetic is float64 in the reference's association order for any box dtype
    (integer boxes are widened exactly as rpn.py:405-406 does)."""
    if len(boxes) == 0:
        return []
    boxes = np.asarray(boxes)
    probs = np.asarray(probs)
    np.testing.assert_array_less(boxes[:, 0], boxes[:, 2])
    np.testing.assert_array_less(boxes[:, 1], boxes[:, 3])
    if boxes.dtype.kind == "i":
        boxes = boxes.astype("float")
    dev = _device()
    b_dev = D.to_device(boxes[:, :4], np.float64, dev)
    p_dev = D.to_device(probs.reshape(-1), np.float64, dev)
    pick, _ = _nms_device(b_dev, p_dev, None, overlap_thresh, max_boxes)
    return boxes[pick].astype("int"), probs[pick]


def nms_with_indices(boxes, probs, overlap_thresh=0.9, max_boxes=300):
    """Like non_max_suppression_fast but returns (pick indices int64, n_score_ties)."""
    boxes = np.asarray(boxes)
    if boxes.dtype.kind == "i":
        boxes = boxes.astype("float")
    dev = _device()
    b_dev = D.to_device(boxes[:, :4], np.float64, dev)
    p_dev = D.to_device(np.asarray(probs).reshape(-1), np.float64, dev)
    return _nms_device(b_dev, p_dev, None, overlap_thresh, max_boxes)


def apply_regr_np(X, T):
    """Box regression for one anchor shape (reference rpn.py:299-344): X (4,H,W) x,y,w,h and
    T (4,H,W) tx,ty,tw,th -> (4,H,W) float64, every plane rounded half-to-even.
    Like the reference it never raises: on any exception X is returned unchanged."""
    try:
        dev = _device()
        Xa = np.asarray(X)
        Ta = np.asarray(T)
        if Xa.shape[0] != 4 or Ta.shape != Xa.shape:
            raise ValueError("apply_regr_np: X and T must both be (4, ...) and equal in shape")
        n = int(np.prod(Xa.shape[1:]))
        x_dev = D.to_device(Xa.reshape(4, n), np.float64, dev)
        t_dev = D.to_device(Ta.reshape(4, n), np.float64, dev)     # float32 -> float64 is exact
        out = D.empty((4, n), np.float64, dev)
        _lib.call("radnet_apply_regr", D.ptr(x_dev), D.ptr(t_dev), n, D.ptr(out), D.stream_ptr(dev))
        return out.cpu().numpy().reshape(Xa.shape)
    except Exception as exc:   # same swallow-and-return as rpn.py:342-344
        print(exc)
        return X


def gt_feature_cells(img_data, C, class_mapping):
    """Figures of one image in feature cells, as calc_iou builds them (reference rpn.py:185-200):
    (gta (G,4) float64 x1,x2,y1,y2 - Python banker's rounding -, class index (G,) int32 with -1 for a
    class name that is not in class_mapping)."""
    bboxes = img_data['bboxes']
    width, height = img_data['width'], img_data['height']
    rw, rh = get_new_img_size(width, height, C.img_size)                      # rpn.py:189
    gta = np.zeros((len(bboxes), 4))
    gcls = np.zeros((len(bboxes),), dtype=np.int32)
    for k, bb in enumerate(bboxes):                                           # rpn.py:193-200
        gta[k, 0] = int(round(bb['x1'] * (rw / float(width)) / C.rpn_stride))
        gta[k, 1] = int(round(bb['x2'] * (rw / float(width)) / C.rpn_stride))
        gta[k, 2] = int(round(bb['y1'] * (rh / float(height)) / C.rpn_stride))
        gta[k, 3] = int(round(bb['y2'] * (rh / float(height)) / C.rpn_stride))
        # a class missing from class_mapping only matters if the figure is some RoI's best match
        # (the reference looks the name up at rpn.py:263, after the match)
        gcls[k] = class_mapping.get(bb['class'], -1)
    return gta, gcls


def _check_bg_last(class_mapping):
    n_cls = len(class_mapping)
    bg = class_mapping['bg']
    if bg != n_cls - 1:
        # the reference sizes the regression block as 4*(n_cls-1) and indexes it with 4*class_num
        # (rpn.py:266-273): with 'bg' anywhere else its rows become ragged lists - not a defined layout
        raise ValueError("calc_iou: class_mapping['bg'] must be the last index")
    return n_cls, bg


class RoiTargetBatch:
    """Batched classifier-head targets (calc_iou, reference rpn.py:209-282) for B panels in one launch,
    device-resident: the training path right after K2.

    `run` takes the figures of every panel in feature cells (`gt_feature_cells`) and the RoIs either as
    the detection records of `ProposalPipeline.sort_nms` (det=...) or as a dense (B,R,4) int32 xyxy
    tensor, and returns CUDA tensors compacted per panel in RoI order (rows >= count[b] are stale):
      x_roi (B,R,4) int32 xywh, y_class (B,R,n_cls) int32 one-hot, y_regr (B,R,8(n_cls-1)) float64
      [labels | coords], ious (B,R) float64, best_gt (B,R) int32 (-1 for 'bg' rows), count (B,) int32."""

    def __init__(self, C, class_mapping, batch, R, Gmax, device=None):
        D.require_cuda()
        _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.C = C
        self.n_cls, self.bg = _check_bg_last(class_mapping)
        self.B, self.R, self.Gmax = int(batch), int(R), int(Gmax)
        dev = self.device
        n_regr = 4 * (self.n_cls - 1)
        self.x_roi = D.empty((self.B, self.R, 4), np.int32, dev)
        self.y_class = D.empty((self.B, self.R, self.n_cls), np.int32, dev)
        self.y_regr = D.empty((self.B, self.R, 2 * n_regr), np.float64, dev)
        self.ious = D.empty((self.B, self.R), np.float64, dev)
        self.best_gt = D.empty((self.B, self.R), np.int32, dev)
        self.count = D.zeros((self.B,), np.int32, dev)
        self._std = D.host_f64(C.classifier_regr_std)

    def run(self, gt_cells, gt_class, gt_count, det=None, rois=None, roi_count=None):
        """gt_cells (B,Gmax,4) float64, gt_class (B,Gmax) int32, gt_count (B,) int32 or None: CUDA tensors.
        det: DetectionRecords (its max_boxes >= R) or rois (B,R,4) int32 xyxy [+ roi_count (B,) int32]."""
        C = self.C
        det_raw = det.raw if det is not None else None
        _lib.call("radnet_roi_targets_batch", D.ptr(det_raw), det.max_boxes if det is not None else 0,
                  D.ptr(rois), D.ptr(roi_count), self.B, self.R, D.ptr(gt_cells), D.ptr(gt_class),
                  D.ptr(gt_count), self.Gmax, self.n_cls, self.bg, float(C.classifier_min_overlap),
                  float(C.classifier_max_overlap), D.ptr(self._std), D.ptr(self.x_roi), D.ptr(self.y_class),
                  D.ptr(self.y_regr), D.ptr(self.ious), D.ptr(self.best_gt), D.ptr(self.count),
                  D.stream_ptr(self.device))
        return self.x_roi, self.y_class, self.y_regr, self.ious, self.best_gt, self.count


def calc_iou(R, img_data, C, class_mapping):
    """Classifier-head targets for proposals R (n,4) x1,y1,x2,y2 (reference rpn.py:176-296).

    Returns (X (1,n',4) int64 xywh, Y1 (1,n',n_cls) int64 one-hot, Y2 (1,n',8(n_cls-1))
    [labels | coords], IoUs list) or (None,)*4 when no RoI reaches classifier_min_overlap.
    KeyError when the best-matching figure of a positive RoI has a class outside class_mapping
    (rpn.py:263); ValueError unless 'bg' is the last class index."""
    bboxes = img_data['bboxes']
    n_cls, bg = _check_bg_last(class_mapping)
    gta, gcls = gt_feature_cells(img_data, C, class_mapping)
    R = np.asarray(R)
    n = int(R.shape[0])
    if n == 0:
        return None, None, None, None
    rois = np.rint(R[:, :4]).astype(np.int32)                                 # int(round()) rpn.py:211-214
    dev = _device()
    G = len(bboxes)
    r_dev = D.to_device(rois, np.int32, dev)
    g_dev = D.to_device(gta, np.float64, dev) if G else None
    c_dev = D.to_device(gcls, np.int32, dev) if G else None
    n_regr = 4 * (n_cls - 1)
    x_roi = D.empty((n, 4), np.int32, dev)
    y_cls = D.empty((n, n_cls), np.int32, dev)
    y_regr = D.empty((n, 2 * n_regr), np.float64, dev)
    ious = D.empty((n,), np.float64, dev)
    best_gt = D.empty((n,), np.int32, dev)
    count = D.zeros((1,), np.int32, dev)
    _lib.call("radnet_roi_targets", D.ptr(r_dev), n, D.ptr(g_dev), D.ptr(c_dev), G, n_cls, int(bg),
              float(C.classifier_min_overlap), float(C.classifier_max_overlap),
              D.ptr(D.host_f64(C.classifier_regr_std)), D.ptr(x_roi), D.ptr(y_cls), D.ptr(y_regr),
              D.ptr(ious), D.ptr(best_gt), D.ptr(count), D.stream_ptr(dev))
    m = int(count.cpu().numpy()[0])
    if m == 0:
        return None, None, None, None                                         # rpn.py:284-285
    if (gcls < 0).any():
        for g in best_gt[:m].cpu().numpy():
            if g >= 0 and gcls[g] < 0:
                raise KeyError(bboxes[int(g)]['class'])                       # class_mapping[cls_name], rpn.py:263
    X = x_roi[:m].cpu().numpy().astype(np.int64)
    Y1 = y_cls[:m].cpu().numpy().astype(np.int64)
    Y2 = y_regr[:m].cpu().numpy()
    if not (Y2[:, :n_regr] != 0).any():
        Y2 = Y2.astype(np.int64)      # the reference builds all-integer lists when nothing is positive
    return (np.expand_dims(X, axis=0), np.expand_dims(Y1, axis=0), np.expand_dims(Y2, axis=0),
            [float(v) for v in ious[:m].cpu().numpy()])
