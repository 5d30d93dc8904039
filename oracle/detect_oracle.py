"""TEST INFRASTRUCTURE ONLY - CPU oracle for the detection post-processing that follows the
RoI pool.

NumPy / pure-Python restatement, written from the behaviour of the reference, of

  f1  RADNet.apply_spatial_pyramid_pooling: RoI padding to chunks of C.n_rois, the per-RoI
      class decision and box decode (reference faster_rcnn/RADNet.py:98-152), the scalar
      apply_regr (rpn.py:346-378), the per-class NMS at 0.2 and get_real_coordinates
      (RADNet.py:44-51, 570-600);
  f2  the tile grid (RADNet.py:511-540), RADNet.final_nms (cluster-and-average,
      RADNet.py:156-240), the cross-image NMS at 0.4 (RADNet.py:695-716) and the
      orchestration of RADNet.predict (RADNet.py:502-718) around two user models.

Parity pinning: PINNED.  The reference ships no tests for these functions, so the restatement
is pinned against the reference class itself, imported unmodified in the build container
(`oracle/reference_import.py: load_radnet`) and driven by deterministic fake models:
`oracle/make_golden.py` stores the reference's outputs in `tests/golden/f*.npz` and
`tests/test_oracle_vs_reference.py` compares live.

NumPy-version note (same situation as row a3' of the survey): the reference mixes NumPy
float32 scalars with Python floats (`tx /= std`, `np.max(P_cls) < 0.7`, `probs.max() < 0.8`).
Under NumPy >= 2 (NEP 50; the only version runnable here) those operations are carried out in
float32; that is the behaviour restated here and on the device.

Nothing in `rock_art_radnet_b200/` imports this file.
"""
import math

import numpy as np

from . import radnet_oracle as O

F32 = np.float32


# --------------------------------------------------------------------------------------
# tile grid
# --------------------------------------------------------------------------------------
def _axis_windows(length, tile_size, step):
    """Start/end pairs along one axis (reference RADNet.py:519-526): every `step` pixels while
    the window fits, plus one window flush with the far edge; duplicates removed, sorted."""
    starts = [s for s in range(0, length, step) if s + tile_size <= length]
    pairs = {(s, s + tile_size) for s in starts}
    pairs.add((max(0, length - tile_size), length))
    return sorted(pairs)


def tile_grid(img_width, img_height, tile_size, step):
    """Tiles [x0, y0, x1, y1], row-major over (y window, x window) (RADNet.py:511-540)."""
    return [[x0, y0, x1, y1] for (y0, y1) in _axis_windows(img_height, tile_size, step)
            for (x0, x1) in _axis_windows(img_width, tile_size, step)]


# --------------------------------------------------------------------------------------
# f1: per-RoI class decision + decode
# --------------------------------------------------------------------------------------
def pad_rois(R, n_rois):
    """RoIs as the detector sees them (RADNet.py:103-118): chunks of n_rois; a short last chunk
    is filled with copies of ITS first RoI.  (n,4) -> (ceil(n/n_rois)*n_rois, 4)."""
    R = np.asarray(R)
    n = R.shape[0]
    tail = n % n_rois
    if tail == 0:
        return R.copy()
    first_of_last = R[n - tail]
    return np.concatenate([R, np.repeat(first_of_last[None], n_rois - tail, axis=0)], axis=0)


def classify_decode(rois, P_cls, P_regr, C, bbox_threshold=0.7):
    """Per-RoI decision of RADNet.py:123-150.

    rois (m,4) int xywh in feature cells, P_cls (m,n_cls) float32, P_regr (m,4(n_cls-1)) float32.
    Returns (cls (m,) int64 with -1 for skipped RoIs, prob (m,) float32, box (m,4) int64
    [x1,y1,x2,y2] in resized-image pixels, n_fallback = RoIs whose regression raised inside
    apply_regr and kept the un-regressed box)."""
    P_cls = np.asarray(P_cls, dtype=F32)
    P_regr = np.asarray(P_regr, dtype=F32)
    m, n_cls = P_cls.shape
    cls = np.full((m,), -1, dtype=np.int64)
    prob = np.zeros((m,), dtype=F32)
    box = np.zeros((m, 4), dtype=np.int64)
    std = [F32(s) for s in C.classifier_regr_std]
    thr = F32(bbox_threshold)                                   # float32 comparison (NEP 50)
    n_fallback = 0
    for i in range(m):
        row = P_cls[i]
        best = int(np.argmax(row))
        top = row.max()
        if top < thr or best == n_cls - 1:                      # RADNet.py:126
            continue
        x, y, w, h = (int(v) for v in rois[i])
        t = [F32(P_regr[i, 4 * best + k]) / std[k] for k in range(4)]     # float32 divide, :139-143
        (x, y, w, h), fell_back = apply_regr_flagged(x, y, w, h, *t)
        n_fallback += int(fell_back)
        s = C.rpn_stride
        cls[i] = best
        prob[i] = top
        box[i] = (s * x, s * y, s * (x + w), s * (y + h))        # RADNet.py:149
    return cls, prob, box, n_fallback


def apply_regr_flagged(x, y, w, h, tx, ty, tw, th):
    """Scalar box regression (rpn.py:346-378) on integer x,y,w,h with float32 deltas; also says
    whether the reference's `except` path was taken (non-finite result: the box is returned
    unchanged, rpn.py:366-378)."""
    try:
        tx, ty = np.float64(tx), np.float64(ty)
        cx = x + w / 2.0
        cy = y + h / 2.0
        cx1 = tx * w + cx
        cy1 = ty * h + cy
        w1 = math.exp(tw) * w
        h1 = math.exp(th) * h
        x1 = cx1 - w1 / 2.0
        y1 = cy1 - h1 / 2.0
        return (int(round(x1)), int(round(y1)), int(round(w1)), int(round(h1))), False
    except (ValueError, OverflowError):
        return (x, y, w, h), True


def group_by_class(cls, prob, box, class_names):
    """The two dicts apply_spatial_pyramid_pooling returns (RADNet.py:130-150): insertion order
    = first appearance in RoI order; values are lists in RoI order."""
    bboxes, probs = {}, {}
    for i in range(len(cls)):
        if cls[i] < 0:
            continue
        name = class_names[int(cls[i])]
        bboxes.setdefault(name, []).append([int(v) for v in box[i]])
        probs.setdefault(name, []).append(prob[i])
    return bboxes, probs


def get_real_coordinates(ratio, x1, y1, x2, y2):
    """Resized-image pixels -> original pixels (RADNet.py:44-51): Python floor division by the
    resize ratio, then round (identity on an integral float)."""
    return tuple(int(round(np.float64(v) // ratio)) for v in (x1, y1, x2, y2))


def tile_detections(rois, P_cls, P_regr, C, ratio, tile_xy, class_names, bbox_threshold=0.7,
                    nms_thresh=0.2):
    """One tile of RADNet.predict (RADNet.py:566-600): decode, per-class NMS at 0.2, real
    coordinates, tile offset.  Returns an insertion-ordered dict name -> (boxes (k,4) int64,
    probs (k,) float32)."""
    cls, prob, box, _ = classify_decode(rois, P_cls, P_regr, C, bbox_threshold)
    bboxes, probs = group_by_class(cls, prob, box, class_names)
    out = {}
    for name in bboxes:
        nb, npb = O.non_max_suppression_fast(np.array(bboxes[name]), np.array(probs[name]),
                                             overlap_thresh=nms_thresh)
        real = np.array([get_real_coordinates(ratio, *nb[j]) for j in range(nb.shape[0])], dtype=np.int64)
        real = real.reshape(-1, 4)
        real[:, 0] += tile_xy[0]
        real[:, 1] += tile_xy[1]
        real[:, 2] += tile_xy[0]
        real[:, 3] += tile_xy[1]
        out[name] = (real, np.asarray(npb, dtype=F32))
    return out


# --------------------------------------------------------------------------------------
# f2: cluster-and-average merge of the tiles of one image
# --------------------------------------------------------------------------------------
def pairwise_sum_f32(a):
    """NumPy's float32 add-reduce order for a contiguous 1-D array (what `probs[p].mean()`
    sums with): blocks of <= 128 with eight running partial sums, halves above that."""
    n = len(a)
    if n < 8:
        r = F32(0.0)
        for v in a:
            r = F32(r + v)
        return r
    if n <= 128:
        acc = [F32(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                acc[j] = F32(acc[j] + a[i + j])
            i += 8
        r = F32(F32(F32(acc[0] + acc[1]) + F32(acc[2] + acc[3])) + F32(F32(acc[4] + acc[5]) + F32(acc[6] + acc[7])))
        while i < n:
            r = F32(r + a[i])
            i += 1
        return r
    half = n // 2
    half -= half % 8
    return F32(pairwise_sum_f32(a[:half]) + pairwise_sum_f32(a[half:]))


def final_nms(boxes, probs, obj_avg_threshold=0.2, obj_confidence_threshold=0.8, n_obj_avg=5,
              return_clusters=False):
    """Cluster-and-average NMS (RADNet.py:156-240).

    Visit boxes by descending score; the top box and every remaining box with
    inter/(union+1e-6) > obj_avg_threshold form a cluster and leave the pool.  The cluster is
    represented by the members scoring above obj_confidence_threshold, or - when even its top
    score is below it - by its n_obj_avg best members: box = rint(mean box), prob = float32
    mean.  Score ties are visited "higher index first" (stable ascending order read from the
    end), as everywhere in this oracle."""
    if len(boxes) == 0:
        return []
    boxes = np.asarray(boxes)
    probs = np.asarray(probs)
    np.testing.assert_array_less(boxes[:, 0], boxes[:, 2])
    np.testing.assert_array_less(boxes[:, 1], boxes[:, 3])
    fb = boxes.astype("float") if boxes.dtype.kind == "i" else boxes
    x1, y1, x2, y2 = fb[:, 0], fb[:, 1], fb[:, 2], fb[:, 3]
    area = (x2 - x1) * (y2 - y1)
    order = np.argsort(probs, kind="stable")
    thr_c = probs.dtype.type(obj_confidence_threshold) if probs.dtype.kind == "f" else obj_confidence_threshold
    clusters = []
    while order.size > 0:
        top = order[-1]
        rest = order[:-1]
        iw = np.maximum(0, np.minimum(x2[top], x2[rest]) - np.maximum(x1[top], x1[rest]))
        ih = np.maximum(0, np.minimum(y2[top], y2[rest]) - np.maximum(y1[top], y1[rest]))
        inter = iw * ih
        ratio = inter / (area[top] + area[rest] - inter + 1e-6)
        member_pos = np.concatenate((np.where(ratio > obj_avg_threshold)[0], [order.size - 1]))
        members = order[member_pos]                               # ascending score, top last
        if probs[members].max() < thr_c:
            rep = members[-n_obj_avg:]
        else:
            rep = members[probs[members] > thr_c]
        clusters.append(rep)
        order = np.delete(order, member_pos)
    out_b = np.array([np.rint(fb[p].mean(axis=0)).astype("int") for p in clusters])
    out_p = np.array([probs[p].mean() for p in clusters])
    if return_clusters:
        return out_b, out_p, clusters
    return out_b, out_p


# --------------------------------------------------------------------------------------
# RADNet.predict around two user models
# --------------------------------------------------------------------------------------
class RADNetOracle:
    """Restatement of the post-processing skeleton of the reference class `RADNet`
    (RADNet.py:22-718).  `model_rpn.predict(X) -> [Y1, Y2, F]`,
    `model_detector.predict([F, ROIs]) -> [P_cls, P_regr]`; `format_img(img) -> (X, ratio)` is
    injected (the reference resizes with OpenCV, which is outside the hot path)."""

    def __init__(self, C, model_rpn, model_detector, format_img):
        self.C = C
        self.model_rpn = model_rpn
        self.model_detector = model_detector
        self.format_img = format_img
        self.bbox_threshold = 0.7
        self.class_mapping = {v: k for k, v in C.class_mapping.items()}

    def apply_spatial_pyramid_pooling(self, R, feature_map):
        rois = pad_rois(R, self.C.n_rois)
        pc, pr = [], []
        for k in range(0, rois.shape[0], self.C.n_rois):
            P_cls, P_regr = self.model_detector.predict([feature_map, rois[None, k:k + self.C.n_rois]])
            pc.append(P_cls[0])
            pr.append(P_regr[0])
        if not pc:
            return {}, {}
        cls, prob, box, _ = classify_decode(rois, np.concatenate(pc), np.concatenate(pr), self.C,
                                            self.bbox_threshold)
        return group_by_class(cls, prob, box, self.class_mapping)

    def _one_view(self, img, origin, bbox_total, probs_total):
        X, ratio = self.format_img(img)
        Y1, Y2, F = self.model_rpn.predict(X)
        R = O.rpn_to_roi(Y1, Y2, self.C, overlap_thresh=0.7)
        R[:, 2] -= R[:, 0]
        R[:, 3] -= R[:, 1]
        bboxes, probs = self.apply_spatial_pyramid_pooling(R, F)
        for name in bboxes:
            nb, npb = O.non_max_suppression_fast(np.array(bboxes[name]), np.array(probs[name]), overlap_thresh=0.2)
            for j in range(nb.shape[0]):
                rx1, ry1, rx2, ry2 = get_real_coordinates(ratio, *nb[j])
                bbox_total.setdefault(name, []).append([origin[0] + rx1, origin[1] + ry1,
                                                        origin[0] + rx2, origin[1] + ry2])
                probs_total.setdefault(name, []).append(npb[j])

    def predict(self, images):
        C = self.C
        all_bbox, all_probs = {}, {}
        for img in images:
            bbox_total, probs_total = {}, {}
            if C.max_n_tiles_train > 0:
                for tile in tile_grid(img.shape[1], img.shape[0], C.tile_size, C.tile_overlap):
                    self._one_view(img[tile[1]:tile[3], tile[0]:tile[2], :], (tile[0], tile[1]),
                                   bbox_total, probs_total)
            if C.include_full_img:
                self._one_view(img, (0, 0), bbox_total, probs_total)
            for name in bbox_total:
                nb, npb = final_nms(np.array(bbox_total[name]), np.array(probs_total[name]))
                for j in range(nb.shape[0]):
                    all_bbox.setdefault(name, []).append([int(v) for v in nb[j]])
                    all_probs.setdefault(name, []).append(npb[j])
        dets = []
        for name in all_bbox:
            nb, npb = O.non_max_suppression_fast(np.array(all_bbox[name]), np.array(all_probs[name]),
                                                 overlap_thresh=0.4)
            for j in range(nb.shape[0]):
                dets.append({'class': name, 'prob': npb[j], 'x1': nb[j, 0], 'y1': nb[j, 1],
                             'x2': nb[j, 2], 'y2': nb[j, 3]})
        return dets
