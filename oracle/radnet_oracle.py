"""TEST INFRASTRUCTURE ONLY - CPU oracle for the RADNet proposal / RoI hot path.

A NumPy restatement, written from the behaviour of the reference, of the five
hot-path entry points listed in SURVEY.md section 8(a).  Each function cites the
reference file:line whose arithmetic it follows.  It is the checker for the CUDA
path and the `cpu_baseline` / `--impl reference` arm of `bench.py`; nothing in
`rock_art_radnet_b200/` imports it and the product never falls back to it.

Parity pinning
--------------
* a1-a4 (`rpn_to_roi`, `apply_regr_np`, `non_max_suppression_fast`,
  `calc_region_props`, `calc_iou`, `iou`, `get_new_img_size`): PINNED.  The
  reference ships no tests or golden vectors, so the
  restatement is pinned against the reference functions themselves, imported
  unmodified in the build container (`oracle/reference_import.py`):
  `tests/test_oracle_vs_reference.py` compares them on seeded inputs and
  `oracle/make_golden.py` stores reference outputs as `tests/golden/*.npz`.
* a5 (`roi_pooling_conv`): PARITY UNPINNED against TensorFlow itself.  The layer
  delegates to TF-1.x `tf.image.resize_images` (RoiPoolingConv.py:75), and
  TensorFlow/Keras cannot be installed here.  The restatement follows the
  published TF-1 legacy `ResizeBilinear` CPU kernel (align_corners=False, no
  half-pixel centres; tensorflow/core/kernels/resize_bilinear_op.cc and
  image_resizer_state.h, TF ~1.13-1.15, version unpinned by the reference) and is
  cross-checked against an independent float64 bilinear evaluation and against
  OpenCV-DNN's import of a hand-encoded TF `ResizeBilinear` GraphDef - an
  independent implementation of the TF op - to 2e-6 relative
  (`tests/test_oracle_roipool.py`).

Score ties: the reference orders candidates with `np.argsort` (default,
unstable, SIMD-implementation dependent; rpn.py:415), so the order among EQUAL
scores is not defined by the reference.  The oracle (and the device) define it
as the stable ascending argsort read from the end (higher flat index first) and
both count ties.
"""
import copy
import math

import numpy as np

# --------------------------------------------------------------------------
# helpers: image / feature-map sizes
# --------------------------------------------------------------------------


def get_new_img_size(width, height, img_min_side=300):
    """Resize rule, min side -> img_min_side (reference utils.py:65-75)."""
    if width <= height:
        scale = float(img_min_side) / width
        return img_min_side, int(scale * height)
    scale = float(img_min_side) / height
    return int(scale * width), img_min_side


def get_img_output_length(width, height):
    """ResNet-50 stride-16 map size (reference base_models/resnet50.py:19-35)."""

    def one(n):
        n += 6
        for k in (7, 3, 1, 1):
            n = (n - k + 2) // 2
        return n

    return one(width), one(height)


# --------------------------------------------------------------------------
# a2: greedy NMS
# --------------------------------------------------------------------------


def non_max_suppression_fast(boxes, probs, overlap_thresh=0.9, max_boxes=300,
                             return_pick=False):
    """Greedy IoU suppression (reference rpn.py:380-456).

    area has no +1 (:412); candidates are visited by descending score (:415,
    :423-425); a candidate is dropped when inter/(union+1e-6) > thresh, strictly
    (:443-447); the loop stops once `max_boxes` are picked (:449-450); boxes come
    back cast to int (:453).  `return_pick=True` additionally returns the picked
    row indices (test convenience, not in the reference signature).
    """
    if len(boxes) == 0:
        return []
    bx1, by1, bx2, by2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    np.testing.assert_array_less(bx1, bx2)
    np.testing.assert_array_less(by1, by2)
    if boxes.dtype.kind == "i":
        boxes = boxes.astype("float")
        bx1, by1, bx2, by2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (bx2 - bx1) * (by2 - by1)
    # ascending; ties resolved "stable", i.e. higher index is visited first
    remaining = np.argsort(probs, kind="stable")
    picked = []
    while remaining.size > 0:
        top = remaining[-1]
        rest = remaining[:-1]
        picked.append(top)
        iw = np.maximum(0, np.minimum(bx2[top], bx2[rest]) - np.maximum(bx1[top], bx1[rest]))
        ih = np.maximum(0, np.minimum(by2[top], by2[rest]) - np.maximum(by1[top], by1[rest]))
        inter = iw * ih
        union = areas[top] + areas[rest] - inter
        ratio = inter / (union + 1e-6)
        remaining = rest[~(ratio > overlap_thresh)]
        if len(picked) >= max_boxes:
            break
    out_boxes = boxes[picked].astype("int")
    out_probs = probs[picked]
    if return_pick:
        return out_boxes, out_probs, np.asarray(picked, dtype=np.int64)
    return out_boxes, out_probs


def count_score_ties(probs):
    """Number of candidates whose score equals another candidate's score."""
    if len(probs) == 0:
        return 0
    s = np.sort(np.asarray(probs).ravel())
    eq = s[1:] == s[:-1]
    dup = np.zeros(s.shape, dtype=bool)
    dup[1:] |= eq
    dup[:-1] |= eq
    return int(dup.sum())


# --------------------------------------------------------------------------
# a1: decode + clip + NMS
# --------------------------------------------------------------------------


def apply_regr_np(X, T):
    """Vectorised box regression for one anchor shape (reference rpn.py:299-344).

    X = (x, y, w, h) planes, T = (tx, ty, tw, th) planes.  Centre shift is
    t*size + centre (:325-328), size is exp(t)*size with the exponent widened to
    float64 (:330-331), then every output is rounded half-to-even (:335-338).
    Any exception returns X untouched (:342-344).
    """
    try:
        ax, ay, aw, ah = X[0], X[1], X[2], X[3]
        ncx = T[0] * aw + (ax + aw / 2.0)
        ncy = T[1] * ah + (ay + ah / 2.0)
        nw = np.exp(T[2].astype(np.float64)) * aw
        nh = np.exp(T[3].astype(np.float64)) * ah
        return np.stack([np.round(ncx - nw / 2.0), np.round(ncy - nh / 2.0),
                         np.round(nw), np.round(nh)])
    except Exception as exc:  # same swallow-and-return as the reference
        print(exc)
        return X


def decode_proposals(rpn_layer, regr_layer, C, use_regr=True):
    """Everything in rpn_to_roi up to the NMS call (reference rpn.py:91-166).

    Returns (all_boxes[N,4] float64 xyxy, all_probs[N], keep_mask[N]) in the
    reference's anchor-major flat order i = a*rows*cols + r*cols + c (:154-155);
    keep_mask is False for the rows the reference deletes (:163-166).
    """
    regr_layer = regr_layer / C.std_scaling                       # :91 (float32 divide)
    assert rpn_layer.shape[0] == 1                                # :96
    rows, cols = rpn_layer.shape[1:3]
    n_anchor = rpn_layer.shape[3]
    planes = np.zeros((4, rows, cols, n_anchor))                  # float64 (:106)
    gx, gy = np.meshgrid(np.arange(cols), np.arange(rows))        # :124
    a = 0
    for scale in C.anchor_box_scales:                             # :108
        for ratio in C.anchor_box_ratios:                         # :109
            aw = (scale * ratio[0]) / C.rpn_stride                # :112
            ah = (scale * ratio[1]) / C.rpn_stride                # :113
            t = np.transpose(regr_layer[0, :, :, 4 * a:4 * a + 4], (2, 0, 1))  # :117-118
            planes[0, :, :, a] = gx - aw / 2                      # :127
            planes[1, :, :, a] = gy - ah / 2                      # :128
            planes[2, :, :, a] = aw                               # :129
            planes[3, :, :, a] = ah                               # :130
            if use_regr:
                planes[:, :, :, a] = apply_regr_np(planes[:, :, :, a], t)  # :134
            planes[2, :, :, a] = np.maximum(1, planes[2, :, :, a])          # :137
            planes[3, :, :, a] = np.maximum(1, planes[3, :, :, a])          # :138
            planes[2, :, :, a] += planes[0, :, :, a]                        # :143
            planes[3, :, :, a] += planes[1, :, :, a]                        # :144
            planes[0, :, :, a] = np.maximum(0, planes[0, :, :, a])          # :147
            planes[1, :, :, a] = np.maximum(0, planes[1, :, :, a])          # :148
            planes[2, :, :, a] = np.minimum(cols - 1, planes[2, :, :, a])   # :149
            planes[3, :, :, a] = np.minimum(rows - 1, planes[3, :, :, a])   # :150
            a += 1
    all_boxes = planes.transpose((0, 3, 1, 2)).reshape(4, -1).T            # :154
    all_probs = rpn_layer.transpose((0, 3, 1, 2)).reshape(-1)              # :155
    drop = (all_boxes[:, 0] - all_boxes[:, 2] >= 0) | (all_boxes[:, 1] - all_boxes[:, 3] >= 0)  # :163
    return all_boxes, all_probs, ~drop


def rpn_to_roi(rpn_layer, regr_layer, C, use_regr=True, max_boxes=300, overlap_thresh=0.9,
               return_debug=False):
    """RPN maps -> kept proposals (reference rpn.py:68-172).

    `return_debug=True` returns a dict with the pre-NMS arrays and the picked
    flat indices (test convenience)."""
    all_boxes, all_probs, keep = decode_proposals(rpn_layer, regr_layer, C, use_regr)
    flat_index = np.nonzero(keep)[0]
    cand_boxes = all_boxes[keep]
    cand_probs = all_probs[keep]
    res = non_max_suppression_fast(cand_boxes, cand_probs, overlap_thresh=overlap_thresh,
                                   max_boxes=max_boxes, return_pick=True)
    boxes, probs, pick = res        # ValueError on empty input, like rpn.py:170
    if return_debug:
        return {"boxes": boxes, "probs": probs, "pick_flat": flat_index[pick],
                "all_boxes": all_boxes, "all_probs": all_probs, "keep_mask": keep}
    return boxes


# --------------------------------------------------------------------------
# a3: RPN target assignment ("calc_rpn")
# --------------------------------------------------------------------------


def iou(a, b):
    """IoU of two (x1,y1,x2,y2) boxes (reference utils.py:77-109).

    Degenerate boxes give 0.0 (:103-104); no +1 in the areas; the denominator
    carries +1e-6 (:109)."""
    if a[0] >= a[2] or a[1] >= a[3] or b[0] >= b[2] or b[1] >= b[3]:
        return 0.0
    ix = max(a[0], b[0])
    iy = max(a[1], b[1])
    iw = min(a[2], b[2]) - ix
    ih = min(a[3], b[3]) - iy
    inter = 0 if (iw < 0 or ih < 0) else iw * ih
    union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return float(inter) / float(union + 1e-6)


def rpn_targets_presample(C, img_data, width, height, width_resized, height_resized,
                          get_feat_map_size):
    """The deterministic part of calc_region_props (reference utils.py:585-766).

    Returns (valid[H,W,A], overlap[H,W,A], regr[H,W,4A], best_anchor[G,4],
    n_anchors_for_bbox[G]) BEFORE the RNG-driven 256-region subsampling.

    Reference semantics kept on purpose:
      * loop order size -> ratio -> ix -> jy -> gt (:616-648);
      * anchors crossing the image are skipped and stay invalid (:629,:638);
      * the per-GT best IoU lives in a float32 array (:603) and, under the
        NumPy >= 2 promotion rules the parity target runs with, the test
        `curr_iou > best[g]` is evaluated in float32; stated explicitly here so
        it does not depend on the NumPy version;
      * the 'neutral' branch assigns a misspelt variable (:720), so anchors with
        0.3 < IoU < 0.7 are negatives;
      * labels are written inside the GT loop (:722-738): no GT, no labels;
      * forced positives take the float32-rounded targets (:605,:766).
    """
    stride = float(C.rpn_stride)
    scales = C.anchor_box_scales
    ratios = C.anchor_box_ratios
    n_ratio = len(ratios)
    n_anchor = len(scales) * n_ratio
    fw, fh = get_feat_map_size(width_resized, height_resized)
    gts = img_data["bboxes"]
    n_gt = len(gts)

    overlap = np.zeros((fh, fw, n_anchor))
    valid = np.zeros((fh, fw, n_anchor))
    regr = np.zeros((fh, fw, 4 * n_anchor))
    hits = np.zeros(n_gt).astype(int)
    best_anchor = -1 * np.ones((n_gt, 4)).astype(int)
    best_iou32 = np.zeros(n_gt).astype(np.float32)
    best_t32 = np.zeros((n_gt, 4)).astype(np.float32)

    gt = np.zeros((n_gt, 4))          # columns x1, x2, y1, y2 (:608-613)
    for g, bb in enumerate(gts):
        gt[g, 0] = bb["x1"] * (width_resized / float(width))
        gt[g, 1] = bb["x2"] * (width_resized / float(width))
        gt[g, 2] = bb["y1"] * (height_resized / float(height))
        gt[g, 3] = bb["y2"] * (height_resized / float(height))

    for si in range(len(scales)):
        for ri in range(n_ratio):
            aw = scales[si] * ratios[ri][0]
            ah = scales[si] * ratios[ri][1]
            ch = ri + n_ratio * si
            for ix in range(fw):
                ax1 = stride * (ix + 0.5) - aw / 2
                ax2 = stride * (ix + 0.5) + aw / 2
                if ax1 < 0 or ax2 > width_resized:
                    continue
                for jy in range(fh):
                    ay1 = stride * (jy + 0.5) - ah / 2
                    ay2 = stride * (jy + 0.5) + ah / 2
                    if ay1 < 0 or ay2 > height_resized:
                        continue
                    is_pos = False
                    loc_best = 0.0
                    loc_t = None
                    for g in range(n_gt):
                        cur = iou([gt[g, 0], gt[g, 2], gt[g, 1], gt[g, 3]], [ax1, ay1, ax2, ay2])
                        beats_gt = bool(np.float32(cur) > best_iou32[g])   # float32 compare (a3')
                        if beats_gt or cur > C.rpn_max_overlap:
                            gcx = (gt[g, 0] + gt[g, 1]) / 2.0
                            gcy = (gt[g, 2] + gt[g, 3]) / 2.0
                            acx = (ax1 + ax2) / 2.0
                            acy = (ay1 + ay2) / 2.0
                            tx = (gcx - acx) / (ax2 - ax1)
                            ty = (gcy - acy) / (ay2 - ay1)
                            tw = np.log((gt[g, 1] - gt[g, 0]) / (ax2 - ax1))
                            th = np.log((gt[g, 3] - gt[g, 2]) / (ay2 - ay1))
                        if gts[g]["class"] != "bg":
                            if beats_gt:
                                best_anchor[g] = [jy, ix, ri, si]
                                best_iou32[g] = cur
                                best_t32[g, :] = [tx, ty, tw, th]
                            if cur > C.rpn_max_overlap:
                                is_pos = True
                                hits[g] += 1
                                if cur > loc_best:
                                    loc_best = cur
                                    loc_t = (tx, ty, tw, th)
                        # label write sits inside the GT loop in the reference
                        valid[jy, ix, ch] = 1
                        if is_pos:
                            overlap[jy, ix, ch] = 1
                            regr[jy, ix, 4 * ch:4 * ch + 4] = loc_t
                        else:
                            overlap[jy, ix, ch] = 0

    for g in range(n_gt):                                          # :741-766
        if hits[g] == 0:
            if best_anchor[g, 0] == -1:
                continue
            jy, ix = best_anchor[g, 0], best_anchor[g, 1]
            ch = best_anchor[g, 2] + n_ratio * best_anchor[g, 3]
            valid[jy, ix, ch] = 1
            overlap[jy, ix, ch] = 1
            regr[jy, ix, 4 * ch:4 * ch + 4] = best_t32[g, :]
    return valid, overlap, regr, best_anchor, hits


def subsample_regions(valid_cf, overlap_cf, max_n_regions=256):
    """RNG-driven balancing, in place on channel-first arrays (reference utils.py:777-813).

    valid_cf / overlap_cf have shape (1, A, H, W).  Uses the legacy global
    `np.random.choice` stream exactly as the reference does.  Returns n_pos."""
    pos = np.where(np.logical_and(overlap_cf[0] == 1, valid_cf[0] == 1))
    neg = np.where(np.logical_and(overlap_cf[0] == 0, valid_cf[0] == 1))
    n_pos = len(pos[0])
    n_neg = len(neg[0])
    if n_pos > max_n_regions / 2:
        # the reference builds this table from the NEGATIVE channel ids (:789)
        uniq, cnt = np.unique(neg[0], return_counts=True)
        frac = cnt / n_pos
        cnt_of = dict(zip(uniq, cnt))
        frac_of = dict(zip(uniq, frac))
        p = [frac_of[ch] / cnt_of[ch] for ch in pos[0]]
        off = np.random.choice(n_pos, n_pos - int(max_n_regions / 2), replace=False, p=p)
        valid_cf[0, pos[0][off], pos[1][off], pos[2][off]] = 0
        n_pos = int(max_n_regions / 2)
    if n_neg + n_pos > max_n_regions:
        uniq, cnt = np.unique(neg[0], return_counts=True)
        frac = cnt / n_neg
        cnt_of = dict(zip(uniq, cnt))
        frac_of = dict(zip(uniq, frac))
        p = [frac_of[ch] / cnt_of[ch] for ch in neg[0]]
        off = np.random.choice(n_neg, n_neg - n_pos, replace=False, p=p)
        valid_cf[0, neg[0][off], neg[1][off], neg[2][off]] = 0
    return n_pos


def calc_region_props(C, img_data, width, height, width_resized, height_resized,
                      get_feat_map_size, verbose=False):
    """RPN anchor targets (reference utils.py:554-821; upstream name calc_rpn)."""
    valid, overlap, regr, best_anchor, _ = rpn_targets_presample(
        C, img_data, width, height, width_resized, height_resized, get_feat_map_size)
    overlap_cf = np.expand_dims(np.transpose(overlap, (2, 0, 1)), axis=0)   # :768-769
    valid_cf = np.expand_dims(np.transpose(valid, (2, 0, 1)), axis=0)       # :771-772
    regr_cf = np.expand_dims(np.transpose(regr, (2, 0, 1)), axis=0)         # :774-775
    n_pos = subsample_regions(valid_cf, overlap_cf)
    y_rpn_cls = np.concatenate([valid_cf, overlap_cf], axis=1)              # :815
    y_rpn_regr = np.concatenate([np.repeat(overlap_cf, 4, axis=1), regr_cf], axis=1)  # :816
    return np.copy(y_rpn_cls), np.copy(y_rpn_regr), best_anchor, n_pos


calc_rpn = calc_region_props


# --------------------------------------------------------------------------
# a4: RoI target assignment
# --------------------------------------------------------------------------


def calc_iou(R, img_data, C, class_mapping):
    """Classifier-head targets for the kept proposals (reference rpn.py:176-296)."""
    gts = img_data["bboxes"]
    width, height = img_data["width"], img_data["height"]
    rw, rh = get_new_img_size(width, height, C.img_size)                    # :189
    gt = np.zeros((len(gts), 4))       # feature cells, columns x1, x2, y1, y2 (:197-200)
    for g, bb in enumerate(gts):
        gt[g, 0] = int(round(bb["x1"] * (rw / float(width)) / C.rpn_stride))
        gt[g, 1] = int(round(bb["x2"] * (rw / float(width)) / C.rpn_stride))
        gt[g, 2] = int(round(bb["y1"] * (rh / float(height)) / C.rpn_stride))
        gt[g, 3] = int(round(bb["y2"] * (rh / float(height)) / C.rpn_stride))
    n_cls = len(class_mapping)
    rois, onehots, coords_all, labels_all, ious = [], [], [], [], []
    for k in range(R.shape[0]):
        x1, y1, x2, y2 = (int(round(v)) for v in R[k, :])                   # :210-214
        best, best_g = 0.0, -1
        for g in range(len(gts)):                                           # :220-226
            cur = iou([gt[g, 0], gt[g, 2], gt[g, 1], gt[g, 3]], [x1, y1, x2, y2])
            if cur > best:
                best, best_g = cur, g
        if best < C.classifier_min_overlap:                                 # :228-230
            continue
        w, h = x2 - x1, y2 - y1
        rois.append([x1, y1, w, h])
        ious.append(best)
        if C.classifier_min_overlap <= best < C.classifier_max_overlap:     # :239-242
            name = "bg"
        elif C.classifier_max_overlap <= best:                              # :244-256
            name = gts[best_g]["class"]
            cxg = (gt[best_g, 0] + gt[best_g, 1]) / 2.0
            cyg = (gt[best_g, 2] + gt[best_g, 3]) / 2.0
            cx = x1 + w / 2.0
            cy = y1 + h / 2.0
            tx = (cxg - cx) / float(w)
            ty = (cyg - cy) / float(h)
            tw = np.log((gt[best_g, 1] - gt[best_g, 0]) / float(w))
            th = np.log((gt[best_g, 3] - gt[best_g, 2]) / float(h))
        else:                                                               # :258-261
            raise RuntimeError("roi = {}".format(best))
        cnum = class_mapping[name]
        onehot = n_cls * [0]
        onehot[cnum] = 1
        onehots.append(onehot)
        coords = [0] * 4 * (n_cls - 1)
        labels = [0] * 4 * (n_cls - 1)
        if name != "bg":                                                    # :270-277
            sx, sy, sw, sh = C.classifier_regr_std
            coords[4 * cnum:4 * cnum + 4] = [sx * tx, sy * ty, sw * tw, sh * th]
            labels[4 * cnum:4 * cnum + 4] = [1, 1, 1, 1]
        coords_all.append(copy.deepcopy(coords))
        labels_all.append(copy.deepcopy(labels))
    if not rois:                                                            # :284-285
        return None, None, None, None
    X = np.array(rois)
    Y1 = np.array(onehots)
    Y2 = np.concatenate([np.array(labels_all), np.array(coords_all)], axis=1)  # :294
    return np.expand_dims(X, 0), np.expand_dims(Y1, 0), np.expand_dims(Y2, 0), ious


# --------------------------------------------------------------------------
# a5: RoI pooling layer (crop + TF-1 legacy bilinear resize)
# --------------------------------------------------------------------------


def _legacy_resize_axis(in_size, out_size):
    """TF-1 ResizeBilinear interpolation table for one axis, float32 arithmetic.

    scale = in/out (float32 divide); in = i*scale; lower = floor(in),
    upper = min(ceil(in), in_size-1); lerp = in - floor(in)
    (image_resizer_state.h CalculateResizeScale / resize_bilinear_op.cc
    compute_interpolation_weights, align_corners=False, legacy scaler)."""
    scale = np.float32(in_size) / np.float32(out_size)
    src = np.arange(out_size, dtype=np.float32) * scale
    flo = np.floor(src)
    lower = np.maximum(flo.astype(np.int64), 0)
    upper = np.minimum(np.ceil(src).astype(np.int64), in_size - 1)
    lerp = (src - flo).astype(np.float32)
    return lower, upper, lerp


def roi_pooling_conv(img, rois, pool_size):
    """RoiPoolingConv.call (reference RoiPoolingConv.py:48-88).

    img (1,H,W,C) float32 NHWC; rois (1,num_rois,4) as (x,y,w,h), truncated to
    int32 (:69-72); crop img[:, y:y+h, x:x+w, :] with the slice end clamped to the
    map (:75, TF strided-slice semantics) and bilinear-resize it to
    (pool,pool); output (1,num_rois,pool,pool,C) float32 (:83).
    value = top + (bottom-top)*y_lerp, top = tl + (tr-tl)*x_lerp, all float32,
    no fused multiply-add."""
    img = np.asarray(img, dtype=np.float32)
    assert img.ndim == 4 and img.shape[0] == 1
    _, H, W, Cn = img.shape
    rois = np.asarray(rois)
    n = rois.shape[1]
    out = np.empty((1, n, pool_size, pool_size, Cn), dtype=np.float32)
    for k in range(n):
        x, y, w, h = (int(np.trunc(v)) for v in rois[0, k, :4])
        if x < 0 or y < 0:
            raise ValueError("negative RoI origin is outside the supported drop-in domain")
        ch = min(y + h, H) - min(y, H)
        cw = min(x + w, W) - min(x, W)
        if ch <= 0 or cw <= 0:
            raise ValueError("RoI %d crops to an empty region" % k)
        crop = img[0, y:y + ch, x:x + cw, :]
        y0, y1, ly = _legacy_resize_axis(ch, pool_size)
        x0, x1, lx = _legacy_resize_axis(cw, pool_size)
        lx_ = lx[None, :, None]
        ly_ = ly[:, None, None]
        tl = crop[y0][:, x0]
        tr = crop[y0][:, x1]
        bl = crop[y1][:, x0]
        br = crop[y1][:, x1]
        top = tl + (tr - tl) * lx_
        bot = bl + (br - bl) * lx_
        out[0, k] = top + (bot - top) * ly_
    return out


class RoiPoolingConv:
    """Callable stand-in with the reference layer's constructor and shapes
    (RoiPoolingConv.py:33-46, :90-94)."""

    def __init__(self, pool_size, num_rois, **kwargs):
        self.pool_size = pool_size
        self.num_rois = num_rois

    def compute_output_shape(self, input_shape):
        return None, self.num_rois, self.pool_size, self.pool_size, input_shape[0][3]

    def __call__(self, x, mask=None):
        assert len(x) == 2
        img, rois = x
        assert np.asarray(rois).shape[1] == self.num_rois
        return roi_pooling_conv(img, rois, self.pool_size)

    call = __call__

    def get_config(self):
        return {"pool_size": self.pool_size, "num_rois": self.num_rois}


def apply_regr(x, y, w, h, tx, ty, tw, th):
    """Scalar box regression of the classifier head (reference rpn.py:346-378)."""
    try:
        cx = x + w / 2.0
        cy = y + h / 2.0
        cx1 = tx * w + cx
        cy1 = ty * h + cy
        w1 = math.exp(tw) * w
        h1 = math.exp(th) * h
        x1 = cx1 - w1 / 2.0
        y1 = cy1 - h1 / 2.0
        return int(round(x1)), int(round(y1)), int(round(w1)), int(round(h1))
    except (ValueError, OverflowError):
        return x, y, w, h
    except Exception as exc:
        print(exc)
        return x, y, w, h
