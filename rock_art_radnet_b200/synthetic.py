"""Seeded synthetic inputs of the shapes named in BASELINE.json / SURVEY.md section 8(d).

NumPy only.  Shared by the tests, the oracle-side golden generator and
`bench.py`, so that the CUDA path, the oracle and the CPU baseline all see the
same bytes for a given seed.
"""
import numpy as np


class HotPathConfig:
    """Duck-typed stand-in for the attributes of the reference `Config` that the hot
    path reads (reference faster_rcnn/config.py:47-108).  `anchor_box_scales`
    defaults to the BASELINE 9-anchor set `[128, 256, 512]` (config.py:46 comment);
    the reference default is the 12-anchor set `[64, 128, 256, 512]` (config.py:47)."""

    def __init__(self, anchor_box_scales=(128, 256, 512)):
        self.anchor_box_scales = list(anchor_box_scales)
        self.anchor_box_ratios = [[1.0, 1.0], [1.0, 2.0], [2.0, 1.0]]   # config.py:52-56
        self.img_size = 600                                             # config.py:70
        self.n_rois = 20                                                # config.py:77
        self.rpn_stride = 16                                            # config.py:81
        self.std_scaling = 4.0                                          # config.py:87
        self.classifier_regr_std = [8.0, 8.0, 4.0, 4.0]                 # config.py:88
        self.rpn_min_overlap = 0.3                                      # config.py:91
        self.rpn_max_overlap = 0.7                                      # config.py:92
        self.classifier_min_overlap = 0.1                               # config.py:95
        self.classifier_max_overlap = 0.5                               # config.py:96
        self.class_mapping = {'boat': 0, 'human': 1, 'other': 2, 'animal': 3,
                              'circle': 4, 'wheel': 5, 'bg': 6}         # config.py:100-108

    @property
    def num_anchors(self):
        return len(self.anchor_box_scales) * len(self.anchor_box_ratios)


def rpn_maps(seed, H=38, W=38, A=9, realistic=False):
    """Objectness and regression maps for one panel.

    Scores are UNIQUE float32 values (a permutation of (k+0.5)/N), so the kept
    order does not depend on the unstable argsort tie order of the reference;
    `realistic=True` bends them through sigmoid(3z-4)-like ranks (still unique).
    regr = 0.5 * N(0,1) float32, i.e. +-0.125 after the std_scaling divide."""
    rng = np.random.default_rng(seed)
    n = H * W * A
    u = (rng.permutation(n) + 0.5) / n
    if realistic:
        u = 1.0 / (1.0 + np.exp(-(6.0 * u - 5.0)))
    cls = u.astype(np.float32).reshape(1, H, W, A)
    regr = (0.5 * rng.standard_normal((1, H, W, 4 * A))).astype(np.float32)
    return cls, regr


def feature_map(seed, H=38, W=38, C=1024):
    """Backbone feature map (1,H,W,C) float32 NHWC, N(0,1)."""
    rng = np.random.default_rng(1_000_003 + seed)
    return rng.standard_normal((1, H, W, C), dtype=np.float32)


def gt_figures(seed, n_gt=20, width=600, height=600, classes=('boat',), lo=48, hi=360):
    """Synthetic ground-truth figures in original-image pixels."""
    rng = np.random.default_rng(2_000_003 + seed)
    boxes = []
    for k in range(n_gt):
        w = int(rng.integers(lo, min(hi, width) + 1))
        h = int(rng.integers(lo, min(hi, height) + 1))
        x1 = int(rng.integers(0, width - w + 1))
        y1 = int(rng.integers(0, height - h + 1))
        boxes.append({'class': classes[k % len(classes)], 'x1': x1, 'x2': x1 + w,
                      'y1': y1, 'y2': y1 + h})
    return {'bboxes': boxes, 'width': width, 'height': height}


def random_rois(seed, n, H=38, W=38):
    """RoIs (1,n,4) int64 as (x,y,w,h) in feature cells, 1<=w,h and inside the map."""
    rng = np.random.default_rng(3_000_003 + seed)
    x = rng.integers(0, W - 1, size=n)
    y = rng.integers(0, H - 1, size=n)
    w = 1 + (rng.integers(0, W, size=n) % (W - 1 - x + 0).clip(1))
    h = 1 + (rng.integers(0, H, size=n) % (H - 1 - y + 0).clip(1))
    w = np.minimum(w, W - x)
    h = np.minimum(h, H - y)
    return np.stack([x, y, w, h], axis=1)[None].astype(np.int64)


def resnet50_map_size(width, height):
    """Stride-16 ResNet-50 feature-map size (reference base_models/resnet50.py:19-35)."""

    def one(n):
        n += 6
        for k in (7, 3, 1, 1):
            n = (n - k + 2) // 2
        return n

    return one(width), one(height)


# ----------------------------------------------------------------------------------------
# Deterministic stand-ins for the two Keras models around the hot path (tests, goldens, bench).
# The reference calls `model_rpn.predict(X) -> [Y1, Y2, F]` and
# `model_detector.predict([F, ROIs]) -> [P_cls, P_regr]` (RADNet.py:560, 120); both networks are
# out of scope here, so these fakes produce maps / head outputs that are pure functions of
# their inputs and lead to overlapping detections across tiles (so that the tile merge has
# real clusters to average).
# ----------------------------------------------------------------------------------------
def coordinate_image(width, height):
    """uint8 BGR image whose pixel values encode their own coordinates, so that a fake RPN
    can recover the tile origin from a crop: ch0 = x % 256, ch1 = y % 256,
    ch2 = 16*(x // 256) + (y // 256)."""
    xs = np.arange(width)[None, :].repeat(height, 0)
    ys = np.arange(height)[:, None].repeat(width, 1)
    return np.stack([xs % 256, ys % 256, 16 * (xs // 256) + (ys // 256)], axis=-1).astype(np.uint8)


def scene_objects(seed, width, height, n_obj=12, n_cls=6, lo=60, hi=260):
    """Synthetic 'true' figures of a scene: (n_obj, 5) int64 [x1, y1, x2, y2, class]."""
    rng = np.random.default_rng(4_000_003 + seed)
    out = []
    for _ in range(n_obj):
        w = int(rng.integers(lo, min(hi, width) + 1))
        h = int(rng.integers(lo, min(hi, height) + 1))
        x1 = int(rng.integers(0, width - w + 1))
        y1 = int(rng.integers(0, height - h + 1))
        out.append([x1, y1, x1 + w, y1 + h, int(rng.integers(0, n_cls))])
    return np.asarray(out, dtype=np.int64)


class FakeRpnModel:
    """`predict(X)`: X (1,h,w,3) float32 RGB view of a `coordinate_image` crop at scale `ratio`
    -> [Y1 (1,H,W,A), Y2 (1,H,W,4A), F (1,H,W,4)].  The maps are `rpn_maps(seed)` with the seed
    derived from the crop origin; F[0,0,0,:3] carries (origin x, origin y, resized width) for
    the fake detector."""

    def __init__(self, num_anchors=9, map_size=resnet50_map_size, seed=0):
        self.num_anchors = num_anchors
        self.map_size = map_size
        self.seed = seed
        self.calls = 0

    def predict(self, X):
        self.calls += 1
        _, h, w, _ = X.shape
        Wm, Hm = self.map_size(w, h)
        r0, g0, b0 = (int(round(float(v))) for v in X[0, 0, 0, :3])      # RGB = ch2, ch1, ch0
        ox = b0 + 256 * (r0 // 16)
        oy = g0 + 256 * (r0 % 16)
        salt = 0
        if h > 1 and w > 1:     # differs between the image types of `predict_case` (odd pixels are flipped)
            salt = (int(round(float(X[0, 1, 1, 1]))) ^ ((oy + 1) % 256)) & 1
        cls, regr = rpn_maps(self.seed * 7919 + ox * 31 + oy + 104729 * salt, Hm, Wm, self.num_anchors)
        F = np.zeros((1, Hm, Wm, 4), dtype=np.float32)
        F[0, 0, 0, :3] = (ox, oy, w)
        return [cls, regr, F]


class FakeDetectorModel:
    """`predict([F, ROIs])`: ROIs (1,n,4) xywh in feature cells -> [P_cls (1,n,n_cls) float32,
    P_regr (1,n,4(n_cls-1)) float32].  Each RoI is scored by its IoU with the scene objects
    (in original-image pixels): the best object's class gets 0.40 + 0.58*iou (+ a hash jitter), 'bg' the
    rest; the regression targets point at the object, scaled by classifier_regr_std and
    perturbed by a hash of the RoI, so equal RoIs give equal outputs."""

    def __init__(self, objects, C, ratio=1.0, n_cls=7):
        self.objects = np.asarray(objects, dtype=np.float64)
        self.C = C
        self.ratio = float(ratio)
        self.n_cls = n_cls
        self.calls = 0

    def predict(self, inputs):
        self.calls += 1
        F, rois = inputs
        ox, oy = float(F[0, 0, 0, 0]), float(F[0, 0, 0, 1])
        n = rois.shape[1]
        P_cls = np.zeros((1, n, self.n_cls), dtype=np.float32)
        P_regr = np.zeros((1, n, 4 * (self.n_cls - 1)), dtype=np.float32)
        s = float(self.C.rpn_stride)
        std = self.C.classifier_regr_std
        ob = self.objects
        for i in range(n):
            x, y, w, h = (float(v) for v in rois[0, i])
            gx1, gy1 = ox + s * x / self.ratio, oy + s * y / self.ratio
            gx2, gy2 = gx1 + s * w / self.ratio, gy1 + s * h / self.ratio
            iw = np.maximum(0.0, np.minimum(gx2, ob[:, 2]) - np.maximum(gx1, ob[:, 0]))
            ih = np.maximum(0.0, np.minimum(gy2, ob[:, 3]) - np.maximum(gy1, ob[:, 1]))
            inter = iw * ih
            iou = inter / ((gx2 - gx1) * (gy2 - gy1) + (ob[:, 2] - ob[:, 0]) * (ob[:, 3] - ob[:, 1]) - inter)
            k = int(np.argmax(iou))
            hsh = (int(x) * 73856093 ^ int(y) * 19349663 ^ int(w) * 83492791 ^ int(h) * 2654435761) & 0xFFFF
            p = 0.40 + 0.58 * float(iou[k]) + 1e-4 * (hsh % 97)          # < 1, never saturates (no planted ties)
            c = int(ob[k, 4])
            P_cls[0, i, c] = p
            P_cls[0, i, self.n_cls - 1] = 1.0 - p
            # regression towards the object, in feature cells of this view
            fx1, fy1 = (ob[k, 0] - ox) * self.ratio / s, (ob[k, 1] - oy) * self.ratio / s
            fw, fh = (ob[k, 2] - ob[k, 0]) * self.ratio / s, (ob[k, 3] - ob[k, 1]) * self.ratio / s
            w_, h_ = max(w, 1.0), max(h, 1.0)
            noise = ((hsh % 13) - 6) * 0.01
            t = [((fx1 + fw / 2) - (x + w / 2)) / w_ + noise, ((fy1 + fh / 2) - (y + h / 2)) / h_ - noise,
                 np.log(max(fw, 0.25) / w_) * 0.8, np.log(max(fh, 0.25) / h_) * 0.8]
            P_regr[0, i, 4 * c:4 * c + 4] = [t[j] * std[j] for j in range(4)]
        return [P_cls, P_regr]


class RandomHeadModel:
    """`predict([F, ROIs])` with head outputs that are a hash-seeded random function of each RoI:
    softmax-like P_cls rows (so every class, 'bg' included, wins somewhere) and N(0,1) regression
    deltas.  `extremes=True` plants what the reference's apply_regr handles through its `except`
    branches (exp overflow, inf, NaN deltas) and, when `nan_cls` is set, NaN class scores."""

    def __init__(self, seed, C, n_cls=7, extremes=False, nan_cls=False, sharp=4.0, regr_scale=0.4):
        self.seed, self.C, self.n_cls = seed, C, n_cls
        self.extremes, self.nan_cls, self.sharp, self.regr_scale = extremes, nan_cls, sharp, regr_scale
        self.calls = 0

    def predict(self, inputs):
        self.calls += 1
        _, rois = inputs
        n = rois.shape[1]
        P_cls = np.zeros((1, n, self.n_cls), dtype=np.float32)
        P_regr = np.zeros((1, n, 4 * (self.n_cls - 1)), dtype=np.float32)
        std = np.tile(np.asarray(self.C.classifier_regr_std, dtype=np.float64), self.n_cls - 1)
        for i in range(n):
            x, y, w, h = (int(v) for v in rois[0, i])
            hsh = (x * 73856093 ^ y * 19349663 ^ w * 83492791 ^ h * 2654435761 ^ self.seed * 40503) & 0x7FFFFFFF
            rng = np.random.default_rng(hsh)
            z = rng.standard_normal(self.n_cls) * self.sharp
            e = np.exp(z - z.max())
            P_cls[0, i] = (e / e.sum()).astype(np.float32)
            t = rng.standard_normal(4 * (self.n_cls - 1)) * self.regr_scale
            if self.extremes:
                k = hsh % 11
                if k == 0:
                    t[2::4] = 800.0          # exp overflow -> OverflowError branch
                elif k == 1:
                    t[0::4] = np.inf         # inf centre -> OverflowError in round()
                elif k == 2:
                    t[3::4] = np.nan         # NaN -> ValueError in round()
                elif k == 3:
                    t[2::4] = -30.0          # width rounds to 0 -> degenerate box
                if self.nan_cls and hsh % 17 == 0:
                    P_cls[0, i, hsh % self.n_cls] = np.nan
            P_regr[0, i] = (t * std).astype(np.float32)
        return [P_cls, P_regr]


def clustered_boxes(seed, n_clusters, per_cluster, score_lo=0.7, score_hi=1.0, extent=1500, ties=False):
    """Integer boxes in `n_clusters` jittered groups with float32 scores - input of the tile merge
    (final_nms).  Returns (boxes (M,4) int64, probs (M,) float32)."""
    rng = np.random.default_rng(5_000_003 + seed)
    b, p = [], []
    for _ in range(n_clusters):
        cx, cy = (int(v) for v in rng.integers(0, extent, 2))
        w, h = (int(v) for v in rng.integers(80, 320, 2))
        for _ in range(per_cluster):
            j = rng.integers(-30, 31, 4)
            b.append([cx + j[0], cy + j[1], cx + w + j[2], cy + h + j[3]])
            p.append(rng.uniform(score_lo, score_hi))
    b = np.asarray(b, dtype=np.int64).reshape(-1, 4)
    p = np.asarray(p, dtype=np.float32)
    if ties and len(p) > 3:
        idx = rng.integers(0, len(p), len(p) // 3)
        p[idx] = p[idx[0]]
    return b, p


# name -> (seed, image width, height, tile_size, step, include_full_img, number of image types)
PREDICT_CASES = {
    "w1000_h800_tiles": (0, 1000, 800, 600, 200, False, 1),
    "w1000_h800_tiles_full": (1, 1000, 800, 600, 200, True, 1),
    "w700_h600_two_types": (2, 700, 600, 600, 200, False, 2),
    "w640_h600_fullonly": (3, 640, 600, 600, 200, True, 1),
}


def predict_case(name, config_cls=HotPathConfig):
    """(C, images, make_models) of one end-to-end post-processing case; `make_models()` returns a
    fresh (model_rpn, model_detector) pair so that the reference, the oracle and the device path
    all see the same call sequence."""
    seed, width, height, tile, step, full, n_types = PREDICT_CASES[name]
    C = config_cls()
    C.anchor_box_scales = [128, 256, 512]
    C.tile_size, C.tile_overlap = tile, step
    C.include_full_img = full
    C.max_n_tiles_train = 0 if name.endswith("fullonly") else 1
    objs = scene_objects(seed, width, height, n_obj=14)
    images = [coordinate_image(width, height) for _ in range(n_types)]
    if n_types > 1:
        images[1] = images[1].copy()
        images[1][1::2, 1::2, :] ^= 1          # a second image type: same geometry, different pixels

    def make_models():
        return FakeRpnModel(seed=seed), FakeDetectorModel(objs, C)

    return C, images, make_models


def tiled_panel_tiles(width, height, tile_size=600, step=200):
    """Tile rectangles [x0,y0,x1,y1] of one panel, same rule as the reference (RADNet.py:511-540)."""
    def axis(length):
        pairs = {(s, s + tile_size) for s in range(0, length, step) if s + tile_size <= length}
        pairs.add((max(0, length - tile_size), length))
        return sorted(pairs)
    return [[x0, y0, x1, y1] for (y0, y1) in axis(height) for (x0, x1) in axis(width)]


def tiled_panel_head_outputs(C, panel_seed, tile, rois_xywh, n_slots=300, n_obj=30):
    """Classifier-head outputs for the kept RoIs of one tile of a synthetic tiled panel: scores from
    the overlap with the panel's `scene_objects` (so neighbouring tiles detect the same figures and
    the tile merge has clusters), padded with zeros to n_slots rows.  Returns (P_cls (n_slots,7),
    P_regr (n_slots,24)) float32."""
    objs = scene_objects(panel_seed, 1600, 1600, n_obj=n_obj, lo=60, hi=300)
    model = FakeDetectorModel(objs, C)
    F = np.zeros((1, 1, 1, 4), dtype=np.float32)
    F[0, 0, 0, :2] = (tile[0], tile[1])
    a, r = model.predict([F, np.asarray(rois_xywh)[None]])
    n = a.shape[1]
    P_cls = np.zeros((n_slots, a.shape[2]), dtype=np.float32)
    P_regr = np.zeros((n_slots, r.shape[2]), dtype=np.float32)
    P_cls[:n], P_regr[:n] = a[0], r[0]
    return P_cls, P_regr


def one_hot_rows(seed, n_pos, n_neg, n_cls=7):
    """Y1 of calc_iou, (1, n_pos+n_neg, n_cls) int64: n_neg 'bg' rows (last class) and n_pos rows of random
    foreground classes, interleaved at random - the input of get_selected_samples (reference train.py:93)."""
    rng = np.random.default_rng(seed)
    n = n_pos + n_neg
    cls = np.full((n,), n_cls - 1, dtype=np.int64)
    where = rng.permutation(n)[:n_pos]
    cls[where] = rng.integers(0, n_cls - 1, n_pos)
    Y1 = np.zeros((1, n, n_cls), dtype=np.int64)
    Y1[0, np.arange(n), cls] = 1
    return Y1


def eval_set(seed, n_gt, n_det, classes=('boat', 'human', 'animal'), extent=1600, hit_rate=0.6, ties=False):
    """A synthetic test set for the mAP evaluation (reference test.py:48-173): `n_gt` figures and `n_det` detections as
    the lists of dicts get_objects takes.  About `hit_rate` of the detections are jittered copies of a figure (some
    figures get several, some the wrong class), the rest are clutter.  ties=True plants equal scores."""
    rng = np.random.default_rng(seed)
    gt, det = [], []
    for _ in range(n_gt):
        w, h = int(rng.integers(40, 300)), int(rng.integers(40, 300))
        x1, y1 = int(rng.integers(0, extent - w)), int(rng.integers(0, extent - h))
        gt.append({'class': classes[int(rng.integers(len(classes)))], 'x1': x1, 'y1': y1, 'x2': x1 + w, 'y2': y1 + h})
    for _ in range(n_det):
        if n_gt and rng.random() < hit_rate:
            g = gt[int(rng.integers(n_gt))]
            j = rng.integers(-25, 26, 4)
            x1, y1 = g['x1'] + int(j[0]), g['y1'] + int(j[1])
            x2, y2 = max(x1 + 1, g['x2'] + int(j[2])), max(y1 + 1, g['y2'] + int(j[3]))
            c = g['class'] if rng.random() < 0.85 else classes[int(rng.integers(len(classes)))]
        else:
            w, h = int(rng.integers(30, 300)), int(rng.integers(30, 300))
            x1, y1 = int(rng.integers(0, extent - w)), int(rng.integers(0, extent - h))
            x2, y2 = x1 + w, y1 + h
            c = classes[int(rng.integers(len(classes)))]
        p = float(np.float32(rng.uniform(0.2, 1.0)))
        if ties:
            p = round(p, 1)
        det.append({'class': c, 'x1': x1, 'y1': y1, 'x2': x2, 'y2': y2, 'prob': p})
    return det, gt
