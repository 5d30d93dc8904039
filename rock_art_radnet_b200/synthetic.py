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
