"""TEST INFRASTRUCTURE ONLY - generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python oracle/make_golden.py

Inputs are not stored: every case is regenerated from its seed by
`rock_art_radnet_b200.synthetic`, so the fixtures hold only the reference's
outputs (small, compressed).  The reference's NMS returns boxes, not indices;
because the synthetic scores are unique, the picked flat indices are recovered
from the returned scores.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_import import load_reference  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# (name, seed, H, W, scales, realistic, use_regr, thr, max_boxes)
A1_CASES = [
    ("p600_s0", 0, 38, 38, (128, 256, 512), False, True, 0.7, 300),
    ("p600_s1", 1, 38, 38, (128, 256, 512), False, True, 0.7, 300),
    ("p600_real_s2", 2, 38, 38, (128, 256, 512), True, True, 0.7, 300),
    ("p600x800_s3", 3, 38, 50, (128, 256, 512), False, True, 0.7, 300),
    ("p600_12anch_s4", 4, 38, 38, (64, 128, 256, 512), False, True, 0.7, 300),
    ("p600_thr09_s5", 5, 38, 38, (128, 256, 512), False, True, 0.9, 300),
    ("small_s6", 6, 18, 25, (128, 256, 512), False, True, 0.5, 50),
    ("tiny_s7", 7, 5, 7, (128, 256, 512), False, True, 0.7, 300),
    ("noregr_s8", 8, 38, 38, (128, 256, 512), False, False, 0.7, 100),
    ("noregr_halfint_s9", 9, 20, 20, (100, 200), False, False, 0.6, 80),
]

# (name, seed, M, integer boxes?, thr, max_boxes)
A2_CASES = [
    ("float_thr02", 10, 300, False, 0.2, 300),
    ("float_thr04", 11, 700, False, 0.4, 300),
    ("float_thr07", 12, 2500, False, 0.7, 300),
    ("float_thr09_default", 13, 1500, False, 0.9, 300),
    ("int16x_thr02", 14, 280, True, 0.2, 300),
    ("int16x_thr04", 15, 900, True, 0.4, 300),
    ("halfint_thr07", 16, 1200, "half", 0.7, 300),
    ("one_box", 17, 1, False, 0.7, 300),
    ("maxboxes_5", 18, 400, False, 0.7, 5),
]

# (name, seed, width, height, n_gt, classes)
A3_CASES = [
    ("sq600_g20", 0, 600, 600, 20, ("boat",)),
    ("sq600_g20_mixed_bg", 1, 600, 600, 20, ("boat", "human", "bg")),
    ("w800_g7", 2, 800, 600, 7, ("boat", "animal")),
    ("h900_g1", 3, 600, 900, 1, ("wheel",)),
    ("sq600_g0", 4, 600, 600, 0, ("boat",)),
    ("sq600_allbg", 5, 600, 600, 6, ("bg",)),
    ("big_g30_manypos", 6, 1000, 700, 30, ("boat", "human")),
    ("small_gt", 7, 600, 600, 12, ("circle",)),
]


def nms_inputs(seed, M, kind):
    rng = np.random.default_rng(seed)
    x1 = rng.uniform(0, 500, M)
    y1 = rng.uniform(0, 500, M)
    b = np.stack([x1, y1, x1 + rng.uniform(1, 200, M), y1 + rng.uniform(1, 200, M)], axis=1)
    if kind is True:
        b = np.round(b / 16).astype(np.int64)
        b[:, 2] = np.maximum(b[:, 2], b[:, 0] + 1)
        b[:, 3] = np.maximum(b[:, 3], b[:, 1] + 1)
        b = b * 16
    elif kind == "half":
        b = np.round(b * 2) / 2
        b[:, 2] = np.maximum(b[:, 2], b[:, 0] + 0.5)
        b[:, 3] = np.maximum(b[:, 3], b[:, 1] + 0.5)
    p = ((rng.permutation(M) + 0.5) / M).astype(np.float32)
    return b, p


def a3_inputs(seed, width, height, n_gt, classes, small=False):
    lo, hi = (16, 90) if small else (48, 360)
    return S.gt_figures(seed, n_gt, width, height, classes=classes, lo=lo, hi=hi)


def main():
    rpn, utils, config = load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    manifest = {"numpy": np.__version__, "a1": [], "a2": [], "a3": [], "a4": []}

    # ---- a1: rpn_to_roi (pre-NMS arrays captured by wrapping the reference's NMS) ----
    out = {}
    for name, seed, H, W, scales, realistic, use_regr, thr, mb in A1_CASES:
        C = S.HotPathConfig(scales)
        A = C.num_anchors
        cls, regr = S.rpn_maps(seed, H, W, A, realistic)
        cap = {}
        orig = rpn.non_max_suppression_fast

        def spy(boxes, probs, **kw):
            cap["boxes"] = boxes.copy()
            cap["probs"] = probs.copy()
            res = orig(boxes, probs, **kw)
            cap["res"] = res
            return res

        rpn.non_max_suppression_fast = spy
        try:
            R = rpn.rpn_to_roi(cls, regr, C, use_regr=use_regr, max_boxes=mb, overlap_thresh=thr)
        finally:
            rpn.non_max_suppression_fast = orig
        kept_probs = cap["res"][1]
        flat_probs = cls.transpose((0, 3, 1, 2)).reshape(-1)
        lookup = {v: i for i, v in enumerate(flat_probs.tolist())}
        assert len(lookup) == flat_probs.size, "scores must be unique"
        pick_flat = np.array([lookup[v] for v in kept_probs.tolist()], dtype=np.int64)
        out[name + "/R"] = R
        out[name + "/pick_flat"] = pick_flat
        out[name + "/n_valid"] = np.int64(cap["boxes"].shape[0])
        out[name + "/pre_nms_checksum"] = np.float64(cap["boxes"].sum())
        # the full pre-NMS boxes only for the small cases (fixture size)
        if H * W * A <= 4100:
            out[name + "/pre_nms_boxes"] = cap["boxes"]
        manifest["a1"].append({"name": name, "seed": seed, "H": H, "W": W, "scales": list(scales),
                               "realistic": realistic, "use_regr": use_regr, "thr": thr, "max_boxes": mb})
    np.savez_compressed(os.path.join(GOLDEN, "a1_rpn_to_roi.npz"), **out)

    # ---- a2: non_max_suppression_fast ------------------------------------------------
    out = {}
    for name, seed, M, kind, thr, mb in A2_CASES:
        b, p = nms_inputs(seed, M, kind)
        boxes, probs = rpn.non_max_suppression_fast(b, p, overlap_thresh=thr, max_boxes=mb)
        lookup = {v: i for i, v in enumerate(p.tolist())}
        out[name + "/boxes"] = boxes
        out[name + "/probs"] = probs
        out[name + "/pick"] = np.array([lookup[v] for v in probs.tolist()], dtype=np.int64)
        manifest["a2"].append({"name": name, "seed": seed, "M": M, "kind": kind, "thr": thr, "max_boxes": mb})
    np.savez_compressed(os.path.join(GOLDEN, "a2_nms.npz"), **out)

    # ---- a3: calc_region_props ---------------------------------------------------------
    out = {}
    C = S.HotPathConfig()
    for name, seed, width, height, n_gt, classes in A3_CASES:
        img = a3_inputs(seed, width, height, n_gt, classes, small=(name == "small_gt"))
        wr, hr = utils.get_new_img_size(width, height, C.img_size)
        np.random.seed(seed)
        y_cls, y_regr, best, n_pos = utils.calc_region_props(C, img, width, height, wr, hr, S.resnet50_map_size)
        out[name + "/y_rpn_cls"] = y_cls.astype(np.int8)          # values are 0/1
        nz = np.flatnonzero(y_regr)
        out[name + "/y_rpn_regr_nz_idx"] = nz.astype(np.int64)
        out[name + "/y_rpn_regr_nz_val"] = y_regr.ravel()[nz]
        out[name + "/y_rpn_regr_shape"] = np.array(y_regr.shape, dtype=np.int64)
        out[name + "/best_anchor"] = best
        out[name + "/n_pos"] = np.int64(n_pos)
        manifest["a3"].append({"name": name, "seed": seed, "width": width, "height": height, "n_gt": n_gt,
                               "classes": list(classes), "resized": [wr, hr]})
    np.savez_compressed(os.path.join(GOLDEN, "a3_calc_region_props.npz"), **out)

    # ---- a4: calc_iou --------------------------------------------------------------------
    out = {}
    for seed in range(4):
        classes = ("boat", "human", "animal", "bg") if seed % 2 else ("boat",)
        img = S.gt_figures(seed, 20 if seed < 3 else 2, 600, 600, classes=classes)
        cls, regr = S.rpn_maps(seed)
        R = rpn.rpn_to_roi(cls, regr, C, max_boxes=300, overlap_thresh=0.7)
        X, Y1, Y2, ious = rpn.calc_iou(R, img, C, C.class_mapping)
        name = "s%d" % seed
        out[name + "/X"] = X
        out[name + "/Y1"] = Y1
        out[name + "/Y2"] = Y2
        out[name + "/ious"] = np.asarray(ious)
        manifest["a4"].append({"name": name, "seed": seed, "classes": list(classes), "n_gt": 20 if seed < 3 else 2})
    # an image whose figures overlap no proposal at all -> (None,)*4
    far = {"bboxes": [{"class": "boat", "x1": 0, "x2": 2, "y1": 0, "y2": 2}], "width": 600, "height": 600}
    cls, regr = S.rpn_maps(0)
    R = rpn.rpn_to_roi(cls, regr, C, max_boxes=20, overlap_thresh=0.7)
    res = rpn.calc_iou(R, far, C, C.class_mapping)
    out["none_case/is_none"] = np.int64(all(r is None for r in res))
    np.savez_compressed(os.path.join(GOLDEN, "a4_calc_iou.npz"), **out)

    with open(os.path.join(GOLDEN, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    for fn in sorted(os.listdir(GOLDEN)):
        print(fn, os.path.getsize(os.path.join(GOLDEN, fn)))


if __name__ == "__main__":
    main()
