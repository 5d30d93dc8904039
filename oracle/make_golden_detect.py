"""TEST INFRASTRUCTURE ONLY - golden vectors for the detection post-processing (SURVEY.md 8(f),
rows f1/f2) generated from the UNMODIFIED reference class `faster_rcnn.RADNet.RADNet`.

Run in the build container (needs /root/reference):

    python oracle/make_golden_detect.py

The two networks are replaced by the deterministic stand-ins of
`rock_art_radnet_b200.synthetic` (the reference only ever calls their `predict`), so every case
is regenerated from its name/seed by the tests and the fixtures hold only reference OUTPUTS.
"""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_import import load_radnet, load_reference  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (seed, number of RoIs, head model kind)
F1_CASES = {
    "scene_300": (0, 300, "scene"),
    "scene_290_padded": (1, 290, "scene"),
    "random_7_short": (2, 7, "random"),
    "random_300": (3, 300, "random"),
    "random_55_extremes": (4, 55, "extremes"),
    "random_120_nan_scores": (5, 120, "nan"),
}

# name -> (seed, clusters, boxes per cluster, score range, ties)
F2_CASES = {
    "mixed_12x9": (0, 12, 9, (0.7, 1.0), False),
    "all_low_8x14": (1, 8, 14, (0.70, 0.79), False),
    "big_clusters_3x150": (2, 3, 150, (0.75, 1.0), False),
    "singles_40x1": (3, 40, 1, (0.7, 1.0), False),
    "one_box": (4, 1, 1, (0.9, 0.95), False),
    "dense_1x300": (5, 1, 300, (0.81, 1.0), False),
}


def f1_inputs(name, C):
    """(R (n,4) int64 xywh, detector model, feature map) of one f1 case."""
    seed, n, kind = F1_CASES[name]
    R = S.random_rois(seed, n)[0]
    F = np.zeros((1, 38, 38, 4), dtype=np.float32)
    if kind == "scene":
        model = S.FakeDetectorModel(S.scene_objects(seed, 600, 600, n_obj=10), C)
    else:
        model = S.RandomHeadModel(seed, C, extremes=(kind in ("extremes", "nan")), nan_cls=(kind == "nan"))
    return R, model, F


def f2_inputs(name):
    seed, ncl, per, (lo, hi), ties = F2_CASES[name]
    return S.clustered_boxes(seed, ncl, per, lo, hi, ties=ties)


def dicts_to_arrays(bboxes, probs, class_to_idx):
    order = [class_to_idx[k] for k in bboxes]
    cls = np.concatenate([np.full((len(bboxes[k]),), class_to_idx[k], dtype=np.int64) for k in bboxes] or
                         [np.zeros((0,), np.int64)])
    box = np.concatenate([np.asarray(bboxes[k], dtype=np.int64).reshape(-1, 4) for k in bboxes] or
                         [np.zeros((0, 4), np.int64)])
    pr = np.concatenate([np.asarray(probs[k], dtype=np.float32).reshape(-1) for k in bboxes] or
                        [np.zeros((0,), np.float32)])
    return np.asarray(order, dtype=np.int64), cls, box, pr


def main():
    _, _, config = load_reference()
    RM = load_radnet()
    os.makedirs(GOLDEN, exist_ok=True)
    manifest = {"numpy": np.__version__, "f1": dict(F1_CASES), "f2": {k: list(v) for k, v in F2_CASES.items()},
                "predict": {k: list(v) for k, v in S.PREDICT_CASES.items()}}
    warnings.simplefilter("ignore")

    # ---- f1: apply_spatial_pyramid_pooling ---------------------------------------------------
    out = {}
    for name in F1_CASES:
        C = config.Config()
        C.anchor_box_scales = [128, 256, 512]
        R, model, F = f1_inputs(name, C)
        net = RM.RADNet(C, None, model, lambda x: x)
        bboxes, probs = net.apply_spatial_pyramid_pooling(R, F)
        order, cls, box, pr = dicts_to_arrays(bboxes, probs, C.class_mapping)
        out[name + "/order"], out[name + "/cls"], out[name + "/box"], out[name + "/prob"] = order, cls, box, pr
    np.savez_compressed(os.path.join(GOLDEN, "f1_classify.npz"), **out)

    # ---- f2: final_nms -----------------------------------------------------------------------
    out = {}
    C = config.Config()
    net = RM.RADNet(C, None, None, lambda x: x)
    for name in F2_CASES:
        b, p = f2_inputs(name)
        nb, npb = net.final_nms(b, p)
        out[name + "/box"] = np.asarray(nb, dtype=np.int64)
        out[name + "/prob"] = np.asarray(npb)
        assert out[name + "/prob"].dtype == np.float32
    np.savez_compressed(os.path.join(GOLDEN, "f2_final_nms.npz"), **out)

    # ---- predict -------------------------------------------------------------------------------
    out = {}
    for name in S.PREDICT_CASES:
        C, images, make_models = S.predict_case(name, config_cls=config.Config)
        m_rpn, m_det = make_models()
        net = RM.RADNet(C, m_rpn, m_det, lambda x: x)
        dets = net.predict(images)
        out[name + "/cls"] = np.asarray([C.class_mapping[d['class']] for d in dets], dtype=np.int64)
        out[name + "/prob"] = np.asarray([d['prob'] for d in dets], dtype=np.float32)
        out[name + "/box"] = np.asarray([[d['x1'], d['y1'], d['x2'], d['y2']] for d in dets], dtype=np.int64).reshape(-1, 4)
        out[name + "/calls"] = np.asarray([m_rpn.calls, m_det.calls], dtype=np.int64)
        for c in set(out[name + "/cls"].tolist()):     # equal scores would make the order implementation-defined
            pc = out[name + "/prob"][out[name + "/cls"] == c]
            assert len(np.unique(pc)) == len(pc), "score tie in golden case %s" % name
        print(name, "detections", len(dets), "rpn calls", m_rpn.calls, "detector calls", m_det.calls)
    np.savez_compressed(os.path.join(GOLDEN, "f2_predict.npz"), **out)

    with open(os.path.join(GOLDEN, "manifest_detect.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    for fn in sorted(os.listdir(GOLDEN)):
        print(fn, os.path.getsize(os.path.join(GOLDEN, fn)))


if __name__ == "__main__":
    main()
