"""CPU, build container only: the oracle restatement against the UNMODIFIED reference functions
imported from /root/reference (skipped on machines without the reference tree)."""
import numpy as np
import pytest

from oracle import radnet_oracle as O
from oracle.make_golden import nms_inputs
from oracle.reference_import import load_reference, reference_available
from rock_art_radnet_b200 import synthetic as S

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return load_reference()


def test_rpn_to_roi_and_nms(ref):
    rpn, _, _ = ref
    C = S.HotPathConfig()
    for seed, (H, W) in enumerate([(38, 38), (38, 50), (18, 25), (5, 7)]):
        cls, regr = S.rpn_maps(20 + seed, H, W, 9, realistic=bool(seed % 2))
        for thr, mb, use_regr in [(0.7, 300, True), (0.9, 300, True), (0.5, 40, True), (0.7, 100, False)]:
            a = rpn.rpn_to_roi(cls, regr, C, use_regr=use_regr, max_boxes=mb, overlap_thresh=thr)
            b = O.rpn_to_roi(cls, regr, C, use_regr=use_regr, max_boxes=mb, overlap_thresh=thr)
            assert a.dtype == b.dtype and np.array_equal(a, b)
    for seed in range(6):
        b, p = nms_inputs(50 + seed, 300 + 100 * seed, [False, True, "half"][seed % 3])
        for thr in (0.2, 0.4, 0.7):
            r = rpn.non_max_suppression_fast(b, p, overlap_thresh=thr)
            o = O.non_max_suppression_fast(b, p, overlap_thresh=thr)
            assert np.array_equal(r[0], o[0]) and np.array_equal(r[1], o[1])
    X = np.random.default_rng(1).uniform(1, 30, (4, 6, 7))
    T = np.random.default_rng(2).standard_normal((4, 6, 7)).astype(np.float32)
    assert np.array_equal(rpn.apply_regr_np(X, T), O.apply_regr_np(X, T))


def test_calc_region_props_and_calc_iou(ref):
    rpn, utils, _ = ref
    C = S.HotPathConfig((64, 128, 256, 512))
    for seed, (w, h, g) in enumerate([(600, 600, 12), (800, 600, 5), (600, 600, 0)]):
        img = S.gt_figures(30 + seed, g, w, h, classes=("boat", "bg", "human"))
        wr, hr = utils.get_new_img_size(w, h, C.img_size)
        np.random.seed(seed)
        a = utils.calc_region_props(C, img, w, h, wr, hr, S.resnet50_map_size)
        np.random.seed(seed)
        b = O.calc_region_props(C, img, w, h, wr, hr, S.resnet50_map_size)
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.asarray(y))
    C9 = S.HotPathConfig()
    img = S.gt_figures(3, 20, 600, 600, classes=("boat", "wheel", "bg"))
    cls, regr = S.rpn_maps(3)
    R = rpn.rpn_to_roi(cls, regr, C9, max_boxes=300, overlap_thresh=0.7)
    a = rpn.calc_iou(R, img, C9, C9.class_mapping)
    b = O.calc_iou(R, img, C9, C9.class_mapping)
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x), np.asarray(y)) and np.asarray(x).dtype == np.asarray(y).dtype
    for (x, y) in [([0, 0, 10, 10], [5, 5, 20, 20]), ([0, 0, 1, 1], [2, 2, 3, 3]), ([3, 3, 3, 9], [0, 0, 5, 5])]:
        assert utils.iou(x, y) == O.iou(x, y)
