"""CPU: the oracle restatement against the committed reference outputs (tests/golden)."""
import numpy as np
import pytest

from conftest import dense_regr
from oracle import radnet_oracle as O
from oracle.make_golden import a3_inputs, nms_inputs
from rock_art_radnet_b200 import synthetic as S


def test_a1_rpn_to_roi_matches_reference_golden(manifest, golden_a1):
    for case in manifest["a1"]:
        C = S.HotPathConfig(case["scales"])
        cls, regr = S.rpn_maps(case["seed"], case["H"], case["W"], C.num_anchors, case["realistic"])
        dbg = O.rpn_to_roi(cls, regr, C, use_regr=case["use_regr"], max_boxes=case["max_boxes"],
                           overlap_thresh=case["thr"], return_debug=True)
        n = case["name"]
        assert dbg["boxes"].dtype == golden_a1[n + "/R"].dtype
        assert np.array_equal(dbg["boxes"], golden_a1[n + "/R"]), n
        assert np.array_equal(dbg["pick_flat"], golden_a1[n + "/pick_flat"]), n
        assert int(dbg["keep_mask"].sum()) == int(golden_a1[n + "/n_valid"]), n
        assert float(dbg["all_boxes"][dbg["keep_mask"]].sum()) == float(golden_a1[n + "/pre_nms_checksum"])
        if n + "/pre_nms_boxes" in golden_a1:
            assert np.array_equal(dbg["all_boxes"][dbg["keep_mask"]], golden_a1[n + "/pre_nms_boxes"])


def test_a2_nms_matches_reference_golden(manifest, golden_a2):
    for case in manifest["a2"]:
        b, p = nms_inputs(case["seed"], case["M"], case["kind"])
        boxes, probs, pick = O.non_max_suppression_fast(b, p, overlap_thresh=case["thr"],
                                                        max_boxes=case["max_boxes"], return_pick=True)
        n = case["name"]
        assert np.array_equal(boxes, golden_a2[n + "/boxes"]) and boxes.dtype == golden_a2[n + "/boxes"].dtype
        assert np.array_equal(probs, golden_a2[n + "/probs"])
        assert np.array_equal(pick, golden_a2[n + "/pick"])


def test_a2_edge_cases():
    assert O.non_max_suppression_fast(np.zeros((0, 4)), np.zeros((0,))) == []
    with pytest.raises(AssertionError):
        O.non_max_suppression_fast(np.array([[5.0, 0, 5.0, 4]]), np.array([0.5]))
    # max_boxes=0 still returns the top box (the stop test runs after the pick)
    b, p = nms_inputs(3, 50, False)
    boxes, _ = O.non_max_suppression_fast(b, p, max_boxes=0)
    assert boxes.shape == (1, 4)
    assert O.count_score_ties(np.array([0.1, 0.2, 0.2, 0.3, 0.3, 0.3])) == 5


def test_a3_calc_region_props_matches_reference_golden(manifest, golden_a3):
    C = S.HotPathConfig()
    for case in manifest["a3"]:
        n = case["name"]
        img = a3_inputs(case["seed"], case["width"], case["height"], case["n_gt"], tuple(case["classes"]),
                        small=(n == "small_gt"))
        wr, hr = O.get_new_img_size(case["width"], case["height"], C.img_size)
        assert [wr, hr] == case["resized"]
        np.random.seed(case["seed"])
        y_cls, y_regr, best, n_pos = O.calc_region_props(C, img, case["width"], case["height"], wr, hr,
                                                         S.resnet50_map_size)
        assert y_cls.dtype == np.float64 and y_regr.dtype == np.float64
        assert np.array_equal(y_cls, golden_a3[n + "/y_rpn_cls"].astype(np.float64)), n
        assert np.array_equal(y_regr, dense_regr(golden_a3, n)), n
        assert np.array_equal(best, golden_a3[n + "/best_anchor"]) and best.dtype == np.int64
        assert int(n_pos) == int(golden_a3[n + "/n_pos"])


def test_a4_calc_iou_matches_reference_golden(manifest, golden_a4):
    C = S.HotPathConfig()
    for case in manifest["a4"]:
        img = S.gt_figures(case["seed"], case["n_gt"], 600, 600, classes=tuple(case["classes"]))
        cls, regr = S.rpn_maps(case["seed"])
        R = O.rpn_to_roi(cls, regr, C, max_boxes=300, overlap_thresh=0.7)
        X, Y1, Y2, ious = O.calc_iou(R, img, C, C.class_mapping)
        n = case["name"]
        for got, key in ((X, "X"), (Y1, "Y1"), (Y2, "Y2")):
            assert got.dtype == golden_a4[n + "/" + key].dtype
            assert np.array_equal(got, golden_a4[n + "/" + key]), (n, key)
        assert np.array_equal(np.asarray(ious), golden_a4[n + "/ious"])
    far = {"bboxes": [{"class": "boat", "x1": 0, "x2": 2, "y1": 0, "y2": 2}], "width": 600, "height": 600}
    cls, regr = S.rpn_maps(0)
    R = O.rpn_to_roi(cls, regr, C, max_boxes=20, overlap_thresh=0.7)
    assert int(golden_a4["none_case/is_none"]) == 1
    assert O.calc_iou(R, far, C, C.class_mapping) == (None, None, None, None)


def test_map_size_and_resize_helpers():
    assert S.resnet50_map_size(600, 600) == (38, 38) == O.get_img_output_length(600, 600)
    assert S.resnet50_map_size(800, 600) == (50, 38)
    assert O.get_new_img_size(1000, 700, 600) == (857, 600)
    assert O.get_new_img_size(600, 900, 600) == (600, 900)
