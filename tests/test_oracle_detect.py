"""CPU: the detection post-processing oracle (oracle/detect_oracle.py)
against the golden vectors generated from the unmodified reference class
(oracle/make_golden_detect.py -> tests/golden/f1_classify.npz, f2_final_nms.npz, f2_predict.npz),
plus the host logic of the mirror class that needs no GPU."""
import numpy as np
import pytest

from conftest import _load_npz
from oracle import detect_oracle as DO
from oracle.make_golden_detect import F1_CASES, F2_CASES, dicts_to_arrays, f1_inputs, f2_inputs
from rock_art_radnet_b200 import synthetic as S
from rock_art_radnet_b200.RADNet import RADNet, tile_grid


@pytest.fixture(scope="module")
def g_f1():
    return _load_npz("f1_classify.npz")


@pytest.fixture(scope="module")
def g_f2():
    return _load_npz("f2_final_nms.npz")


@pytest.fixture(scope="module")
def g_pred():
    return _load_npz("f2_predict.npz")


def _config():
    C = S.HotPathConfig()
    C.tile_size, C.tile_overlap, C.include_full_img, C.max_n_tiles_train = 600, 200, False, 1
    return C


@pytest.mark.parametrize("name", sorted(F1_CASES))
def test_classify_decode_oracle_matches_reference_golden(g_f1, name):
    C = _config()
    R, model, F = f1_inputs(name, C)
    net = DO.RADNetOracle(C, None, model, None)
    bboxes, probs = net.apply_spatial_pyramid_pooling(R, F)
    order, cls, box, pr = dicts_to_arrays(bboxes, probs, C.class_mapping)
    assert np.array_equal(order, g_f1[name + "/order"])
    assert np.array_equal(cls, g_f1[name + "/cls"])
    assert np.array_equal(box, g_f1[name + "/box"])
    assert np.array_equal(pr, g_f1[name + "/prob"], equal_nan=True)
    assert len(cls) > 0


def test_pad_rois_rule():
    R = np.arange(4 * 47).reshape(47, 4)
    P = DO.pad_rois(R, 20)
    assert P.shape == (60, 4) and np.array_equal(P[:47], R)
    assert (P[47:] == R[40]).all()                     # copies of the last chunk's FIRST RoI (RADNet.py:112)
    assert DO.pad_rois(R[:40], 20).shape == (40, 4)


@pytest.mark.parametrize("name", sorted(F2_CASES))
def test_final_nms_oracle_matches_reference_golden(g_f2, name):
    b, p = f2_inputs(name)
    nb, npb, clusters = DO.final_nms(b, p, return_clusters=True)
    assert nb.dtype == g_f2[name + "/box"].dtype and np.array_equal(nb, g_f2[name + "/box"])
    assert npb.dtype == np.float32 and np.array_equal(npb, g_f2[name + "/prob"])
    # the float32 summation order the device reproduces
    for q in clusters:
        s = DO.pairwise_sum_f32(p[q])
        assert np.float32(np.float64(s) / np.float64(len(q))) == p[q].mean()


def test_pairwise_sum_restates_numpy():
    rng = np.random.default_rng(0)
    for n in list(range(1, 140)) + [255, 256, 257, 1000, 4096]:
        a = rng.random(n).astype(np.float32)
        assert DO.pairwise_sum_f32(a) == a.sum()


def test_final_nms_branches():
    # top below the confidence threshold: the 5 best members are averaged
    b = np.array([[0, 0, 100, 100]] * 7, dtype=np.int64) + np.arange(7)[:, None]
    p = np.array([0.71, 0.72, 0.73, 0.74, 0.75, 0.76, 0.77], dtype=np.float32)
    nb, npb = DO.final_nms(b, p)
    assert nb.shape == (1, 4) and np.array_equal(nb[0], np.rint(b[2:].mean(axis=0)).astype(int))
    assert npb[0] == p[2:].mean()
    # top above it: only members above the threshold
    p2 = p.copy()
    p2[6], p2[0] = 0.95, 0.9
    nb2, npb2 = DO.final_nms(b, p2)
    assert np.array_equal(nb2[0], np.rint(b[[0, 6]].mean(axis=0)).astype(int)) and npb2[0] == p2[[0, 6]].mean()
    assert DO.final_nms(np.zeros((0, 4)), np.zeros((0,))) == []
    with pytest.raises(AssertionError):
        DO.final_nms(np.array([[5, 5, 5, 9]]), np.array([0.9], dtype=np.float32))


def test_get_real_coordinates_floor_division():
    rng = np.random.default_rng(3)
    for ratio in (1.0, 0.75, 0.3, 600 / 799.0, 600 / 1000.0, 1.7, 2.0):
        for v in rng.integers(0, 3000, 50):
            got = DO.get_real_coordinates(ratio, v, v, v, v)[0]
            assert got == int(round(np.int64(v) // ratio))


def test_tile_grid():
    for (w, h, t, s) in [(1000, 800, 600, 200), (1600, 1600, 600, 200), (500, 400, 600, 200), (600, 600, 600, 200),
                         (2500, 2100, 2000, 400), (1601, 700, 600, 200)]:
        a = DO.tile_grid(w, h, t, s)
        assert a == tile_grid(w, h, t, s)
        assert all(0 <= x0 < x1 <= w and 0 <= y0 < y1 <= h for x0, y0, x1, y1 in a)
        assert len({tuple(v) for v in a}) == len(a)
        covered = np.zeros((h, w), dtype=bool)
        for x0, y0, x1, y1 in a:
            covered[y0:y1, x0:x1] = True
        assert covered.all()
    assert len(DO.tile_grid(1600, 1600, 600, 200)) == 36          # BASELINE configs[3]: 6 x 6 windows


@pytest.mark.parametrize("name", sorted(S.PREDICT_CASES))
def test_predict_oracle_matches_reference_golden(g_pred, name):
    pytest.importorskip("cv2")
    C, images, make_models = S.predict_case(name)
    m_rpn, m_det = make_models()
    fmt = RADNet(C, None, None, lambda x: x).format_img          # OpenCV resize, host glue shared with the mirror
    dets = DO.RADNetOracle(C, m_rpn, m_det, fmt).predict(images)
    cls = np.asarray([C.class_mapping[d['class']] for d in dets], dtype=np.int64)
    prob = np.asarray([d['prob'] for d in dets], dtype=np.float32)
    box = np.asarray([[d['x1'], d['y1'], d['x2'], d['y2']] for d in dets], dtype=np.int64).reshape(-1, 4)
    assert np.array_equal(cls, g_pred[name + "/cls"])
    assert np.array_equal(prob, g_pred[name + "/prob"])
    assert np.array_equal(box, g_pred[name + "/box"])
    assert [m_rpn.calls, m_det.calls] == list(g_pred[name + "/calls"])
    assert len(dets) > 0
