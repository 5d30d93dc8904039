"""GPU: seeded random-shape property tests of the whole drop-in surface against the oracle
(the role hypothesis plays in SURVEY.md 8(c); seeds are fixed so failures reproduce)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import radnet_oracle as O  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    import rock_art_radnet_b200 as R
    return R


def _cfg(rng):
    scales = sorted(rng.choice([32, 48, 64, 100, 128, 200, 256, 384, 512], size=int(rng.integers(1, 5)), replace=False).tolist())
    C = S.HotPathConfig(scales)
    if rng.random() < 0.3:
        C.anchor_box_ratios = [[1.0, 1.0], [1.0, 2.0], [2.0, 1.0], [1.0, 3.0], [3.0, 1.0]]   # config.py:58-64 variant
    C.std_scaling = float(rng.choice([4.0, 1.0, 2.5]))
    return C


@pytest.mark.parametrize("seed", range(12))
def test_rpn_to_roi_random(pkg, seed):
    rng = np.random.default_rng(1000 + seed)
    C = _cfg(rng)
    H, W = int(rng.integers(1, 60)), int(rng.integers(1, 60))
    cls, regr = S.rpn_maps(seed, H, W, C.num_anchors, realistic=bool(seed % 2))
    if seed % 3 == 0:
        cls = (np.round(cls * 50) / 50).astype(np.float32)          # heavy score ties
    regr = (regr * float(rng.choice([0.2, 1.0, 3.0]))).astype(np.float32)
    thr = float(rng.choice([0.05, 0.3, 0.5, 0.7, 0.9, 1.0]))
    mb = int(rng.choice([1, 7, 300, 5000]))
    use_regr = bool(rng.random() < 0.8)
    try:
        ref = O.rpn_to_roi(cls, regr, C, use_regr=use_regr, max_boxes=mb, overlap_thresh=thr)
    except ValueError:
        with pytest.raises(ValueError):
            pkg.rpn_to_roi(cls, regr, C, use_regr=use_regr, max_boxes=mb, overlap_thresh=thr)
        return
    got = pkg.rpn_to_roi(cls, regr, C, use_regr=use_regr, max_boxes=mb, overlap_thresh=thr)
    assert got.dtype == ref.dtype and np.array_equal(got, ref), (H, W, C.anchor_box_scales, thr, mb, use_regr)


@pytest.mark.parametrize("seed", range(8))
def test_nms_random(pkg, seed):
    rng = np.random.default_rng(2000 + seed)
    M = int(rng.integers(1, 6000))
    scale = float(rng.choice([1.0, 16.0, 1e-3, 1e4]))
    x1 = rng.uniform(0, 300, M); y1 = rng.uniform(0, 300, M)
    b = np.stack([x1, y1, x1 + rng.uniform(0.5, 150, M), y1 + rng.uniform(0.5, 150, M)], 1) * scale
    if seed % 2:
        b = np.floor(b) ; b[:, 2] = np.maximum(b[:, 2], b[:, 0] + 1); b[:, 3] = np.maximum(b[:, 3], b[:, 1] + 1)
        b = b.astype(np.int64)
    p = rng.random(M) if seed % 3 else rng.integers(0, 30, M) / 30.0           # float64 scores / many ties
    thr = float(rng.choice([0.2, 0.4, 0.7, 0.9]))
    mb = int(rng.choice([3, 300, 10000]))
    ref = O.non_max_suppression_fast(b, p, overlap_thresh=thr, max_boxes=mb)
    got = pkg.non_max_suppression_fast(b, p, overlap_thresh=thr, max_boxes=mb)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


@pytest.mark.parametrize("seed", range(8))
def test_roi_pool_random(pkg, seed):
    rng = np.random.default_rng(3000 + seed)
    H, W = int(rng.integers(2, 70)), int(rng.integers(2, 70))
    Cn = int(rng.choice([1, 3, 4, 12, 32, 96, 256, 1024]))
    pool = int(rng.choice([1, 2, 7, 14, 20]))
    n = int(rng.integers(1, 40))
    feat = rng.standard_normal((1, H, W, Cn), dtype=np.float32)
    x = rng.integers(0, W, n); y = rng.integers(0, H, n)
    rois = np.stack([x, y, rng.integers(1, W + 5, n), rng.integers(1, H + 5, n)], 1)[None]
    got = pkg.RoiPoolingConv(pool, n)([feat, rois])
    assert np.array_equal(got, O.roi_pooling_conv(feat, rois, pool)), (H, W, Cn, pool)


@pytest.mark.parametrize("seed", range(6))
def test_calc_region_props_and_calc_iou_random(pkg, seed):
    rng = np.random.default_rng(4000 + seed)
    C = _cfg(rng)
    C.std_scaling = 4.0
    w, h = int(rng.integers(300, 1400)), int(rng.integers(300, 1400))
    g = int(rng.integers(0, 26))
    img = S.gt_figures(seed, g, w, h, classes=("boat", "human", "bg", "wheel"), lo=12, hi=min(w, h) - 1)
    wr, hr = O.get_new_img_size(w, h, C.img_size)
    np.random.seed(seed)
    try:
        ref = O.calc_region_props(C, img, w, h, wr, hr, S.resnet50_map_size)
        err = None
    except (KeyError, ValueError) as e:          # the reference's own subsampling failures
        err = type(e)
    np.random.seed(seed)
    if err is not None:
        with pytest.raises(err):
            pkg.calc_region_props(C, img, w, h, wr, hr, S.resnet50_map_size)
    else:
        got = pkg.calc_region_props(C, img, w, h, wr, hr, S.resnet50_map_size)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[2], ref[2]) and got[3] == ref[3]
        assert np.array_equal(got[1] != 0, ref[1] != 0)
        np.testing.assert_allclose(got[1], ref[1], rtol=1e-12, atol=0)
    fw, fh = S.resnet50_map_size(wr, hr)
    R = np.stack([rng.integers(0, fw - 1, 200), rng.integers(0, fh - 1, 200)], 1)
    R = np.concatenate([R, R + rng.integers(1, 12, (200, 2))], 1)
    ref = O.calc_iou(R, img, C, C.class_mapping)
    got = pkg.calc_iou(R, img, C, C.class_mapping)
    if ref[0] is None:
        assert got == (None, None, None, None)
    else:
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2].dtype == ref[2].dtype
        np.testing.assert_allclose(got[2], ref[2], rtol=1e-12, atol=0)
        assert np.array_equal(np.asarray(got[3]), np.asarray(ref[3]))


@pytest.mark.parametrize("seed", range(24))
def test_decode_random_with_extremes(pkg, seed):
    """K1 alone against the oracle's decode: random map shapes, regression spreads from 0.5 to 20 (so the
    float32 shortcut, its float64 fallback near rounding boundaries and the large-magnitude path all run),
    NaN and inf deltas planted in every fourth case."""
    from rock_art_radnet_b200.pipeline import ProposalPipeline
    rng = np.random.default_rng(1000 + seed)
    C = S.HotPathConfig((64, 128, 256, 512) if seed % 3 == 0 else (128, 256, 512))
    A = C.num_anchors
    H, W = int(rng.integers(1, 70)), int(rng.integers(1, 70))
    cls = rng.random((1, H, W, A)).astype(np.float32)
    regr = ([0.5, 2.0, 6.0, 20.0][seed % 4] * rng.standard_normal((1, H, W, 4 * A))).astype(np.float32)
    if seed % 4 == 1:
        regr.flat[rng.integers(0, regr.size, 25)] = np.float32(np.nan)
        regr.flat[rng.integers(0, regr.size, 25)] = np.float32(np.inf)
        regr.flat[rng.integers(0, regr.size, 25)] = np.float32(-np.inf)
    pipe = ProposalPipeline(C, 1, H, W, alloc_pooled=False)
    pipe.decode(torch.from_numpy(cls).cuda(), torch.from_numpy(regr).cuda())
    boxes = pipe.boxes.cpu().numpy()[0]
    keys = pipe.keys.cpu().numpy()[0]
    stats = pipe.stats.cpu().numpy()[0]
    with np.errstate(all="ignore"):
        all_boxes, _, keep = O.decode_proposals(cls, regr, C)[:3]
    finite = np.isfinite(all_boxes).all(axis=1)
    ok = keep & finite
    assert np.array_equal(keys != 0, ok)
    assert np.array_equal(boxes[ok].astype(np.float64), all_boxes[ok])
    assert stats[0] == keep.sum() and stats[1] == (keep & ~finite).sum() and stats[2] == 0
