"""GPU parity tests: the CUDA path (through the ctypes C-ABI and the drop-in functions)
against the CPU oracle and the committed reference golden vectors.

Bars (BASELINE.json north_star): kept-proposal indices and anchor labels bit-exact;
decoded boxes bit-exact here (integer valued; exp() near-ties are counted by the kernel);
regression targets within 1e-12 relative (device log() vs NumPy log(), <= 1-2 ulp);
pooled features bit-exact against the float32 restatement (tolerance allowed 1e-5 rel).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from conftest import dense_regr  # noqa: E402
from oracle import radnet_oracle as O  # noqa: E402
from oracle.make_golden import a3_inputs, nms_inputs  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402

pytestmark = pytest.mark.gpu

REGR_RTOL = 1e-12      # float64 targets: log() ulp differences only
POOL_RTOL = 1e-5       # north_star tolerance for pooled features (we expect bit-exact)


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    import rock_art_radnet_b200 as R
    from rock_art_radnet_b200 import _lib
    _lib.load()
    return R


# ----------------------------------------------------------------------------- a1 / K1+K2
def test_rpn_to_roi_matches_reference_golden(pkg, manifest, golden_a1):
    for case in manifest["a1"]:
        C = S.HotPathConfig(case["scales"])
        cls, regr = S.rpn_maps(case["seed"], case["H"], case["W"], C.num_anchors, case["realistic"])
        R = pkg.rpn_to_roi(cls, regr, C, use_regr=case["use_regr"], max_boxes=case["max_boxes"],
                           overlap_thresh=case["thr"])
        ref = golden_a1[case["name"] + "/R"]
        assert R.dtype == ref.dtype and R.flags.writeable
        assert np.array_equal(R, ref), case["name"]


def test_decode_boxes_and_kept_indices_bit_exact(pkg):
    from rock_art_radnet_b200.pipeline import ProposalPipeline
    C = S.HotPathConfig()
    B, H, W = 6, 38, 38
    maps = [S.rpn_maps(100 + s, H, W, 9, realistic=bool(s % 2)) for s in range(B)]
    cls = torch.from_numpy(np.concatenate([m[0] for m in maps])).cuda()
    regr = torch.from_numpy(np.concatenate([m[1] for m in maps])).cuda()
    pipe = ProposalPipeline(C, B, H, W, alloc_pooled=False)
    pipe.decode(cls, regr)
    pipe.sort_nms()
    stats = pipe.check_stats()
    boxes = pipe.boxes.cpu().numpy()
    keys = pipe.keys.cpu().numpy().view(np.uint32)
    dets = pipe.records.to_numpy()
    for b in range(B):
        dbg = O.rpn_to_roi(maps[b][0], maps[b][1], C, max_boxes=300, overlap_thresh=0.7, return_debug=True)
        keep = dbg["keep_mask"]
        assert stats[b, 0] == keep.sum() and stats[b, 1] == 0
        assert np.array_equal((keys[b] != 0), keep)
        assert np.array_equal(boxes[b][keep].astype(np.float64), dbg["all_boxes"][keep])   # decoded boxes
        assert np.array_equal(dets[b]["index"], dbg["pick_flat"])                         # kept indices
        assert np.array_equal(dets[b]["boxes"], dbg["boxes"])
        assert np.array_equal(dets[b]["scores"], dbg["probs"])
        assert dets[b]["n_candidates"] == keep.sum() and dets[b]["score_ties"] == 0
        assert 1700 <= dets[b]["n_sorted"] <= 3264            # sampled select: only about the top 2.4k keys are sorted


@pytest.mark.parametrize("H,W,scales,thr,mb", [
    (38, 50, (128, 256, 512), 0.7, 300),      # 600x800 panel
    (38, 38, (64, 128, 256, 512), 0.7, 300),  # reference default 12 anchors -> global-memory sort path
    (100, 100, (128, 256, 512), 0.7, 300),    # un-tiled 1600-px stress: 90,000 candidates in one segment
    (38, 38, (128, 256, 512), 0.3, 2000),     # kept list larger than the shared-memory list
    (38, 38, (128, 256, 512), 0.05, 300),     # heavy suppression: top slice exhausted -> full-sort round
    (38, 50, (64, 128, 256, 512), 0.1, 300),  # same on the global-memory sort path
    (7, 5, (128, 256, 512), 0.9, 300),
    (1, 1, (128,), 0.7, 300),
])
def test_rpn_to_roi_shapes_and_paths(pkg, H, W, scales, thr, mb):
    C = S.HotPathConfig(scales)
    cls, regr = S.rpn_maps(7, H, W, C.num_anchors)
    try:
        ref = O.rpn_to_roi(cls, regr, C, max_boxes=mb, overlap_thresh=thr)
    except ValueError:
        with pytest.raises(ValueError):
            pkg.rpn_to_roi(cls, regr, C, max_boxes=mb, overlap_thresh=thr)
        return
    got = pkg.rpn_to_roi(cls, regr, C, max_boxes=mb, overlap_thresh=thr)
    assert np.array_equal(got, ref)


def test_rpn_to_roi_score_ties_follow_documented_rule(pkg):
    C = S.HotPathConfig()
    cls, regr = S.rpn_maps(11)
    cls = (np.floor(cls * 40) / 40).astype(np.float32)          # ~40 distinct scores -> many ties
    ref = O.rpn_to_roi(cls, regr, C, max_boxes=300, overlap_thresh=0.7)
    got = pkg.rpn_to_roi(cls, regr, C, max_boxes=300, overlap_thresh=0.7)
    assert np.array_equal(got, ref)
    from rock_art_radnet_b200.rpn import _single_panel_pipeline
    pipe = _single_panel_pipeline(C, 38, 38, 300, 0.7)
    flat = cls.transpose((0, 3, 1, 2)).reshape(-1)
    dbg = O.decode_proposals(cls, regr, C)
    det = pipe.records.to_numpy()[0]
    # ties are counted among the n_sorted top-scoring candidates the kernel had to sort
    top = np.sort(flat[dbg[2]])[::-1][:det["n_sorted"]]
    assert det["n_sorted"] >= 1700 and det["score_ties"] == O.count_score_ties(top)


def test_rpn_to_roi_errors(pkg):
    C = S.HotPathConfig()
    cls, regr = S.rpn_maps(0, 4, 4, 9)
    with pytest.raises(AssertionError):
        pkg.rpn_to_roi(np.concatenate([cls, cls]), np.concatenate([regr, regr]), C)
    bad = regr.copy()
    bad[0, 1, 1, 2] = np.float32(4000.0)     # exp overflow -> inf - inf = NaN survives to the NMS assert
    with pytest.raises(AssertionError):
        O.rpn_to_roi(cls, bad, C)
    with pytest.raises(AssertionError):
        pkg.rpn_to_roi(cls, bad, C)


def test_apply_regr_np(pkg):
    rng = np.random.default_rng(3)
    X = np.stack([rng.uniform(-5, 40, (18, 25)), rng.uniform(-5, 40, (18, 25)),
                  rng.uniform(1, 30, (18, 25)), rng.uniform(1, 30, (18, 25))])
    T = (0.3 * rng.standard_normal((4, 18, 25))).astype(np.float32)
    got = pkg.apply_regr_np(X, T)
    ref = O.apply_regr_np(X, T)
    assert got.dtype == np.float64 and got.shape == ref.shape
    assert np.array_equal(got, ref)
    assert pkg.apply_regr_np(X, T[:3]) is X      # swallowed exception returns X (rpn.py:342-344)


# ----------------------------------------------------------------------------- a2 general NMS
def test_nms_matches_reference_golden(pkg, manifest, golden_a2):
    for case in manifest["a2"]:
        b, p = nms_inputs(case["seed"], case["M"], case["kind"])
        boxes, probs = pkg.non_max_suppression_fast(b, p, overlap_thresh=case["thr"], max_boxes=case["max_boxes"])
        n = case["name"]
        assert boxes.dtype == golden_a2[n + "/boxes"].dtype
        assert np.array_equal(boxes, golden_a2[n + "/boxes"]), n
        assert np.array_equal(probs, golden_a2[n + "/probs"]), n
        from rock_art_radnet_b200.rpn import nms_with_indices
        pick, ties = nms_with_indices(b, p, case["thr"], case["max_boxes"])
        assert np.array_equal(pick, golden_a2[n + "/pick"]) and ties == 0


def test_nms_edge_cases(pkg):
    assert pkg.non_max_suppression_fast(np.zeros((0, 4)), np.zeros((0,))) == []
    with pytest.raises(AssertionError):
        pkg.non_max_suppression_fast(np.array([[5.0, 0, 5.0, 4]]), np.array([0.5]))
    b, p = nms_inputs(3, 50, False)
    boxes, _ = pkg.non_max_suppression_fast(b, p, max_boxes=0)
    assert boxes.shape == (1, 4)
    # large M, float64 scores that do not fit float32, default threshold
    rng = np.random.default_rng(9)
    b, _ = nms_inputs(21, 20000, False)
    p = rng.random(20000)
    ref = O.non_max_suppression_fast(b, p, return_pick=True)
    got = pkg.non_max_suppression_fast(b, p)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    # near-threshold IoUs: identical translated boxes (IoU lattice), every threshold on the lattice
    base = np.array([[0.0, 0.0, 10.0, 10.0]])
    shifts = np.arange(0, 40)[:, None] * np.array([[1.0, 0.0, 1.0, 0.0]])
    bb = base + shifts
    pp = np.linspace(1.0, 0.1, 40)
    for thr in (9.0 / 11.0, 8.0 / 12.0, 0.5, 1.0 / 3.0):
        ref = O.non_max_suppression_fast(bb, pp, overlap_thresh=thr, return_pick=True)
        got = pkg.non_max_suppression_fast(bb, pp, overlap_thresh=thr)
        assert np.array_equal(got[0], ref[0]), thr


# ----------------------------------------------------------------------------- a3 / K3
def test_calc_region_props_matches_reference_golden(pkg, manifest, golden_a3):
    C = S.HotPathConfig()
    for case in manifest["a3"]:
        n = case["name"]
        img = a3_inputs(case["seed"], case["width"], case["height"], case["n_gt"], tuple(case["classes"]),
                        small=(n == "small_gt"))
        wr, hr = case["resized"]
        np.random.seed(case["seed"])
        y_cls, y_regr, best, n_pos = pkg.calc_region_props(C, img, case["width"], case["height"], wr, hr,
                                                           S.resnet50_map_size)
        assert y_cls.dtype == np.float64 and y_regr.dtype == np.float64 and best.dtype == np.int64
        assert y_cls.flags.writeable and y_regr.flags.writeable
        assert np.array_equal(y_cls, golden_a3[n + "/y_rpn_cls"].astype(np.float64)), n      # labels bit-exact
        ref_regr = dense_regr(golden_a3, n)
        assert np.array_equal(y_regr != 0, ref_regr != 0), n
        np.testing.assert_allclose(y_regr, ref_regr, rtol=REGR_RTOL, atol=0)
        assert np.array_equal(best, golden_a3[n + "/best_anchor"]), n
        assert int(n_pos) == int(golden_a3[n + "/n_pos"])
    assert pkg.calc_rpn is pkg.calc_region_props


@pytest.mark.parametrize("fill_bulk,two_launches", [(0, 0), (1, 0), (4096, 0), (0, 1), (1, 1)])
def test_rpn_targets_batched_presample_vs_oracle(pkg, lib_option, fill_bulk, two_launches):
    """fill_bulk > 0: the regression zeros are streamed by TMA bulk copies that every CTA of the launch shares out;
    two_launches: the fill and the panels as two launches (no co-residency assumed)."""
    from rock_art_radnet_b200.utils import rpn_targets_device
    lib_option("targets_fill_bulk", fill_bulk)
    lib_option("targets_two_launches", two_launches)
    C = S.HotPathConfig((64, 128, 256, 512))          # 12 anchors
    sizes = [(600, 600, 20), (800, 600, 9), (600, 750, 33), (600, 600, 0)]
    B, Gmax = len(sizes), 33
    gt = np.zeros((B, Gmax, 4)); bg = np.zeros((B, Gmax), np.uint8); cnt = np.zeros(B, np.int32)
    wh = np.zeros((B, 2)); imgs = []
    for b, (w, h, g) in enumerate(sizes):
        img = S.gt_figures(40 + b, g, w, h, classes=("boat", "bg", "human"))
        imgs.append(img)
        for k, bb in enumerate(img["bboxes"]):
            gt[b, k] = [bb["x1"], bb["x2"], bb["y1"], bb["y2"]]
            bg[b, k] = bb["class"] == "bg"
        cnt[b] = g
        wh[b] = [w, h]
    # batched call needs one map size: use the 600x600 map for every panel (anchors outside the
    # smaller image are skipped by the img_wh test, as in the reference)
    fw, fh = 50, 47
    y_cls, y_regr, best, hits = rpn_targets_device(C, gt, bg, cnt, fh, fw, wh)
    y_cls, y_regr, best, hits = (t.cpu().numpy() for t in (y_cls, y_regr, best, hits))
    A = 12
    for b, (w, h, g) in enumerate(sizes):
        valid, overlap, regr, ba, nh = O.rpn_targets_presample(C, imgs[b], w, h, w, h, lambda *_: (fw, fh))
        assert np.array_equal(y_cls[b, :A], valid.transpose(2, 0, 1))
        assert np.array_equal(y_cls[b, A:], overlap.transpose(2, 0, 1))
        assert np.array_equal(y_regr[b, :4 * A], np.repeat(overlap.transpose(2, 0, 1), 4, axis=0))
        np.testing.assert_allclose(y_regr[b, 4 * A:], regr.transpose(2, 0, 1), rtol=REGR_RTOL, atol=0)
        assert np.array_equal(best[b, :g], ba) and (best[b, g:] == -1).all()
        assert np.array_equal(hits[b, :g], nh)


def test_rpn_targets_hit_list_overflow_replay(pkg, lib_option):
    """With a 3-entry positive-cell list the kernel must fall back to the figure-by-figure replay."""
    lib_option("targets_hit_cap", 3)
    C = S.HotPathConfig()
    img = S.gt_figures(6, 30, 1000, 700, classes=("boat", "human"))
    wr, hr = O.get_new_img_size(1000, 700, C.img_size)
    np.random.seed(6)
    got = pkg.calc_region_props(C, img, 1000, 700, wr, hr, S.resnet50_map_size)
    np.random.seed(6)
    ref = O.calc_region_props(C, img, 1000, 700, wr, hr, S.resnet50_map_size)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[2], ref[2]) and got[3] == ref[3]
    np.testing.assert_allclose(got[1], ref[1], rtol=REGR_RTOL, atol=0)


def test_rpn_targets_many_seeds_best_anchor_and_hits(pkg):
    """Stresses the float32 pre-filter of the IoU kernel: tiny / huge / touching / duplicate figures."""
    from rock_art_radnet_b200.utils import rpn_targets_device
    C = S.HotPathConfig()
    rng = np.random.default_rng(77)
    imgs = []
    for s in range(10):
        lo, hi = [(8, 40), (48, 360), (200, 600), (16, 600)][s % 4]
        img = S.gt_figures(300 + s, 14, 600, 600, classes=("boat", "human"), lo=lo, hi=min(hi, 600))
        bbs = img["bboxes"]
        bbs.append(dict(bbs[0]))                                       # exact duplicate figure
        bbs.append({"class": "boat", "x1": 0, "x2": 600, "y1": 0, "y2": 600})   # whole image
        bbs.append({"class": "boat", "x1": 64, "x2": 192, "y1": 64, "y2": 192})  # coincides with an anchor
        bbs.append({"class": "boat", "x1": 300, "x2": 300, "y1": 10, "y2": 50})  # degenerate
        imgs.append(img)
    B, Gmax = len(imgs), max(len(i["bboxes"]) for i in imgs)
    gt = np.zeros((B, Gmax, 4)); bg = np.zeros((B, Gmax), np.uint8); cnt = np.zeros(B, np.int32)
    for b, img in enumerate(imgs):
        for k, bb in enumerate(img["bboxes"]):
            gt[b, k] = [bb["x1"], bb["x2"], bb["y1"], bb["y2"]]
        cnt[b] = len(img["bboxes"])
    wh = np.tile(np.array([[600.0, 600.0]]), (B, 1))
    y_cls, y_regr, best, hits = (t.cpu().numpy() for t in rpn_targets_device(C, gt, bg, cnt, 38, 38, wh))
    for b, img in enumerate(imgs):
        valid, overlap, regr, ba, nh = O.rpn_targets_presample(C, img, 600, 600, 600, 600, S.resnet50_map_size)
        g = len(img["bboxes"])
        assert np.array_equal(best[b, :g], ba), b
        assert np.array_equal(hits[b, :g], nh), b
        assert np.array_equal(y_cls[b, :9], valid.transpose(2, 0, 1)) and np.array_equal(y_cls[b, 9:], overlap.transpose(2, 0, 1))
        np.testing.assert_allclose(y_regr[b, 36:], regr.transpose(2, 0, 1), rtol=REGR_RTOL, atol=0)


# ----------------------------------------------------------------------------- a4
def test_calc_iou_matches_reference_golden(pkg, manifest, golden_a4):
    C = S.HotPathConfig()
    for case in manifest["a4"]:
        img = S.gt_figures(case["seed"], case["n_gt"], 600, 600, classes=tuple(case["classes"]))
        cls, regr = S.rpn_maps(case["seed"])
        R = pkg.rpn_to_roi(cls, regr, C, max_boxes=300, overlap_thresh=0.7)
        X, Y1, Y2, ious = pkg.calc_iou(R, img, C, C.class_mapping)
        n = case["name"]
        assert X.dtype == np.int64 and Y1.dtype == np.int64 and Y2.dtype == golden_a4[n + "/Y2"].dtype
        assert np.array_equal(X, golden_a4[n + "/X"]) and np.array_equal(Y1, golden_a4[n + "/Y1"])
        ref = golden_a4[n + "/Y2"]
        assert np.array_equal(Y2 != 0, ref != 0)
        np.testing.assert_allclose(Y2, ref, rtol=REGR_RTOL, atol=0)
        assert np.array_equal(np.asarray(ious), golden_a4[n + "/ious"]) and isinstance(ious, list)
    far = {"bboxes": [{"class": "boat", "x1": 0, "x2": 2, "y1": 0, "y2": 2}], "width": 600, "height": 600}
    cls, regr = S.rpn_maps(0)
    R = pkg.rpn_to_roi(cls, regr, C, max_boxes=20, overlap_thresh=0.7)
    assert pkg.calc_iou(R, far, C, C.class_mapping) == (None, None, None, None)
    # many RoIs (> one block pass) and no GT at all
    rng = np.random.default_rng(5)
    x1 = rng.integers(0, 30, 2500); y1 = rng.integers(0, 30, 2500)
    Rbig = np.stack([x1, y1, x1 + rng.integers(1, 8, 2500), y1 + rng.integers(1, 8, 2500)], 1)
    img = S.gt_figures(3, 15, 600, 600, classes=("boat", "wheel"))
    ref = O.calc_iou(Rbig, img, C, C.class_mapping)
    got = pkg.calc_iou(Rbig, img, C, C.class_mapping)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    np.testing.assert_allclose(got[2], ref[2], rtol=REGR_RTOL, atol=0)
    assert pkg.calc_iou(Rbig, {"bboxes": [], "width": 600, "height": 600}, C, C.class_mapping) == (None,) * 4


# ----------------------------------------------------------------------------- a5 / K4
def _all_size_rois(H, W):
    rois = []
    for w in range(1, W):
        h = 1 + (w * 7) % (H - 1)
        rois.append([(w * 3) % (W - w), (w * 5) % (H - h), w, h])
    for h in (1, 14, 28, H - 1):
        for w in (1, 14, 28, W - 1):
            rois.append([0, 0, w, h])
    return np.array(rois, dtype=np.int64)[None]


@pytest.mark.parametrize("H,W,Cn,pool", [(38, 38, 1024, 14), (38, 50, 1024, 14), (38, 38, 512, 7),
                                         (20, 16, 36, 3), (9, 11, 7, 5)])
def test_roi_pooling_conv_vs_oracle(pkg, H, W, Cn, pool):
    feat = S.feature_map(1, H, W, Cn)
    rois = _all_size_rois(H, W)
    layer = pkg.RoiPoolingConv(pool, rois.shape[1])
    got = layer([feat, rois])
    ref = O.roi_pooling_conv(feat, rois, pool)
    assert got.shape == ref.shape == (1, rois.shape[1], pool, pool, Cn) and got.dtype == np.float32
    assert layer.compute_output_shape([feat.shape, rois.shape]) == (None, rois.shape[1], pool, pool, Cn)
    assert layer.get_config() == {"pool_size": pool, "num_rois": rois.shape[1]}
    np.testing.assert_allclose(got, ref, rtol=POOL_RTOL, atol=1e-6)
    assert np.array_equal(got, ref), "expected bit-exact float32 (no FMA contraction)"


@pytest.mark.parametrize("H,W,Cn,pool", [(38, 38, 1024, 14), (38, 50, 1024, 14), (38, 38, 512, 7), (37, 41, 64, 14),
                                         (2, 3, 32, 2), (75, 75, 32, 7)])
@pytest.mark.parametrize("bands", [0, 2, 3, 5])
def test_roi_pooling_band_form_matches(pkg, lib_option, H, W, Cn, pool, bands):
    """The band form (a CTA holds a band of map rows and emits the output rows that sample it) on even / odd row
    counts and band counts that do not divide H (kept as a selectable form; the whole-map form measured faster on every shape)."""
    lib_option("roipool_form", 2)
    lib_option("roipool_bands", bands)
    feat = S.feature_map(4, H, W, Cn)
    rois = _all_size_rois(H, W)
    got = pkg.RoiPoolingConv(pool, rois.shape[1])([feat, rois])
    assert np.array_equal(got, O.roi_pooling_conv(feat, rois, pool))


def test_roi_pooling_band_form_batch_with_empty_slots(pkg, lib_option):
    """Batched call through the detection records (slots beyond the kept count must be zero-filled)."""
    from rock_art_radnet_b200.pipeline import ProposalPipeline
    C = S.HotPathConfig()
    B, H, W = 3, 38, 50
    maps = [S.rpn_maps(60 + s, H, W, 9) for s in range(B)]
    feat = torch.from_numpy(np.concatenate([S.feature_map(60 + s, H, W, 64) for s in range(B)])).cuda()
    outs = []
    for form in (1, 2):
        lib_option("roipool_form", form)
        pipe = ProposalPipeline(C, B, H, W, channels=64, pool_size=14, max_boxes=40, overlap_thresh=0.3)
        pipe.decode(torch.from_numpy(np.concatenate([m[0] for m in maps])).cuda(),
                    torch.from_numpy(np.concatenate([m[1] for m in maps])).cuda())
        pipe.sort_nms()
        pipe.pooled.fill_(7.0)
        outs.append(pipe.pool(feat).clone())
        counts = pipe.records.counts.cpu().numpy()
    assert torch.equal(outs[0], outs[1])
    for b in range(B):
        assert bool((outs[1][b, int(counts[b]):] == 0).all())


@pytest.mark.parametrize("cluster,every,ctas", [(0, 0, 0), (0, 1, 0), (0, 3, 0), (2, 1, 0), (2, 2, 0), (4, 2, 0), (8, 1, 0),
                                                (0, 0, 3), (0, 2, 5), (-1, 0, 0)])
def test_roi_pooling_lockstep_and_persistent_variants_match(pkg, lib_option, cluster, every, ctas):
    """Whole-map form: CTA / cluster barriers every few column rounds (the lockstep the launcher tunes per shape) and a
    grid smaller than the number of (panel, slice) work items never change a bit of the result."""
    from rock_art_radnet_b200.pipeline import ProposalPipeline
    lib_option("roipool_form", 1)
    lib_option("roipool_cluster", cluster)
    lib_option("roipool_sync_every", every)
    lib_option("roipool_ctas", ctas)
    C = S.HotPathConfig()
    B, H, W, Cn = 3, 38, 38, 256
    maps = [S.rpn_maps(40 + s, H, W, 9) for s in range(B)]
    feat_np = np.concatenate([S.feature_map(40 + s, H, W, Cn) for s in range(B)])
    pipe = ProposalPipeline(C, B, H, W, channels=Cn, pool_size=14, max_boxes=60, overlap_thresh=0.5)
    pipe.decode(torch.from_numpy(np.concatenate([m[0] for m in maps])).cuda(),
                torch.from_numpy(np.concatenate([m[1] for m in maps])).cuda())
    pipe.sort_nms()
    pipe.pooled.fill_(3.0)
    got = pipe.pool(torch.from_numpy(feat_np).cuda()).cpu().numpy()
    counts = pipe.records.counts.cpu().numpy()
    boxes = pipe.records.boxes.cpu().numpy()
    for b in range(B):
        n = int(counts[b])
        rois = np.zeros((1, 60, 4), np.float64)
        rois[0, :n, 0], rois[0, :n, 1] = boxes[b, :n, 0], boxes[b, :n, 1]
        rois[0, :n, 2], rois[0, :n, 3] = boxes[b, :n, 2] - boxes[b, :n, 0], boxes[b, :n, 3] - boxes[b, :n, 1]
        want = O.roi_pooling_conv(feat_np[b:b + 1], rois[:, :n], 14)
        assert np.array_equal(got[b, :n], want[0]), (b, cluster, every, ctas)
        assert (got[b, n:] == 0).all()


def test_roi_pooling_direct_kernel_matches(pkg, lib_option):
    lib_option("roipool_force_direct", 1)
    feat = S.feature_map(2, 38, 38, 256)
    rois = _all_size_rois(38, 38)
    got = pkg.RoiPoolingConv(14, rois.shape[1])([feat, rois])
    assert np.array_equal(got, O.roi_pooling_conv(feat, rois, 14))


def test_roi_pooling_clamp_float_and_errors(pkg):
    feat = S.feature_map(3, 12, 12, 64)
    rois = np.array([[[2.9, 3.2, 20.7, 4.0], [0, 0, 12, 12], [11, 11, 5, 5]]])   # float rois truncate, ends clamp
    got = pkg.RoiPoolingConv(7, 3)([feat, rois])
    assert np.array_equal(got, O.roi_pooling_conv(feat, rois, 7))
    with pytest.raises(ValueError):
        pkg.RoiPoolingConv(7, 1)([feat, np.array([[[12, 0, 3, 3]]])])
    with pytest.raises(ValueError):
        pkg.RoiPoolingConv(7, 1)([feat, np.array([[[-1, 0, 3, 3]]])])
    t = torch.from_numpy(feat).cuda()
    out = pkg.RoiPoolingConv(7, 3)([t, rois])
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), got)


# ----------------------------------------------------------------------------- full path
def test_pipeline_batch_end_to_end(pkg):
    from rock_art_radnet_b200.pipeline import ProposalPipeline
    C = S.HotPathConfig()
    B, H, W, Cn = 3, 38, 38, 1024
    maps = [S.rpn_maps(200 + s) for s in range(B)]
    feats = [S.feature_map(200 + s) for s in range(B)]
    # panel with < 300 candidates: push every anchor outside row 0 / cols < 20 off the map (degenerate)
    far = maps[2][1].copy()
    far[0, 1:, :, 0::4] = 4000.0
    far[0, 0, 20:, 0::4] = 4000.0
    maps[2] = (maps[2][0], far)
    pipe = ProposalPipeline(C, B, H, W, channels=Cn, pool_size=14, max_boxes=300, overlap_thresh=0.7)
    rec, pooled = pipe(torch.from_numpy(np.concatenate([m[0] for m in maps])).cuda(),
                       torch.from_numpy(np.concatenate([m[1] for m in maps])).cuda(),
                       torch.from_numpy(np.concatenate(feats)).cuda())
    pipe.check_stats()
    dets = rec.to_numpy()
    pooled = pooled.cpu().numpy()
    for b in range(B):
        R = O.rpn_to_roi(maps[b][0], maps[b][1], C, max_boxes=300, overlap_thresh=0.7)
        assert np.array_equal(dets[b]["boxes"], R)
        xywh = R.copy()
        xywh[:, 2] -= xywh[:, 0]
        xywh[:, 3] -= xywh[:, 1]                                     # RADNet.py:564-565
        ref = O.roi_pooling_conv(feats[b], xywh[None], 14)[0]
        k = R.shape[0]
        assert np.array_equal(pooled[b, :k], ref)
        assert not pooled[b, k:].any()                               # unused slots are zero-filled
    assert dets[2]["boxes"].shape[0] < 300


# ----------------------------------------------------------------------------- full-size properties
def _device_panels(n, seed):
    """n synthetic 600-px panels generated on the device: unique scores per panel, N(0,0.5) regression."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    N = 38 * 38 * 9
    perm = torch.stack([torch.randperm(N, device="cuda", generator=g) for _ in range(n)])
    cls = ((perm.float() + 0.5) / N).reshape(n, 38, 38, 9).contiguous()
    regr = (0.5 * torch.randn((n, 38, 38, 36), device="cuda", generator=g)).contiguous()
    return cls, regr


def test_full_size_batch_properties(pkg):
    """BASELINE configs[2] at full size (64 panels, 300 RoIs, 14x14x1024 pooling = 15.4 GB): the batch result
    is (1) identical run to run, (2) identical, panel by panel, to single-panel launches (which take the
    cluster form of sort+NMS) and to the reference's call pattern for the pooling (RoiPoolingConv on chunks of
    20 RoIs), (3) equal to the CPU oracle on sampled panels."""
    from rock_art_radnet_b200.pipeline import ProposalPipeline
    from rock_art_radnet_b200.RoiPoolingConv import RoiPoolingConv
    C = S.HotPathConfig()
    B, H, W, Cn = 64, 38, 38, 1024
    maps = [S.rpn_maps(300 + s, realistic=bool(s % 3 == 0)) for s in range(B)]
    cls = torch.from_numpy(np.concatenate([m[0] for m in maps])).cuda()
    regr = torch.from_numpy(np.concatenate([m[1] for m in maps])).cuda()
    feat = torch.randn((B, H, W, Cn), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    pipe = ProposalPipeline(C, B, H, W, channels=Cn, pool_size=14, max_boxes=300, overlap_thresh=0.7)
    rec, pooled = pipe(cls, regr, feat)
    pipe.check_stats()
    raw1 = rec.raw.clone()
    sums1 = pooled.view(B, -1).double().sum(dim=1).clone()
    rec, pooled = pipe(cls, regr, feat)
    assert torch.equal(rec.raw, raw1) and torch.equal(pooled.view(B, -1).double().sum(dim=1), sums1)      # (1)
    dets = rec.to_numpy()
    assert all(d["boxes"].shape[0] == 300 for d in dets)
    single = ProposalPipeline(C, 1, H, W, alloc_pooled=False)
    layer = RoiPoolingConv(14, 20)
    for b in (0, 17, 63):
        single.decode(cls[b:b + 1], regr[b:b + 1])
        single.sort_nms()
        assert torch.equal(single.records.raw[0], rec.raw[b])                                               # (2)
        R = dets[b]["boxes"].copy()
        R[:, 2] -= R[:, 0]
        R[:, 3] -= R[:, 1]
        rois = torch.from_numpy(R).cuda()
        for k in (0, 140, 280):
            out = layer([feat[b:b + 1], rois[None, k:k + 20]])
            assert torch.equal(out[0], pooled[b, k:k + 20])
    for b in (5, 40):                                                                                       # (3)
        want = O.rpn_to_roi(maps[b][0], maps[b][1], C, max_boxes=300, overlap_thresh=0.7)
        assert np.array_equal(dets[b]["boxes"], want)
        xywh = want.copy()
        xywh[:, 2] -= xywh[:, 0]
        xywh[:, 3] -= xywh[:, 1]
        ref = O.roi_pooling_conv(feat[b:b + 1].cpu().numpy(), xywh[None, :40], 14)[0]
        assert np.array_equal(pooled[b, :40].cpu().numpy(), ref)


def test_sweep_records_do_not_depend_on_batching(pkg):
    """BASELINE configs[4] in small: a sweep over 192 device-generated panels gives byte-identical detection
    records whether it is cut into batches of 64, 16, 8 or 4 panels - i.e. whichever form of sort+NMS a launch
    takes (one CTA per panel, clusters of 8, clusters of 16) - and whichever rank a panel lands on."""
    from rock_art_radnet_b200 import sharding
    from rock_art_radnet_b200.pipeline import ProposalPipeline
    C = S.HotPathConfig()
    n = 192
    cls, regr = _device_panels(n, 11)
    results = {}
    for B in (64, 16, 8, 4):
        pipe = ProposalPipeline(C, B, 38, 38, alloc_pooled=False)
        out = []
        for s in range(0, n, B):
            pipe.decode(cls[s:s + B], regr[s:s + B])
            pipe.sort_nms()
            out.append(pipe.records.raw.clone())
        pipe.check_stats()
        results[B] = torch.cat(out)
    for B in (16, 8, 4):
        assert torch.equal(results[B], results[64]), "batch size %d" % B
    assert int(results[64].view(torch.int32)[:, 0].min()) == 300
    # sharded over 3 ranks (simulated): each rank's panels, gathered and reordered, give the same bytes
    world, per = 3, 64
    parts = []
    for r in range(world):
        ids = torch.from_numpy(sharding.shard_indices(n, r, world)).cuda()
        pipe = ProposalPipeline(C, per, 38, 38, alloc_pooled=False)
        pipe.decode(cls[ids], regr[ids])
        pipe.sort_nms()
        parts.append(pipe.records.raw.clone())
    glob = sharding.gathered_to_global(torch.stack(parts), n)
    assert torch.equal(glob, results[64])


def test_pipelined_stream_equals_sequential_pipeline(pkg):
    """PipelinedProposalStream (decode+NMS of batch k+1 under the pool of batch k) returns, batch by batch,
    the bytes of the plain ProposalPipeline."""
    from rock_art_radnet_b200.pipeline import PipelinedProposalStream, ProposalPipeline
    C = S.HotPathConfig()
    B, Cn = 6, 256
    cls, regr = _device_panels(5 * B, 21)
    feat = torch.randn((5 * B, 38, 38, Cn), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    seq = ProposalPipeline(C, B, 38, 38, channels=Cn, pool_size=7)
    want = []
    for k in range(5):
        rec, pooled = seq(cls[k * B:(k + 1) * B], regr[k * B:(k + 1) * B], feat[k * B:(k + 1) * B])
        want.append((rec.raw.clone(), pooled.clone()))
    pipe = PipelinedProposalStream(C, B, 38, 38, channels=Cn, pool_size=7)
    pipe.begin()
    got = []
    for k in range(5):
        rec, pooled = pipe.submit(cls[k * B:(k + 1) * B], regr[k * B:(k + 1) * B], feat[k * B:(k + 1) * B])
        # the slot's records stay valid until two submissions later; the shared pooled buffer until the next pool
        with torch.cuda.stream(pipe.pool_stream):
            got.append((rec.raw.clone(), pooled.clone()))
    pipe.end()
    torch.cuda.synchronize()
    for (a, b), (c, d) in zip(got, want):
        assert torch.equal(a, c) and torch.equal(b, d)
