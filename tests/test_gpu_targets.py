"""GPU parity of the training-side target kernels in their batched, device-resident forms:
K3 `radnet_rpn_targets` (one launch: both output layouts, the self-cleaning workspace) and the batched
a4 `radnet_roi_targets_batch`, against the CPU oracle (reference utils.py:554-775, rpn.py:176-296)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import radnet_oracle as O  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402

pytestmark = pytest.mark.gpu

REGR_RTOL = 1e-12      # float64 targets: log() ulp differences only


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    import rock_art_radnet_b200 as R
    from rock_art_radnet_b200 import _lib
    _lib.load()
    return R


def _gt_batch(imgs, Gmax=None):
    B = len(imgs)
    Gmax = Gmax or max(1, max(len(i["bboxes"]) for i in imgs))
    gt = np.zeros((B, Gmax, 4)); bg = np.zeros((B, Gmax), np.uint8); cnt = np.zeros(B, np.int32)
    for b, img in enumerate(imgs):
        for k, bb in enumerate(img["bboxes"]):
            gt[b, k] = [bb["x1"], bb["x2"], bb["y1"], bb["y2"]]
            bg[b, k] = bb["class"] == "bg"
        cnt[b] = len(img["bboxes"])
    return gt, bg, cnt


def _oracle_nhwc(C, img, scale):
    """The two tensors the reference generator yields BEFORE subsampling (utils.py:768-775, 815-816,
    475-478): NHWC, regr half times std_scaling."""
    valid, overlap, regr, ba, nh = O.rpn_targets_presample(C, img, 600, 600, 600, 600, S.resnet50_map_size)
    y_cls = np.concatenate([valid, overlap], axis=2)
    y_regr = np.concatenate([np.repeat(overlap, 4, axis=2), regr * scale], axis=2)
    return y_cls, y_regr, ba, nh


def test_rpn_targets_nhwc_scaled_layout_vs_oracle(pkg):
    from rock_art_radnet_b200.utils import LAYOUT_NHWC, rpn_targets_device
    C = S.HotPathConfig()
    imgs = [S.gt_figures(500 + s, g, 600, 600, classes=("boat", "bg", "human")) for s, g in enumerate((20, 1, 0, 33, 7))]
    gt, bg, cnt = _gt_batch(imgs)
    wh = np.tile(np.array([[600.0, 600.0]]), (len(imgs), 1))
    y_cls, y_regr, best, hits = (t.cpu().numpy() for t in rpn_targets_device(
        C, gt, bg, cnt, 38, 38, wh, layout=LAYOUT_NHWC, regr_scale=C.std_scaling))
    assert y_cls.shape == (5, 38, 38, 18) and y_regr.shape == (5, 38, 38, 72)
    for b, img in enumerate(imgs):
        rc, rr, ba, nh = _oracle_nhwc(C, img, C.std_scaling)
        g = len(img["bboxes"])
        assert np.array_equal(y_cls[b], rc), b
        assert np.array_equal(y_regr[b] != 0, rr != 0), b
        np.testing.assert_allclose(y_regr[b], rr, rtol=REGR_RTOL, atol=0)
        assert np.array_equal(best[b, :g], ba) and np.array_equal(hits[b, :g], nh)


def test_rpn_targets_workspace_is_left_clean_between_launches(pkg):
    """The single-launch kernel accumulates per-panel state in its workspace and the last CTA of a panel
    zeroes it again: back-to-back launches on different inputs must not see each other."""
    from rock_art_radnet_b200.utils import RpnTargetBatch
    C = S.HotPathConfig()
    B, G = 16, 24
    tb = RpnTargetBatch(C, B, G, 38, 38)
    wh = torch.tensor([[600.0, 600.0]] * B, dtype=torch.float64, device="cuda")
    outs = []
    for rnd in range(3):
        imgs = [S.gt_figures(900 + 31 * rnd + b, (b * 5 + rnd) % (G + 1), 600, 600) for b in range(B)]
        gt, bg, cnt = _gt_batch(imgs, G)
        res = tb.run(torch.from_numpy(gt).cuda(), torch.from_numpy(bg).cuda(), torch.from_numpy(cnt).cuda(), wh)
        outs.append((imgs, [t.cpu().numpy().copy() for t in res]))
    assert int(tb.ws.to(torch.int64).sum()) == 0, "workspace counters not left at zero"
    for imgs, (y_cls, y_regr, best, hits) in outs:
        for b in (0, 5, 15):
            valid, overlap, regr, ba, nh = O.rpn_targets_presample(C, imgs[b], 600, 600, 600, 600, S.resnet50_map_size)
            g = len(imgs[b]["bboxes"])
            assert np.array_equal(y_cls[b, :9], valid.transpose(2, 0, 1))
            assert np.array_equal(y_cls[b, 9:], overlap.transpose(2, 0, 1))
            assert np.array_equal(y_regr[b, :36], np.repeat(overlap.transpose(2, 0, 1), 4, axis=0))
            np.testing.assert_allclose(y_regr[b, 36:], regr.transpose(2, 0, 1), rtol=REGR_RTOL, atol=0)
            assert np.array_equal(best[b, :g], ba) and (best[b, g:] == -1).all()
            assert np.array_equal(hits[b, :g], nh) and (hits[b, g:] == 0).all()


def test_rpn_targets_64_panel_batch_equals_single_panel_launches(pkg):
    """BASELINE configs[1] at the batch the bench times: 64 panels x 20 figures in one launch are byte-identical
    to 64 single-panel launches (size-independent property), and sampled panels equal the oracle."""
    from rock_art_radnet_b200.utils import RpnTargetBatch
    C = S.HotPathConfig()
    B, G = 64, 20
    imgs = [S.gt_figures(b, G, 600, 600) for b in range(B)]
    gt, bg, cnt = _gt_batch(imgs, G)
    gt_d, bg_d, cnt_d = torch.from_numpy(gt).cuda(), torch.from_numpy(bg).cuda(), torch.from_numpy(cnt).cuda()
    wh = torch.tensor([[600.0, 600.0]] * B, dtype=torch.float64, device="cuda")
    big = RpnTargetBatch(C, B, G, 38, 38)
    y_cls, y_regr, best, hits = [t.clone() for t in big.run(gt_d, bg_d, cnt_d, wh)]
    one = RpnTargetBatch(C, 1, G, 38, 38)
    for b in range(B):
        r = one.run(gt_d[b:b + 1], bg_d[b:b + 1], cnt_d[b:b + 1], wh[b:b + 1])
        assert torch.equal(r[0][0], y_cls[b]) and torch.equal(r[1][0], y_regr[b]), b
        assert torch.equal(r[2][0], best[b]) and torch.equal(r[3][0], hits[b]), b
    for b in (0, 17, 63):
        valid, overlap, regr, ba, nh = O.rpn_targets_presample(C, imgs[b], 600, 600, 600, 600, S.resnet50_map_size)
        assert np.array_equal(y_cls[b, :9].cpu().numpy(), valid.transpose(2, 0, 1))
        assert np.array_equal(best[b].cpu().numpy(), ba) and np.array_equal(hits[b].cpu().numpy(), nh)


def test_roi_targets_batch_matches_per_panel_calc_iou(pkg):
    """Batched a4 fed straight from the detection records of K2 == calc_iou per panel (oracle)."""
    from rock_art_radnet_b200.pipeline import ProposalPipeline
    from rock_art_radnet_b200.rpn import RoiTargetBatch, gt_feature_cells
    C = S.HotPathConfig()
    B, H, W, Gmax = 8, 38, 38, 24
    maps = [S.rpn_maps(700 + s, H, W, 9) for s in range(B)]
    cls = torch.from_numpy(np.concatenate([m[0] for m in maps])).cuda()
    regr = torch.from_numpy(np.concatenate([m[1] for m in maps])).cuda()
    pipe = ProposalPipeline(C, B, H, W, alloc_pooled=False)
    pipe.decode(cls, regr)
    pipe.sort_nms()
    imgs = [S.gt_figures(700 + s, (3 * s) % (Gmax + 1), 600, 600, classes=("boat", "human", "wheel")) for s in range(B)]
    imgs[2] = {"bboxes": [{"class": "boat", "x1": 0, "x2": 2, "y1": 0, "y2": 2}], "width": 600, "height": 600}
    gt = np.zeros((B, Gmax, 4)); gc = np.zeros((B, Gmax), np.int32); cnt = np.zeros(B, np.int32)
    for b, img in enumerate(imgs):
        a, c = gt_feature_cells(img, C, C.class_mapping)
        gt[b, :len(a)] = a
        gc[b, :len(a)] = c
        cnt[b] = len(a)
    rt = RoiTargetBatch(C, C.class_mapping, B, 300, Gmax)
    x, y1, y2, ious, best, count = (t.cpu().numpy() for t in rt.run(
        torch.from_numpy(gt).cuda(), torch.from_numpy(gc).cuda(), torch.from_numpy(cnt).cuda(), det=pipe.records))
    dets = pipe.records.to_numpy()
    for b, img in enumerate(imgs):
        ref = O.calc_iou(dets[b]["boxes"], img, C, C.class_mapping)
        n = int(count[b])
        if ref[0] is None:
            assert n == 0, b
            continue
        assert n == ref[0].shape[1], b
        assert np.array_equal(x[b, :n], ref[0][0]) and np.array_equal(y1[b, :n], ref[1][0])
        assert np.array_equal(y2[b, :n] != 0, ref[2][0] != 0)
        np.testing.assert_allclose(y2[b, :n], ref[2][0], rtol=REGR_RTOL, atol=0)
        assert np.array_equal(ious[b, :n], np.asarray(ref[3]))
        pos = y1[b, :n, -1] == 0
        assert ((best[b, :n] >= 0) == pos).all()
    # dense-RoI addressing with per-panel counts gives the same rows
    rois = torch.zeros((B, 300, 4), dtype=torch.int32, device="cuda")
    rc = torch.zeros((B,), dtype=torch.int32, device="cuda")
    for b in range(B):
        k = dets[b]["boxes"].shape[0]
        rois[b, :k] = torch.from_numpy(dets[b]["boxes"].astype(np.int32)).cuda()
        rc[b] = k
    rt2 = RoiTargetBatch(C, C.class_mapping, B, 300, Gmax)
    r2 = rt2.run(torch.from_numpy(gt).cuda(), torch.from_numpy(gc).cuda(), torch.from_numpy(cnt).cuda(), rois=rois, roi_count=rc)
    assert np.array_equal(r2[5].cpu().numpy(), count)
    for b in range(B):
        n = int(count[b])
        assert np.array_equal(r2[0][b, :n].cpu().numpy(), x[b, :n]) and np.array_equal(r2[2][b, :n].cpu().numpy(), y2[b, :n])


def test_calc_iou_unknown_class_only_matters_for_the_best_match(pkg):
    """The reference looks a figure's class up only once it is the best match of a positive RoI (rpn.py:263)."""
    C = S.HotPathConfig()
    cls, regr = S.rpn_maps(11)
    R = pkg.rpn_to_roi(cls, regr, C, max_boxes=300, overlap_thresh=0.7)
    img = S.gt_figures(11, 12, 600, 600, classes=("boat",))
    ref = O.calc_iou(R, img, C, C.class_mapping)
    # a figure of an unknown class far away from every proposal: never the best match -> same result
    far = dict(img, bboxes=img["bboxes"] + [{"class": "unicorn", "x1": 0, "x2": 1, "y1": 0, "y2": 1}])
    got = pkg.calc_iou(R, far, C, C.class_mapping)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    # the same unknown class on a figure that IS matched -> KeyError, like class_mapping[cls_name]
    bad = dict(img, bboxes=[dict(bb, **{"class": "unicorn"}) for bb in img["bboxes"]])
    if (ref[1][0, :, -1] == 0).any():
        with pytest.raises(KeyError):
            pkg.calc_iou(R, bad, C, C.class_mapping)
    with pytest.raises(ValueError):
        pkg.calc_iou(R, img, C, {"bg": 0, "boat": 1})


def test_rpn_targets_two_launch_form_and_role_split_agree(pkg, lib_option):
    """The persistent launch (fill SMs + compute SMs), the same with a single compute CTA, and the
    two-launch form (no co-residency assumed) write byte-identical tensors."""
    from rock_art_radnet_b200.utils import RpnTargetBatch
    C = S.HotPathConfig()
    B, G = 24, 20
    imgs = [S.gt_figures(1200 + b, G - (b % 3), 600, 600, classes=("boat", "bg")) for b in range(B)]
    gt, bg, cnt = _gt_batch(imgs, G)
    args = (torch.from_numpy(gt).cuda(), torch.from_numpy(bg).cuda(), torch.from_numpy(cnt).cuda(),
            torch.tensor([[600.0, 600.0]] * B, dtype=torch.float64, device="cuda"))
    tb = RpnTargetBatch(C, B, G, 38, 38)
    ref = [t.clone() for t in tb.run(*args)]
    for name, value in (("targets_compute_ctas", 1), ("targets_compute_ctas", 147), ("targets_two_launches", 1)):
        lib_option(name, value)
        for t in (tb.y_cls, tb.y_regr, tb.best, tb.hits):
            t.fill_(-7)
        got = tb.run(*args)
        assert all(torch.equal(a, b) for a, b in zip(ref, got)), (name, value)
        lib_option(name, 0)
    assert int(tb.ws.to(torch.int64).sum()) == 0
