"""GPU parity of f4: the losses (faster_rcnn/losses.py:16-95) against the oracle restatement -
bar 1e-5 relative, the north_star tolerance, parity UNPINNED against Keras itself - and the mAP evaluation
(test.py:48-173) against the goldens of the unmodified reference functions - match flags and orders bit-exact,
precision / recall arrays bit-exact, AP within 1e-12 relative (tree sum instead of the reference's left-to-right)."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from conftest import GOLDEN, _load_npz  # noqa: E402
from oracle import train_oracle as T  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402
from test_oracle_train import loss_inputs  # noqa: E402

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-5


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    import rock_art_radnet_b200 as R
    from rock_art_radnet_b200 import _lib
    _lib.load()
    return R


@pytest.fixture(scope="module")
def golden_f4():
    return _load_npz("f4_eval.npz")


@pytest.fixture(scope="module")
def manifest_f4():
    with open(os.path.join(GOLDEN, "manifest_f4.json")) as f:
        return json.load(f)


def test_dropin_losses_match_oracle(pkg):
    from rock_art_radnet_b200 import losses as L
    for seed in range(4):
        y_cls, y_regr, p_cls, p_regr, Y1, Y2, q_cls, q_regr = loss_inputs(seed)
        pairs = [(L.rpn_loss_regr(9)(y_regr, p_regr), T.rpn_loss_regr(9)(y_regr, p_regr)),
                 (L.rpn_loss_cls(9)(y_cls, p_cls), T.rpn_loss_cls(9)(y_cls, p_cls)),
                 (L.class_loss_regr(6)(Y2, q_regr), T.class_loss_regr(6)(Y2, q_regr)),
                 (L.class_loss_cls(Y1, q_cls), T.class_loss_cls(Y1, q_cls))]
        for got, want in pairs:
            assert got.dtype == np.float32 and np.isfinite(got)
            assert abs(float(got) - float(want)) <= LOSS_RTOL * abs(float(want)), (seed, got, want)


def test_batched_losses_on_the_device_pipeline(pkg):
    """K3 (NHWC, regr x std_scaling) -> subsampler -> rpn losses, and K2 records -> a4 -> selection -> class losses,
    all resident; every panel's four values equal the oracle's on the same tensors; repeated launches are bit-equal."""
    from rock_art_radnet_b200.losses import RpnLossBatch, class_losses_device
    from rock_art_radnet_b200.pipeline import ProposalPipeline
    from rock_art_radnet_b200.rpn import RoiTargetBatch, gt_feature_cells
    from rock_art_radnet_b200.sampling import RpnSubsampler, SampleSelector, seed_states
    from rock_art_radnet_b200.utils import LAYOUT_NHWC, RpnTargetBatch
    C = S.HotPathConfig()
    B, G, H, W, A = 6, 20, 38, 38, 9
    imgs = [S.gt_figures(3000 + b, G, 600, 600, classes=("boat", "human", "wheel")) for b in range(B)]
    gt = np.zeros((B, G, 4)); gtc = np.zeros((B, G, 4)); gcl = np.zeros((B, G), np.int32)
    for b, img in enumerate(imgs):
        for k, bb in enumerate(img["bboxes"]):
            gt[b, k] = [bb["x1"], bb["x2"], bb["y1"], bb["y2"]]
        gtc[b], gcl[b] = gt_feature_cells(img, C, C.class_mapping)
    cnt = torch.full((B,), G, dtype=torch.int32, device="cuda")
    wh = torch.tensor([[600.0, 600.0]] * B, dtype=torch.float64, device="cuda")
    tb = RpnTargetBatch(C, B, G, H, W, layout=LAYOUT_NHWC, regr_scale=C.std_scaling)
    y_cls, y_regr, _, _ = tb.run(torch.from_numpy(gt).cuda(), torch.zeros((B, G), dtype=torch.uint8, device="cuda"), cnt, wh)
    RpnSubsampler(B, H, W, A, layout=LAYOUT_NHWC).run(y_cls, seed_states(np.arange(B)))
    g = torch.Generator(device="cuda").manual_seed(1)
    p_cls = torch.sigmoid(3 * torch.randn((B, H, W, A), device="cuda", generator=g))
    p_regr = torch.randn((B, H, W, 4 * A), device="cuda", generator=g)
    lb = RpnLossBatch(B, H, W, A)
    loss = lb.run(y_cls, y_regr, p_cls, p_regr).cpu().numpy()
    assert np.array_equal(loss, lb.run(y_cls, y_regr, p_cls, p_regr).cpu().numpy())
    yc, yr, pc, pr = (t.cpu().numpy() for t in (y_cls, y_regr, p_cls, p_regr))
    for b in range(B):
        want = (T.rpn_loss_cls(A)(yc[b:b + 1], pc[b:b + 1]), T.rpn_loss_regr(A)(yr[b:b + 1], pr[b:b + 1]))
        assert abs(loss[b, 0] - want[0]) <= LOSS_RTOL * abs(want[0]) and abs(loss[b, 1] - want[1]) <= LOSS_RTOL * abs(want[1]), b
    # classifier side
    maps = [S.rpn_maps(3000 + b) for b in range(B)]
    pipe = ProposalPipeline(C, B, H, W, alloc_pooled=False)
    pipe.decode(torch.from_numpy(np.concatenate([m[0] for m in maps])).cuda(),
                torch.from_numpy(np.concatenate([m[1] for m in maps])).cuda())
    pipe.sort_nms()
    rt = RoiTargetBatch(C, C.class_mapping, B, 300, G)
    x_roi, y_class, y2, _, _, count = rt.run(torch.from_numpy(gtc).cuda(), torch.from_numpy(gcl).cuda(), cnt, det=pipe.records)
    n_rois, n_cls = int(C.n_rois), 7
    sel, rep = SampleSelector(B, 300, n_cls, n_rois).run(y_class, count, seed_states(np.arange(B) + 50))
    q_cls = torch.softmax(2 * torch.randn((B, n_rois, n_cls), device="cuda", generator=g), dim=-1)
    q_regr = torch.randn((B, n_rois, 4 * (n_cls - 1)), device="cuda", generator=g)
    closs = class_losses_device(y_class, y2, q_cls, q_regr, sel=sel, n_sel_per_panel=rep[:, 0].contiguous()).cpu().numpy()
    selh, reph = sel.cpu().numpy(), rep.cpu().numpy()
    Y1, Y2h, qc, qr = (t.cpu().numpy() for t in (y_class, y2, q_cls, q_regr))
    for b in range(B):
        k = int(reph[b, 0])
        rows = selh[b, :k]
        want_c = T.class_loss_cls(Y1[b:b + 1, rows], qc[b:b + 1, :k])
        want_r = T.class_loss_regr(n_cls - 1)(Y2h[b:b + 1, rows], qr[b:b + 1, :k])
        assert abs(closs[b, 0] - want_c) <= LOSS_RTOL * abs(want_c), b
        assert abs(closs[b, 1] - want_r) <= LOSS_RTOL * max(abs(want_r), 1e-30), b


def test_get_objects_and_calc_class_ap_match_reference_golden(pkg, golden_f4, manifest_f4):
    from rock_art_radnet_b200 import evaluation as E
    for case in manifest_f4["eval"]:
        det, gt = S.eval_set(case["seed"], case["n_gt"], case["n_det"], ties=case["ties"])
        Tm, Pm = E.get_objects(det, gt, 0.5)
        n = case["name"]
        assert list(Tm.keys()) == golden_f4[n + "/keys"].tolist(), n
        assert [int(g['bbox_matched']) for g in gt] == golden_f4[n + "/matched"].tolist(), n
        aps = []
        for k in Tm:
            assert Tm[k] == golden_f4["%s/T/%s" % (n, k)].tolist(), (n, k)
            assert np.array_equal(np.array(Pm[k], dtype=np.float64), golden_f4["%s/P/%s" % (n, k)]), (n, k)
            ap, prec, rec, ip, ir = E.calc_class_ap(Tm[k], Pm[k])
            want = float(golden_f4["%s/ap/%s" % (n, k)])
            assert abs(ap - want) <= 1e-12 * max(abs(want), 1e-300), (n, k, ap, want)
            assert np.array_equal(prec, golden_f4["%s/prec/%s" % (n, k)]) and np.array_equal(rec, golden_f4["%s/rec/%s" % (n, k)])
            assert np.array_equal(np.array(ip), golden_f4["%s/iprec/%s" % (n, k)]) and isinstance(ip, list)
            assert np.array_equal(np.array(ir), golden_f4["%s/irec/%s" % (n, k)])
            aps.append(ap)
        if aps:
            assert abs(np.mean(aps) - float(golden_f4[n + "/mAP"])) <= 1e-12


def test_matching_with_tied_scores_and_large_sets_matches_oracle(pkg):
    from rock_art_radnet_b200 import evaluation as E
    for seed, n_gt, n_det, ties in [(20, 40, 200, True), (21, 2500, 6000, False), (22, 33, 1, False)]:
        det, gt = S.eval_set(seed, n_gt, n_det, ties=ties)
        det2, gt2 = S.eval_set(seed, n_gt, n_det, ties=ties)
        got = E.get_objects(det, gt, 0.5)
        want = T.get_objects(det2, gt2, 0.5)
        assert list(got[0].keys()) == list(want[0].keys())
        for k in want[0]:
            assert got[0][k] == want[0][k] and got[1][k] == want[1][k], (seed, k)
            a = E.calc_class_ap(got[0][k], got[1][k])
            b = T.calc_class_ap(want[0][k], want[1][k])
            assert abs(a[0] - b[0]) <= 1e-12 * max(abs(b[0]), 1e-300) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
