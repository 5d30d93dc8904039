"""CPU: the f4 oracle (oracle/train_oracle.py) - evaluation functions against the goldens of the unmodified reference
(oracle/make_golden_f4.py) and the loss restatements against independent float64 evaluations of the same formulas."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, _load_npz
from oracle import train_oracle as T
from rock_art_radnet_b200 import synthetic as S


@pytest.fixture(scope="module")
def golden_f4():
    return _load_npz("f4_eval.npz")


@pytest.fixture(scope="module")
def manifest_f4():
    with open(os.path.join(GOLDEN, "manifest_f4.json")) as f:
        return json.load(f)


def test_get_objects_and_class_ap_match_reference_golden(golden_f4, manifest_f4):
    for case in manifest_f4["eval"]:
        det, gt = S.eval_set(case["seed"], case["n_gt"], case["n_det"], ties=case["ties"])
        Tm, Pm = T.get_objects(det, gt, 0.5)
        n = case["name"]
        assert list(Tm.keys()) == golden_f4[n + "/keys"].tolist(), n
        for k in Tm:
            assert Tm[k] == golden_f4["%s/T/%s" % (n, k)].tolist(), (n, k)
            assert np.array_equal(np.array(Pm[k], dtype=np.float64), golden_f4["%s/P/%s" % (n, k)]), (n, k)
            ap, prec, rec, ip, ir = T.calc_class_ap(Tm[k], Pm[k])
            assert ap == float(golden_f4["%s/ap/%s" % (n, k)])
            assert np.array_equal(prec, golden_f4["%s/prec/%s" % (n, k)]) and np.array_equal(rec, golden_f4["%s/rec/%s" % (n, k)])
            assert np.array_equal(np.array(ip), golden_f4["%s/iprec/%s" % (n, k)])
            assert np.array_equal(np.array(ir), golden_f4["%s/irec/%s" % (n, k)])


def loss_inputs(seed, H=38, W=38, A=9, n_rows=20, n_cls=7):
    """Targets in the layouts of K3 (NHWC, regr half x std_scaling) / a4 and plausible network outputs."""
    rng = np.random.default_rng(seed)
    valid = (rng.random((1, H, W, A)) < 0.02).astype(np.float64)
    overlap = valid * (rng.random((1, H, W, A)) < 0.4)
    y_cls = np.concatenate([valid, overlap], axis=3)
    regr = np.repeat(overlap, 4, axis=3) * rng.standard_normal((1, H, W, 4 * A)) * 2.0
    y_regr = np.concatenate([np.repeat(overlap, 4, axis=3), regr], axis=3)
    p_cls = (1.0 / (1.0 + np.exp(-3 * rng.standard_normal((1, H, W, A))))).astype(np.float32)
    p_cls[0, 0, 0, :3] = [0.0, 1.0, 0.5]                        # saturated predictions
    p_regr = rng.standard_normal((1, H, W, 4 * A)).astype(np.float32)
    cls_id = rng.integers(0, n_cls, n_rows)
    Y1 = np.zeros((1, n_rows, n_cls), dtype=np.int64)
    Y1[0, np.arange(n_rows), cls_id] = 1
    n4 = 4 * (n_cls - 1)
    Y2 = np.zeros((1, n_rows, 2 * n4))
    for r, c in enumerate(cls_id):
        if c < n_cls - 1:
            Y2[0, r, 4 * c:4 * c + 4] = 1
            Y2[0, r, n4 + 4 * c:n4 + 4 * c + 4] = rng.standard_normal(4) * 3
    z = rng.standard_normal((1, n_rows, n_cls)) * 2
    q_cls = (np.exp(z) / np.exp(z).sum(-1, keepdims=True)).astype(np.float32)
    q_regr = rng.standard_normal((1, n_rows, n4)).astype(np.float32)
    return y_cls, y_regr, p_cls, p_regr, Y1, Y2, q_cls, q_regr


def test_loss_restatements_against_float64_formulas():
    y_cls, y_regr, p_cls, p_regr, Y1, Y2, q_cls, q_regr = loss_inputs(0)
    A, n_cls = 9, 7
    # smooth-L1 (losses.py:31-42) in float64
    def sl1(mask, t, pred):
        x = t.astype(np.float32).astype(np.float64) - pred.astype(np.float64)
        xa = np.abs(x)
        return (mask * np.where(xa <= 1.0, 0.5 * x * x, xa - 0.5)).sum() / (1e-4 + mask).sum()
    got = T.rpn_loss_regr(A)(y_regr, p_regr)
    assert got.dtype == np.float32 and abs(got - sl1(y_regr[..., :4 * A], y_regr[..., 4 * A:], p_regr)) <= 1e-5 * abs(got)
    n4 = 4 * (n_cls - 1)
    got = T.class_loss_regr(n_cls - 1)(Y2, q_regr)
    assert abs(got - sl1(Y2[..., :n4], Y2[..., n4:], q_regr)) <= 1e-5 * abs(got)
    # BCE with the reference's argument order: logits from the clipped LABEL, prediction as target
    lab = np.clip(y_cls[..., A:], 1e-7, 1 - 1e-7).astype(np.float32).astype(np.float64)
    x = np.log(lab / (1 - lab))
    bce = np.maximum(x, 0) - x * p_cls.astype(np.float64) + np.log1p(np.exp(-np.abs(x)))
    want = (y_cls[..., :A] * bce).sum() / (1e-4 + y_cls[..., :A]).sum()
    got = T.rpn_loss_cls(A)(y_cls, p_cls)
    assert abs(got - want) <= 2e-5 * abs(want)
    # categorical cross-entropy
    o = q_cls[0].astype(np.float64)
    o = np.clip(o / o.sum(-1, keepdims=True), 1e-7, 1 - 1e-7)
    want = (-(Y1[0] * np.log(o)).sum(-1)).mean()
    got = T.class_loss_cls(Y1, q_cls)
    assert abs(got - want) <= 1e-5 * abs(want)


def test_loss_restatements_against_torch_library_functions():
    """The Keras/TF backend calls `losses.py` composes have library counterparts in torch with the same published
    definitions: tf.nn.sigmoid_cross_entropy_with_logits = F.binary_cross_entropy_with_logits (max(x,0) - x*z +
    log(1+exp(-|x|))), the reference's hand-written smooth-L1 = F.smooth_l1_loss(beta=1), categorical cross-entropy =
    NLL of the log of the re-normalised, clipped probabilities.  Not Keras itself (not installable: the losses stay
    'parity unpinned'), but an independent implementation of every building block."""
    import torch
    import torch.nn.functional as F
    y_cls, y_regr, p_cls, p_regr, Y1, Y2, q_cls, q_regr = loss_inputs(3)
    A, n_cls = 9, 7
    t64 = lambda a: torch.from_numpy(np.asarray(a)).double()
    # rpn_loss_cls (losses.py:63-65): K.binary_crossentropy(y_pred, label) -> logits from the CLIPPED LABEL
    lab = torch.from_numpy(np.clip(y_cls[..., A:], 1e-7, 1 - 1e-7).astype(np.float32)).double()
    logits = torch.log(lab / (1 - lab))
    bce = F.binary_cross_entropy_with_logits(logits, t64(p_cls), reduction="none")
    valid = t64(y_cls[..., :A])
    want = float((valid * bce).sum() / (1e-4 + valid).sum())
    got = float(T.rpn_loss_cls(A)(y_cls, p_cls))
    assert abs(got - want) <= 2e-5 * abs(want)
    # smooth-L1 (losses.py:31-42, 79-86)
    for y_true, pred, n4 in ((y_regr, p_regr, 4 * A), (Y2, q_regr, 4 * (n_cls - 1))):
        mask, tgt = t64(y_true[..., :n4]), t64(y_true[..., n4:].astype(np.float32))
        sl1 = F.smooth_l1_loss(t64(pred), tgt, beta=1.0, reduction="none")
        want = float((mask * sl1).sum() / (1e-4 + mask).sum())
        fn = T.rpn_loss_regr(A) if n4 == 4 * A else T.class_loss_regr(n_cls - 1)
        got = float(fn(y_true, pred))
        assert abs(got - want) <= 1e-5 * abs(want)
    # class_loss_cls (losses.py:93-95)
    o = t64(q_cls[0])
    o = torch.clamp(o / o.sum(-1, keepdim=True), 1e-7, 1 - 1e-7)
    want = float(F.nll_loss(torch.log(o), torch.from_numpy(Y1[0].argmax(-1)), reduction="mean"))
    got = float(T.class_loss_cls(Y1, q_cls))
    assert abs(got - want) <= 1e-5 * abs(want)
