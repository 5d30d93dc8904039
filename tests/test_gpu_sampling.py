"""GPU parity of the training-side sampling: the device replay of NumPy's legacy generator
against the goldens of the unmodified reference functions (oracle/make_golden_f3.py), against the MT19937
oracle and against np.random itself - picks bit-identical, generator left at the same stream position."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from conftest import GOLDEN, _load_npz, dense_regr  # noqa: E402
from oracle import mt19937_oracle as M  # noqa: E402
from oracle.make_golden import A3_CASES, a3_inputs  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    import rock_art_radnet_b200 as R
    from rock_art_radnet_b200 import _lib
    _lib.load()
    return R


@pytest.fixture(scope="module")
def golden_f3():
    return _load_npz("f3_sampling.npz")


@pytest.fixture(scope="module")
def manifest_f3():
    with open(os.path.join(GOLDEN, "manifest_f3.json")) as f:
        return json.load(f)


def test_seed_states_match_numpy(pkg):
    from rock_art_radnet_b200.sampling import seed_states
    seeds = [0, 1, 1234, 2 ** 32 - 1, 77]
    st = seed_states(seeds).cpu().numpy().view(np.uint32)
    for s, row in zip(seeds, st):
        ref = np.random.RandomState(s).get_state()
        assert np.array_equal(row[:624], ref[1]) and row[624] == ref[2]


def test_calc_region_props_leaves_np_random_where_the_reference_does(pkg, manifest, golden_a3, golden_f3):
    """Reference goldens: labels after the 256-region balancing AND the next draw of the global generator."""
    C = S.HotPathConfig()
    for case in manifest["a3"]:
        img = a3_inputs(case["seed"], case["width"], case["height"], case["n_gt"], tuple(case["classes"]),
                        small=(case["name"] == "small_gt"))
        wr, hr = case["resized"]
        np.random.seed(case["seed"])
        y_cls, y_regr, best, n_pos = pkg.calc_region_props(C, img, case["width"], case["height"], wr, hr,
                                                           S.resnet50_map_size)
        n = case["name"]
        assert np.array_equal(y_cls, golden_a3[n + "/y_rpn_cls"].astype(np.float64)), n
        assert int(n_pos) == int(golden_a3[n + "/n_pos"])
        assert np.random.random_sample() == float(golden_f3["a3_state/%s/next_draw" % n]), n
        assert y_cls.flags.writeable and y_regr.flags.writeable and y_cls.base is None


def _label_tensor(seed, n_pos, n_neg, A, H, W, nhwc):
    rng = np.random.default_rng(seed)
    flat = rng.permutation(A * H * W)
    valid = np.zeros(A * H * W); overlap = np.zeros(A * H * W)
    valid[flat[:n_pos + n_neg]] = 1
    overlap[flat[:n_pos]] = 1
    # a few invalid anchors that carry an overlap flag must be ignored by both branches
    overlap[flat[n_pos + n_neg:n_pos + n_neg + 5]] = 1
    v, o = valid.reshape(A, H, W), overlap.reshape(A, H, W)
    y = np.concatenate([v, o], axis=0)
    return v, o, (np.ascontiguousarray(y.transpose(1, 2, 0)) if nhwc else y)


@pytest.mark.parametrize("force_exact", [0, 1])
def test_batched_subsampler_matches_oracle_draw_for_draw(pkg, lib_option, force_exact):
    from rock_art_radnet_b200.sampling import RpnSubsampler, seed_states
    lib_option("sampler_force_exact", force_exact)
    A, H, W = 9, 20, 23
    cases = [(40, 3000), (200, 3000), (150, 60), (10, 100), (0, 500), (700, 2000), (129, 127), (128, 128), (3000, 1000)]
    for nhwc in (False, True):
        B = len(cases)
        tens, refs = [], []
        for i, (n_pos, n_neg) in enumerate(cases):
            v, o, y = _label_tensor(100 + i, n_pos, n_neg, A, H, W, nhwc)
            tens.append(y)
            refs.append((v, o))
        y_d = torch.from_numpy(np.stack(tens)).cuda()
        seeds = [500 + i for i in range(B)]
        states = seed_states(seeds)
        sub = RpnSubsampler(B, H, W, A, layout=1 if nhwc else 0)
        rep = sub.run(y_d, states).cpu().numpy()
        got = y_d.cpu().numpy()
        st = states.cpu().numpy().view(np.uint32)
        for i, (v, o) in enumerate(refs):
            g = M.MT19937.from_seed(seeds[i])
            v2 = v.copy()
            try:
                n_pos = M.subsample_regions(g, v2, o)
                err = False
            except KeyError:
                err = True
            gv = got[i][..., :A].transpose(2, 0, 1) if nhwc else got[i][:A]
            if err:
                assert rep[i, 3] == 1 and np.array_equal(gv, v), i          # nothing drawn, nothing changed
                ref_state = np.random.RandomState(seeds[i]).get_state()
                assert np.array_equal(st[i, :624], ref_state[1]) and st[i, 624] == ref_state[2]
                continue
            assert rep[i, 3] == 0 and rep[i, 0] == n_pos, (i, rep[i])
            assert np.array_equal(gv, v2), (i, cases[i])
            assert np.array_equal(st[i, :624], np.array(g.key, dtype=np.uint32)) and st[i, 624] == g.pos, i
            if force_exact:
                assert rep[i, 4] == 0


def test_subsampler_on_the_headline_map_and_many_seeds(pkg):
    """38x38x9 label tensors as K3 writes them (20 figures), > 256 regions: every panel loses negatives; the picks
    equal the oracle's for 24 panels with independent streams, and the report counts the draws."""
    from rock_art_radnet_b200.sampling import RpnSubsampler, seed_states
    from rock_art_radnet_b200.utils import rpn_targets_device
    from oracle import radnet_oracle as O
    C = S.HotPathConfig()
    B, G = 24, 20
    imgs = [S.gt_figures(2000 + b, G, 600, 600) for b in range(B)]
    gt = np.zeros((B, G, 4)); bg = np.zeros((B, G), np.uint8)
    for b, img in enumerate(imgs):
        for k, bb in enumerate(img["bboxes"]):
            gt[b, k] = [bb["x1"], bb["x2"], bb["y1"], bb["y2"]]
    cnt = np.full((B,), G, np.int32)
    wh = np.tile(np.array([[600.0, 600.0]]), (B, 1))
    y_cls, _, _, _ = rpn_targets_device(C, gt, bg, cnt, 38, 38, wh)
    pre = y_cls.cpu().numpy().copy()
    states = seed_states(np.arange(B) + 9000)
    sub = RpnSubsampler(B, 38, 38, 9)
    rep = sub.run(y_cls, states).cpu().numpy()
    post = y_cls.cpu().numpy()
    for b in range(B):
        g = M.MT19937.from_seed(9000 + b)
        v = pre[b, :9].copy()
        n_pos = M.subsample_regions(g, v, pre[b, 9:])
        assert np.array_equal(post[b, :9], v) and np.array_equal(post[b, 9:], pre[b, 9:]), b
        assert rep[b, 0] == n_pos and rep[b, 5] > 0
        assert int((post[b, :9] == 1).sum()) <= 256


def test_get_selected_samples_matches_reference_golden(pkg, golden_f3, manifest_f3):
    from rock_art_radnet_b200.sampling import get_selected_samples
    for case in manifest_f3["select"]:
        Y1 = S.one_hot_rows(case["seed"], case["n_pos"], case["n_neg"])
        C = S.HotPathConfig()
        C.n_rois = case["n_rois"]
        np.random.seed(case["seed"])
        sel, n_pos = get_selected_samples(Y1, C)
        n = "select/%s/" % case["name"]
        assert sel == golden_f3[n + "sel"].tolist(), case["name"]
        assert n_pos == int(golden_f3[n + "n_pos"])
        assert np.random.random_sample() == float(golden_f3[n + "next_draw"]), case["name"]
    with pytest.raises(ValueError):
        get_selected_samples(np.zeros((1, 0, 7), dtype=np.int64), S.HotPathConfig())


def test_batched_sample_selector_matches_oracle(pkg):
    from rock_art_radnet_b200.sampling import SampleSelector, seed_states
    rng = np.random.default_rng(3)
    B, R, n_cls, n_rois = 32, 300, 7, 20
    y = np.zeros((B, R, n_cls), dtype=np.int32)
    counts = np.zeros((B,), dtype=np.int32)
    rows = []
    for b in range(B):
        n_pos = int(rng.integers(0, 60)) if b % 5 else 0
        n_neg = int(rng.integers(0, 240)) if b % 7 else 0
        if n_pos + n_neg == 0:
            n_neg = 3
        Y1 = S.one_hot_rows(40 + b, n_pos, n_neg, n_cls)
        rows.append(Y1)
        y[b, :n_pos + n_neg] = Y1[0]
        y[b, n_pos + n_neg:, 2] = 1            # stale rows beyond count must be ignored
        counts[b] = n_pos + n_neg
    states = seed_states(np.arange(B) + 70)
    sel, rep = SampleSelector(B, R, n_cls, n_rois).run(torch.from_numpy(y).cuda(), torch.from_numpy(counts).cuda(), states)
    sel, rep = sel.cpu().numpy(), rep.cpu().numpy()
    st = states.cpu().numpy().view(np.uint32)
    for b in range(B):
        g = M.MT19937.from_seed(70 + b)
        try:
            ref, n_pos = M.get_selected_samples(g, rows[b], n_rois)
        except ValueError:
            assert rep[b, 3] == 2, b
            continue
        assert rep[b, 3] == 0 and rep[b, 0] == len(ref) and rep[b, 1] == n_pos, (b, rep[b])
        assert sel[b, :len(ref)].tolist() == ref, b
        assert np.array_equal(st[b, :624], np.array(g.key, dtype=np.uint32)) and st[b, 624] == g.pos, b
