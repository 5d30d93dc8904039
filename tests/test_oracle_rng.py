"""CPU: the MT19937 / legacy np.random.choice restatement (oracle/mt19937_oracle.py) is pinned against NumPy's
own RandomState (values and the state left behind) and against the f3 goldens produced by the unmodified
reference functions (oracle/make_golden_f3.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, _load_npz
from oracle import mt19937_oracle as M
from oracle import radnet_oracle as O
from oracle.make_golden import A3_CASES, a3_inputs
from rock_art_radnet_b200 import synthetic as S


@pytest.fixture(scope="module")
def golden_f3():
    return _load_npz("f3_sampling.npz")


@pytest.fixture(scope="module")
def manifest_f3():
    with open(os.path.join(GOLDEN, "manifest_f3.json")) as f:
        return json.load(f)


def _same_state(rs, g):
    st = rs.get_state()
    return np.array_equal(st[1], np.array(g.key, dtype=np.uint32)) and st[2] == g.pos


@pytest.mark.parametrize("seed", range(12))
def test_generator_and_choice_variants_match_numpy(seed):
    rs = np.random.RandomState(seed)
    rs.random_sample(seed * 37 % 700)                      # move the position inside the key block
    g = M.MT19937.from_numpy_state(rs.get_state())
    n = 50 + seed * 97
    k = n // 2 + seed
    p = np.random.default_rng(seed).random(n)
    p /= p.sum()
    assert np.array_equal(rs.choice(n, k, replace=False, p=p), M.choice_noreplace_p(g, n, k, p))
    pop = np.arange(1000, 1000 + n)
    assert np.array_equal(rs.choice(pop, k, replace=False), M.choice_noreplace(g, pop, k))
    assert np.array_equal(rs.choice(pop, k + 5, replace=True), M.choice_replace(g, pop, k + 5))
    assert np.array_equal(rs.choice(pop[:1], 4, replace=True), M.choice_replace(g, pop[:1], 4))
    assert np.array_equal(rs.random_sample(700), g.random_sample(700))
    assert _same_state(rs, g)


def test_seeding_matches_numpy():
    for seed in (0, 1, 1234, 2 ** 32 - 1):
        g = M.MT19937.from_seed(seed)
        rs = np.random.RandomState(seed)
        assert np.array_equal(rs.random_sample(3), g.random_sample(3)) and _same_state(rs, g)


@pytest.mark.parametrize("n_pos,n_neg", [(40, 3000), (200, 3000), (150, 60), (10, 100), (0, 500), (700, 2000)])
def test_subsample_regions_replays_the_global_generator(n_pos, n_neg):
    rng = np.random.default_rng(n_pos * 7 + n_neg)
    A, H, W = 9, 20, 20
    flat = rng.permutation(A * H * W)
    valid = np.zeros(A * H * W); overlap = np.zeros(A * H * W)
    valid[flat[:n_pos + n_neg]] = 1
    overlap[flat[:n_pos]] = 1
    v1 = valid.reshape(1, A, H, W).copy(); o1 = overlap.reshape(1, A, H, W).copy()
    v2, o2 = v1[0].copy(), o1[0].copy()
    np.random.seed(123)
    g = M.MT19937.from_numpy_state(np.random.get_state())
    try:
        ref = O.subsample_regions(v1, o1)                  # the oracle port that draws from np.random
        err = None
    except (KeyError, ValueError) as e:                    # the reference's own failure modes (utils.py:789-797)
        ref, err = None, type(e)
    if err is not None:
        with pytest.raises(err):
            M.subsample_regions(g, v2, o2)
        return
    got = M.subsample_regions(g, v2, o2)
    assert got == ref and np.array_equal(v1[0], v2)
    assert np.random.random_sample() == g.next_double()    # same position in the stream afterwards


def test_get_selected_samples_matches_reference_golden(golden_f3, manifest_f3):
    for case in manifest_f3["select"]:
        Y1 = S.one_hot_rows(case["seed"], case["n_pos"], case["n_neg"])
        g = M.MT19937.from_seed(case["seed"])
        sel, n_pos = M.get_selected_samples(g, Y1, case["n_rois"])
        n = "select/%s/" % case["name"]
        assert sel == golden_f3[n + "sel"].tolist(), case["name"]
        assert n_pos == int(golden_f3[n + "n_pos"])
        assert g.next_double() == float(golden_f3[n + "next_draw"]), case["name"]


def test_calc_region_props_stream_position_matches_reference_golden(golden_f3, manifest_f3):
    C = S.HotPathConfig()
    by_name = {c[0]: c for c in A3_CASES}
    for case in manifest_f3["a3_state"]:
        name, seed, width, height, n_gt, classes = by_name[case["name"]]
        img = a3_inputs(seed, width, height, n_gt, classes, small=(name == "small_gt"))
        wr, hr = O.get_new_img_size(width, height, C.img_size)
        valid, overlap, regr, ba, nh = O.rpn_targets_presample(C, img, width, height, wr, hr, S.resnet50_map_size)
        v = np.ascontiguousarray(valid.transpose(2, 0, 1)); o = np.ascontiguousarray(overlap.transpose(2, 0, 1))
        g = M.MT19937.from_seed(seed)
        M.subsample_regions(g, v, o)
        assert int(v.sum()) == int(golden_f3["a3_state/%s/n_valid" % name]), name
        assert g.next_double() == float(golden_f3["a3_state/%s/next_draw" % name]), name
