import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _load_npz(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_a1():
    return _load_npz("a1_rpn_to_roi.npz")


@pytest.fixture(scope="session")
def golden_a2():
    return _load_npz("a2_nms.npz")


@pytest.fixture(scope="session")
def golden_a3():
    return _load_npz("a3_calc_region_props.npz")


@pytest.fixture(scope="session")
def golden_a4():
    return _load_npz("a4_calc_iou.npz")


def dense_regr(g, name):
    """Rebuild the dense y_rpn_regr of an a3 golden case from its sparse storage."""
    shape = tuple(int(v) for v in g[name + "/y_rpn_regr_shape"])
    out = np.zeros(int(np.prod(shape)))
    out[g[name + "/y_rpn_regr_nz_idx"]] = g[name + "/y_rpn_regr_nz_val"]
    return out.reshape(shape)


@pytest.fixture
def lib_option():
    """lib_option(name, value): set a tuning option of libradnet_b200 for the duration of one test."""
    from rock_art_radnet_b200 import _lib
    saved = {}

    def setter(name, value):
        saved.setdefault(name, _lib.get_option(name))
        _lib.set_option(name, value)

    yield setter
    for name, value in saved.items():
        _lib.set_option(name, value)
