"""TEST INFRASTRUCTURE ONLY - golden vectors for the training-side sampling from the
UNMODIFIED reference, run in the build container (needs /root/reference):

    python oracle/make_golden_f3.py

* get_selected_samples (reference train.py:93-129): train.py is a script whose imports (tensorflow, keras)
  are absent here, so the function definition is taken from its source with `ast` - the function's own code
  runs unmodified against this container's NumPy - and called after `np.random.seed(seed)`.
* calc_region_props (reference faster_rcnn/utils.py:554-821): the cases of `make_golden.A3_CASES` again, this
  time recording the generator state left behind (as the next `np.random.random_sample()`), which pins how many
  words the 256-region balancing consumed.
Every case also stores the next draw after the call, so that a replay has to leave the generator exactly where
the reference leaves it.
"""
import ast
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.make_golden import A3_CASES, a3_inputs  # noqa: E402
from oracle.reference_import import REFERENCE_ROOT, load_reference  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# (name, seed, n_pos, n_neg, n_rois)
SELECT_CASES = [
    ("typical", 0, 33, 211, 20),
    ("few_pos", 1, 3, 120, 20),
    ("exactly_half", 2, 10, 50, 20),
    ("no_pos", 3, 0, 40, 20),
    ("neg_too_few", 4, 30, 4, 20),            # replace=False raises, replace=True is used (train.py:115-118)
    ("one_neg", 5, 2, 1, 20),                 # randint over a population of one: no word consumed
    ("no_neg_many_pos", 6, 15, 0, 20),        # first selection drawn and discarded, then train.py:124-125
    ("no_neg_few_pos", 7, 4, 0, 20),
    ("n_rois_4", 8, 9, 17, 4),
    ("big", 9, 140, 160, 64),
    ("long_stream", 10, 290, 10, 300),        # more than 624 words: the generator regenerates mid-call
]


def load_get_selected_samples():
    src = open(os.path.join(REFERENCE_ROOT, "train.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_selected_samples")
    code = compile(ast.Module(body=[fn], type_ignores=[]), os.path.join(REFERENCE_ROOT, "train.py"), "exec")
    ns = {"np": np}
    exec(code, ns)
    return ns["get_selected_samples"]


def main():
    rpn, utils, config = load_reference()
    get_selected_samples = load_get_selected_samples()
    out, manifest = {}, {"numpy": np.__version__, "select": [], "a3_state": []}
    for name, seed, n_pos, n_neg, n_rois in SELECT_CASES:
        Y1 = S.one_hot_rows(seed, n_pos, n_neg)
        C = S.HotPathConfig()
        C.n_rois = n_rois
        np.random.seed(seed)
        sel, npos = get_selected_samples(Y1, C)
        out["select/%s/sel" % name] = np.asarray(sel, dtype=np.int64)
        out["select/%s/n_pos" % name] = np.int64(npos)
        out["select/%s/next_draw" % name] = np.float64(np.random.random_sample())
        manifest["select"].append({"name": name, "seed": seed, "n_pos": n_pos, "n_neg": n_neg, "n_rois": n_rois})
    C = S.HotPathConfig()
    for name, seed, width, height, n_gt, classes in A3_CASES:
        img = a3_inputs(seed, width, height, n_gt, classes, small=(name == "small_gt"))
        wr, hr = utils.get_new_img_size(width, height, C.img_size)
        np.random.seed(seed)
        y_cls, y_regr, best, n_pos = utils.calc_region_props(C, img, width, height, wr, hr, S.resnet50_map_size)
        out["a3_state/%s/next_draw" % name] = np.float64(np.random.random_sample())
        out["a3_state/%s/n_valid" % name] = np.int64(y_cls[0, :y_cls.shape[1] // 2].sum())
        manifest["a3_state"].append({"name": name, "seed": seed})
    np.savez_compressed(os.path.join(GOLDEN, "f3_sampling.npz"), **out)
    with open(os.path.join(GOLDEN, "manifest_f3.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("f3_sampling.npz", os.path.getsize(os.path.join(GOLDEN, "f3_sampling.npz")))


if __name__ == "__main__":
    main()
