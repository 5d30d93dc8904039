"""TEST INFRASTRUCTURE ONLY - golden vectors for the mAP evaluation from the UNMODIFIED reference
functions `get_objects` and `calc_class_ap` (reference test.py:48-173), run in the build container:

    python oracle/make_golden_f4.py

test.py is a script whose imports (tensorflow, keras, matplotlib) are absent here, so the two function definitions
are taken from its source with `ast` and executed against this container's NumPy and the reference's own `iou`
(faster_rcnn/utils.py:77-109).  Inputs are regenerated from seeds by `synthetic.eval_set`.
"""
import ast
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_import import REFERENCE_ROOT, load_reference  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# (name, seed, n_gt, n_det, ties)
EVAL_CASES = [
    ("small", 0, 12, 30, False),
    ("typical", 1, 80, 260, False),
    ("no_det", 2, 9, 0, False),
    ("no_gt", 3, 0, 25, False),
    ("crowded", 4, 300, 1500, False),
]


def load_test_functions(utils):
    path = os.path.join(os.path.dirname(REFERENCE_ROOT.rstrip("/")) if not os.path.isfile(os.path.join(REFERENCE_ROOT, "test.py"))
                        else REFERENCE_ROOT, "test.py")
    src = open(path).read()
    tree = ast.parse(src)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("get_objects", "calc_class_ap")]
    ns = {"np": np, "iou": utils.iou}
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
    return ns["get_objects"], ns["calc_class_ap"]


def main():
    rpn, utils, config = load_reference()
    get_objects, calc_class_ap = load_test_functions(utils)
    out, manifest = {}, {"numpy": np.__version__, "eval": []}
    for name, seed, n_gt, n_det, ties in EVAL_CASES:
        det, gt = S.eval_set(seed, n_gt, n_det, ties=ties)
        T, P = get_objects(det, gt, 0.5)
        keys = list(T.keys())
        out["%s/keys" % name] = np.array(keys, dtype="U16")
        out["%s/matched" % name] = np.array([int(g['bbox_matched']) for g in gt], dtype=np.int8)
        aps = []
        for k in keys:
            out["%s/T/%s" % (name, k)] = np.array(T[k], dtype=np.int64)
            out["%s/P/%s" % (name, k)] = np.array(P[k], dtype=np.float64)
            ap, prec, rec, ip, ir = calc_class_ap(T[k], P[k])
            out["%s/ap/%s" % (name, k)] = np.float64(ap)
            out["%s/prec/%s" % (name, k)] = np.asarray(prec, dtype=np.float64)
            out["%s/rec/%s" % (name, k)] = np.asarray(rec, dtype=np.float64)
            out["%s/iprec/%s" % (name, k)] = np.asarray(ip, dtype=np.float64)
            out["%s/irec/%s" % (name, k)] = np.asarray(ir, dtype=np.float64)
            aps.append(ap)
        out["%s/mAP" % name] = np.float64(np.mean(np.array(aps))) if aps else np.float64(np.nan)
        manifest["eval"].append({"name": name, "seed": seed, "n_gt": n_gt, "n_det": n_det, "ties": ties})
    np.savez_compressed(os.path.join(GOLDEN, "f4_eval.npz"), **out)
    with open(os.path.join(GOLDEN, "manifest_f4.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("f4_eval.npz", os.path.getsize(os.path.join(GOLDEN, "f4_eval.npz")))


if __name__ == "__main__":
    main()
