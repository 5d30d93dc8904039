"""TEST INFRASTRUCTURE ONLY - loader for the *unmodified* reference modules.

The reference (`/root/reference`, read-only) is pure Python; its hot-path
functions (`faster_rcnn/rpn.py`, `faster_rcnn/utils.py`) run under this
container's NumPy once the absent third-party imports (`keras`, `skimage`) are
stubbed in `sys.modules`.  This loader exists for two jobs, both done in the
build container only (the GPU box has no `/root/reference`):

  * `oracle/make_golden.py` imports the reference through it to generate the
    committed golden vectors under `tests/golden/`;
  * `tests/test_oracle_vs_reference.py` validates the NumPy restatement in
    `oracle/radnet_oracle.py` against the real functions (skipped when the
    reference tree is absent).

Nothing in the product package imports this file.
"""
import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")      # oracle/build_ref.py
REFERENCE_ROOT = os.environ.get("RADNET_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "faster_rcnn", "rpn.py")) and \
        os.path.isfile(os.path.join(_STAGED, "faster_rcnn", "rpn.py")):
    REFERENCE_ROOT = _STAGED             # the GPU box: only the staged copy of the hot-path modules exists


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "faster_rcnn", "rpn.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load_reference():
    """Return (rpn, utils, config) modules of the reference, imported unmodified."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    # rpn.py:9 does `from keras.layers import Conv2D`; utils.py:17 imports
    # augmentation.py, which imports skimage (augmentation.py:10).
    keras = _stub("keras")
    layers = _stub("keras.layers", Conv2D=object)
    keras.layers = layers
    sk = _stub("skimage")
    sk.exposure = _stub("skimage.exposure")
    sk.util = _stub("skimage.util", random_noise=None)
    sk.transform = _stub("skimage.transform")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    rpn = importlib.import_module("faster_rcnn.rpn")
    utils = importlib.import_module("faster_rcnn.utils")
    config = importlib.import_module("faster_rcnn.config")
    return rpn, utils, config


def load_radnet():
    """Return the reference module `faster_rcnn.RADNet` (class RADNet), imported unmodified.
    RADNet.py:14-15 imports `keras.layers.Input` and `keras.models.Model`, stubbed here; cv2,
    pandas and tqdm are present in the build container."""
    load_reference()
    layers = sys.modules["keras.layers"]
    if not hasattr(layers, "Input"):
        layers.Input = object
    models = _stub("keras.models", Model=object)
    sys.modules["keras"].models = models
    return importlib.import_module("faster_rcnn.RADNet")
