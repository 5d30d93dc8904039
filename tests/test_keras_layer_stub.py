"""CPU: with a (stub) Keras importable, RoiPoolingConv IS a keras Layer - constructor kwargs reach Layer.__init__,
build() marks it built, get_config() merges the base config, Layer.__call__ drives call() - and graph tensors are
routed through tf.numpy_function with the static output shape restored (reference RoiPoolingConv.py:8-46, 90-95;
instantiated in the graph at resnet50.py:249-252)."""
import importlib
import sys
import types

import numpy as np
import pytest


class _StubLayer:
    def __init__(self, **kwargs):
        self.name = kwargs.get("name", "layer")
        self.built = False
        self.base_init_called = True

    def build(self, input_shape):
        self.built = True

    def get_config(self):
        return {"name": self.name, "trainable": True}

    def __call__(self, inputs, **kw):
        if not self.built:
            self.build([getattr(t, "shape", None) for t in inputs])
        return self.call(inputs, **kw)


class _GraphTensor:
    """Stands for a tf.Tensor: comes from a module called `tensorflow`, carries a static shape."""
    def __init__(self, value):
        self.value = value
        self.shape = value.shape
        self.static = None

    def set_shape(self, shape):
        self.static = tuple(shape)


_GraphTensor.__module__ = "tensorflow.python.framework.ops"


@pytest.fixture
def keras_stubs(monkeypatch):
    keras = types.ModuleType("keras")
    engine = types.ModuleType("keras.engine")
    topology = types.ModuleType("keras.engine.topology")
    topology.Layer = _StubLayer
    keras.engine = engine
    engine.topology = topology
    tf = types.ModuleType("tensorflow")
    tf.float32 = "float32"
    calls = []

    def numpy_function(fn, inputs, dtype):
        calls.append((inputs, dtype))
        return _GraphTensor(np.zeros((0,), np.float32))          # the real op is deferred; shape is set by the layer
    tf.numpy_function = numpy_function
    for name, mod in (("keras", keras), ("keras.engine", engine), ("keras.engine.topology", topology), ("tensorflow", tf)):
        monkeypatch.setitem(sys.modules, name, mod)
    # the package re-exports the class under the submodule's name, so `import a.b as M` would bind the class
    M = importlib.reload(importlib.import_module("rock_art_radnet_b200.RoiPoolingConv"))
    yield M, calls
    for name in ("keras", "keras.engine", "keras.engine.topology", "tensorflow"):
        monkeypatch.delitem(sys.modules, name, raising=False)
    importlib.reload(M)


def test_roi_pooling_conv_is_a_keras_layer_when_keras_imports(keras_stubs):
    M, calls = keras_stubs
    layer = M.RoiPoolingConv(14, 20, name="roi_pool")
    assert isinstance(layer, _StubLayer) and layer.base_init_called and layer.name == "roi_pool"
    layer.build([(None, 38, 38, 1024), (None, 20, 4)])
    assert layer.built and layer.nb_channels == 1024
    assert layer.compute_output_shape([(None, 38, 38, 1024), (None, 20, 4)]) == (None, 20, 14, 14, 1024)
    assert layer.get_config() == {"name": "roi_pool", "trainable": True, "pool_size": 14, "num_rois": 20}
    img = _GraphTensor(np.zeros((1, 38, 38, 1024), np.float32))
    rois = _GraphTensor(np.zeros((1, 20, 4), np.float32))
    out = layer([img, rois])                                        # Layer.__call__ -> call -> tf.numpy_function
    assert len(calls) == 1 and calls[0][1] == "float32"
    assert out.static == (1, 20, 14, 14, 1024)


def test_roi_pooling_conv_without_keras_is_a_plain_callable():
    M = importlib.import_module("rock_art_radnet_b200.RoiPoolingConv")
    if M._LayerBase is not object:
        pytest.skip("a real Keras is installed")
    layer = M.RoiPoolingConv(7, 4)
    assert layer.get_config() == {"pool_size": 7, "num_rois": 4} and callable(layer)
