"""Drop-in for the reference Keras layer `faster_rcnn/RoiPoolingConv.py`.

Same constructor, `call([img, rois])`, `compute_output_shape` and `get_config`; the
per-RoI crop + TF-1 legacy bilinear resize runs in `radnet_roi_pool` on the GPU
for all RoIs in one launch instead of `num_rois` tiny TF sub-graphs.
"""
import numpy as np
import torch

from . import _device as D
from . import _lib


def _keras_layer_base():
    """`keras.layers.Layer` (the reference subclasses `keras.engine.topology.Layer`, RoiPoolingConv.py:6-8) when a
    Keras is importable, else `object`: without Keras the class is a plain callable with the same methods."""
    for mod, attr in (("keras.engine.topology", "Layer"), ("keras.layers", "Layer"), ("tensorflow.keras.layers", "Layer")):
        try:
            m = __import__(mod, fromlist=[attr])
            base = getattr(m, attr)
            if isinstance(base, type):
                return base
        except Exception:
            continue
    return object


_LayerBase = _keras_layer_base()


def _is_graph_tensor(x):
    """A symbolic / eager tensor of TensorFlow (anything from a `tensorflow` module that is neither NumPy nor torch)."""
    return type(x).__module__.split(".")[0] in ("tensorflow", "keras") and not isinstance(x, np.ndarray)


def roi_pool_device(feat, rois, roi_count, pool_size, out=None):
    """feat (B,H,W,C) float32 CUDA, rois (B,R,4) int32 CUDA as (x,y,w,h), roi_count (B,) int32
    CUDA or None -> (B,R,pool,pool,C) float32 CUDA.  Asynchronous."""
    B, H, W, C = (int(v) for v in feat.shape)
    R = int(rois.shape[1])
    if out is None:
        out = D.empty((B, R, pool_size, pool_size, C), np.float32, feat.device)
    _lib.call("radnet_roi_pool", D.ptr(feat), B, H, W, C, None, 0, D.ptr(rois), D.ptr(roi_count), R,
              int(pool_size), D.ptr(out), D.stream_ptr(feat.device))
    return out


class RoiPoolingConv(_LayerBase):
    """ROI pooling layer for 2D inputs (reference RoiPoolingConv.py:8-95).  A `keras.layers.Layer` when Keras is
    importable (so it can sit inside the classifier graph exactly where the reference instantiates it,
    resnet50.py:249-252 / vgg16.py:85); called on graph tensors it runs the CUDA kernel through
    `tf.numpy_function` with the static output shape restored.

    # Arguments
        pool_size: int, side of the pooled output (14 for ResNet-50, 7 for VGG-16).
        num_rois: number of regions of interest per call.
    # Input
        [X_img (1,rows,cols,channels) float32, X_roi (1,num_rois,4) as (x,y,w,h)]
        NumPy arrays or CUDA tensors.
    # Output
        (1,num_rois,pool_size,pool_size,channels) float32; NumPy in -> NumPy out,
        CUDA tensor in -> CUDA tensor out.
    """

    def __init__(self, pool_size, num_rois, **kwargs):
        self.pool_size = pool_size
        self.num_rois = num_rois
        self.nb_channels = None
        if _LayerBase is not object:
            super().__init__(**kwargs)                              # RoiPoolingConv.py:38

    def build(self, input_shape):
        self.nb_channels = input_shape[0][3]                       # RoiPoolingConv.py:42
        if _LayerBase is not object:
            super().build(input_shape)

    def compute_output_shape(self, input_shape):
        nb = self.nb_channels if self.nb_channels is not None else input_shape[0][3]
        return None, self.num_rois, self.pool_size, self.pool_size, nb   # RoiPoolingConv.py:45-46

    def _call_graph(self, img, rois):
        """Graph / eager TensorFlow tensors: the kernel runs as a `tf.numpy_function`; the static shape
        (1, num_rois, pool, pool, channels) that `compute_output_shape` promises is set on the result."""
        import tensorflow as tf
        fn = getattr(tf, "numpy_function", None) or tf.py_func
        out = fn(lambda a, b: self.call([np.asarray(a), np.asarray(b)]), [img, rois], tf.float32)
        nb = self.nb_channels if self.nb_channels is not None else img.shape[3]
        out.set_shape((1, self.num_rois, self.pool_size, self.pool_size, nb))
        return out

    def call(self, x, mask=None):
        assert (len(x) == 2)                                       # RoiPoolingConv.py:50
        img, rois = x[0], x[1]
        if _is_graph_tensor(img):
            return self._call_graph(img, rois)
        D.require_cuda()
        on_device = D.is_cuda_tensor(img)
        dev = img.device if on_device else torch.device("cuda:%d" % torch.cuda.current_device())
        if tuple(img.shape[:1]) != (1,) or len(img.shape) != 4:
            raise ValueError("RoiPoolingConv expects X_img of shape (1, rows, cols, channels)")
        H, W = int(img.shape[1]), int(img.shape[2])
        self.nb_channels = int(img.shape[3])
        # K.cast(.., 'int32') truncates toward zero (RoiPoolingConv.py:69-72)
        r = rois.detach().cpu().numpy() if isinstance(rois, torch.Tensor) else np.asarray(rois)
        if r.shape[0] != 1 or r.shape[1] < self.num_rois or r.shape[2] != 4:
            raise ValueError("RoiPoolingConv expects X_roi of shape (1, num_rois, 4)")
        r = np.trunc(r[:, :self.num_rois, :]).astype(np.int32)
        x0, y0, w0, h0 = r[0, :, 0], r[0, :, 1], r[0, :, 2], r[0, :, 3]
        if (x0 < 0).any() or (y0 < 0).any():
            raise ValueError("RoiPoolingConv: negative RoI origin is outside the supported domain")
        ch = np.minimum(y0 + h0, H) - np.minimum(y0, H)            # TF slice end clamps to the map
        cw = np.minimum(x0 + w0, W) - np.minimum(x0, W)
        if (ch <= 0).any() or (cw <= 0).any():
            raise ValueError("RoiPoolingConv: RoI crops to an empty region (TF resize would fail)")
        feat = D.to_device(img, np.float32, dev)
        out = roi_pool_device(feat, torch.from_numpy(r).to(dev), None, self.pool_size)
        # (1,num_rois,pool,pool,C); the reference's final permute is the identity (RoiPoolingConv.py:86)
        return out if on_device else out.cpu().numpy()

    if _LayerBase is object:
        __call__ = call                                            # without Keras the instance itself is the op

    def get_config(self):
        config = {'pool_size': self.pool_size, 'num_rois': self.num_rois}
        if _LayerBase is not object:
            base_config = super().get_config()                     # RoiPoolingConv.py:90-95
            return dict(list(base_config.items()) + list(config.items()))
        return config
