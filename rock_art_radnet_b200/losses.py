"""Drop-in for the reference `faster_rcnn/losses.py`: the same four names; each returns the
loss value (float32) the reference's Keras graph computes for `y_true`, `y_pred`, evaluated on the GPU by
`radnet_rpn_losses` / `radnet_class_losses` (fused masked reductions over the target layouts of K3 / a4).

Parity is UNPINNED against Keras / TensorFlow (not installable here): the kernels follow the public definitions of the
Keras-2.2 / TF-1 backend calls that losses.py composes, element by element in float32; bar 1e-5 relative.
`RpnLossBatch` / `class_losses_device` are the batched, device-resident forms used after K3 / a4.
"""
import numpy as np
import torch

from . import _device as D
from . import _lib

lambda_rpn_regr = 1.0
lambda_rpn_class = 1.0
lambda_cls_regr = 1.0
lambda_cls_class = 1.0
epsilon = 1e-4


class RpnLossBatch:
    """rpn_loss_cls and rpn_loss_regr (losses.py:16-67) for B panels: `run(y_rpn_cls, y_rpn_regr, p_cls, p_regr)` with
    the NHWC float64 targets of `RpnTargetBatch(layout=LAYOUT_NHWC, regr_scale=C.std_scaling)` (after the subsampler)
    and the float32 RPN outputs p_cls (B,H,W,A), p_regr (B,H,W,4A).  Returns the resident (B,2) float32 tensor
    [:,0] = rpn_loss_cls, [:,1] = rpn_loss_regr."""

    def __init__(self, batch, H, W, A, device=None):
        D.require_cuda()
        lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.B, self.H, self.W, self.A = int(batch), int(H), int(W), int(A)
        self.ws_bytes = int(lib.radnet_rpn_losses_workspace_bytes(self.B))
        self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=self.device)
        _lib.call("radnet_rpn_losses_workspace_init", D.ptr(self.ws), self.ws_bytes, self.B, D.stream_ptr(self.device))
        self.loss = torch.zeros((self.B, 2), dtype=torch.float32, device=self.device)

    def run(self, y_rpn_cls, y_rpn_regr, p_cls, p_regr):
        _lib.call("radnet_rpn_losses", D.ptr(y_rpn_cls), D.ptr(y_rpn_regr), D.ptr(p_cls), D.ptr(p_regr), self.B,
                  self.H, self.W, self.A, D.ptr(self.loss), D.ptr(self.ws), self.ws_bytes, D.stream_ptr(self.device))
        return self.loss


def class_losses_device(y_class, y_regr, p_cls, p_regr, sel=None, n_sel_per_panel=None, out=None):
    """class_loss_cls and class_loss_regr (losses.py:70-95) for B panels.  y_class (B,R,n_cls) int32 and y_regr
    (B,R,8(n_cls-1)) float64 as `RoiTargetBatch` writes them; sel (B,n_sel) int32 row indices (`SampleSelector`)
    or None; p_cls (B,n_sel,n_cls), p_regr (B,n_sel,4(n_cls-1)) float32.  Returns (B,2) float32 CUDA tensor."""
    B, R, n_cls = (int(v) for v in y_class.shape)
    n_sel = int(p_cls.shape[1])
    out = out if out is not None else torch.zeros((B, 2), dtype=torch.float32, device=y_class.device)
    _lib.call("radnet_class_losses", D.ptr(y_class), D.ptr(y_regr), D.ptr(sel), D.ptr(n_sel_per_panel), B, R, n_cls,
              n_sel, D.ptr(p_cls), D.ptr(p_regr), D.ptr(out), D.stream_ptr(y_class.device))
    return out


def _dev():
    D.require_cuda()
    return torch.device("cuda:%d" % torch.cuda.current_device())


def _rpn_pair(y_true, y_pred, num_anchors, which):
    """(1,H,W,8A) / (1,H,W,2A) targets + the matching prediction -> one of the two RPN losses."""
    dev = _dev()
    y_true = np.asarray(y_true)
    y_pred = np.asarray(y_pred)
    n, H, W = (int(v) for v in y_true.shape[:3])
    A = int(num_anchors)
    # Keras feeds float32: the double -> float32 cast of the kernel is the same rounding
    if which == "cls":
        yc = D.to_device(y_true, np.float64, dev)
        yr = torch.zeros((n, H, W, 8 * A), dtype=torch.float64, device=dev)
        pc = D.to_device(y_pred, np.float32, dev)
        pr = torch.zeros((n, H, W, 4 * A), dtype=torch.float32, device=dev)
    else:
        yc = torch.zeros((n, H, W, 2 * A), dtype=torch.float64, device=dev)
        yr = D.to_device(y_true, np.float64, dev)
        pc = torch.full((n, H, W, A), 0.5, dtype=torch.float32, device=dev)
        pr = D.to_device(y_pred, np.float32, dev)
    res = RpnLossBatch(n, H, W, A, device=dev).run(yc, yr, pc, pr).cpu().numpy()
    if n != 1:
        raise ValueError("the reference trains with batch size 1 (K.sum over the whole batch); use RpnLossBatch "
                         "for per-panel losses of a batch")
    return np.float32(res[0, 0 if which == "cls" else 1])


def rpn_loss_regr(num_anchors):
    """losses.py:16-44: smooth-L1 over the positive anchors' regression targets."""
    def rpn_loss_regr_fixed_num(y_true, y_pred):
        return np.float32(lambda_rpn_regr) * _rpn_pair(y_true, y_pred, num_anchors, "regr")
    return rpn_loss_regr_fixed_num


def rpn_loss_cls(num_anchors):
    """losses.py:47-67: binary cross-entropy over the valid anchors (argument order of the reference kept)."""
    def rpn_loss_cls_fixed_num(y_true, y_pred):
        return np.float32(lambda_rpn_class) * _rpn_pair(y_true, y_pred, num_anchors, "cls")
    return rpn_loss_cls_fixed_num


def _class_pair(y_true_cls, y_true_regr, p_cls, p_regr, num_classes):
    dev = _dev()
    if y_true_cls is None:          # regression loss only
        yr = np.asarray(y_true_regr)
        n = int(yr.shape[1])
        y_class = torch.zeros((1, n, num_classes + 1), dtype=torch.int32, device=dev)
        pc = torch.full((1, n, num_classes + 1), 1.0 / (num_classes + 1), dtype=torch.float32, device=dev)
        out = class_losses_device(y_class, D.to_device(yr, np.float64, dev), pc, D.to_device(p_regr, np.float32, dev))
        return np.float32(out.cpu().numpy()[0, 1])
    yc = np.asarray(y_true_cls)
    n, n_cls = int(yc.shape[1]), int(yc.shape[2])
    yr = torch.zeros((1, n, 8 * (n_cls - 1)), dtype=torch.float64, device=dev)
    pr = torch.zeros((1, n, 4 * (n_cls - 1)), dtype=torch.float32, device=dev)
    out = class_losses_device(D.to_device(yc, np.int32, dev), yr, D.to_device(p_cls, np.float32, dev), pr)
    return np.float32(out.cpu().numpy()[0, 0])


def class_loss_regr(num_classes):
    """losses.py:70-88: smooth-L1 over the regression block of the matched class (num_classes excludes 'bg')."""
    def class_loss_regr_fixed_num(y_true, y_pred):
        return np.float32(lambda_cls_regr) * _class_pair(None, y_true, None, y_pred, num_classes)
    return class_loss_regr_fixed_num


def class_loss_cls(y_true, y_pred):
    """losses.py:93-95: mean categorical cross-entropy over the rows of the first (only) image.  y_true must hold
    integer-valued one-hot rows, as calc_iou builds them."""
    yt = np.asarray(y_true)
    if not np.array_equal(yt, np.rint(yt)):
        raise ValueError("class_loss_cls: y_true must be the one-hot rows of calc_iou")
    return np.float32(lambda_cls_class) * _class_pair(yt[:1], None, np.asarray(y_pred)[:1], None, int(yt.shape[2]) - 1)
