"""Training-side sampling on the device, draw for draw identical to the reference's calls into NumPy's
legacy global generator:

  * the 256-region balancing at the end of `calc_region_props` (reference faster_rcnn/utils.py:777-813)
    -> `RpnSubsampler` (`radnet_rpn_subsample`), used by `utils.calc_region_props`;
  * `get_selected_samples(Y1, C)` (reference train.py:93-129) -> `get_selected_samples` here (drop-in, uses and
    advances `np.random`'s own state) and `SampleSelector` (batched, device-resident).

A generator state on the device is uint32[625] = MT19937 key[624] + pos, the layout of
`np.random.get_state()[1:3]`; `numpy_state_words` / `set_numpy_state` convert.
"""
import numpy as np
import torch

from . import _device as D
from . import _lib

STATE_WORDS = 625


def numpy_state_words(state=None):
    """np.random.get_state() -> uint32[625] (key + pos)."""
    st = np.random.get_state() if state is None else state
    if st[0] != "MT19937":
        raise ValueError("legacy NumPy generator expected (MT19937), got %r" % (st[0],))
    out = np.empty((STATE_WORDS,), dtype=np.uint32)
    out[:624] = st[1]
    out[624] = st[2]
    return out


def set_numpy_state(words):
    """uint32[625] -> np.random.set_state (cached Gaussian cleared, as no Gaussian draw is involved)."""
    words = np.asarray(words, dtype=np.uint32)
    np.random.set_state(("MT19937", words[:624].copy(), int(words[624]), 0, 0.0))


def seed_states(seeds, device=None):
    """np.random.seed(s) for every s in `seeds` (32-bit integers) -> (B,625) uint32 CUDA tensor
    (`radnet_mt19937_seed`), one independent stream per panel."""
    D.require_cuda()
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    s = torch.from_numpy(np.ascontiguousarray(np.asarray(seeds, dtype=np.uint32)).view(np.int32)).to(dev)
    states = torch.empty((int(s.shape[0]), STATE_WORDS), dtype=torch.int32, device=dev)
    _lib.call("radnet_mt19937_seed", D.ptr(s), int(s.shape[0]), D.ptr(states), D.stream_ptr(dev))
    return states


class RpnSubsampler:
    """Batched 256-region balancing (reference utils.py:777-813) in place on the label tensor written by
    `RpnTargetBatch` (either layout).  `run(y_cls, states)` launches `radnet_rpn_subsample` on the current
    stream; states (B,625) int32/uint32 CUDA tensor, advanced in place.  Returns the resident (B,8) int32
    report: [:,0] n_pos as calc_region_props returns it, [:,3] status (1 = the reference raises KeyError)."""

    def __init__(self, batch, H, W, A, device=None, layout=0, max_regions=256):
        D.require_cuda()
        lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.B, self.H, self.W, self.A = int(batch), int(H), int(W), int(A)
        self.layout, self.max_regions = int(layout), int(max_regions)
        self.ws_bytes = int(lib.radnet_rpn_subsample_workspace_bytes(self.B, self.H, self.W, self.A))
        self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=self.device)
        self.out = torch.zeros((self.B, 8), dtype=torch.int32, device=self.device)

    def run(self, y_cls, states, out=None):
        out = self.out if out is None else out
        _lib.call("radnet_rpn_subsample", D.ptr(y_cls), self.B, self.H, self.W, self.A, self.layout,
                  self.max_regions, D.ptr(states), D.ptr(out), D.ptr(self.ws), self.ws_bytes,
                  D.stream_ptr(self.device))
        return out


class SampleSelector:
    """Batched get_selected_samples (reference train.py:93-129): `run(y_class, count, states)` with
    y_class (B,R,n_cls) int32 one-hot rows (RoiTargetBatch.y_class), count (B,) int32 or None.  Returns the
    resident tensors sel (B,n_rois) int32 and report (B,4) int32 {rows selected, n_pos, n_neg, status}."""

    def __init__(self, batch, R, n_cls, n_rois, device=None):
        D.require_cuda()
        _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.B, self.R, self.n_cls, self.n_rois = int(batch), int(R), int(n_cls), int(n_rois)
        self.sel = torch.zeros((self.B, self.n_rois), dtype=torch.int32, device=self.device)
        self.out = torch.zeros((self.B, 4), dtype=torch.int32, device=self.device)

    def run(self, y_class, count, states):
        _lib.call("radnet_select_samples", D.ptr(y_class), D.ptr(count), self.B, self.R, self.n_cls, self.n_rois,
                  D.ptr(states), D.ptr(self.sel), D.ptr(self.out), D.stream_ptr(self.device))
        return self.sel, self.out


def get_selected_samples(Y1, C):
    """Drop-in for the reference's train.get_selected_samples(Y1, C) (train.py:93-129): Y1 (1,R,n_cls)
    one-hot rows of calc_iou ('bg' last) -> (list of C.n_rois selected row indices, number of positives).
    Draws from - and advances - NumPy's global legacy generator exactly as the reference does."""
    D.require_cuda()
    dev = torch.device("cuda:%d" % torch.cuda.current_device())
    Y1 = np.asarray(Y1)
    R, n_cls = int(Y1.shape[1]), int(Y1.shape[2])
    if R == 0:       # np.random.choice([], n_rois, replace=True) (train.py:125)
        raise ValueError("'a' cannot be empty unless no samples are taken")
    y = torch.from_numpy(np.ascontiguousarray(Y1[0], dtype=np.int32)).to(dev).unsqueeze(0)
    states = torch.from_numpy(numpy_state_words().view(np.int32)).to(dev).unsqueeze(0)
    selr = SampleSelector(1, R, n_cls, int(C.n_rois), device=dev)
    sel, out = selr.run(y, None, states)
    rep = out.cpu().numpy()[0]
    set_numpy_state(states.cpu().numpy().view(np.uint32)[0])
    if rep[3] != 0:
        # empty population or negative sample count: np.random.choice raises (train.py:125)
        raise ValueError("a must be non-empty / negative dimensions are not allowed")
    return [int(v) for v in sel.cpu().numpy()[0, :int(rep[0])]], int(rep[1])
