"""Drop-in for the evaluation functions of the reference `test.py`: `get_objects(pred, gt,
treshold)` (test.py:48-113) and `calc_class_ap(y_true, y_pred)` (test.py:117-173), same signatures and return
values.  The O(n_det x n_gt) IoU matching and the sort / scan of the AP run on the GPU (`radnet_match_detections`,
`radnet_class_ap`); the host only turns the lists of dicts into arrays and the result arrays back into the
reference's dicts.  `iou_matrix`-style helpers that used to launch one kernel per scalar pair are gone: callers that
need many IoUs use `utils.iou_pairs`.
"""
import numpy as np
import torch

from . import _device as D
from . import _lib


def _device():
    D.require_cuda()
    return torch.device("cuda:%d" % torch.cuda.current_device())


def match_detections(det_box, det_cls, det_prob, gt_box, gt_cls, threshold):
    """Arrays in, arrays out: (visit (n_det,) detection index visited at each rank, match (n_det,) matched figure or
    -1 at each rank).  det_box (n_det,4) / gt_box (n_gt,4) x1,y1,x2,y2; classes as int32 codes."""
    dev = _device()
    n_det, n_gt = int(len(det_prob)), int(len(gt_cls))
    if n_det == 0:
        return np.zeros((0,), np.int64), np.zeros((0,), np.int64)
    lib = _lib.load()
    db = D.to_device(np.asarray(det_box, dtype=np.float64).reshape(n_det, 4), np.float64, dev)
    dc = D.to_device(det_cls, np.int32, dev)
    dp = D.to_device(det_prob, np.float64, dev)
    gb = D.to_device(np.asarray(gt_box, dtype=np.float64).reshape(n_gt, 4), np.float64, dev) if n_gt else None
    gc = D.to_device(gt_cls, np.int32, dev) if n_gt else None
    ws_bytes = int(lib.radnet_match_detections_workspace_bytes(n_det, n_gt))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    visit = D.empty((n_det,), np.int32, dev)
    match = D.empty((n_det,), np.int32, dev)
    _lib.call("radnet_match_detections", D.ptr(db), D.ptr(dc), D.ptr(dp), n_det, D.ptr(gb), D.ptr(gc), n_gt,
              float(threshold), D.ptr(visit), D.ptr(match), D.ptr(ws), ws_bytes, D.stream_ptr(dev))
    return visit.cpu().numpy().astype(np.int64), match.cpu().numpy().astype(np.int64)


def get_objects(pred, gt, treshold):
    """test.py:48-113.  pred / gt: lists of dicts with 'class', 'x1', 'y1', 'x2', 'y2' (+ 'prob' for pred).  Sets
    gt[k]['bbox_matched'] like the reference and returns (T, P): per class, the match flags and scores in visiting
    order (descending score), then a (1, 0) pair for every unmatched figure (dict keys in first-appearance order)."""
    for g in gt:
        g['bbox_matched'] = False
    names = {}
    for d in list(pred) + list(gt):
        names.setdefault(d['class'], len(names))
    det_cls = np.array([names[d['class']] for d in pred], dtype=np.int32)
    gt_cls = np.array([names[g['class']] for g in gt], dtype=np.int32)
    det_box = np.array([[d['x1'], d['y1'], d['x2'], d['y2']] for d in pred], dtype=np.float64).reshape(-1, 4)
    gt_box = np.array([[g['x1'], g['y1'], g['x2'], g['y2']] for g in gt], dtype=np.float64).reshape(-1, 4)
    det_prob = np.array([d['prob'] for d in pred], dtype=np.float64)
    visit, match = match_detections(det_box, det_cls, det_prob, gt_box, gt_cls, treshold)
    T, P = {}, {}
    for r in range(len(visit)):
        d = pred[int(visit[r])]
        c = d['class']
        if c not in P:
            P[c], T[c] = [], []
        P[c].append(d['prob'])
        T[c].append(int(match[r] >= 0))
        if match[r] >= 0:
            gt[int(match[r])]['bbox_matched'] = True
    for g in gt:
        if not g['bbox_matched']:
            if g['class'] not in P:
                P[g['class']], T[g['class']] = [], []
            T[g['class']].append(1)
            P[g['class']].append(0)
    return T, P


def calc_class_ap(y_true, y_pred):
    """test.py:117-173: (ap, precision ndarray, recall ndarray, interpolated_precision list, interpolated_recall list)."""
    dev = _device()
    yt = np.asarray(y_true)
    yp = np.asarray(y_pred, dtype=np.float64)
    n = int(yt.shape[0])
    if n == 0:
        return 0, np.array([]), np.array([]), [], []
    lib = _lib.load()
    t_d = D.to_device(yt, np.int32, dev)
    p_d = D.to_device(yp, np.float64, dev)
    out = D.empty((4, n), np.float64, dev)
    ap = D.empty((1,), np.float64, dev)
    ws_bytes = int(lib.radnet_class_ap_workspace_bytes(n))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    _lib.call("radnet_class_ap", D.ptr(t_d), D.ptr(p_d), n, D.ptr(out[0]), D.ptr(out[1]), D.ptr(out[2]), D.ptr(out[3]),
              D.ptr(ap), D.ptr(ws), ws_bytes, D.stream_ptr(dev))
    o = out.cpu().numpy()
    return float(ap.cpu().numpy()[0]), o[0].copy(), o[1].copy(), o[2].tolist(), o[3].tolist()
