"""Detection post-processing on the device: classifier-head decode, per-class NMS, tile merge.

Thin host side of `csrc/detect.cu` (C ABI: `radnet_classify_decode`, `radnet_classify_nms`,
`radnet_class_nms`, `radnet_final_nms`).  The unit that moves between these calls - and between
ranks, when the tiles of one panel were processed on different GPUs - is the labelled detection
record described in `include/radnet_b200.h`; `ClassRecords` gives typed views of a batch of them.

Reference functions taken over:
  RADNet.apply_spatial_pyramid_pooling decode loop   faster_rcnn/RADNet.py:123-150
  rpn.apply_regr                                     faster_rcnn/rpn.py:346-378
  per-class / cross-image NMS + get_real_coordinates faster_rcnn/RADNet.py:44-51, 570-600, 695-716
  RADNet.final_nms                                   faster_rcnn/RADNet.py:156-240
"""
import numpy as np
import torch

from . import _device as D
from . import _lib

MAX_CLASSES = 32
HEADER_WORDS = 72            # int32 header[8] + class_order[32] + class_count[32]

H_NDET, H_NIN, H_NDEGEN, H_NTIES, H_NCLASSES, H_NFALLBACK, H_NNEARTIE, H_NRANGE = range(8)

ENTRY_DTYPE = np.dtype([("cls", "<i4"), ("prob", "<f4"), ("box", "<i4", (4,)), ("src", "<i4"), ("aux", "<i4")])


def record_bytes(max_det):
    return int(_lib.load().radnet_cls_record_bytes(int(max_det)))


def _record_dtype(max_det):
    return np.dtype([("header", "<i4", (8,)), ("order", "<i4", (MAX_CLASSES,)), ("count", "<i4", (MAX_CLASSES,)),
                     ("entry", ENTRY_DTYPE, (max_det,))])


class ClassRecords:
    """`n` labelled detection records of capacity `max_det` in one CUDA uint8 tensor (n, stride)."""

    def __init__(self, n, max_det, device, raw=None):
        self.n, self.max_det = int(n), int(max_det)
        self.stride = record_bytes(max_det)
        self.raw = raw if raw is not None else torch.zeros((self.n, self.stride), dtype=torch.uint8, device=device)
        assert tuple(self.raw.shape) == (self.n, self.stride)

    @property
    def header(self):
        """(n, 8) int32 view of the headers (device)."""
        return self.raw.view(torch.int32)[:, :8]

    def to_numpy(self):
        """Structured host copy: fields header (8,), order (32,), count (32,), entry (max_det,)."""
        host = self.raw.cpu().numpy()
        return host.view(_record_dtype(self.max_det)).reshape(self.n)

    @staticmethod
    def from_arrays(groups, max_det, device):
        """Build records on the device from host data: `groups` is a list (one per record) of
        (cls (k,), prob (k,), boxes (k,4)) triples in entry order."""
        rec = np.zeros((len(groups),), dtype=_record_dtype(max_det))
        for r, (cls, prob, boxes) in enumerate(groups):
            cls = np.asarray(cls, dtype=np.int32).reshape(-1)
            k = len(cls)
            if k > max_det:
                raise ValueError("record capacity %d < %d entries" % (max_det, k))
            rec["header"][r, H_NDET] = k
            rec["header"][r, H_NIN] = k
            rec["entry"]["cls"][r, :k] = cls
            rec["entry"]["prob"][r, :k] = np.asarray(prob, dtype=np.float32).reshape(-1)
            rec["entry"]["box"][r, :k] = np.asarray(boxes, dtype=np.int64).reshape(-1, 4)
            rec["entry"]["src"][r, :k] = np.arange(k)
            seen = []
            for c in cls:
                if c not in seen:
                    seen.append(int(c))
            rec["order"][r] = -1
            rec["order"][r, :len(seen)] = seen
            for c in seen:
                rec["count"][r, c] = int((cls == c).sum())
        raw = torch.from_numpy(rec.view(np.uint8).reshape(len(groups), -1).copy()).to(device)
        return ClassRecords(len(groups), max_det, device, raw=raw)


def check_records(host_records, what):
    """Raise what the reference would have raised for the conditions the kernels only count."""
    hdr = host_records["header"]
    if (hdr[:, H_NDET] == -1).any():
        raise RuntimeError("%s: an input record carried a fault" % what)
    if (hdr[:, H_NDET] == -2).any():
        raise _lib.RadnetError(-4, "RADNET_E_UNSUPPORTED", "%s: more boxes in one segment than the kernel holds" % what)
    if (hdr[:, H_NRANGE] > 0).any():
        raise OverflowError("%s: box coordinate beyond 2**25 (the reference would build Python big ints)" % what)
    if (hdr[:, H_NDEGEN] > 0).any():
        # np.testing.assert_array_less(x1, x2) in the reference's NMS (rpn.py:400-401, RADNet.py:176-177)
        raise AssertionError("\nArrays are not strictly ordered `x < y` (%s: %d degenerate boxes)"
                             % (what, int(hdr[:, H_NDEGEN].sum())))


def _head_args(P_cls, P_regr, rois, roi_count, det, C, dev):
    P_cls = D.to_device(P_cls, np.float32, dev)
    P_regr = D.to_device(P_regr, np.float32, dev)
    if P_cls.dim() != 3 or P_regr.dim() != 3:
        raise ValueError("P_cls / P_regr must be (B, R, n_cls) / (B, R, 4*(n_cls-1))")
    B, R, n_cls = (int(v) for v in P_cls.shape)
    if tuple(P_regr.shape) != (B, R, 4 * (n_cls - 1)):
        raise ValueError("P_regr shape %s does not match P_cls %s" % (tuple(P_regr.shape), tuple(P_cls.shape)))
    stride = int(C.rpn_stride)
    if stride != C.rpn_stride:
        raise ValueError("rpn_stride must be an integer")
    det_raw, det_k = None, 0
    if det is not None:
        det_raw, det_k = det.raw, det.max_boxes
        if det.batch != B:
            raise ValueError("detection records: batch %d != %d" % (det.batch, B))
    else:
        rois = D.to_device(rois, np.int32, dev)
        if tuple(rois.shape) != (B, R, 4):
            raise ValueError("rois must be (B, R, 4) xywh; got %s" % (tuple(rois.shape),))
        if roi_count is not None:
            roi_count = D.to_device(roi_count, np.int32, dev)
    std = D.host_f64(C.classifier_regr_std)
    return P_cls, P_regr, B, R, n_cls, det_raw, det_k, rois, roi_count, std, stride


def classify_decode(P_cls, P_regr, C, rois=None, roi_count=None, det=None, bbox_threshold=0.7, out=None):
    """Per-RoI class decision + box decode for B tiles (RADNet.py:123-150).  Returns ClassRecords
    (B, capacity R) with entries in RoI order.  Asynchronous."""
    D.require_cuda()
    dev = torch.device("cuda:%d" % torch.cuda.current_device())
    P_cls, P_regr, B, R, n_cls, det_raw, det_k, rois, roi_count, std, stride = _head_args(
        P_cls, P_regr, rois, roi_count, det, C, dev)
    out = out if out is not None else ClassRecords(B, R, dev)
    _lib.call("radnet_classify_decode", D.ptr(P_cls), D.ptr(P_regr), B, R, n_cls, D.ptr(det_raw), det_k,
              D.ptr(rois), D.ptr(roi_count), float(bbox_threshold), D.ptr(std), stride, D.ptr(out.raw),
              out.max_det, D.stream_ptr(dev))
    return out


def classify_nms(P_cls, P_regr, C, rois=None, roi_count=None, det=None, bbox_threshold=0.7, nms_thresh=0.2,
                 max_boxes=300, ratio=None, origin=None, out=None):
    """Decode + per-class NMS + real coordinates + tile offset in one launch (RADNet.py:566-600).
    ratio (B,) float64 and origin (B,2) int32 are optional.  Returns ClassRecords (B, capacity R)."""
    D.require_cuda()
    dev = torch.device("cuda:%d" % torch.cuda.current_device())
    P_cls, P_regr, B, R, n_cls, det_raw, det_k, rois, roi_count, std, stride = _head_args(
        P_cls, P_regr, rois, roi_count, det, C, dev)
    ratio_d = _ratio_tensor(ratio, B, dev)
    origin_d = D.to_device(origin, np.int32, dev).reshape(B, 2) if origin is not None else None
    out = out if out is not None else ClassRecords(B, R, dev)
    if out.n != B or out.max_det < 1:
        raise ValueError("output records: %d records for %d tiles" % (out.n, B))
    _lib.call("radnet_classify_nms", D.ptr(P_cls), D.ptr(P_regr), B, R, n_cls, D.ptr(det_raw), det_k,
              D.ptr(rois), D.ptr(roi_count), float(bbox_threshold), D.ptr(std), stride, float(nms_thresh),
              int(max_boxes), D.ptr(ratio_d), D.ptr(origin_d), D.ptr(out.raw), out.max_det, D.stream_ptr(dev))
    return out


def _ratio_tensor(ratio, n, dev):
    if ratio is None:
        return None
    if D.is_cuda_tensor(ratio):          # resident (batched pipeline): the caller vouches for positive finite values
        assert ratio.dtype == torch.float64 and ratio.numel() == n
        return ratio.contiguous()
    r = np.asarray(ratio, dtype=np.float64).reshape(n)
    if not (np.isfinite(r).all() and (r > 0).all()):
        raise ZeroDivisionError("resize ratio must be finite and positive")
    return torch.from_numpy(r).to(dev)


def class_nms(records, n_segments, n_in, n_cls, thresh, max_boxes=300, ratio=None, origin=None, in_count=None,
              out_max_det=None, out=None, ws=None):
    """Per-class greedy NMS over the concatenation of `n_in` consecutive records per segment
    (rpn.py:380-456 applied per class: RADNet.py:574, 639, 698).  Returns ClassRecords (n_segments,)."""
    dev = records.raw.device
    assert records.n == n_segments * n_in
    cap = int(out_max_det) if out_max_det is not None else n_in * records.max_det
    out = out if out is not None else ClassRecords(n_segments, cap, dev)
    lib = _lib.load()
    ws_bytes = int(lib.radnet_class_nms_workspace_bytes(n_segments, n_in, records.max_det, n_cls))
    if ws is None or ws.numel() < ws_bytes:
        ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    ratio_d = _ratio_tensor(ratio, n_segments, dev)
    origin_d = D.to_device(origin, np.int32, dev).reshape(n_segments, 2) if origin is not None else None
    cnt_d = D.to_device(in_count, np.int32, dev) if in_count is not None else None
    _lib.call("radnet_class_nms", D.ptr(records.raw), records.max_det, n_segments, n_in, D.ptr(cnt_d), int(n_cls),
              float(thresh), int(max_boxes), D.ptr(ratio_d), D.ptr(origin_d), D.ptr(out.raw), out.max_det,
              D.ptr(ws), ws_bytes, D.stream_ptr(dev))
    return out


def final_nms_records(records, n_segments, n_in, n_cls, obj_avg_threshold=0.2, obj_confidence_threshold=0.8,
                      n_obj_avg=5, in_count=None, out_max_det=None, out=None, ws=None):
    """Cluster-and-average merge (RADNet.final_nms, RADNet.py:156-240) of `n_in` consecutive tile
    records per image, every class at once.  Returns ClassRecords (n_segments,)."""
    dev = records.raw.device
    assert records.n == n_segments * n_in
    cap = int(out_max_det) if out_max_det is not None else n_in * records.max_det
    out = out if out is not None else ClassRecords(n_segments, cap, dev)
    lib = _lib.load()
    ws_bytes = int(lib.radnet_final_nms_workspace_bytes(n_segments, n_in, records.max_det, n_cls))
    if ws is None or ws.numel() < ws_bytes:
        ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    cnt_d = D.to_device(in_count, np.int32, dev) if in_count is not None else None
    _lib.call("radnet_final_nms", D.ptr(records.raw), records.max_det, n_segments, n_in, D.ptr(cnt_d), int(n_cls),
              float(obj_avg_threshold), float(obj_confidence_threshold), int(n_obj_avg), D.ptr(out.raw),
              out.max_det, D.ptr(ws), ws_bytes, D.stream_ptr(dev))
    return out


def record_to_dicts(rec, class_names):
    """One host record -> (bboxes, probs) dicts keyed by class name in first-appearance order, the
    shape in which the reference carries detections around (RADNet.py:100-152)."""
    n = int(rec["header"][H_NDET])
    ent = rec["entry"][:n]
    bboxes, probs = {}, {}
    for c in rec["order"]:
        if c < 0:
            break
        sel = ent["cls"] == c
        name = class_names[int(c)]
        bboxes[name] = ent["box"][sel].astype(np.int64)
        probs[name] = ent["prob"][sel].copy()
    return bboxes, probs
