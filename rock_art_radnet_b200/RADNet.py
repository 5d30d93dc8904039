"""Drop-in for the detection post-processing of the reference class `faster_rcnn/RADNet.py`.

`RADNet(C, model_rpn, model_detector, preprocess_func)` keeps the reference's constructor,
attributes and the methods that sit on or right after the hot path:

  get_real_coordinates                                                       RADNet.py:44-51
  format_img_size / format_img_channels / format_img                         (OpenCV on the host, as the reference)
  apply_spatial_pyramid_pooling(R, feature_map) -> (bboxes, probs)           RADNet.py:98-152
  final_nms(boxes, probs, ...) -> (boxes, probs)                             RADNet.py:156-240
  predict(images) -> list of detection dicts                                 RADNet.py:502-718
  tile_grid(width, height)                                                   RADNet.py:511-540

The two networks stay what they are in the reference - objects with a Keras-style
`predict` - and are out of scope; everything between and after their calls (proposal decode +
NMS, head decode, per-class NMS, real coordinates, tile merge, cross-image NMS) runs in
libradnet_b200.so.  Detections of a tile never come back to the host before the final result:
each tile leaves one labelled detection record in HBM and the merge kernels read those.

Training, data loading, evaluation (`get_map`, `predict_region_proposals`' plotting) are not
part of the path and are not provided.
"""
import numpy as np
import torch

from . import _device as D
from . import detect as DT
from . import rpn


def tile_grid(img_width, img_height, tile_size, step_size):
    """Tiles [x0, y0, x1, y1] of the sliding window (reference RADNet.py:511-540): one window every
    `step_size` pixels while it fits, plus a window flush with the far edge; unique, row-major."""

    def axis(length):
        pairs = {(s, s + tile_size) for s in range(0, length, step_size) if s + tile_size <= length}
        pairs.add((max(0, length - tile_size), length))
        return sorted(pairs)

    return [[x0, y0, x1, y1] for (y0, y1) in axis(img_height) for (x0, x1) in axis(img_width)]


class RADNet:
    """Detection of rock art figures: post-processing around two user-supplied networks."""

    def __init__(self, C, model_rpn, model_detector, preprocess_func):
        self.is_object_threshold = 0.5
        self.bbox_threshold = 0.7
        self.C = C
        self.model_rpn = model_rpn
        self.model_detector = model_detector
        self.preprocess_func = preprocess_func
        self.class_mapping = {v: k for k, v in C.class_mapping.items()}

    # ------------------------------------------------------------------ helpers
    def get_real_coordinates(self, ratio, x1, y1, x2, y2):
        """Resized-image pixels -> original pixels, `int(round(v // ratio))` (RADNet.py:44-51), evaluated by
        `radnet_real_coordinates`; `predict` applies the same rule fused into the per-class NMS kernel."""
        if not ratio > 0:
            raise ZeroDivisionError("float floor division by zero")
        D.require_cuda()
        dev = torch.device("cuda:%d" % torch.cuda.current_device())
        v = D.to_device(np.asarray([x1, y1, x2, y2]), np.int32, dev)
        out = D.empty((4,), np.int32, dev)
        from . import _lib
        _lib.call("radnet_real_coordinates", D.ptr(v), 4, float(ratio), D.ptr(out), D.stream_ptr(dev))
        return tuple(int(t) for t in out.cpu().numpy())

    def format_img_size(self, img):
        """Resize so that the short side is C.img_size (RADNet.py:53-74); OpenCV, on the host."""
        import cv2
        img_min_side = float(self.C.img_size)
        (height, width, _) = img.shape
        if width <= height:
            ratio = img_min_side / width
            new_height = int(ratio * height)
            new_width = int(img_min_side)
        else:
            ratio = img_min_side / height
            new_width = int(ratio * width)
            new_height = int(img_min_side)
        img = cv2.resize(img, (new_width, new_height), interpolation=cv2.INTER_CUBIC)
        return img, ratio

    def format_img_channels(self, img):
        """BGR -> RGB, float32, batch axis, network preprocessing (RADNet.py:76-90)."""
        img = img[:, :, (2, 1, 0)]
        img = img.astype(np.float32)
        img = np.expand_dims(img, axis=0)
        return self.preprocess_func(img)

    def format_img(self, img):
        img, ratio = self.format_img_size(img)
        return self.format_img_channels(img), ratio

    def tile_grid(self, img_width, img_height):
        return tile_grid(img_width, img_height, self.C.tile_size, self.C.tile_overlap)

    # ------------------------------------------------------------------ f1
    def _head_outputs(self, R, feature_map):
        """Run the detector head chunk by chunk as the reference does (RADNet.py:103-120) and return
        (padded RoIs (m,4) int32, P_cls (m,n_cls), P_regr (m,4(n_cls-1))) or None for no RoIs."""
        n_rois = self.C.n_rois
        R = np.asarray(R)
        n = R.shape[0]
        if n == 0:
            return None
        tail = n % n_rois
        if tail:                                   # short last chunk: repeat ITS first RoI (RADNet.py:106-112)
            R = np.concatenate([R, np.repeat(R[n - tail][None], n_rois - tail, axis=0)], axis=0)
        pc, pr = [], []
        for k in range(0, R.shape[0], n_rois):
            P_cls, P_regr = self.model_detector.predict([feature_map, np.expand_dims(R[k:k + n_rois], axis=0)])
            pc.append(np.asarray(P_cls)[0])
            pr.append(np.asarray(P_regr)[0])
        return R.astype(np.int32), np.concatenate(pc), np.concatenate(pr)

    def apply_spatial_pyramid_pooling(self, R, feature_map):
        """R (n,4) x,y,w,h in feature cells -> (bboxes, probs): dicts class name -> list of
        [x1,y1,x2,y2] in resized-image pixels / list of float32 scores (RADNet.py:98-152)."""
        head = self._head_outputs(R, feature_map)
        if head is None:
            return {}, {}
        rois, P_cls, P_regr = head
        rec = DT.classify_decode(P_cls[None], P_regr[None], self.C, rois=rois[None],
                                 bbox_threshold=self.bbox_threshold).to_numpy()
        if rec["header"][0, DT.H_NRANGE] > 0:
            raise OverflowError("apply_spatial_pyramid_pooling: decoded box beyond 2**25")
        b, p = DT.record_to_dicts(rec[0], self.class_mapping)
        return ({k: [[int(x) for x in row] for row in v] for k, v in b.items()},
                {k: list(v) for k, v in p.items()})

    # ------------------------------------------------------------------ f2
    def final_nms(self, boxes, probs, obj_avg_threshold=0.2, obj_confidence_threshold=0.8, n_obj_avg=5):
        """Cluster-and-average NMS of one class (RADNet.py:156-240): (M,4) boxes, (M,) probs ->
        (boxes (K,4) int64, probs (K,) float32); `[]` for no boxes.

        Domain of the device path = what `predict` feeds it: integer pixel boxes and float32 scores.  The reference
        sorts, thresholds and averages in the dtype it is given; scores that are not exactly float32 values (a
        float64 array with more precision) and non-integer boxes would be silently narrowed here, so they are
        refused with ValueError instead."""
        if len(boxes) == 0:
            return []
        boxes = np.asarray(boxes)
        probs = np.asarray(probs)
        if probs.dtype != np.float32 and not np.array_equal(probs.astype(np.float32).astype(probs.dtype), probs):
            raise ValueError("final_nms: the device path takes float32 scores (as RADNet.predict produces)")
        np.testing.assert_array_less(boxes[:, 0], boxes[:, 2])
        np.testing.assert_array_less(boxes[:, 1], boxes[:, 3])
        if boxes.dtype.kind != "i" and not np.array_equal(boxes, np.rint(boxes)):
            raise ValueError("final_nms: the device path takes integer pixel boxes (as RADNet.predict produces)")
        D.require_cuda()
        dev = torch.device("cuda:%d" % torch.cuda.current_device())
        m = boxes.shape[0]
        rec_in = DT.ClassRecords.from_arrays([(np.zeros((m,), np.int32), probs, boxes)], m, dev)
        out = DT.final_nms_records(rec_in, 1, 1, 1, obj_avg_threshold, obj_confidence_threshold, n_obj_avg)
        host = out.to_numpy()
        DT.check_records(host, "final_nms")
        k = int(host["header"][0, DT.H_NDET])
        ent = host["entry"][0, :k]
        return ent["box"].astype("int"), ent["prob"].copy()

    # ------------------------------------------------------------------ predict
    def _tile_record(self, img, origin):
        """One view (tile or full image) -> ClassRecords(1) on the device (RADNet.py:553-600)."""
        X, ratio = self.format_img(img)
        [Y1, Y2, F] = self.model_rpn.predict(X)
        R = rpn.rpn_to_roi(Y1, Y2, self.C, overlap_thresh=0.7)
        R[:, 2] -= R[:, 0]                                   # (x1,y1,x2,y2) -> (x,y,w,h), RADNet.py:564-565
        R[:, 3] -= R[:, 1]
        rois, P_cls, P_regr = self._head_outputs(R, F)
        # every view gets a record of the same capacity (300 proposals at most, padded to whole chunks), so the
        # records of all views of an image form one array for the merge kernels
        n_rois = self.C.n_rois
        cap = -(-300 // n_rois) * n_rois
        out = DT.ClassRecords(1, max(cap, rois.shape[0]), torch.device("cuda:%d" % torch.cuda.current_device()))
        return DT.classify_nms(P_cls[None], P_regr[None], self.C, rois=rois[None],
                               bbox_threshold=self.bbox_threshold, nms_thresh=0.2, max_boxes=300,
                               ratio=[ratio], origin=[[int(origin[0]), int(origin[1])]], out=out)

    def predict(self, images):
        """images: list of HxWx3 arrays (the image types of one panel).  Returns the reference's list
        of dicts {'class','prob','x1','y1','x2','y2'} (RADNet.py:502-718)."""
        C = self.C
        n_cls = len(self.class_mapping)
        per_image = []
        for img in images:
            views = []
            if C.max_n_tiles_train > 0:
                for tile in self.tile_grid(img.shape[1], img.shape[0]):
                    views.append(self._tile_record(np.copy(img[tile[1]:tile[3], tile[0]:tile[2], :]),
                                                   (tile[0], tile[1])))
            if C.include_full_img:
                views.append(self._tile_record(img, (0, 0)))
            if not views:
                continue
            tiles = DT.ClassRecords(len(views), views[0].max_det, views[0].raw.device,
                                    raw=torch.cat([v.raw for v in views], dim=0))
            per_image.append((tiles, DT.final_nms_records(tiles, 1, len(views), n_cls, obj_avg_threshold=0.2,
                                                          obj_confidence_threshold=0.8, n_obj_avg=5)))
        if not per_image:
            return []
        for tiles, _ in per_image:                            # conditions the reference turns into exceptions
            DT.check_records(tiles.to_numpy(), "per-class NMS (RADNet.py:574)")
        cap = max(m.max_det for _, m in per_image)
        merged_raw = []
        for _, m in per_image:
            DT.check_records(m.to_numpy(), "final_nms (RADNet.py:672)")
            if m.max_det != cap:                              # views per image differ: re-house in a common capacity
                h = m.to_numpy()[0]
                k = int(h["header"][DT.H_NDET])
                m = DT.ClassRecords.from_arrays([(h["entry"]["cls"][:k], h["entry"]["prob"][:k], h["entry"]["box"][:k])],
                                                cap, m.raw.device)
            merged_raw.append(m.raw)
        merged = DT.ClassRecords(len(merged_raw), cap, merged_raw[0].device, raw=torch.cat(merged_raw, dim=0))
        final = DT.class_nms(merged, 1, len(merged_raw), n_cls, 0.4, max_boxes=300).to_numpy()
        DT.check_records(final, "cross-image NMS (RADNet.py:698)")
        rec = final[0]
        k = int(rec["header"][DT.H_NDET])
        dets = []
        for e in rec["entry"][:k]:
            dets.append({'class': self.class_mapping[int(e["cls"])], 'prob': np.float32(e["prob"]),
                         'x1': np.int64(e["box"][0]), 'y1': np.int64(e["box"][1]),
                         'x2': np.int64(e["box"][2]), 'y2': np.int64(e["box"][3])})
        return dets
