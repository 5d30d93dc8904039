"""CPU: the TF-1 legacy-bilinear restatement (a5) against independent evaluations.

TensorFlow cannot be installed here, so the restatement is checked (i) against a float64
re-derivation of the same published kernel written independently of the vectorised code,
(ii) against torch.nn.functional.grid_sample(align_corners=True) on an explicitly built grid of
src = i*scale coordinates, and (iii) on closed-form cases."""
import numpy as np
import pytest

from oracle import radnet_oracle as O
from rock_art_radnet_b200 import synthetic as S


def _scalar_reference(img, roi, pool):
    """Loop-level float64 evaluation of crop + legacy bilinear (no vectorisation shared with the oracle)."""
    _, H, W, C = img.shape
    x, y, w, h = [int(v) for v in roi]
    ch = min(y + h, H) - y
    cw = min(x + w, W) - x
    out = np.zeros((pool, pool, C))
    sy = np.float32(ch) / np.float32(pool)
    sx = np.float32(cw) / np.float32(pool)
    for i in range(pool):
        fy = float(np.float32(i) * sy)
        y0 = int(np.floor(fy)); y1 = min(int(np.ceil(fy)), ch - 1); ly = fy - np.floor(fy)
        for j in range(pool):
            fx = float(np.float32(j) * sx)
            x0 = int(np.floor(fx)); x1 = min(int(np.ceil(fx)), cw - 1); lx = fx - np.floor(fx)
            tl = img[0, y + y0, x + x0].astype(np.float64); tr = img[0, y + y0, x + x1].astype(np.float64)
            bl = img[0, y + y1, x + x0].astype(np.float64); br = img[0, y + y1, x + x1].astype(np.float64)
            top = tl + (tr - tl) * lx
            bot = bl + (br - bl) * lx
            out[i, j] = top + (bot - top) * ly
    return out


@pytest.mark.parametrize("pool", [14, 7, 3])
def test_restatement_matches_scalar_float64_evaluation(pool):
    img = S.feature_map(5, 20, 24, 16)
    rng = np.random.default_rng(0)
    rois = []
    for _ in range(40):
        x = int(rng.integers(0, 23)); y = int(rng.integers(0, 19))
        rois.append([x, y, int(rng.integers(1, 24 - x + 1)), int(rng.integers(1, 20 - y + 1))])
    rois += [[0, 0, 24, 20], [23, 19, 1, 1], [3, 4, 40, 40]]      # full map, 1x1, end clamped
    rois = np.array(rois)[None]
    got = O.roi_pooling_conv(img, rois, pool)
    assert got.dtype == np.float32 and got.shape == (1, rois.shape[1], pool, pool, 16)
    for k in range(rois.shape[1]):
        ref = _scalar_reference(img, rois[0, k], pool)
        np.testing.assert_allclose(got[0, k], ref, rtol=2e-6, atol=2e-6)


def test_restatement_matches_grid_sample():
    torch = pytest.importorskip("torch")
    img = S.feature_map(6, 16, 16, 8)
    roi = np.array([[[2, 3, 11, 9]]])
    pool = 7
    got = O.roi_pooling_conv(img, roi, pool)[0, 0]
    x, y, w, h = roi[0, 0]
    crop = torch.from_numpy(img[0, y:y + h, x:x + w]).permute(2, 0, 1)[None].double()
    sy = np.float32(h) / np.float32(pool); sx = np.float32(w) / np.float32(pool)
    fy = (np.arange(pool, dtype=np.float32) * sy).astype(np.float64)
    fx = (np.arange(pool, dtype=np.float32) * sx).astype(np.float64)
    gy = 2 * fy / (h - 1) - 1
    gx = 2 * fx / (w - 1) - 1
    grid = torch.from_numpy(np.stack(np.meshgrid(gx, gy), axis=-1))[None]
    ref = torch.nn.functional.grid_sample(crop, grid, mode="bilinear", padding_mode="border", align_corners=True)
    np.testing.assert_allclose(got, ref[0].permute(1, 2, 0).numpy(), rtol=1e-5, atol=1e-5)


def test_closed_form_cases():
    # linear ramp: bilinear sampling of a ramp at src = i*scale is exact
    H = W = 10
    ramp = np.zeros((1, H, W, 2), np.float32)
    ramp[0, :, :, 0] = np.arange(W)[None, :]
    ramp[0, :, :, 1] = np.arange(H)[:, None]
    out = O.roi_pooling_conv(ramp, np.array([[[2, 1, 4, 8]]]), 4)[0, 0]
    np.testing.assert_array_equal(out[:, :, 0], np.tile(2 + np.arange(4), (4, 1)).astype(np.float32))
    np.testing.assert_array_equal(out[:, :, 1], np.tile((1 + 2 * np.arange(4))[:, None], (1, 4)).astype(np.float32))
    # a 1x1 RoI replicates its cell; equal in/out size is the identity crop
    img = S.feature_map(7, 9, 9, 3)
    one = O.roi_pooling_conv(img, np.array([[[4, 5, 1, 1]]]), 5)[0, 0]
    assert (one == img[0, 5, 4]).all()
    same = O.roi_pooling_conv(img, np.array([[[1, 2, 6, 6]]]), 6)[0, 0]
    np.testing.assert_array_equal(same, img[0, 2:8, 1:7])
    # float rois truncate toward zero like K.cast(int32)
    a = O.roi_pooling_conv(img, np.array([[[1.9, 2.9, 6.9, 6.2]]]), 6)
    np.testing.assert_array_equal(a[0, 0], same)
    with pytest.raises(ValueError):
        O.roi_pooling_conv(img, np.array([[[9, 0, 3, 3]]]), 4)
    layer = O.RoiPoolingConv(6, 1)
    assert layer.compute_output_shape([img.shape, (1, 1, 4)]) == (None, 1, 6, 6, 3)


# ---- an independent implementation of TensorFlow's op: OpenCV's DNN module imports TF GraphDefs and implements
#      ResizeBilinear(align_corners, half_pixel_centers) to reproduce TF's results.  A three-node GraphDef
#      (Placeholder -> ResizeBilinear <- Const size) is hand-encoded in protobuf wire format (TF itself is not
#      installable here) and run by cv2.dnn; the restatement must agree with it to float32 rounding (OpenCV orders
#      the four-tap sum differently from TF's kernel, so agreement is to ~1e-6, not bit for bit) - and a half-pixel
#      interpretation of the same op must NOT (that is the geometry question the unpinned restatement leaves open).
def _varint(n):
    out = b""
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out += bytes([b | 0x80])
        else:
            return out + bytes([b])


def _ld(num, data):                       # length-delimited field
    return _varint((num << 3) | 2) + _varint(len(data)) + data


def _vi(num, n):                          # varint field
    return _varint(num << 3) + _varint(n)


def _attr(key, value):                    # NodeDef.attr map entry {key = 1, value = 2}
    return _ld(5, _ld(1, key.encode()) + _ld(2, value))


def _node(name, op, inputs=(), attrs=()):  # GraphDef.node = 1: NodeDef {name = 1, op = 2, input = 3, attr = 5}
    body = _ld(1, name.encode()) + _ld(2, op.encode())
    for i in inputs:
        body += _ld(3, i.encode())
    return _ld(1, body + b"".join(attrs))


def _tf_resize_bilinear_graph(out_h, out_w):
    import struct
    DT_FLOAT, DT_INT32 = 1, 3
    shape = _ld(2, _vi(1, 2))                                                  # TensorShapeProto.dim {size = 2}
    tensor = _vi(1, DT_INT32) + _ld(2, shape) + _ld(4, struct.pack("<2i", out_h, out_w))   # TensorProto
    g = _node("input", "Placeholder", attrs=[_attr("dtype", _vi(6, DT_FLOAT))])
    g += _node("size", "Const", attrs=[_attr("dtype", _vi(6, DT_INT32)), _attr("value", _ld(8, tensor))])
    g += _node("resize", "ResizeBilinear", ["input", "size"],
               [_attr("T", _vi(6, DT_FLOAT)), _attr("align_corners", _vi(5, 0)), _attr("half_pixel_centers", _vi(5, 0))])
    return np.frombuffer(g, np.uint8)


@pytest.mark.parametrize("pool", [14, 7])
def test_restatement_matches_opencv_dnn_import_of_tf_resize_bilinear(pool):
    cv2 = pytest.importorskip("cv2")
    if not hasattr(cv2, "dnn"):
        pytest.skip("OpenCV without the dnn module")
    net = cv2.dnn.readNetFromTensorflow(_tf_resize_bilinear_graph(pool, pool))
    rng = np.random.default_rng(5)
    C = 6
    worst, worst_half = 0.0, np.inf
    for (h, w) in [(1, 1), (1, 9), (2, 3), (5, 9), (7, 7), (13, 20), (14, 14), (15, 29), (28, 28), (38, 38), (38, 50)]:
        crop = rng.standard_normal((1, h, w, C)).astype(np.float32)
        # the layer's path for one RoI that covers the whole (h, w) map: crop = img[:, 0:h, 0:w, :], resized to pool x pool
        ours = O.roi_pooling_conv(crop, np.array([[[0, 0, w, h]]], np.float64), pool)[0, 0]          # (pool, pool, C)
        net.setInput(np.ascontiguousarray(crop.transpose(0, 3, 1, 2)))                               # NCHW blob
        theirs = net.forward()[0].transpose(1, 2, 0)
        assert theirs.shape == ours.shape
        scale = max(1.0, float(np.abs(crop).max()))
        worst = max(worst, float(np.abs(ours - theirs).max()) / scale)
        if h > 2 and w > 2 and (h != pool or w != pool):
            # the half-pixel interpretation (TF2's default) of the same op, for contrast
            ys = np.clip((np.arange(pool) + 0.5) * h / pool - 0.5, 0, None)
            xs = np.clip((np.arange(pool) + 0.5) * w / pool - 0.5, 0, None)
            y0, x0 = np.floor(ys).astype(int), np.floor(xs).astype(int)
            y1, x1 = np.minimum(y0 + 1, h - 1), np.minimum(x0 + 1, w - 1)
            ly, lx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
            img = crop[0].astype(np.float64)
            top = img[y0][:, x0] + (img[y0][:, x1] - img[y0][:, x0]) * lx
            bot = img[y1][:, x0] + (img[y1][:, x1] - img[y1][:, x0]) * lx
            half = top + (bot - top) * ly
            worst_half = min(worst_half, float(np.abs(half - theirs).max()))
    assert worst < 2e-6, worst
    assert worst_half > 1e-2, worst_half      # the two geometries are far apart: the op under test is the legacy one
