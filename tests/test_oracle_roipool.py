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
