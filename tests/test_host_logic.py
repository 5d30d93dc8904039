"""CPU: host-side logic of the drop-in package (no kernel launches)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import radnet_oracle as O
from rock_art_radnet_b200 import sharding
from rock_art_radnet_b200 import synthetic as S
from rock_art_radnet_b200.pipeline import DetectionRecords, anchor_cells, anchor_pixels
from rock_art_radnet_b200.utils import get_new_img_size, iou


def test_anchor_tables_follow_reference_order():
    C = S.HotPathConfig((64, 128, 256, 512))
    cells, px = anchor_cells(C), anchor_pixels(C)
    assert cells.shape == (12, 2) and px.shape == (12, 2)
    a = 0
    for scale in C.anchor_box_scales:              # size-major, ratio-minor (rpn.py:108-109)
        for ratio in C.anchor_box_ratios:
            assert cells[a, 0] == (scale * ratio[0]) / C.rpn_stride
            assert px[a, 1] == scale * ratio[1]
            a += 1


def test_scalar_helpers_match_oracle():
    for w, h in [(600, 600), (1000, 700), (333, 900), (1600, 1600)]:
        assert get_new_img_size(w, h, 600) == O.get_new_img_size(w, h, 600)


def test_scalar_arithmetic_helpers_have_no_cpu_path():
    """iou() and RADNet.get_real_coordinates() are evaluated by the library; without a GPU they fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    from rock_art_radnet_b200.RADNet import RADNet
    with pytest.raises(RuntimeError):
        iou([0, 0, 4, 4], [1, 1, 5, 5])
    net = RADNet(S.HotPathConfig(), None, None, None)
    with pytest.raises(RuntimeError):
        net.get_real_coordinates(0.75, 1, 2, 3, 4)


def test_detection_record_views_on_cpu():
    rec = DetectionRecords(3, 5, "cpu")
    assert rec.stride == 144 and rec.raw.shape == (3, rec.stride)      # 16 + 5*24 rounded up to 16 B
    rec.header[1, 0] = 2
    rec.header[1, 1] = 77
    rec.boxes[1, 0] = torch.tensor([1, 2, 3, 4], dtype=torch.int32)
    rec.boxes[1, 1] = torch.tensor([5, 6, 7, 8], dtype=torch.int32)
    rec.scores[1, 0] = 0.5
    rec.index[1, 1] = 42
    d = rec.to_numpy()
    assert d[0]["boxes"].shape == (0, 4) and d[1]["boxes"].dtype == np.int64
    assert d[1]["boxes"].tolist() == [[1, 2, 3, 4], [5, 6, 7, 8]] and d[1]["index"].tolist() == [0, 42]
    assert d[1]["scores"][0] == np.float32(0.5) and d[1]["n_candidates"] == 77
    # byte layout promised by include/radnet_b200.h
    raw = rec.raw[1].numpy()
    assert raw[:4].view(np.int32)[0] == 2
    assert raw[16:32].view(np.int32).tolist() == [1, 2, 3, 4]
    assert raw[16 + 5 * 16:16 + 5 * 16 + 4].view(np.float32)[0] == np.float32(0.5)
    assert raw[16 + 5 * 20 + 4:16 + 5 * 20 + 8].view(np.int32)[0] == 42
    rec.header[2, 0] = -1
    with pytest.raises(RuntimeError):
        rec.to_numpy()


def test_shard_indices_and_global_order():
    for n, world in [(10, 1), (10, 3), (64, 8), (7, 8), (10000, 8)]:
        seen = np.concatenate([sharding.shard_indices(n, r, world) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(n))
        assert sum(sharding.shard_sizes(n, world)) == n
        per = max(sharding.shard_sizes(n, world))
        order = sharding.global_order(n, world, per)
        layout = np.full(world * per, -1)
        for r in range(world):
            ids = sharding.shard_indices(n, r, world)
            layout[r * per:r * per + len(ids)] = ids
        assert np.array_equal(layout[order], np.arange(n))
    with pytest.raises(ValueError):
        sharding.shard_indices(5, 3, 3)


def test_compute_entry_points_refuse_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import rock_art_radnet_b200 as R
    C = S.HotPathConfig()
    cls, regr = S.rpn_maps(0, 4, 4, 9)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        R.rpn_to_roi(cls, regr, C)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        R.non_max_suppression_fast(np.array([[0.0, 0, 1, 1]]), np.array([0.5]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        R.RoiPoolingConv(7, 1)([np.zeros((1, 4, 4, 8), np.float32), np.array([[[0, 0, 2, 2]]])])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        R.calc_region_props(C, S.gt_figures(0, 2), 600, 600, 600, 600, S.resnet50_map_size)
    assert R.non_max_suppression_fast(np.zeros((0, 4)), np.zeros(0)) == []   # rpn.py:391-392 needs no device
