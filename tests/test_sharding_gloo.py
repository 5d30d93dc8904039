"""CPU: the world_size > 1 path (per-panel sharding + detection gather) on the gloo backend."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from conftest import ROOT  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_panels, max_boxes, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rock_art_radnet_b200 import sharding
    from rock_art_radnet_b200.pipeline import DetectionRecords
    ids = sharding.shard_indices(n_panels, rank, world)
    per = max(sharding.shard_sizes(n_panels, world))
    rec = DetectionRecords(per, max_boxes, "cpu")
    # fake "detections": panel p keeps (p % max_boxes) + 1 boxes whose coordinates encode p
    for slot, p in enumerate(ids):
        k = int(p) % max_boxes + 1
        rec.header[slot, 0] = k
        rec.header[slot, 1] = 1000 + int(p)
        for j in range(k):
            rec.boxes[slot, j] = torch.tensor([int(p), j, int(p) + 1, j + 1], dtype=torch.int32)
            rec.scores[slot, j] = float(p) + j / 100.0
            rec.index[slot, j] = int(p) * 10 + j
    gathered, work = sharding.gather_detections(rec.raw, async_op=True)
    work.wait()
    assert gathered.shape == (world, per, rec.stride)
    dets = sharding.split_gathered(gathered, n_panels, max_boxes)
    ok = len(dets) == n_panels
    for p, d in enumerate(dets):
        k = p % max_boxes + 1
        ok &= d["boxes"].shape == (k, 4) and d["n_candidates"] == 1000 + p
        ok &= d["boxes"][:, 0].tolist() == [p] * k and d["index"].tolist() == [p * 10 + j for j in range(k)]
    np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.array([int(ok)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_panels", [(2, 9), (2, 64)])
def test_gather_detections_world2_gloo(tmp_path, world, n_panels):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_panels, 6, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / ("ok_%d.npy" % r))[0] == 1


def test_gather_without_process_group_is_identity():
    from rock_art_radnet_b200 import sharding
    raw = torch.arange(2 * 40, dtype=torch.uint8).reshape(2, 40)
    g, work = sharding.gather_detections(raw)
    assert work is None and g.shape == (1, 2, 40) and torch.equal(g[0], raw)


def _tile_worker(rank, world, port, n_panels, tiles_per_panel, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rock_art_radnet_b200 import sharding
    shard = sharding.TiledPanelSharder(n_panels, tiles_per_panel)
    ids, pad = shard.local_slots()
    stride = 48
    raw = torch.zeros((shard.per_rank, stride), dtype=torch.uint8)
    for slot, t in enumerate(ids):
        raw[slot] = int(t) % 251                      # record bytes encode the global tile id
    raw[len(ids):] = 255                              # padding slots must be dropped by the reorder
    glob, _ = shard.gather_tiles(raw)
    g2, work = shard.gather_tiles(raw, async_op=True)
    glob2 = shard.finish(g2, work)
    ok = tuple(glob.shape) == (n_panels * tiles_per_panel, stride) and torch.equal(glob, glob2)
    ok &= glob[:, 0].tolist() == [t % 251 for t in range(n_panels * tiles_per_panel)]
    ok &= len(ids) + pad == shard.per_rank
    np.save(os.path.join(out_dir, "tile_ok_%d.npy" % rank), np.array([int(ok)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_panels,tiles_per_panel", [(1, 36), (3, 7)])
def test_tiled_panel_sharder_world2_gloo(tmp_path, n_panels, tiles_per_panel):
    """Tiles of tiled panels (BASELINE configs[3]: 36 tiles of a 1600-px panel) over 2 ranks: gather +
    reorder to global tile order, padding slots dropped, blocking and asynchronous forms agree."""
    port = _free_port()
    mp.spawn(_tile_worker, args=(2, port, n_panels, tiles_per_panel, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(tmp_path / ("tile_ok_%d.npy" % r))[0] == 1


def _owner_worker(rank, world, port, n_panels, T, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rock_art_radnet_b200 import sharding
    stride, final_stride = 32, 80            # merged panel records are wider than tile records, as in the real path
    router = sharding.OwnerRoutedTiles(n_panels, T, final_stride=final_stride)
    raw = torch.zeros((len(router.local_ids), stride), dtype=torch.uint8)
    for slot, g in enumerate(router.local_ids):
        raw[slot, 0] = int(g) // T            # panel
        raw[slot, 1] = int(g) % T             # tile
        raw[slot, 2] = rank
    got = router.exchange(raw)
    ok = tuple(got.shape) == (len(router.owned) * T, stride)
    want = [(int(p), t) for p in router.owned for t in range(T)]
    ok &= [(int(r[0]), int(r[1])) for r in got] == want                     # (owned panel, tile) order
    ok &= all(int(r[2]) == (int(r[0]) * T + int(r[1])) % world for r in got)   # each came from the rank that held it
    final = torch.zeros((len(router.owned), final_stride), dtype=torch.uint8)
    for i, p in enumerate(router.owned):
        final[i, 0] = int(p)
        final[i, 1] = 100 + rank
    try:                                      # rows of the tile-record width (the bug class that hangs a real gather)
        router.gather_final(torch.zeros((len(router.owned), stride), dtype=torch.uint8))
        ok = False
    except ValueError:
        pass
    # a rank that owns no panel has no merged record to take the width from: it passes None
    glob = router.gather_final(final if len(router.owned) else None)
    ok &= tuple(glob.shape) == (n_panels, final_stride) and glob[:, 0].tolist() == list(range(n_panels))
    ok &= glob[:, 1].tolist() == [100 + p % world for p in range(n_panels)]
    np.save(os.path.join(out_dir, "owner_ok_%d.npy" % rank), np.array([int(ok)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_panels,T", [(2, 1, 36), (2, 5, 7), (3, 4, 36), (4, 3, 5)])
def test_owner_routed_tiles_gloo(tmp_path, world, n_panels, T):
    """Owner routing of the tile merge: every tile record reaches exactly the rank that owns its panel
    (p % world), in (panel, tile) order; final records come back in global panel order.  Uneven splits,
    ranks that own no panel, more ranks than panels."""
    port = _free_port()
    mp.spawn(_owner_worker, args=(world, port, n_panels, T, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / ("owner_ok_%d.npy" % r))[0] == 1


def test_owner_routing_without_process_group_is_identity():
    from rock_art_radnet_b200 import sharding
    router = sharding.OwnerRoutedTiles(2, 5, rank=0, world=1)
    raw = torch.arange(10 * 8, dtype=torch.uint8).reshape(10, 8)
    assert torch.equal(router.exchange(raw), raw) and torch.equal(router.gather_final(raw[:2]), raw[:2])


def _counter_worker(rank, world, port, total, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rock_art_radnet_b200.pipeline import SharedBatchCounter
    q = SharedBatchCounter(total)
    mine = []
    while True:
        i = q.next()
        if i is None:
            break
        mine.append(i)
        if rank == 0:
            import time
            time.sleep(0.002)                 # a slow link: the other rank should take more batches
    np.save(os.path.join(out_dir, "cnt_%d.npy" % rank), np.array(mine, dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


def test_shared_batch_counter_hands_out_every_batch_once_gloo(tmp_path):
    """Rate-aware streaming: every batch index is handed out exactly once across ranks, and the rank that is slower
    per batch ends up with fewer of them."""
    port = _free_port()
    total = 200
    mp.spawn(_counter_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "cnt_0.npy"), np.load(tmp_path / "cnt_1.npy")
    assert sorted(a.tolist() + b.tolist()) == list(range(total))
    assert len(b) > len(a)


def test_shared_batch_counter_without_process_group():
    from rock_art_radnet_b200.pipeline import SharedBatchCounter
    q = SharedBatchCounter(3)
    assert [q.next(), q.next(), q.next(), q.next()] == [0, 1, 2, None]
