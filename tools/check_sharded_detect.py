#!/usr/bin/env python
"""Multi-GPU check of the tiled-panel path (BASELINE configs[3]): tiles of 1600-px panels are sharded
tile i -> rank i % world, each rank runs decode -> NMS -> head decode -> per-class NMS on its tiles,
the labelled detection records are all-gathered over NCCL and merged (final_nms + NMS 0.4) on every
rank.  The result must be byte-identical to the same panels processed by one GPU alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_sharded_detect.py [--panels 2]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rock_art_radnet_b200 import detect as DT  # noqa: E402
from rock_art_radnet_b200 import sharding  # noqa: E402
from rock_art_radnet_b200 import synthetic as S  # noqa: E402
from rock_art_radnet_b200.pipeline import DetectionPipeline  # noqa: E402


def run(C, tiles, n_panels, rank, world, group_ready):
    T = len(tiles)
    shard = sharding.TiledPanelSharder(n_panels, T, rank=rank, world=world) if not group_ready else \
        sharding.TiledPanelSharder(n_panels, T)
    ids, pad = shard.local_slots()
    B = shard.per_rank
    pipe = DetectionPipeline(C, B, 38, 38, alloc_pooled=False)
    slots = list(ids) + [int(ids[0])] * pad                    # padding slots recompute a real tile, then get emptied
    maps = [S.rpn_maps(1000 * (g // T) + g % T) for g in slots]
    pipe.decode(torch.from_numpy(np.concatenate([m[0] for m in maps])).cuda(),
                torch.from_numpy(np.concatenate([m[1] for m in maps])).cuda())
    pipe.sort_nms()
    dets = pipe.records.to_numpy()
    pcs, prs = [], []
    for i, g in enumerate(slots):
        R = dets[i]["boxes"].copy()
        R[:, 2] -= R[:, 0]
        R[:, 3] -= R[:, 1]
        pc, pr = S.tiled_panel_head_outputs(C, g // T, tiles[g % T], R)
        if i >= len(ids):
            pc[:] = 0                                            # padding: nothing clears the threshold
        pcs.append(pc)
        prs.append(pr)
    origin = torch.tensor([[tiles[g % T][0], tiles[g % T][1]] for g in slots], dtype=torch.int32, device="cuda")
    ratio = torch.ones((B,), dtype=torch.float64, device="cuda")
    pipe.classify(torch.from_numpy(np.stack(pcs)).cuda(), torch.from_numpy(np.stack(prs)).cuda(), ratio=ratio, origin=origin)
    if group_ready:
        g, work = shard.gather_tiles(pipe.class_records.raw, async_op=True)
        glob = shard.finish(g, work)
    else:
        glob = sharding.gathered_to_global(pipe.class_records.raw.unsqueeze(0), shard.n_tiles)
    tile_rec = DT.ClassRecords(n_panels * T, pipe.max_boxes, glob.device, raw=glob)
    merged = DT.final_nms_records(tile_rec, n_panels, T, pipe.n_cls)
    final = DT.class_nms(merged, n_panels, 1, pipe.n_cls, 0.4)
    return tile_rec.raw.cpu().numpy(), final.raw.cpu().numpy(), final.to_numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--panels", type=int, default=2)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    C = S.HotPathConfig()
    tiles = S.tiled_panel_tiles(1600, 1600)
    tiles_sh, final_sh, rec = run(C, tiles, args.panels, rank, world, world > 1)
    tiles_1, final_1, _ = run(C, tiles, args.panels, 0, 1, False)
    ok = np.array_equal(tiles_sh, tiles_1) and np.array_equal(final_sh, final_1)
    n_det = [int(r["header"][0]) for r in rec]
    flag = torch.tensor([int(ok)], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("sharded tiled-panel check: world=%d panels=%d tiles/panel=%d detections/panel=%s identical=%s"
              % (world, args.panels, len(tiles), n_det, bool(flag.item())))
    if world > 1:
        dist.destroy_process_group()
    return 0 if flag.item() else 1


if __name__ == "__main__":
    sys.exit(main())
