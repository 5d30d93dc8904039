#!/usr/bin/env python
"""Multi-GPU check of the tiled-panel path (BASELINE configs[3]): tiles of 1600-px panels are sharded
tile g -> rank g % world, each rank runs decode -> NMS -> head decode -> per-class NMS on its tiles, the labelled
records of panel p are routed to its owner p % world (one all_to_all), which alone merges the panel (final_nms +
NMS 0.4), and the final records are all-gathered.  The result must be byte-identical to the same panels processed
by one GPU alone, and every panel must be merged on exactly one rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_sharded_detect.py [--panels 5]
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
from rock_art_radnet_b200 import synthetic as S  # noqa: E402
from rock_art_radnet_b200.tiled import TiledPanelRunner, synthetic_tile_inputs  # noqa: E402


def run(C, tiles, n_panels, rank, world):
    runner = TiledPanelRunner(C, n_panels, tiles, rank=rank, world=world, alloc_pooled=False)
    cls, regr, _, P_cls, P_regr = synthetic_tile_inputs(runner)
    final = runner.run(cls, regr, None, P_cls, P_regr)
    return runner, final.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--panels", type=int, default=5)
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
    runner, final_sh = run(C, tiles, args.panels, rank, world)
    _, final_1 = run(C, tiles, args.panels, 0, 1)
    ok = final_sh.shape == final_1.shape and np.array_equal(final_sh, final_1)
    n_det = np.ascontiguousarray(final_sh).view(np.int32)[:, DT.H_NDET].tolist()
    flag = torch.tensor([int(ok)], device="cuda")
    owned = torch.tensor([len(runner.router.owned)], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    if rank == 0:
        print("sharded tiled-panel check: world=%d panels=%d tiles/panel=%d detections/panel=%s merged once each=%s "
              "identical=%s" % (world, args.panels, len(tiles), n_det, int(owned.item()) == args.panels, bool(flag.item())))
    if world > 1:
        dist.destroy_process_group()
    return 0 if flag.item() and int(owned.item()) == args.panels else 1


if __name__ == "__main__":
    sys.exit(main())
