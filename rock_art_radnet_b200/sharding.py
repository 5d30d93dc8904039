"""Per-image sharding across the GPUs of one box and the single collective of the path.

The reference is single-process, batch 1 (rpn.py:96); every hot-path function is
per image, so panels (or tiles of one panel, RADNet.py:543-604) shard with no data
exchange: panel i is processed by rank i % world.  The only collective is a gather
of the fixed-size detection records written by K2; pooled features
stay on the GPU that produced them (they feed that GPU's classifier head).

One process per GPU (`torchrun`), `torch.distributed` with the NCCL backend over
NVLink/NVSwitch; the same code runs on the gloo backend with CPU tensors, which is
how the CPU test-suite covers the world_size > 1 path.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_panels, rank, world):
    """Global panel ids owned by `rank`: i % world == rank (interleaved, balanced to +-1)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return np.arange(rank, n_panels, world, dtype=np.int64)


def shard_sizes(n_panels, world):
    return [len(range(r, n_panels, world)) for r in range(world)]


def global_order(n_panels, world, per_rank):
    """Index array that maps the rank-major gathered layout [world][per_rank] back to global
    panel order; padded slots (rank has fewer than per_rank panels) are dropped."""
    order = np.full((n_panels,), -1, dtype=np.int64)
    for r in range(world):
        ids = shard_indices(n_panels, r, world)
        if len(ids) > per_rank:
            raise ValueError("per_rank=%d too small for %d panels on rank %d" % (per_rank, len(ids), r))
        order[ids] = r * per_rank + np.arange(len(ids))
    assert (order >= 0).all()
    return order


def gather_detections(raw, group=None, async_op=False, out=None):
    """All-gather the per-panel detection records of every rank.

    raw: (B, stride) uint8 tensor on this rank (DetectionRecords.raw).  Returns
    (gathered (world, B, stride) uint8, work-or-None).  With NCCL this is one
    ncclAllGather of B*stride bytes per rank over NVLink; it is latency-bound (tens of
    microseconds for 64 panels), so callers issue it once per batch, asynchronously, and
    overlap it with the RoI-pool kernel of the same batch."""
    if not dist.is_initialized():
        g = raw.unsqueeze(0) if out is None else out.copy_(raw.unsqueeze(0))
        return g, None
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world,) + tuple(raw.shape), dtype=raw.dtype, device=raw.device)
    work = dist.all_gather_into_tensor(out.view(-1), raw.contiguous().view(-1), group=group, async_op=async_op)
    return out, work


def split_gathered(gathered, n_panels, max_boxes):
    """(world, per_rank, stride) uint8 -> list over GLOBAL panel ids of dicts
    {boxes int64 (k,4), scores float32 (k,), index int64 (k,)}."""
    from .pipeline import DetectionRecords
    world, per_rank = int(gathered.shape[0]), int(gathered.shape[1])
    flat = gathered.reshape(world * per_rank, -1).contiguous()
    rec = DetectionRecords(world * per_rank, max_boxes, flat.device, raw=flat)
    dets = rec.to_numpy()
    order = global_order(n_panels, world, per_rank)
    return [dets[int(j)] for j in order]


def gathered_to_global(gathered, n_items):
    """(world, per_rank, stride) gathered records -> (n_items, stride) in GLOBAL item order
    (item i lives on rank i % world, slot i // world); padded slots are dropped.  Stays on the
    device of `gathered` (one index_select), so the merge kernels can consume it directly."""
    world, per_rank = int(gathered.shape[0]), int(gathered.shape[1])
    order = torch.from_numpy(global_order(n_items, world, per_rank)).to(gathered.device)
    return gathered.reshape(world * per_rank, -1).index_select(0, order).contiguous()


class TiledPanelSharder:
    """Tiles of tiled panels (reference RADNet.py:511-604: every tile is an independent 600-px
    image until `final_nms`) spread over the ranks, tile i -> rank i % world.

    Each rank runs decode -> NMS -> RoI pool -> [head] -> head decode + per-class NMS on its own
    tiles (`DetectionPipeline`), the fixed-size labelled detection records are all-gathered over
    NCCL (9.9 KB per tile for 300 slots), brought into global tile order on the device and
    merged per panel (K7 final_nms + per-class NMS at 0.4) - every rank ends up with every
    panel's detections, which is what an all-gather means.  With world == 1 the same code runs
    without a process group."""

    def __init__(self, n_panels, tiles_per_panel, rank=None, world=None, group=None):
        self.group = group
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rank, self.world = int(rank), int(world)
        self.n_panels, self.tiles_per_panel = int(n_panels), int(tiles_per_panel)
        self.n_tiles = self.n_panels * self.tiles_per_panel
        self.local_ids = shard_indices(self.n_tiles, self.rank, self.world)
        self.per_rank = max(shard_sizes(self.n_tiles, self.world))

    def local_slots(self):
        """(global tile ids of this rank, number of padding slots at the end of its batch)."""
        return self.local_ids, self.per_rank - len(self.local_ids)

    def gather_tiles(self, class_records_raw, async_op=False):
        """class_records_raw: (per_rank, stride) uint8 of this rank (padding slots must hold empty
        records).  Returns ((n_tiles, stride) tensor-or-None, work): with async_op the tensor is
        produced by `finish`."""
        gathered, work = gather_detections(class_records_raw, group=self.group, async_op=async_op)
        if work is not None and async_op:
            return gathered, work
        return gathered_to_global(gathered, self.n_tiles), None

    def finish(self, gathered, work):
        work.wait()
        return gathered_to_global(gathered, self.n_tiles)


class OwnerRoutedTiles:
    """Tiles of tiled panels with the merge sharded too: tile g = p*T + t is processed on rank g % world
    (as `TiledPanelSharder`), and the tile records of panel p are routed to ONE rank, its owner p % world,
    which alone runs the merge (K7 final_nms + per-class NMS) for that panel.  Per rank the exchange moves
    only its own tiles' records (one `all_to_all_single` with per-destination split sizes over NVLink) and
    the merge work is n_panels / world panels - the all-gather-to-everyone of `TiledPanelSharder` costs
    world times the bytes and merges every panel on every rank.  `gather_final` then collects the final
    records (one per panel) on every rank in global panel order.

    The schedule (who sends which slots to whom, in which order they arrive) is a pure function of
    (n_panels, T, world) and is built once on the host.  world == 1 needs no process group."""

    def __init__(self, n_panels, tiles_per_panel, rank=None, world=None, group=None, device=None, final_stride=None):
        self.group = group
        # width in bytes of one merged (final) panel record.  Every rank must gather rows of the SAME width, also a
        # rank that owns no panel (fewer panels than ranks) and therefore has no merged record to take it from: give
        # it here and a mismatch is an error instead of a hung collective
        self.final_stride = None if final_stride is None else int(final_stride)
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rank, self.world = int(rank), int(world)
        self.n_panels, self.T = int(n_panels), int(tiles_per_panel)
        self.n_tiles = self.n_panels * self.T
        self.local_ids = shard_indices(self.n_tiles, self.rank, self.world)
        self.owned = shard_indices(self.n_panels, self.rank, self.world)          # panels merged here
        self.max_owned = max(shard_sizes(self.n_panels, self.world))
        T, w = self.T, self.world
        # send side: my slots ordered by (destination, global tile id)
        dest = (self.local_ids // T) % w
        send_perm = np.argsort(dest, kind="stable")
        self.in_splits = [int((dest == d).sum()) for d in range(w)]
        # receive side: from source s arrive, ascending, the tiles g with g % w == s whose panel I own
        recv_ids = []
        self.out_splits = []
        for s in range(w):
            ids = shard_indices(self.n_tiles, s, w)
            mine = ids[(ids // T) % w == self.rank]
            recv_ids.append(mine)
            self.out_splits.append(len(mine))
        recv_ids = np.concatenate(recv_ids) if recv_ids else np.zeros((0,), dtype=np.int64)
        assert len(recv_ids) == len(self.owned) * T
        # position of every arrived record in (owned panel, tile) order
        key = (recv_ids // T // w) * T + recv_ids % T
        to_panel_order = np.argsort(key, kind="stable")
        self._send_perm = torch.from_numpy(send_perm.astype(np.int64))
        self._to_panel_order = torch.from_numpy(to_panel_order.astype(np.int64))
        self._final_order = torch.from_numpy(global_order(self.n_panels, w, self.max_owned))
        if device is not None:
            self.to(device)

    def to(self, device):
        self._send_perm = self._send_perm.to(device)
        self._to_panel_order = self._to_panel_order.to(device)
        self._final_order = self._final_order.to(device)
        return self

    def exchange(self, tile_records):
        """tile_records: (len(local_ids), stride) uint8, the records of this rank's tiles in local slot order.
        Returns (len(owned) * T, stride): the tile records of the panels this rank owns, in (panel, tile)
        order - the input layout of `DetectionPipeline.merge`."""
        if self.world == 1:
            return tile_records
        send = tile_records.index_select(0, self._send_perm)
        recv = torch.empty((sum(self.out_splits), tile_records.shape[1]), dtype=tile_records.dtype,
                           device=tile_records.device)
        dist.all_to_all_single(recv, send, output_split_sizes=self.out_splits, input_split_sizes=self.in_splits,
                               group=self.group)
        return recv.index_select(0, self._to_panel_order)

    def gather_final(self, final_records):
        """final_records: (len(owned), stride) uint8, one merged record per owned panel, ascending panel id (None on a
        rank that owns no panel, when `final_stride` was given).  Returns (n_panels, stride) in global panel order on
        every rank (one all-gather of max_owned records)."""
        if final_records is None:
            if self.final_stride is None:
                raise ValueError("gather_final(None) needs OwnerRoutedTiles(final_stride=...)")
            final_records = torch.zeros((0, self.final_stride), dtype=torch.uint8, device=self._final_order.device)
        if self.final_stride is not None and int(final_records.shape[1]) != self.final_stride:
            raise ValueError("gather_final: records of %d bytes, every rank gathers %d" % (int(final_records.shape[1]),
                                                                                          self.final_stride))
        if int(final_records.shape[0]) != len(self.owned):
            raise ValueError("gather_final: %d records for %d owned panels" % (int(final_records.shape[0]), len(self.owned)))
        if self.world == 1:
            return final_records
        pad = self.max_owned - int(final_records.shape[0])
        if pad:
            final_records = torch.cat([final_records, final_records.new_zeros((pad, final_records.shape[1]))])
        gathered, _ = gather_detections(final_records, group=self.group)
        return gathered.reshape(self.world * self.max_owned, -1).index_select(0, self._final_order)
