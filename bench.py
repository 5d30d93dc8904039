#!/usr/bin/env python
"""bench.py - headline benchmark of the RADNet proposal -> NMS -> RoI-pool hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[2]): a batch of 64 synthetic 600-px panels per GPU
(38x38 ResNet-50 stride-16 maps, 9 anchors, 1024 feature channels), decode + sort + NMS
(overlap 0.7, 300 kept boxes) + 14x14 RoI pool.  One "step" = one pass of the hot path over
one batch per GPU.  Metric: panels/sec, whole job (all GPUs), weak scaling.

  value  device-resident throughput: inputs already in HBM, CUDA events on the launching stream
  e2e    same metric through the public host API (HostPanelStream): pinned host buffers ->
         H2D -> kernels -> detection records D2H, every step
  roofline   the dominant kernel (roi_pool_slice_kernel) against the MEASURED HBM copy peak
  cpu_baseline  the reference's own rpn_to_roi (staged copy, oracle/_ref) + the NumPy restatement of the TF
             RoI-pooling op, on this box's host cores
Further objects on the same JSON line (the other BASELINE configs and north-star kernels):
  targets    configs[1] at the bench batch: K3 radnet_rpn_targets on 64 panels x 20 figures (+ the device replay
             of the 256-region balancing) and the batched calc_iou + get_selected_samples, ms / bytes / frac
  sweep      configs[4]: 10,000 device-generated panels, panel i on rank i % world, batches of 64, NCCL gather of
             the records, global order - STRONG scaling panels/s
  tiled      configs[3]: 1600-px panels as 36 tiles through K1 -> K2 -> K4 -> K5/6 -> owner routing -> K7 + K6

`--impl reference` times the CPU path only (rank 0), same metric/config, bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PANELS_PER_GPU = 64
H = W = 38
A = 9
CH = 1024
POOL = 14
MAX_BOXES = 300
THR = 0.7
FALLBACK_HBM_GBS = 6650.0          # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def config_dict(n_gpus):
    return {
        "workload": "BASELINE configs[2]: batch of 64 synthetic 600-px panels per GPU, decode+NMS+14x14 RoI pool",
        "panels_per_gpu": PANELS_PER_GPU, "map": [H, W], "anchors": A, "feature_channels": CH,
        "pool": POOL, "max_boxes": MAX_BOXES, "overlap_thresh": THR,
        "parallelism": "per-panel sharding x%d, NCCL all-gather of detection records" % n_gpus,
        "l2_policy": "inputs larger than L2 (378 MB of feature maps in, 15.4 GB of pooled features out per step)",
    }


# ------------------------------------------------------------------------------ CPU arm
_REF_RPN = None


def _reference_rpn():
    """The reference's own faster_rcnn.rpn module (staged copy under oracle/_ref, or /root/reference in the build
    container), or None when neither exists."""
    global _REF_RPN
    if _REF_RPN is None:
        try:
            from oracle.reference_import import load_reference, reference_available
            _REF_RPN = load_reference()[0] if reference_available() else False
        except Exception:
            _REF_RPN = False
    return _REF_RPN or None


def cpu_kind():
    return "reference+restated-pool" if _reference_rpn() is not None else "port"


def _cpu_one_panel(seed):
    """Reference path for one panel on one core: rpn_to_roi (the reference's own function when its staged copy is
    present, else the oracle port) + RoiPoolingConv over the kept boxes (NumPy restatement of the TF-1 op: TF is
    not installable).  Returns (total s, kept, checksum, s spent in the restated pooling op)."""
    from oracle import radnet_oracle as O
    from rock_art_radnet_b200 import synthetic as S
    C = S.HotPathConfig()
    cls, regr = S.rpn_maps(seed, H, W, A)
    feat = S.feature_map(seed, H, W, CH)
    ref = _reference_rpn()
    t0 = time.perf_counter()
    R = (ref or O).rpn_to_roi(cls, regr, C, max_boxes=MAX_BOXES, overlap_thresh=THR)
    R[:, 2] -= R[:, 0]
    R[:, 3] -= R[:, 1]
    t1 = time.perf_counter()
    out = O.roi_pooling_conv(feat, R[None], POOL)
    t2 = time.perf_counter()
    return t2 - t0, int(R.shape[0]), float(out[0, 0, 0, 0, 0]), t2 - t1


def cpu_sample(n_panels, workers, seed0=0):
    """Run n_panels reference panels over `workers` processes; returns (panels/s, wall s)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_one_panel, range(seed0, seed0 + workers))          # warm the workers (imports, caches)
        t0 = time.perf_counter()
        res = pool.map(_cpu_one_panel, range(seed0 + 1000, seed0 + 1000 + n_panels), chunksize=1)
        wall = time.perf_counter() - t0
    assert all(r[1] > 0 for r in res)
    cpu_sample.pool_share = sum(r[3] for r in res) / max(sum(r[0] for r in res), 1e-12)
    return n_panels / wall, wall


def host_workers():
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 32))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    workers = host_workers()
    per_step = workers                       # one panel per worker per step (about 1.2 s of wall each)
    for _ in range(args.warmup):
        cpu_sample(per_step, workers)
    t_total, n_total = 0.0, 0
    for k in range(args.steps):
        v, wall = cpu_sample(per_step, workers, seed0=5000 + k * per_step)
        t_total += wall
        n_total += per_step
    value = n_total / t_total
    line = {
        "impl": "reference", "metric": "panels_per_sec", "value": value, "unit": "panels/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 decode/NMS + f32 pooling (NumPy)", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": value, "unit": "panels/s", "cores": workers, "kind": cpu_kind(),
                         "sample": "%d panels per step (one per worker process), %d steps; the reference's rpn_to_roi "
                                   "+ NumPy restatement of RoiPoolingConv's TF-1 bilinear op" % (per_step, args.steps),
                         "restated_pool_share_of_cpu_time": getattr(cpu_sample, "pool_share", None)},
        "e2e": {"value": value, "unit": "panels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}



# ------------------------------------------------------------------------------ further sections
_FLUSH = {}


def flush_l2(dev):
    """Write a buffer twice the size of L2 (inputs / outputs of the next launch start cold)."""
    import torch
    if dev not in _FLUSH:
        _FLUSH[dev] = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    _FLUSH[dev].fill_(1)


def _median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def event_ms(fn, dev, iters=15, flush=True, before=None):
    """Median CUDA-event time of fn() on the current stream; L2 flushed (and `before()` run) ahead of every call."""
    import torch
    ts = []
    for i in range(iters + 2):
        if flush:
            flush_l2(dev)
        if before is not None:
            before()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    return _median(ts)


def stream_ms(fns, reps):
    """Per-call time of `reps` back-to-back calls cycling through `fns` (each writes its own output set; together
    the sets are larger than L2), between two events."""
    import torch
    for fn in fns:
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fns[i % len(fns)]()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


def bench_targets(dev, C, pipe, peak):
    """BASELINE configs[1] at the bench batch: 64 panels (600 px, 20 figures, seeds 0..63).
    K3 is timed two ways: `ms` one launch with a cold L2 (the figures are uploaded right before it, as the data
    generator would), `ms_stream` back-to-back launches cycling over output sets that together exceed L2."""
    import torch
    from rock_art_radnet_b200 import synthetic as S
    from rock_art_radnet_b200.rpn import RoiTargetBatch, gt_feature_cells
    from rock_art_radnet_b200.sampling import RpnSubsampler, SampleSelector, seed_states
    from rock_art_radnet_b200.utils import LAYOUT_NHWC, RpnTargetBatch
    B, G = PANELS_PER_GPU, 20
    gt = np.zeros((B, G, 4)); gtc = np.zeros((B, G, 4)); gcl = np.zeros((B, G), np.int32)
    for b in range(B):
        img = S.gt_figures(b, G, 600, 600, classes=("boat", "human"))
        for k, bb in enumerate(img["bboxes"]):
            gt[b, k] = [bb["x1"], bb["x2"], bb["y1"], bb["y2"]]
        gtc[b], gcl[b] = gt_feature_cells(img, C, C.class_mapping)
    gt_h = torch.from_numpy(gt).pin_memory()
    gt_d = torch.empty_like(gt_h, device=dev)
    bg_d = torch.zeros((B, G), dtype=torch.uint8, device=dev)
    cnt_d = torch.full((B,), G, dtype=torch.int32, device=dev)
    wh_d = torch.tensor([[600.0, 600.0]] * B, dtype=torch.float64, device=dev)
    per_panel = G * 32 + 10 * A * H * W * 8                       # 1,040,320 B
    out = {"workload": "BASELINE configs[1] x %d panels: 600-px panel, %d figures, %dx%dx%d anchors" % (B, G, H, W, A),
           "bytes_per_panel": per_panel}

    def upload():
        gt_d.copy_(gt_h, non_blocking=True)

    for tag, kw in (("rpn", {}), ("rpn_nhwc", dict(layout=LAYOUT_NHWC, regr_scale=float(C.std_scaling)))):
        sets = [RpnTargetBatch(C, B, G, H, W, device=dev, **kw) for _ in range(4)]      # 4 x 66.6 MB > L2
        ms = event_ms(lambda: sets[0].run(gt_d, bg_d, cnt_d, wh_d), dev, before=upload)
        ms_s = stream_ms([(lambda t=t: t.run(gt_d, bg_d, cnt_d, wh_d)) for t in sets], 40)
        nbytes = B * per_panel
        out[tag] = {"kernel": "rpn_targets_kernel (one launch)", "ms": ms, "ms_stream": ms_s,
                    "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak,
                    "gbs_stream": nbytes / ms_s / 1e6, "frac_stream": nbytes / ms_s / 1e6 / peak,
                    "l2": "ms: L2 flushed before every launch; ms_stream: 40 launches over 4 output sets (266 MB > L2)"}
        if tag == "rpn":
            # context for a 66.6 MB burst: the driver's own memset of the same two tensors, timed the same two ways
            # (a burst this short never sees the steady-state HBM rate the roofline peak is quoted at)
            ms_m = event_ms(lambda: (sets[0].y_cls.zero_(), sets[0].y_regr.zero_()), dev, before=upload)
            ms_ms = stream_ms([(lambda t=t: (t.y_cls.zero_(), t.y_regr.zero_())) for t in sets], 40)
            out[tag].update(memset_same_tensors_ms=ms_m, memset_same_tensors_ms_stream=ms_ms,
                            frac_of_memset=ms_m / ms, frac_of_memset_stream=ms_ms / ms_s)
            y_cls = sets[0].y_cls
            sub = RpnSubsampler(B, H, W, A, device=dev)
            states = seed_states(np.arange(B), device=dev)
            keep = y_cls.clone()

            def restore():
                y_cls.copy_(keep)
            out["subsample"] = {"kernel": "rpn_subsample_kernel (np.random.choice replay, one CTA per panel)",
                                "ms": event_ms(lambda: sub.run(y_cls, states), dev, before=restore)}
        del sets
    # 512 panels: the same launch at a size where the stream is 533 MB
    B2 = 512
    idx = torch.arange(B2, device=dev) % B
    gt2, bg2, cnt2, wh2 = gt_d[idx].contiguous(), bg_d[idx].contiguous(), cnt_d[idx].contiguous(), wh_d[idx].contiguous()
    big = RpnTargetBatch(C, B2, G, H, W, device=dev)
    ms = event_ms(lambda: big.run(gt2, bg2, cnt2, wh2), dev, iters=8)
    out["rpn_512_panels"] = {"ms": ms, "algorithmic_bytes": B2 * per_panel, "gbs": B2 * per_panel / ms / 1e6,
                             "frac": B2 * per_panel / ms / 1e6 / peak}
    del big
    # a4 batched, straight from the detection records of the headline pipeline (K2's output of the last step)
    rt = RoiTargetBatch(C, C.class_mapping, B, MAX_BOXES, G, device=dev)
    gtc_d, gcl_d = torch.from_numpy(gtc).to(dev), torch.from_numpy(gcl).to(dev)
    ms = event_ms(lambda: rt.run(gtc_d, gcl_d, cnt_d, det=pipe.records), dev)
    kept = int(pipe.records.counts.sum())
    n_out = int(rt.count.sum())
    n_cls = len(C.class_mapping)
    nbytes = kept * 16 + B * G * 32 + n_out * (32 + 8 * n_cls + 8 * 8 * (n_cls - 1))
    out["roi"] = {"kernel": "roi_targets_kernel (radnet_roi_targets_batch, one CTA per panel)", "ms": ms,
                  "rois_in": kept, "rows_out": n_out, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6,
                  "frac": nbytes / ms / 1e6 / peak, "bound": "latency (0.15 MB per panel)"}
    sel = SampleSelector(B, MAX_BOXES, n_cls, int(C.n_rois), device=dev)
    states = seed_states(np.arange(B) + 100, device=dev)
    out["select"] = {"kernel": "select_samples_kernel (get_selected_samples, one CTA per panel)",
                     "ms": event_ms(lambda: sel.run(rt.y_class, rt.count, states), dev)}
    out["gpu_launches"] = 1
    return out


def bench_sweep(dev, C, rank, world, barrier, n_panels):
    """BASELINE configs[4]: the archive sweep, strong scaling (total work fixed)."""
    import torch
    from rock_art_radnet_b200.sweep import ArchiveSweep
    sweep = ArchiveSweep(C, n_panels, batch=PANELS_PER_GPU, seed=2024, H=H, W=W, channels=CH, pool_size=POOL,
                         max_boxes=MAX_BOXES, overlap_thresh=THR, rank=rank, world=world, device=dev)
    warm = ArchiveSweep(C, min(n_panels, 2 * PANELS_PER_GPU * world), batch=PANELS_PER_GPU, seed=7, H=H, W=W,
                        channels=CH, pool_size=POOL, max_boxes=MAX_BOXES, overlap_thresh=THR, rank=rank, world=world,
                        device=dev)
    warm.pipe.pooled = sweep.pipe.pooled            # one pooled buffer
    warm.run()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rec = sweep.run()
    b.record()
    barrier()
    ms = a.elapsed_time(b)
    hdr = rec.view(torch.int32)[:, :4]
    kept = int(hdr[:, 0].sum())
    bad = int((hdr[:, 0] <= 0).sum())
    checksum = int(rec.view(torch.int32).to(torch.int64).sum() % (1 << 61))
    return ms, {"workload": "BASELINE configs[4]: %d device-generated 600-px panels, panel i on rank i %% world, batches "
                            "of %d, NCCL all-gather of the records per batch, global order" % (n_panels, PANELS_PER_GPU),
                "panels": n_panels, "scaling": "strong", "steps_per_rank": sweep.steps, "tail_batch": sweep.tail,
                "kept_boxes": kept, "panels_without_proposals": bad, "records_checksum": checksum,
                "gather_bytes_per_rank": sweep.steps * PANELS_PER_GPU * sweep.stride,
                "gpu_launches_per_rank": sweep.launches,
                "note": "the generator kernel (seed -> maps on device) runs inside the timed region"}


def bench_tiled(dev, C, rank, world, barrier, panels_per_gpu, steps):
    """BASELINE configs[3]: 1600-px panels, 36 tiles each, every stage of the tiled path."""
    import torch
    from rock_art_radnet_b200 import synthetic as S
    from rock_art_radnet_b200.tiled import TiledPanelRunner, synthetic_tile_inputs
    tiles = S.tiled_panel_tiles(1600, 1600)
    n_panels = panels_per_gpu * world
    runner = TiledPanelRunner(C, n_panels, tiles, chunk=PANELS_PER_GPU, H=H, W=W, channels=CH, pool_size=POOL,
                              max_boxes=MAX_BOXES, overlap_thresh=THR, rank=rank, world=world, device=dev)
    cls, regr, feat, P_cls, P_regr = synthetic_tile_inputs(runner, with_features=True, channels=CH)
    for _ in range(2):
        final = runner.run(cls, regr, feat, P_cls, P_regr)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        final = runner.run(cls, regr, feat, P_cls, P_regr)
    b.record()
    barrier()
    ms = a.elapsed_time(b)
    n_det = final.view(torch.int32)[:, 0]
    return ms, {"workload": "BASELINE configs[3]: 1600-px panels as %d tiles of 600 px (12,996 pre-NMS proposals per tile), "
                            "K1 -> K2 -> K4 -> K5/K6 -> owner routing (all_to_all) -> K7 + K6 -> all-gather of the final "
                            "records; the classifier head is a resident stand-in" % len(tiles),
                "panels": n_panels, "tiles_per_panel": len(tiles), "tiles_per_gpu": runner.n_local, "scaling": "weak",
                "steps": steps, "detections_per_panel": [int(v) for v in n_det[:4].tolist()],
                "merge": "each panel merged on one rank (p % world)", "gpu_launches_per_step_per_rank": runner.launches}

# ------------------------------------------------------------------------------ GPU arm
def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def run_b200_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # CPU baseline first: worker processes are forked before this process touches CUDA
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        workers = host_workers()
        n_cpu = 16 * workers                 # about 15-30 s of CPU work on this box
        v, wall = cpu_sample(n_cpu, workers)
        cpu_baseline = {"value": v, "unit": "panels/s", "cores": workers, "kind": cpu_kind(),
                        "sample": "%d panels over %d worker processes (%.1f s wall); the reference's rpn_to_roi + "
                                  "NumPy restatement of RoiPoolingConv's TF-1 bilinear op" % (n_cpu, workers, wall),
                        "restated_pool_share_of_cpu_time": getattr(cpu_sample, "pool_share", None)}

    import torch
    import torch.distributed as dist
    from rock_art_radnet_b200 import synthetic as S
    from rock_art_radnet_b200 import sharding
    from rock_art_radnet_b200.pipeline import HostPanelStream, ProposalPipeline

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if args.k4_lockstep:
        from rock_art_radnet_b200 import _lib
        c, e = (int(v) for v in args.k4_lockstep.split(","))
        _lib.set_option("roipool_cluster", c)
        _lib.set_option("roipool_sync_every", e)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    C = S.HotPathConfig()
    B = args.panels
    # synthetic panels of this rank (global panel id = rank + world*i), generated on the host
    ids = [rank + world * i for i in range(B)]
    cls_h = torch.empty((B, H, W, A), dtype=torch.float32).pin_memory()
    regr_h = torch.empty((B, H, W, 4 * A), dtype=torch.float32).pin_memory()
    feat_h = torch.empty((B, H, W, CH), dtype=torch.float32).pin_memory()
    for i, pid in enumerate(ids):
        c, r = S.rpn_maps(pid, H, W, A)
        cls_h[i] = torch.from_numpy(c[0])
        regr_h[i] = torch.from_numpy(r[0])
        feat_h[i] = torch.from_numpy(S.feature_map(pid, H, W, CH)[0])
    cls_d, regr_d, feat_d = cls_h.to(dev), regr_h.to(dev), feat_h.to(dev)

    pipe = ProposalPipeline(C, B, H, W, channels=CH, pool_size=POOL, max_boxes=MAX_BOXES, overlap_thresh=THR,
                            device=dev)
    gathered = torch.empty((world, B, pipe.records.stride), dtype=torch.uint8, device=dev) if world > 1 else None

    def step(evs=None):
        s = torch.cuda.current_stream()
        if evs:
            evs[0].record(s)
        pipe.decode(cls_d, regr_d)
        if evs:
            evs[1].record(s)
        pipe.sort_nms()
        if evs:
            evs[2].record(s)
        work = None
        if world > 1:
            _, work = sharding.gather_detections(pipe.records.raw, async_op=True, out=gathered)
        pipe.pool(feat_d)
        if evs:
            evs[3].record(s)
        if work is not None:
            work.wait()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    pipe.check_stats()
    counts = pipe.records.counts.cpu().numpy()

    # ---- timed region: K steps, device-resident inputs ------------------------------
    K = args.steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(K):
        step(evs[k])
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = t_start.elapsed_time(t_end)
    dec_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / K
    nms_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / K
    pool_ms = sum(e[2].elapsed_time(e[3]) for e in evs) / K

    # ---- NMS latency: single-panel launches, p50 / p95 ------------------------------
    # (a) device latency: the single-panel call captured in a CUDA graph and replayed between two events,
    #     so that the Python / ctypes launch path is not inside the timed region;
    # (b) the same call issued from Python (what a caller of the drop-in function pays on top).
    single = ProposalPipeline(C, 1, H, W, alloc_pooled=False, device=dev)
    single.decode(cls_d[0:1], regr_d[0:1])
    single.sort_nms()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        single.sort_nms()
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            single.sort_nms()
    torch.cuda.synchronize()
    lat, lat_call = [], []
    for i in range(48):
        single.decode(cls_d[i % B:i % B + 1], regr_d[i % B:i % B + 1])
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b_.record()
        b_.synchronize()
        c, d_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c.record()
        single.sort_nms()
        d_.record()
        d_.synchronize()
        if i >= 8:
            lat.append(a.elapsed_time(b_) * 1e3)
            lat_call.append(c.elapsed_time(d_) * 1e3)
    lat.sort()
    lat_call.sort()

    # ---- e2e: public host API, H2D + kernels + D2H of the records every step --------
    stream = HostPanelStream(pipe)
    for _ in range(2):
        stream.submit(cls_h, regr_h, feat_h)
        stream.collect()
    # the job's K * world batches are handed out through a shared counter: a rank with a faster host link takes more
    # of them (rate-aware streaming); the number of panels of the job is what a static split would process
    from rock_art_radnet_b200.pipeline import SharedBatchCounter
    queue = SharedBatchCounter(K * world, name="radnet_e2e_%d" % os.getpid() if world == 1 else "radnet_e2e")
    barrier()
    t0 = time.perf_counter()
    inflight = 0
    checksum = 0
    my_batches = 0
    while True:
        if queue.next() is None:
            break
        stream.submit(cls_h, regr_h, feat_h)
        my_batches += 1
        inflight += 1
        if inflight == 2:
            checksum += int(stream.collect()[0, 0])
            inflight -= 1
    while inflight:
        checksum += int(stream.collect()[0, 0])
        inflight -= 1
    torch.cuda.synchronize()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    batches = torch.tensor([my_batches], dtype=torch.int64, device=dev)
    if world > 1:
        gathered_b = [torch.zeros_like(batches) for _ in range(world)]
        dist.all_gather(gathered_b, batches)
        batches_per_rank = [int(t.item()) for t in gathered_b]
    else:
        batches_per_rank = [my_batches]
    assert sum(batches_per_rank) == K * world, batches_per_rank

    # ---- the other BASELINE configs and north-star kernels ---------------------------
    peak, peak_src = measured_hbm_peak()
    targets = sweep = tiled = None
    sweep_ms = tiled_ms = 0.0
    h2d_bytes, d2h_bytes = stream.h2d_bytes_per_batch, stream.d2h_bytes_per_batch
    if not args.skip_extra:
        del stream
        targets = bench_targets(dev, C, pipe, peak)
        del single, graph
        pipe.pooled = None
        torch.cuda.empty_cache()
        sweep_ms, sweep = bench_sweep(dev, C, rank, world, barrier, args.sweep_panels)
        torch.cuda.empty_cache()
        tiled_ms, tiled = bench_tiled(dev, C, rank, world, barrier, args.tiled_panels, args.tiled_steps)

    # ---- max over ranks ------------------------------------------------------------
    times = torch.tensor([elapsed_ms, e2e_ms, sweep_ms, tiled_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms, sweep_ms, tiled_ms = (float(v) for v in times)

    if rank == 0:
        panels = world * B * K
        value = panels / (elapsed_ms * 1e-3)
        kept = int(counts.sum())
        # algorithmic bytes of the RoI-pool launch: map once + rois + pooled output
        pool_bytes = B * (H * W * CH * 4) + kept * 16 + kept * POOL * POOL * CH * 4
        achieved = pool_bytes / (pool_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None      # DRAM bytes of one launch: NOT measured in this run
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("panels_per_launch") == B:
                traffic = tj["traffic_bytes_per_launch"]
                traffic_src = "from the committed ncu --set full capture %s (same launch shape), not measured in this run" % tj.get("source", "profiles/")
        except (OSError, ValueError, KeyError):
            pass
        line = {
            "metric": "panels_per_sec", "value": value, "unit": "panels/s", "n_gpus": world, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 decode, int32 exact NMS, f32 pooling", "data": "synthetic",
            "config": dict(config_dict(world), panels_per_gpu=B), "impl": "b200",
            "clocks": clocks,
            "e2e": {"value": panels / (e2e_ms * 1e-3), "unit": "panels/s",
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "batches_per_rank": batches_per_rank,
                    "note": "HostPanelStream: pinned host maps -> H2D -> decode+NMS+pool -> detection records D2H; "
                            "pooled features stay in HBM for the classifier head; the job's batches are handed out "
                            "through a shared counter (a faster host link takes more)"},
            "gpu_launches": 3 * K,
            "kernels_ms_per_step": {"decode_clip": dec_ms, "sort_nms": nms_ms, "roi_pool": pool_ms},
            "nms_latency_us": {"p50": lat[len(lat) // 2], "p95": lat[int(len(lat) * 0.95) - 1], "n": len(lat),
                               "p50_python_call": lat_call[len(lat_call) // 2],
                               "what": "radnet_sort_nms_i32, one 600-px panel (12,996 candidates) per launch; p50/p95 = "
                                       "CUDA-graph replay between two events (device latency), p50_python_call = the "
                                       "same launch issued through the ctypes binding"},
            "roofline": {"kernel": "roi_pool_slice_kernel<8, 14>", "launch_form": pipe.pool_form(),
                         "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": pool_bytes, "avg_launch_ms": pool_ms},
            "kept_boxes_per_step": kept,
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if targets is not None:
            line["targets"] = targets
            sweep["ms"] = sweep_ms
            sweep["value"] = sweep["panels"] / (sweep_ms * 1e-3)
            sweep["unit"] = "panels/s"
            sweep["n_gpus"] = world
            line["sweep"] = sweep
            tiled["ms_per_step"] = tiled_ms / tiled["steps"]
            tiled["value"] = tiled["panels"] * tiled["steps"] / (tiled_ms * 1e-3)
            tiled["unit"] = "panels/s (1600-px panels; x %d for tiles/s)" % tiled["tiles_per_panel"]
            tiled["n_gpus"] = world
            line["tiled"] = tiled
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-extra", action="store_true", help="headline only (profiling runs)")
    ap.add_argument("--k4-lockstep", default=None, metavar="CLUSTER,EVERY",
                    help="fix K4's launch form instead of letting the first call time the candidates (profiling runs: "
                         "under ncu the timing of the candidates means nothing)")
    ap.add_argument("--sweep-panels", type=int, default=10000)
    ap.add_argument("--tiled-panels", type=int, default=4, help="1600-px panels per GPU in the tiled section")
    ap.add_argument("--tiled-steps", type=int, default=5)
    ap.add_argument("--panels", type=int, default=PANELS_PER_GPU,
                    help="panels per GPU per step (default = the BASELINE workload; smaller only for profiling)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
