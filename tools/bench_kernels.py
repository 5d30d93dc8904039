#!/usr/bin/env python
"""Per-kernel micro-benchmarks on one B200 (CUDA events, warm-up, inputs > L2 or L2 flush).

    python tools/bench_kernels.py [--out gpurun_out/kernels.json]

Reports, against the measured HBM copy peak (MEASURED_PEAKS.json):
  K1 decode_clip   GB/s on 64 and 512 panels          (algorithmic bytes 36*N)
  K2 sort_nms      latency p50/p95 per panel, batch time for 64 panels, 90k-candidate stress
  K3 rpn_targets   GB/s on 64 and 512 panels, 20 GT   (algorithmic bytes G*32 + 10*A*H*W*8)
  a4 roi_targets   latency for 300 RoIs x 20 GT
  K4 roi_pool      GB/s for 64 panels (38x38x1024, 300 RoIs, pool 14) and for VGG-like 512ch/pool 7
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rock_art_radnet_b200 import synthetic as S  # noqa: E402
from rock_art_radnet_b200.pipeline import ProposalPipeline  # noqa: E402
from rock_art_radnet_b200.utils import RpnTargetBatch  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.fill_(1)


def time_ms(fn, iters=20, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return {"p50_ms": ts[len(ts) // 2], "min_ms": ts[0], "p95_ms": ts[max(0, int(len(ts) * 0.95) - 1)]}


def time_graph_ms(fn, reps=20, iters=10, warmup=2):
    """GPU time of one call for kernels shorter than the Python launch path: `reps` calls are captured
    into a CUDA graph on a side stream and the graph is replayed; returns per-call figures."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn()
        st.synchronize()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(reps):
                fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    ts.sort()
    return {"p50_ms": ts[len(ts) // 2], "min_ms": ts[0], "how": "CUDA graph of %d back-to-back calls (L2 warm)" % reps}


def tile_maps(B, H=38, W=38, A=9):
    base = [S.rpn_maps(s, H, W, A) for s in range(8)]
    cls = np.concatenate([base[i % 8][0] for i in range(B)])
    regr = np.concatenate([base[i % 8][1] for i in range(B)])
    return torch.from_numpy(cls).cuda(), torch.from_numpy(regr).cuda()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernels.json"))
    args = ap.parse_args()
    torch.cuda.set_device(0)
    pk = peak()
    C = S.HotPathConfig()
    res = {"hbm_peak_gbs": pk, "gpu": torch.cuda.get_device_name(0)}

    # ---- K1 + K2 ----------------------------------------------------------------------
    for B in (64, 512):
        cls, regr = tile_maps(B)
        pipe = ProposalPipeline(C, B, 38, 38, alloc_pooled=False)
        t = time_ms(lambda: pipe.decode(cls, regr))
        nbytes = 36 * pipe.N * B
        res["decode_clip_B%d" % B] = dict(t, algorithmic_bytes=nbytes, gbs=nbytes / t["p50_ms"] / 1e6,
                                           frac_of_measured_peak=nbytes / t["p50_ms"] / 1e6 / pk)
        t = time_ms(lambda: pipe.sort_nms())
        res["sort_nms_B%d" % B] = dict(t, us_per_panel_amortised=1e3 * t["p50_ms"] / B)
        del pipe
    cls, regr = tile_maps(1)
    pipe = ProposalPipeline(C, 1, 38, 38, alloc_pooled=False)
    pipe.decode(cls, regr)
    res["sort_nms_single_panel"] = time_ms(lambda: pipe.sort_nms(), iters=50, flush=False)
    res["sort_nms_single_panel_cold_l2"] = time_ms(lambda: pipe.sort_nms(), iters=30, flush=True)
    cls, regr = tile_maps(1, 100, 100)
    pipe = ProposalPipeline(C, 1, 100, 100, alloc_pooled=False)
    pipe.decode(cls, regr)
    res["sort_nms_90k_candidates"] = time_ms(lambda: pipe.sort_nms(), iters=20, flush=False)
    del pipe

    # secondary shapes: the reference's DEFAULT 12-anchor config and a 600x800 panel (38x50 map)
    for tag, (Hh, Ww, scales) in {"12anchors_38x38": (38, 38, (64, 128, 256, 512)),
                                  "9anchors_38x50": (38, 50, (128, 256, 512))}.items():
        C2 = S.HotPathConfig(scales)
        cls, regr = tile_maps(64, Hh, Ww, C2.num_anchors)
        pipe = ProposalPipeline(C2, 64, Hh, Ww, alloc_pooled=False)
        pipe.decode(cls, regr)
        res["sort_nms_B64_" + tag] = time_ms(lambda: pipe.sort_nms())
        cls1, regr1 = tile_maps(1, Hh, Ww, C2.num_anchors)
        pipe1 = ProposalPipeline(C2, 1, Hh, Ww, alloc_pooled=False)
        pipe1.decode(cls1, regr1)
        res["sort_nms_single_" + tag] = time_ms(lambda: pipe1.sort_nms(), iters=30, flush=False)
        del pipe, pipe1

    # ---- K3 ----------------------------------------------------------------------------
    G = 20
    for B in (64, 512):
        gt = np.zeros((B, G, 4)); bg = np.zeros((B, G), np.uint8)
        for b in range(B):
            img = S.gt_figures(b, G)
            for k, bb in enumerate(img["bboxes"]):
                gt[b, k] = [bb["x1"], bb["x2"], bb["y1"], bb["y2"]]
        gt_d = torch.from_numpy(gt).cuda(); bg_d = torch.from_numpy(bg).cuda()
        cnt_d = torch.full((B,), G, dtype=torch.int32, device="cuda")
        wh_d = torch.tensor([[600.0, 600.0]] * B, dtype=torch.float64, device="cuda")
        nbytes = B * (G * 32 + 10 * 9 * 38 * 38 * 8)
        for tag, kw in (("", {}), ("_nhwc", dict(layout=1, regr_scale=4.0))):
            tb = RpnTargetBatch(C, B, G, 38, 38, **kw)
            t = time_ms(lambda: tb.run(gt_d, bg_d, cnt_d, wh_d))
            res["rpn_targets%s_B%d" % (tag, B)] = dict(
                t, algorithmic_bytes=nbytes, gbs=nbytes / t["p50_ms"] / 1e6,
                frac_of_measured_peak=nbytes / t["p50_ms"] / 1e6 / pk,
                note="one launch (fill team + compute team + last-CTA finalize), pre-allocated outputs, L2 flushed")
            del tb

    # ---- a4, batched: 64 panels x 300 RoIs x 20 figures straight from the detection records ----
    from rock_art_radnet_b200.rpn import RoiTargetBatch, gt_feature_cells
    B = 64
    cls, regr = tile_maps(B)
    pipe = ProposalPipeline(C, B, 38, 38, alloc_pooled=False)
    pipe.decode(cls, regr)
    pipe.sort_nms()
    gtc = np.zeros((B, G, 4)); gcl = np.zeros((B, G), np.int32)
    for b in range(B):
        a_, c_ = gt_feature_cells(S.gt_figures(b, G, classes=("boat", "human")), C, C.class_mapping)
        gtc[b], gcl[b] = a_, c_
    rt = RoiTargetBatch(C, C.class_mapping, B, 300, G)
    gtc_d, gcl_d = torch.from_numpy(gtc).cuda(), torch.from_numpy(gcl).cuda()
    cnt_d = torch.full((B,), G, dtype=torch.int32, device="cuda")
    res["roi_targets_batch_B64"] = dict(time_ms(lambda: rt.run(gtc_d, gcl_d, cnt_d, det=pipe.records)),
                                        note="radnet_roi_targets_batch, one CTA per panel, L2 flushed")
    res["roi_targets_batch_B64_graph"] = time_graph_ms(lambda: rt.run(gtc_d, gcl_d, cnt_d, det=pipe.records))
    del pipe, rt

    # ---- a4 ----------------------------------------------------------------------------
    import rock_art_radnet_b200 as R
    img = S.gt_figures(0, 20, classes=("boat", "human"))
    cls1, regr1 = S.rpn_maps(0)
    Rb = R.rpn_to_roi(cls1, regr1, C, max_boxes=300, overlap_thresh=0.7)
    import time
    R.calc_iou(Rb, img, C, C.class_mapping)
    t0 = time.perf_counter()
    for _ in range(20):
        R.calc_iou(Rb, img, C, C.class_mapping)
    res["calc_iou_dropin_wall_ms"] = (time.perf_counter() - t0) / 20 * 1e3
    t0 = time.perf_counter()
    for _ in range(20):
        R.rpn_to_roi(cls1, regr1, C, max_boxes=300, overlap_thresh=0.7)
    res["rpn_to_roi_dropin_wall_ms"] = (time.perf_counter() - t0) / 20 * 1e3
    np.random.seed(0)
    R.calc_region_props(C, img, 600, 600, 600, 600, S.resnet50_map_size)
    t0 = time.perf_counter()
    for _ in range(10):
        R.calc_region_props(C, img, 600, 600, 600, 600, S.resnet50_map_size)
    res["calc_region_props_dropin_wall_ms"] = (time.perf_counter() - t0) / 10 * 1e3

    # ---- f1/f2: detection post-processing (K5+K6 fused, K7, cross-image NMS) -------------------
    from rock_art_radnet_b200 import detect as DT
    from rock_art_radnet_b200.pipeline import DetectionPipeline

    def head_outputs(B, R, seed, frac_fg=0.25):
        """Random classifier-head outputs: about frac_fg of the RoIs clear the 0.7 threshold."""
        g = torch.Generator(device="cuda").manual_seed(seed)
        logits = torch.randn((B, R, 7), device="cuda", generator=g) * 1.5
        logits[..., 6] += 1.0
        boost = torch.rand((B, R, 1), device="cuda", generator=g) < frac_fg
        pick = torch.randint(0, 6, (B, R, 1), device="cuda", generator=g)
        logits.scatter_add_(2, pick, boost.float() * 6.0)
        return torch.softmax(logits, dim=-1).contiguous(), (torch.randn((B, R, 24), device="cuda", generator=g) * 0.8).contiguous()

    for B in (1, 64, 512):
        cls, regr = tile_maps(B)
        dp = DetectionPipeline(C, B, 38, 38, alloc_pooled=False)
        dp.decode(cls, regr)
        dp.sort_nms()
        P_cls, P_regr = head_outputs(B, 300, B)
        ratio = torch.ones((B,), dtype=torch.float64, device="cuda")
        origin = torch.zeros((B, 2), dtype=torch.int32, device="cuda")
        t = time_graph_ms(lambda: dp.classify(P_cls, P_regr, ratio=ratio, origin=origin))
        hdr = dp.class_records.header.cpu().numpy()
        res["classify_nms_B%d" % B] = dict(t, kept_per_tile=float(hdr[:, 0].mean()), candidates_per_tile=300,
                                            degenerate=int(hdr[:, 2].sum()))
        dec_out = DT.ClassRecords(B, 300, "cuda")
        t = time_graph_ms(lambda: DT.classify_decode(P_cls, P_regr, C, det=dp.records, out=dec_out))
        res["classify_decode_B%d" % B] = t
        if B == 64:
            # BASELINE configs[3]: 36 tiles per 1600-px panel -> one merge per panel; here 1 panel of 36 and 16 of 4
            tiles36 = DT.ClassRecords(36, 300, "cuda", raw=dp.class_records.raw[:36].clone())
            ws = torch.empty((64 << 20,), dtype=torch.uint8, device="cuda")

            def bench_merge(tag, fn_make):
                out_holder = {}

                def fn():
                    out_holder["o"] = fn_make(out_holder.get("o"))
                fn()
                res[tag] = time_graph_ms(fn, reps=10)
                return out_holder["o"]

            o = bench_merge("final_nms_1x36_tiles", lambda o: DT.final_nms_records(tiles36, 1, 36, 7, out=o, ws=ws))
            res["final_nms_1x36_tiles"].update(boxes_in=int(hdr[:36, 0].sum()), clusters_out=int(o.header[0, 0].item()))
            merged = bench_merge("final_nms_16x4_tiles", lambda o: DT.final_nms_records(dp.class_records, 16, 4, 7, out=o, ws=ws))
            res["final_nms_16x4_tiles"].update(boxes_in=int(hdr[:, 0].sum()), clusters_out=int(merged.header[:, 0].sum().item()))
            o = bench_merge("class_nms_cross_image_16", lambda o: DT.class_nms(merged, 1, 16, 7, 0.4, out=o, ws=ws))
            res["class_nms_cross_image_16"].update(kept=int(o.header[0, 0].item()), capacity=16 * merged.max_det)
            small = DT.ClassRecords(3, 300, "cuda", raw=dp.class_records.raw[:3].clone())
            o = bench_merge("class_nms_matrix_3x300", lambda o: DT.class_nms(small, 1, 3, 7, 0.4, out=o, ws=ws))
            res["class_nms_matrix_3x300"].update(kept=int(o.header[0, 0].item()))
        del dp

    # ---- K4 ----------------------------------------------------------------------------
    for (B, Cn, pool, tag, Hh, Ww) in ((64, 1024, 14, "resnet50", 38, 38), (64, 512, 7, "vgg16", 38, 38),
                                       (32, 1024, 14, "resnet50_38x50", 38, 50)):
        cls, regr = tile_maps(B, Hh, Ww)
        feat = torch.randn((B, Hh, Ww, Cn), dtype=torch.float32, device="cuda")
        pipe = ProposalPipeline(C, B, Hh, Ww, channels=Cn, pool_size=pool)
        pipe.decode(cls, regr)
        pipe.sort_nms()
        kept = int(pipe.records.counts.sum().item())
        nbytes = B * Hh * Ww * Cn * 4 + kept * 16 + kept * pool * pool * Cn * 4
        from rock_art_radnet_b200 import _lib
        # automatic choice first (form 0, lockstep tuned at the first call), then the fixed variants it chooses between
        for form, bands, cluster, every, ftag in ((0, 0, -1, 0, ""), (1, 0, 0, 0, "_free_running"), (1, 0, 0, 1, "_cta_lockstep"),
                                                  (1, 0, 2, 2, "_pair_lockstep"), (2, 2, -1, 0, "_2_bands"), (2, 3, -1, 0, "_3_bands")):
            _lib.set_option("roipool_form", form)
            _lib.set_option("roipool_bands", bands)
            _lib.set_option("roipool_cluster", cluster)
            _lib.set_option("roipool_sync_every", every)
            t = time_ms(lambda: pipe.pool(feat), iters=10)
            res["roi_pool_%s_B%d%s" % (tag, B, ftag)] = dict(t, algorithmic_bytes=nbytes, gbs=nbytes / t["p50_ms"] / 1e6,
                                                           frac_of_measured_peak=nbytes / t["p50_ms"] / 1e6 / pk)
        _lib.set_option("roipool_cluster", -1)
        _lib.set_option("roipool_sync_every", 0)
        _lib.set_option("roipool_form", 0)
        _lib.set_option("roipool_bands", 0)
        del pipe, feat
        torch.cuda.empty_cache()

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
