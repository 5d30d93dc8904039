"""GPU parity of the detection post-processing through the C ABI:
radnet_classify_decode / radnet_classify_nms / radnet_class_nms / radnet_final_nms and the
`RADNet` mirror class, against the reference-generated goldens and the CPU oracle.
Bar: bit-exact (classes, int boxes, float32 scores, order)."""
import numpy as np
import pytest

from conftest import _load_npz
from oracle import detect_oracle as DO
from oracle import radnet_oracle as O
from oracle.make_golden_detect import F1_CASES, F2_CASES, dicts_to_arrays, f1_inputs, f2_inputs
from rock_art_radnet_b200 import synthetic as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device; there is no CPU fallback"
    from rock_art_radnet_b200 import RADNet as RN
    from rock_art_radnet_b200 import detect as DT
    return RN, DT, torch


def _config():
    C = S.HotPathConfig()
    C.tile_size, C.tile_overlap, C.include_full_img, C.max_n_tiles_train = 600, 200, False, 1
    return C


NAMES = {v: k for k, v in S.HotPathConfig().class_mapping.items()}


@pytest.mark.parametrize("name", sorted(F1_CASES))
def test_apply_spatial_pyramid_pooling_matches_reference_golden(mods, name):
    RN, DT, _ = mods
    g = _load_npz("f1_classify.npz")
    C = _config()
    R, model, F = f1_inputs(name, C)
    net = RN.RADNet(C, None, model, lambda x: x)
    bboxes, probs = net.apply_spatial_pyramid_pooling(R, F)
    order, cls, box, pr = dicts_to_arrays(bboxes, probs, C.class_mapping)
    assert np.array_equal(order, g[name + "/order"])
    assert np.array_equal(cls, g[name + "/cls"])
    assert np.array_equal(box, g[name + "/box"])
    assert np.array_equal(pr, g[name + "/prob"], equal_nan=True)
    assert all(isinstance(v, np.float32) for k in probs for v in probs[k])


def test_classify_decode_counters(mods):
    RN, DT, _ = mods
    C = _config()
    R, model, F = f1_inputs("random_55_extremes", C)
    rois = DO.pad_rois(R, C.n_rois)
    pc, pr = [], []
    for k in range(0, len(rois), C.n_rois):
        a, b = model.predict([F, rois[None, k:k + C.n_rois]])
        pc.append(a[0]); pr.append(b[0])
    pc, pr = np.concatenate(pc), np.concatenate(pr)
    rec = DT.classify_decode(pc[None], pr[None], C, rois=rois[None].astype(np.int32)).to_numpy()[0]
    cls, prob, box, n_fallback = DO.classify_decode(rois, pc, pr, C)
    keep = cls >= 0
    n = int(rec["header"][DT.H_NDET])
    assert n == keep.sum() and rec["header"][DT.H_NIN] == len(rois)
    assert np.array_equal(rec["entry"]["cls"][:n], cls[keep])
    assert np.array_equal(rec["entry"]["box"][:n], box[keep])
    assert np.array_equal(rec["entry"]["prob"][:n], prob[keep])
    assert np.array_equal(rec["entry"]["src"][:n], np.flatnonzero(keep))
    assert rec["header"][DT.H_NFALLBACK] == n_fallback and n_fallback > 0
    assert rec["header"][DT.H_NNEARTIE] == 0 and rec["header"][DT.H_NRANGE] == 0


@pytest.mark.parametrize("seed,n,ratio,origin", [(0, 300, 1.0, (0, 0)), (1, 290, 0.75, (400, 200)),
                                                 (2, 64, 600 / 799.0, (1234, 77)), (3, 300, 2.0, (0, 16))])
def test_classify_nms_matches_oracle(mods, seed, n, ratio, origin):
    """Fused decode + per-class NMS (0.2) + get_real_coordinates + tile offset, B = 3 tiles at once."""
    RN, DT, _ = mods
    C = _config()
    B = 3
    rois, pcs, prs = [], [], []
    for b in range(B):
        R = S.random_rois(10 * seed + b, n)[0]
        R[:, 2:] = np.maximum(R[:, 2:], 3)            # with regr_scale 0.2 no box can shrink to zero width
        model = S.RandomHeadModel(seed + 100 * b, C, sharp=2.0 + b, regr_scale=0.2) if b else \
            S.FakeDetectorModel(S.scene_objects(seed, 600, 600, n_obj=25, lo=40, hi=200), C)
        a, r = model.predict([np.zeros((1, 1, 1, 4), np.float32), R[None]])
        rois.append(R); pcs.append(a[0]); prs.append(r[0])
    rec = DT.classify_nms(np.stack(pcs), np.stack(prs), C, rois=np.stack(rois).astype(np.int32),
                          ratio=[ratio] * B, origin=[origin] * B).to_numpy()
    DT.check_records(rec, "test")
    total = 0
    for b in range(B):
        want = DO.tile_detections(rois[b], pcs[b], prs[b], C, ratio, origin, NAMES)
        got_b, got_p = DT.record_to_dicts(rec[b], NAMES)
        assert list(got_b) == list(want)
        for k in want:
            assert np.array_equal(got_b[k], want[k][0]) and np.array_equal(got_p[k], want[k][1])
            total += len(got_p[k])
    assert total > 0


def test_classify_nms_from_detection_records(mods):
    """RoIs taken straight from the K2 detection records (xyxy -> xywh), as in the batched pipeline."""
    RN, DT, torch = mods
    from rock_art_radnet_b200.pipeline import ProposalPipeline
    C = _config()
    B = 2
    pipe = ProposalPipeline(C, B, 38, 38, alloc_pooled=False)
    maps = [S.rpn_maps(40 + b) for b in range(B)]
    cls = torch.from_numpy(np.concatenate([m[0] for m in maps])).cuda()
    regr = torch.from_numpy(np.concatenate([m[1] for m in maps])).cuda()
    pipe.decode(cls, regr)
    pipe.sort_nms()
    dets = pipe.records.to_numpy()
    pcs, prs, rois = [], [], []
    for b in range(B):
        R = dets[b]["boxes"].copy()
        R[:, 2] -= R[:, 0]
        R[:, 3] -= R[:, 1]
        a, r = S.RandomHeadModel(b, C, sharp=3.0, regr_scale=0.1).predict([None, R[None]])
        pc = np.zeros((300, 7), np.float32); pr = np.zeros((300, 24), np.float32)
        pc[:len(R)] = a[0]; pr[:len(R)] = r[0]
        pcs.append(pc); prs.append(pr); rois.append(R)
    rec = DT.classify_nms(np.stack(pcs), np.stack(prs), C, det=pipe.records).to_numpy()
    for b in range(B):
        n = len(rois[b])
        want = DO.tile_detections(rois[b], pcs[b][:n], prs[b][:n], C, 1.0, (0, 0), NAMES)
        got_b, got_p = DT.record_to_dicts(rec[b], NAMES)
        assert list(got_b) == list(want)
        for k in want:
            assert np.array_equal(got_b[k], want[k][0]) and np.array_equal(got_p[k], want[k][1])


@pytest.mark.parametrize("name", sorted(F2_CASES))
def test_final_nms_matches_reference_golden(mods, name):
    RN, DT, _ = mods
    g = _load_npz("f2_final_nms.npz")
    b, p = f2_inputs(name)
    net = RN.RADNet(_config(), None, None, lambda x: x)
    nb, npb = net.final_nms(b, p)
    assert nb.dtype == g[name + "/box"].dtype and np.array_equal(nb, g[name + "/box"])
    assert npb.dtype == np.float32 and np.array_equal(npb, g[name + "/prob"])


@pytest.mark.parametrize("seed", range(6))
def test_final_nms_random_and_ties_vs_oracle(mods, seed):
    RN, DT, _ = mods
    rng = np.random.default_rng(seed)
    ncl, per = int(rng.integers(1, 30)), int(rng.integers(1, 60))
    b, p = S.clustered_boxes(100 + seed, ncl, per, 0.6, 1.0, ties=bool(seed % 2))
    thr, conf, navg = [(0.2, 0.8, 5), (0.5, 0.9, 3), (0.05, 0.7, 1)][seed % 3]
    net = RN.RADNet(_config(), None, None, lambda x: x)
    nb, npb = net.final_nms(b, p, obj_avg_threshold=thr, obj_confidence_threshold=conf, n_obj_avg=navg)
    ob, op = DO.final_nms(b, p, obj_avg_threshold=thr, obj_confidence_threshold=conf, n_obj_avg=navg)
    assert np.array_equal(nb, ob) and np.array_equal(npb, op)
    assert net.final_nms(np.zeros((0, 4)), np.zeros((0,))) == []
    with pytest.raises(AssertionError):
        net.final_nms(np.array([[5, 5, 5, 9]]), np.array([0.9], dtype=np.float32))


@pytest.mark.parametrize("n_in,per", [(3, 100), (5, 400)])
def test_class_nms_over_concatenated_records(mods, n_in, per):
    """Cross-image NMS at 0.4 over several records (RADNet.py:695-716): the bit-matrix kernel up to
    1024 boxes, the cluster kernel in plain-NMS mode beyond."""
    RN, DT, torch = mods
    rng = np.random.default_rng(n_in)
    groups, all_c, all_p, all_b = [], [], [], []
    for j in range(n_in):
        b, p = S.clustered_boxes(200 + j, 10, per // 10, 0.5, 1.0)
        c = rng.integers(0, 6, len(p)).astype(np.int32)
        c = np.sort(c) if j % 2 else c            # grouped and ungrouped inputs
        groups.append((c, p, b))
        all_c.append(c); all_p.append(p); all_b.append(b)
    rec_in = DT.ClassRecords.from_arrays(groups, per, torch.device("cuda"))
    out = DT.class_nms(rec_in, 1, n_in, 7, 0.4, max_boxes=50).to_numpy()
    DT.check_records(out, "test")
    c, p, b = np.concatenate(all_c), np.concatenate(all_p), np.concatenate(all_b)
    got_b, got_p = DT.record_to_dicts(out[0], NAMES)
    first = []
    for v in c:
        if v not in first:
            first.append(int(v))
    assert [NAMES[v] for v in first] == list(got_b)
    for v in first:
        wb, wp = O.non_max_suppression_fast(b[c == v], p[c == v], overlap_thresh=0.4, max_boxes=50)
        assert np.array_equal(got_b[NAMES[v]], wb) and np.array_equal(got_p[NAMES[v]], wp)


@pytest.mark.parametrize("name", sorted(S.PREDICT_CASES))
def test_predict_matches_reference_golden(mods, name):
    pytest.importorskip("cv2")
    RN, DT, _ = mods
    g = _load_npz("f2_predict.npz")
    C, images, make_models = S.predict_case(name)
    m_rpn, m_det = make_models()
    dets = RN.RADNet(C, m_rpn, m_det, lambda x: x).predict(images)
    cls = np.asarray([C.class_mapping[d['class']] for d in dets], dtype=np.int64)
    prob = np.asarray([d['prob'] for d in dets], dtype=np.float32)
    box = np.asarray([[d['x1'], d['y1'], d['x2'], d['y2']] for d in dets], dtype=np.int64).reshape(-1, 4)
    assert np.array_equal(cls, g[name + "/cls"])
    assert np.array_equal(prob, g[name + "/prob"])
    assert np.array_equal(box, g[name + "/box"])
    assert [m_rpn.calls, m_det.calls] == list(g[name + "/calls"])
    assert all(isinstance(d['prob'], np.float32) and isinstance(d['x1'], np.int64) for d in dets)


def test_predict_degenerate_box_raises_like_the_reference(mods):
    """A head that shrinks boxes to zero width makes the reference's NMS assert (rpn.py:400-401)."""
    RN, DT, _ = mods
    C = _config()
    R = S.random_rois(0, 40)[0]
    pc = np.zeros((1, 40, 7), np.float32); pc[..., 0] = 0.9; pc[..., 6] = 0.1
    pr = np.zeros((1, 40, 24), np.float32); pr[..., 2] = -30.0 * 4.0          # tw = -30 -> w rounds to 0
    rec = DT.classify_nms(pc, pr, C, rois=R[None].astype(np.int32)).to_numpy()
    assert rec["header"][0, DT.H_NDEGEN] == 40
    with pytest.raises(AssertionError):
        DT.check_records(rec, "test")


def test_repeated_launches_are_bitwise_identical(mods):
    """The cluster kernel pipelines a finaliser warp against the search warps and packs through a
    last-CTA-done counter; any race there would show up as run-to-run differences."""
    RN, DT, torch = mods
    dev = torch.device("cuda")
    groups = []
    rng = np.random.default_rng(7)
    for j in range(12):
        b, p = S.clustered_boxes(300 + j, 25, 8, 0.6, 1.0)
        groups.append((rng.integers(0, 6, len(p)).astype(np.int32), p, b))
    rec_in = DT.ClassRecords.from_arrays(groups, 200, dev)
    first_a = first_b = first_c = None
    for _ in range(25):
        a = DT.final_nms_records(rec_in, 3, 4, 7).raw.cpu().numpy()
        b = DT.class_nms(rec_in, 2, 6, 7, 0.4).raw.cpu().numpy()            # 1200 boxes: cluster kernel, NMS mode
        c = DT.class_nms(rec_in, 3, 4, 7, 0.3).raw.cpu().numpy()            # 800 boxes: bit-matrix kernel
        if first_a is None:
            first_a, first_b, first_c = a, b, c
        assert np.array_equal(a, first_a) and np.array_equal(b, first_b) and np.array_equal(c, first_c)
    assert (first_a.view(np.int32)[:, 0] > 0).all()


def _oracle_panel(C, tiles, rois, pcs, prs):
    """Reference flow for one tiled panel: per-tile detections -> final_nms per class -> NMS 0.4."""
    bbox_total, probs_total = {}, {}
    for t, tile in enumerate(tiles):
        n = len(rois[t])
        for name, (bx, pb) in DO.tile_detections(rois[t], pcs[t][:n], prs[t][:n], C, 1.0, (tile[0], tile[1]), NAMES).items():
            bbox_total.setdefault(name, []).extend(bx.tolist())
            probs_total.setdefault(name, []).extend(list(pb))
    out = {}
    for name in bbox_total:
        nb, npb = DO.final_nms(np.array(bbox_total[name]), np.array(probs_total[name]))
        out[name] = O.non_max_suppression_fast(nb, npb, overlap_thresh=0.4)
    return out


@pytest.mark.parametrize("width,height,n_panels", [(1000, 800, 2), (1600, 1600, 1)])
def test_tiled_panel_pipeline_end_to_end(mods, width, height, n_panels):
    """BASELINE configs[3] shape: tiles of a tiled panel through decode -> NMS -> head decode ->
    per-class NMS -> (gather) -> final_nms -> NMS 0.4, all on the device, against the oracle."""
    RN, DT, torch = mods
    from rock_art_radnet_b200 import sharding
    from rock_art_radnet_b200.pipeline import DetectionPipeline
    C = _config()
    tiles = S.tiled_panel_tiles(width, height)
    T = len(tiles)
    shard = sharding.TiledPanelSharder(n_panels, T, rank=0, world=1)
    B = shard.per_rank
    pipe = DetectionPipeline(C, B, 38, 38, alloc_pooled=False)
    maps = [S.rpn_maps(1000 * p + t) for p in range(n_panels) for t in range(T)]
    pipe.decode(torch.from_numpy(np.concatenate([m[0] for m in maps])).cuda(),
                torch.from_numpy(np.concatenate([m[1] for m in maps])).cuda())
    pipe.sort_nms()
    dets = pipe.records.to_numpy()
    rois, pcs, prs = [], [], []
    for i in range(B):
        R = dets[i]["boxes"].copy()
        R[:, 2] -= R[:, 0]
        R[:, 3] -= R[:, 1]
        pc, pr = S.tiled_panel_head_outputs(C, i // T, tiles[i % T], R)
        rois.append(R); pcs.append(pc); prs.append(pr)
    origin = torch.tensor([[tiles[i % T][0], tiles[i % T][1]] for i in range(B)], dtype=torch.int32, device="cuda")
    ratio = torch.ones((B,), dtype=torch.float64, device="cuda")
    pipe.classify(torch.from_numpy(np.stack(pcs)).cuda(), torch.from_numpy(np.stack(prs)).cuda(), ratio=ratio, origin=origin)
    glob, _ = shard.gather_tiles(pipe.class_records.raw)
    tile_rec = DT.ClassRecords(n_panels * T, pipe.max_boxes, glob.device, raw=glob)
    merged = DT.final_nms_records(tile_rec, n_panels, T, pipe.n_cls)
    final = DT.class_nms(merged, n_panels, 1, pipe.n_cls, 0.4).to_numpy()
    DT.check_records(tile_rec.to_numpy(), "tiles")
    DT.check_records(final, "final")
    n_clustered = 0
    for p in range(n_panels):
        sl = slice(p * T, (p + 1) * T)
        want = _oracle_panel(C, tiles, rois[sl], pcs[sl], prs[sl])
        got_b, got_p = DT.record_to_dicts(final[p], NAMES)
        assert list(got_b) == list(want) and len(want) > 0
        for k in want:
            assert np.array_equal(got_b[k], want[k][0]) and np.array_equal(got_p[k], want[k][1])
        n_clustered += int((merged.to_numpy()[p]["entry"]["aux"] > 1).sum())
    assert n_clustered > 0          # neighbouring tiles really produced multi-member clusters


def test_scalar_helpers_on_the_device(mods):
    """utils.iou and RADNet.get_real_coordinates (API-parity helpers) evaluated by the library."""
    RN, DT, _ = mods
    from rock_art_radnet_b200.utils import iou, iou_pairs
    rng = np.random.default_rng(0)
    a = np.sort(rng.uniform(0, 50, (300, 2, 2)), axis=1).transpose(0, 2, 1).reshape(300, 4)[:, [0, 2, 1, 3]]
    b = np.sort(rng.uniform(0, 50, (300, 2, 2)), axis=1).transpose(0, 2, 1).reshape(300, 4)[:, [0, 2, 1, 3]]
    got = iou_pairs(a, b)
    assert all(got[i] == O.iou(list(a[i]), list(b[i])) for i in range(300)) and (got > 0).any()
    assert iou([0, 0, 0, 5], [0, 0, 5, 5]) == 0.0 and iou([0, 0, 4, 4], [2, 2, 6, 6]) == O.iou([0, 0, 4, 4], [2, 2, 6, 6])
    net = RN.RADNet(_config(), None, None, lambda x: x)
    for ratio in (1.0, 0.75, 0.3, 600 / 799.0, 1.7, 2.0):
        for v in rng.integers(0, 3000, (20, 4)):
            assert net.get_real_coordinates(ratio, *[np.int64(t) for t in v]) == DO.get_real_coordinates(ratio, *v)


def test_capacity_limits_are_flagged_not_truncated(mods):
    RN, DT, torch = mods
    C = _config()
    dev = torch.device("cuda")
    # (1) decode into a record with fewer slots than detections
    R = S.random_rois(3, 300)[0]
    a, r = S.RandomHeadModel(3, C, sharp=4.0).predict([None, R[None]])
    small = DT.ClassRecords(1, 8, dev)
    rec = DT.classify_decode(a, r, C, rois=R[None].astype(np.int32), out=small).to_numpy()
    assert rec["header"][0, DT.H_NDET] == -2
    with pytest.raises(Exception) as ei:
        DT.check_records(rec, "test")
    assert "RADNET_E_UNSUPPORTED" in str(ei.value)
    # (2) tile merge / per-class NMS into too small an output
    b, p = S.clustered_boxes(1, 12, 6, 0.7, 1.0)
    rec_in = DT.ClassRecords.from_arrays([(np.zeros(len(p), np.int32), p, b)], len(p), dev)
    assert DT.final_nms_records(rec_in, 1, 1, 7, out_max_det=3).to_numpy()["header"][0, DT.H_NDET] == -2
    assert DT.class_nms(rec_in, 1, 1, 7, 0.4, out_max_det=3).to_numpy()["header"][0, DT.H_NDET] == -2
    assert DT.final_nms_records(rec_in, 1, 1, 7, out_max_det=12).to_numpy()["header"][0, DT.H_NDET] == 12
    # (3) more than 4096 boxes of one class in one image
    b, p = S.clustered_boxes(2, 50, 100, 0.7, 1.0)
    big = DT.ClassRecords.from_arrays([(np.zeros(len(p), np.int32), p, b)], len(p), dev)
    assert DT.final_nms_records(big, 1, 1, 7).to_numpy()["header"][0, DT.H_NDET] == -2
    # 4096 exactly is fine
    ok = DT.ClassRecords.from_arrays([(np.zeros(4096, np.int32), p[:4096], b[:4096])], 4096, dev)
    got = DT.final_nms_records(ok, 1, 1, 7).to_numpy()[0]
    ob, op = DO.final_nms(b[:4096], p[:4096])
    n = int(got["header"][DT.H_NDET])
    assert n == len(op) and np.array_equal(got["entry"]["box"][:n], ob) and np.array_equal(got["entry"]["prob"][:n], op)


def test_maximum_sizes_1024_rois_32_classes(mods):
    RN, DT, _ = mods
    C = _config()
    n_cls = 32
    names = {i: "c%d" % i for i in range(n_cls)}
    C.classifier_regr_std = [8.0, 8.0, 4.0, 4.0]
    R = S.random_rois(9, 1024)[0]
    R[:, 2:] = np.maximum(R[:, 2:], 3)
    a, r = S.RandomHeadModel(9, C, n_cls=n_cls, sharp=5.0, regr_scale=0.2).predict([None, R[None]])
    rec = DT.classify_nms(a, r, C, rois=R[None].astype(np.int32), ratio=[0.6], origin=[(7, 9)]).to_numpy()
    DT.check_records(rec, "test")
    want = DO.tile_detections(R, a[0], r[0], C, 0.6, (7, 9), names)
    got_b, got_p = DT.record_to_dicts(rec[0], names)
    assert list(got_b) == list(want) and len(want) >= 20
    for k in want:
        assert np.array_equal(got_b[k], want[k][0]) and np.array_equal(got_p[k], want[k][1])


def test_in_count_limits_the_records_used(mods):
    RN, DT, torch = mods
    dev = torch.device("cuda")
    groups = []
    for j in range(6):
        b, p = S.clustered_boxes(400 + j, 6, 5, 0.7, 1.0)
        groups.append((np.full(len(p), j % 3, np.int32), p, b))
    rec_in = DT.ClassRecords.from_arrays(groups, 30, dev)
    # two segments of three records; the second segment uses only its first record
    out = DT.final_nms_records(rec_in, 2, 3, 7, in_count=[3, 1]).to_numpy()
    for seg, used in ((0, [0, 1, 2]), (1, [3])):
        got_b, got_p = DT.record_to_dicts(out[seg], NAMES)
        cls = np.concatenate([groups[j][0] for j in used])
        pb = np.concatenate([groups[j][1] for j in used])
        bx = np.concatenate([groups[j][2] for j in used])
        order = []
        for c in cls:
            if c not in order:
                order.append(int(c))
        assert [NAMES[c] for c in order] == list(got_b)
        for c in order:
            ob, op = DO.final_nms(bx[cls == c], pb[cls == c])
            assert np.array_equal(got_b[NAMES[c]], ob) and np.array_equal(got_p[NAMES[c]], op)
    nm = DT.class_nms(rec_in, 2, 3, 7, 0.4, in_count=[2, 3]).to_numpy()
    assert nm["header"][0, DT.H_NIN] == 60 and nm["header"][1, DT.H_NIN] == 90


def test_predict_views_with_different_proposal_counts(mods):
    """Small feature maps: every tile keeps a different number (< 300) of proposals, so the padded RoI
    count differs from view to view; the merge still sees one array of equally sized records."""
    pytest.importorskip("cv2")
    RN, DT, _ = mods
    C, images, _ = S.predict_case("w1000_h800_tiles")
    objs = S.scene_objects(7, 1000, 800, n_obj=14)

    def models():
        return S.FakeRpnModel(seed=7, map_size=lambda w, h: (9, 8)), S.FakeDetectorModel(objs, C)

    m_rpn, m_det = models()
    got = RN.RADNet(C, m_rpn, m_det, lambda x: x).predict(images)
    o_rpn, o_det = models()
    fmt = RN.RADNet(C, None, None, lambda x: x).format_img
    want = DO.RADNetOracle(C, o_rpn, o_det, fmt).predict(images)
    assert m_det.calls == o_det.calls and m_det.calls < 6 * 15          # fewer chunks than 300 proposals would need
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a['class'] == b['class'] and a['prob'] == b['prob']
        assert (a['x1'], a['y1'], a['x2'], a['y2']) == (b['x1'], b['y1'], b['x2'], b['y2'])


def test_empty_inputs(mods):
    """Nothing clears the threshold / no boxes at all: every stage returns empty records, the mirror class
    returns what the reference returns ({} / [] / [])."""
    RN, DT, torch = mods
    C = _config()
    dev = torch.device("cuda")
    R = S.random_rois(1, 40)[0]
    pc = np.zeros((1, 40, 7), np.float32)
    pc[..., 6] = 1.0                                            # everything is 'bg'
    pr = np.zeros((1, 40, 24), np.float32)
    rec = DT.classify_nms(pc, pr, C, rois=R[None].astype(np.int32))
    host = rec.to_numpy()
    DT.check_records(host, "empty")
    assert host["header"][0, DT.H_NDET] == 0 and host["header"][0, DT.H_NCLASSES] == 0 and (host["order"][0] == -1).all()
    pc[..., 6], pc[..., 0] = 0.4, 0.6                           # best class below bbox_threshold
    assert DT.classify_decode(pc, pr, C, rois=R[None].astype(np.int32)).to_numpy()["header"][0, DT.H_NDET] == 0
    tiles = DT.ClassRecords(6, 300, dev)                        # six tiles without detections
    merged = DT.final_nms_records(tiles, 2, 3, 7)
    final = DT.class_nms(merged, 1, 2, 7, 0.4).to_numpy()
    DT.check_records(merged.to_numpy(), "empty")
    assert (merged.to_numpy()["header"][:, DT.H_NDET] == 0).all() and final["header"][0, DT.H_NDET] == 0
    assert DT.record_to_dicts(final[0], NAMES) == ({}, {})

    class NoRois:
        def predict(self, inputs):
            raise AssertionError("the detector must not be called without RoIs")

    net = RN.RADNet(C, None, NoRois(), lambda x: x)
    assert net.apply_spatial_pyramid_pooling(np.zeros((0, 4), np.int64), None) == ({}, {})
    assert net.final_nms(np.zeros((0, 4)), np.zeros((0,))) == []
    from rock_art_radnet_b200.rpn import non_max_suppression_fast
    assert non_max_suppression_fast(np.zeros((0, 4)), np.zeros((0,))) == []
