"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * NMS kept rows / kept indices / class ids: bit-exact when both sides get the same decoded input;
  * decoded boxes: |err| <= 1e-5*|ref| + 1e-5*stride (fp32; SURVEY §7 hard part 3), scores likewise 1e-5 relative;
  * end-to-end from raw logits: mismatching detections < 1e-4 of all detections.
"""
import pytest
import torch

from oracle import postprocess_ref as R

pytestmark = pytest.mark.gpu


def _nms_both(sarpost, y_cpu, dev, **kw):
    rows, idx = sarpost.non_max_suppression(y_cpu.to(dev), return_index=True, **kw)
    ref_rows, ref_idx = R.non_max_suppression_ref(y_cpu, return_index=True, **kw)
    return rows, idx, ref_rows, ref_idx


def _assert_same(rows, idx, ref_rows, ref_idx, nc):
    assert len(rows) == len(ref_rows)
    for b, (r, i, rr, ri) in enumerate(zip(rows, idx, ref_rows, ref_idx)):
        assert tuple(r.shape) == tuple(rr.shape), f"image {b}: {tuple(r.shape)} vs {tuple(rr.shape)}"
        i = i.cpu().long()
        assert torch.equal(i // nc, ri[:, 0]), f"image {b}: kept anchor indices differ"
        assert torch.equal(i % nc, ri[:, 1]), f"image {b}: kept class ids differ"
        assert torch.equal(r.cpu(), rr), f"image {b}: rows differ"


@pytest.mark.parametrize("bs,na,nc,nm,kw", [
    (1, 8400, 1, 0, dict(conf_thres=0.25, iou_thres=0.7)),
    (3, 8400, 6, 0, dict(conf_thres=0.25, iou_thres=0.7)),
    (2, 8400, 1, 262, dict(conf_thres=0.25, iou_thres=0.7)),                      # JDE layout
    (2, 8400, 6, 5, dict(conf_thres=0.001, iou_thres=0.7, multi_label=True)),      # > max_nms after expansion
    (2, 6000, 6, 0, dict(conf_thres=0.001, iou_thres=0.7, multi_label=True, max_nms=3000)),
    (2, 5000, 4, 0, dict(conf_thres=0.05, iou_thres=0.6, agnostic=True)),          # thr whose float rounds up
    (2, 5000, 5, 3, dict(conf_thres=0.05, iou_thres=0.45, classes=[0, 3])),
    (2, 5000, 5, 0, dict(conf_thres=0.05, iou_thres=0.45, classes=[1], multi_label=True)),
    (1, 1000, 2, 0, dict(conf_thres=0.9999, iou_thres=0.5)),                       # (almost) nothing passes
    (1, 37, 3, 0, dict(conf_thres=0.0, iou_thres=0.5, max_det=5)),                 # tiny, ragged tile
    (2, 3000, 1, 0, dict(conf_thres=0.01, iou_thres=0.7, max_det=1000, nc=1)),
    (1, 70000, 1, 0, dict(conf_thres=0.001, iou_thres=0.7)),                       # top-k cut engaged (n > 30000)
])
def test_nms_decoded_bit_exact(sarpost, cuda, bs, na, nc, nm, kw):
    y = sarpost.synth.decoded_prediction(bs, na, nc, nm, seed=na + nc, score_pow=2.0 if kw["conf_thres"] < 0.01 else 4.0)
    kw = dict(kw, nc=nc)
    rows, idx, ref_rows, ref_idx = _nms_both(sarpost, y, cuda, **kw)
    _assert_same(rows, idx, ref_rows, ref_idx, nc)


@pytest.mark.parametrize("seed", [0, 1])
def test_nms_decoded_clustered_heavy_suppression(sarpost, cuda, seed):
    """Clustered boxes: thousands of candidates, most suppressed, many chunks walked by K4."""
    y = sarpost.synth.decoded_prediction(2, 12000, 3, 0, seed=seed, clustered=True, score_pow=1.0)
    for kw in (dict(conf_thres=0.05, iou_thres=0.7), dict(conf_thres=0.05, iou_thres=0.3, agnostic=True, max_det=100),
               dict(conf_thres=0.001, iou_thres=0.5, multi_label=True)):
        rows, idx, ref_rows, ref_idx = _nms_both(sarpost, y, cuda, **kw)
        _assert_same(rows, idx, ref_rows, ref_idx, 3)


def test_nms_ties_stable_order(sarpost, cuda):
    """Heavy score ties (scores quantised to 1/16): equal scores must resolve lower-index-first,
    including at the max_nms cut."""
    y = sarpost.synth.decoded_prediction(2, 9000, 2, 0, seed=11)
    y[:, 4:6] = (y[:, 4:6] * 16).floor() / 16 + 1 / 32
    for kw in (dict(conf_thres=0.01, iou_thres=0.7), dict(conf_thres=0.01, iou_thres=0.7, multi_label=True, max_nms=2500),
               dict(conf_thres=0.01, iou_thres=0.7, max_nms=100, max_det=50)):
        rows, idx, ref_rows, ref_idx = _nms_both(sarpost, y, cuda, **kw)
        _assert_same(rows, idx, ref_rows, ref_idx, 2)


def test_nms_all_scores_equal(sarpost, cuda):
    """Pathological whole-model random-init case: every score identical."""
    y = sarpost.synth.decoded_prediction(1, 40000, 1, 0, seed=5)
    y[:, 4] = 0.0124
    kw = dict(conf_thres=0.001, iou_thres=0.7)
    rows, idx, ref_rows, ref_idx = _nms_both(sarpost, y, cuda, **kw)
    _assert_same(rows, idx, ref_rows, ref_idx, 1)


def test_nms_known_answers(sarpost, cuda):
    """SURVEY §8c known-answer vectors (probed behaviour of the reference)."""
    def run(boxes_xyxy, scores, thr):
        b = torch.tensor(boxes_xyxy, dtype=torch.float32)
        xywh = torch.stack(((b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]), 1)
        y = torch.cat((xywh, torch.tensor(scores, dtype=torch.float32)[:, None]), 1).t()[None].contiguous()
        rows, idx = sarpost.non_max_suppression(y.to(cuda), conf_thres=0.1, iou_thres=thr, return_index=True)
        ref = R.non_max_suppression_ref(y, conf_thres=0.1, iou_thres=thr, return_index=True)
        assert torch.equal(idx[0].cpu().long(), ref[1][0][:, 0])
        return idx[0].cpu().tolist()
    assert run([[0, 0, 10, 10]] * 3, [0.5, 0.5, 0.5], 0.5) == [0]                       # identical boxes, equal scores
    assert run([[0, 0, 6, 1], [0, 0, 3.6000001, 1]], [0.9, 0.8], 0.6) == [0]            # IoU == 0.6f, double compare suppresses
    assert run([[0, 0, 6, 1], [0, 0, 3.6000001, 1]], [0.9, 0.8], 0.7) == [0, 1]
    assert run([[5, 5, 5, 5], [5, 5, 5, 5]], [0.9, 0.8], 0.5) == [0, 1]                 # zero-area: 0/0 never suppresses


def test_nms_fma_discriminating_pairs(sarpost, cuda):
    """Pairs whose IoU sits within an ulp of the threshold: an FMA-contracted IoU would flip them."""
    g = torch.Generator().manual_seed(123)
    n = 4000
    a = torch.rand(n, 2, generator=g) * 100
    wh = 10 + torch.rand(n, 2, generator=g) * 50
    boxes = []
    scores = []
    for i in range(n):
        x1, y1 = a[i].tolist()
        w, h = wh[i].tolist()
        boxes.append([x1 + 1000.0 * i, y1, x1 + w + 1000.0 * i, y1 + h])
        # partner shifted so IoU is close to 0.7: overlap fraction f with f/(2-f) = 0.7 -> f = 0.8235
        dx = w * (1 - 0.823529411)
        boxes.append([x1 + dx + 1000.0 * i, y1, x1 + dx + w + 1000.0 * i, y1 + h])
        scores += [0.9, 0.8]
    b = torch.tensor(boxes, dtype=torch.float32)
    xywh = torch.stack(((b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]), 1)
    y = torch.cat((xywh, torch.tensor(scores)[:, None]), 1).t()[None].contiguous()
    kw = dict(conf_thres=0.1, iou_thres=0.7, max_det=4096, max_wh=0)
    rows, idx, ref_rows, ref_idx = _nms_both(sarpost, y, cuda, **kw)
    _assert_same(rows, idx, ref_rows, ref_idx, 1)
    assert 2 * n * 0.2 < rows[0].shape[0] <= 4096


HEADS = [
    # (imgsz, strides, nc, embed_dim, state_classes, batch)
    (640, (8, 16, 32), 1, 256, 6, 2),      # cfg1/cfg2 SAR posture JDE head (no = 327)
    (640, (8, 16, 32), 6, 0, 0, 3),        # posture-as-class Detect head (no = 70)
    (320, (4, 8, 16, 32), 1, 16, 0, 2),    # P2 head, JDE without state
    ((96, 160), (8, 16, 32), 3, 0, 0, 2),  # rect
    ((88, 88), (8, 16, 32), 2, 8, 6, 2),   # 11x11 / odd levels: H*W*4 not 16-byte aligned -> LDG path
]


@pytest.mark.parametrize("imgsz,strides,nc,ed,sc,bs", HEADS)
def test_decode_matches_oracle(sarpost, cuda, imgsz, strides, nc, ed, sc, bs):
    shapes = sarpost.synth.level_shapes(imgsz, strides)
    levels = sarpost.synth.head_outputs(bs, shapes, nc, ed, sc, seed=3)
    spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
    y = sarpost.decode([x.to(cuda) for x in levels], spec).cpu()
    y_ref = R.decode_ref(levels, strides, nc, 16, ed, sc)
    assert y.shape == y_ref.shape
    st = torch.cat([torch.full((h * w,), float(s)) for (h, w), s in zip(shapes, strides)])
    # boxes: rtol 1e-5 / atol 1e-5*stride (tolerance stated by north_star, evaluated per SURVEY §7.3)
    err = (y[:, :4] - y_ref[:, :4]).abs()
    assert bool((err <= 1e-5 * y_ref[:, :4].abs() + 1e-5 * st).all()), err.max().item()
    # class probabilities and sigmoid state: 1e-5 relative (+1e-7 absolute floor)
    assert torch.allclose(y[:, 4:4 + nc], y_ref[:, 4:4 + nc], rtol=1e-5, atol=1e-7)
    if ed:
        assert torch.equal(y[:, 4 + nc:4 + nc + ed], y_ref[:, 4 + nc:4 + nc + ed])  # raw embedding: pure copy
    if sc:
        assert torch.allclose(y[:, 4 + nc + ed:], y_ref[:, 4 + nc + ed:], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("imgsz,strides,nc,ed,sc,bs", HEADS)
@pytest.mark.parametrize("mode", ["predict", "val", "val_multi"])
def test_fused_equals_decode_then_nms(sarpost, cuda, imgsz, strides, nc, ed, sc, bs, mode):
    """Fused kernels vs oracle NMS fed OUR decoded y: rows, indices and extras bit-exact
    (isolates candidate generation / compaction / extras gather from decode rounding)."""
    shapes = sarpost.synth.level_shapes(imgsz, strides)
    levels = [x.to(cuda) for x in sarpost.synth.head_outputs(bs, shapes, nc, ed, sc, seed=9)]
    spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
    kw = {"predict": dict(conf_thres=0.25, iou_thres=0.7), "val": dict(conf_thres=0.001, iou_thres=0.7),
          "val_multi": dict(conf_thres=0.001, iou_thres=0.7, multi_label=True)}[mode]
    rows, idx = sarpost.postprocess_fused(levels, spec, return_index=True, **kw)
    y = sarpost.decode(levels, spec).cpu()
    ref_rows, ref_idx = R.non_max_suppression_ref(y, nc=nc, return_index=True, **kw)
    _assert_same(rows, idx, ref_rows, ref_idx, nc)


def test_fused_tma_and_ldg_paths_agree(sarpost, cuda, monkeypatch):
    strides = (8, 16, 32)
    shapes = sarpost.synth.level_shapes(640, strides)
    levels = [x.to(cuda) for x in sarpost.synth.head_outputs(2, shapes, 6, 0, 0, seed=21)]
    spec = sarpost.HeadSpec(nc=6, strides=strides)
    kw = dict(conf_thres=0.001, iou_thres=0.7, multi_label=True)
    a = sarpost.postprocess_fused(levels, spec, **kw)
    monkeypatch.setenv("SARPOST_K1_FORCE_LDG", "1")
    b = sarpost.postprocess_fused(levels, spec, **kw)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


@pytest.mark.parametrize("imgsz,strides,nc,bs,conf", [
    (640, (8, 16, 32), 1, 8, 0.25), (640, (8, 16, 32), 6, 4, 0.25), (1280, (4, 8, 16, 32), 1, 1, 0.001)])
def test_end_to_end_mismatch_budget(sarpost, cuda, imgsz, strides, nc, bs, conf):
    """Raw logits -> detections: CUDA fused path vs oracle decode + oracle NMS.  Detections are matched
    by (anchor, class); a detection counts as a mismatch if it is missing on either side or its box /
    score is outside the decode tolerance.  Budget: < 1e-4 of all detections (north_star)."""
    shapes = sarpost.synth.level_shapes(imgsz, strides)
    total = bad = 0
    for seed in range(3):
        levels = sarpost.synth.head_outputs(bs, shapes, nc, 0, 0, seed=100 + seed, blobs=20)
        spec = sarpost.HeadSpec(nc=nc, strides=strides)
        rows, idx = sarpost.postprocess_fused([x.to(cuda) for x in levels], spec, conf_thres=conf, iou_thres=0.7,
                                              return_index=True)
        y_ref = R.decode_ref(levels, strides, nc)
        ref_rows, ref_idx = R.non_max_suppression_ref(y_ref, conf_thres=conf, iou_thres=0.7, nc=nc, return_index=True)
        for b in range(bs):
            ours = {int(k): r for k, r in zip(idx[b].cpu().tolist(), rows[b].cpu())}
            ref = {int(a) * nc + int(c): r for (a, c), r in zip(ref_idx[b].tolist(), ref_rows[b])}
            total += max(len(ours), len(ref))
            for k in set(ours) | set(ref):
                if k not in ours or k not in ref:
                    bad += 1
                elif not torch.allclose(ours[k], ref[k], rtol=1e-5, atol=1e-5 * max(strides)):
                    bad += 1
    assert total > 0
    assert bad <= max(1e-4 * total, 0), f"{bad} mismatching detections out of {total}"


def test_full_size_properties_cfg3(sarpost, cuda):
    """BASELINE cfg3 at full size (1280x1280 P2, 136 000 anchors, val thresholds, B=4): size-independent
    properties — descending scores, counts <= max_det, idempotence of NMS on its own output, no kept
    pair above the IoU threshold within a class."""
    strides = (4, 8, 16, 32)
    shapes = sarpost.synth.level_shapes(1280, strides)
    levels = sarpost.synth.head_outputs(4, shapes, 1, 256, 6, seed=3000, device=cuda, blobs=30)
    spec = sarpost.HeadSpec(nc=1, strides=strides, embed_dim=256, state_classes=6)
    rows = sarpost.postprocess_fused(levels, spec, conf_thres=0.001, iou_thres=0.7)
    assert len(rows) == 4
    for r in rows:
        assert 0 < r.shape[0] <= 300 and r.shape[1] == 268
        s = r[:, 4]
        assert bool((s[:-1] >= s[1:]).all())
        assert bool((s > 0.001).all())
        b = r[:, :4].cpu()
        keep = R.nms_ref(b, s.cpu(), 0.7)
        assert keep.tolist() == list(range(r.shape[0])), "NMS output is not a fixed point of NMS"


def test_merge_tiles_matches_oracle(sarpost, cuda):
    """Cross-tile merge = shift by origin + class-offset NMS per frame (SURVEY §8c definition)."""
    g = torch.Generator().manual_seed(4)
    origins = sarpost.dist.sahi_grid(4000, 3000)
    tpf, d, row_len, nf = origins.shape[0], 60, 8, 3
    dets = torch.zeros(nf * tpf, d, row_len)
    cnt = torch.randint(0, d + 1, (nf * tpf,), generator=g, dtype=torch.int32)
    xy = torch.rand(nf * tpf, d, 2, generator=g) * 560
    wh = 20 + torch.rand(nf * tpf, d, 2, generator=g) * 120
    dets[..., 0:2] = xy
    dets[..., 2:4] = xy + wh
    dets[..., 4] = torch.rand(nf * tpf, d, generator=g).sort(dim=1, descending=True).values
    dets[..., 5] = torch.randint(0, 3, (nf * tpf, d), generator=g).float()
    dets[..., 6:] = torch.randn(nf * tpf, d, 2, generator=g)
    org = origins.repeat(nf, 1)
    rows, idx = sarpost.merge_tiles(dets.to(cuda), cnt.to(cuda), org.to(cuda), tpf, iou_thres=0.5, max_det=500, return_index=True)
    for f in range(nf):
        cand, src = [], []
        for t in range(tpf):
            k = f * tpf + t
            n = int(cnt[k])
            r = dets[k, :n].clone()
            r[:, 0] += org[k, 0]; r[:, 2] += org[k, 0]; r[:, 1] += org[k, 1]; r[:, 3] += org[k, 1]
            cand.append(r)
            src.append(torch.arange(n) + t * d)
        x = torch.cat(cand)
        src = torch.cat(src)
        keep = R.nms_ref(x[:, :4] + x[:, 5:6] * 7680, x[:, 4], 0.5)[:500]
        assert torch.equal(rows[f].cpu(), x[keep])
        assert torch.equal(idx[f].cpu().long(), src[keep])


def test_host_entry_matches_device_entry(sarpost, cuda):
    strides = (8, 16, 32)
    shapes = sarpost.synth.level_shapes(640, strides)
    spec = sarpost.HeadSpec(nc=1, strides=strides, embed_dim=256, state_classes=6)
    levels = [x.pin_memory() for x in sarpost.synth.head_outputs(4, shapes, 1, 256, 6, seed=77)]
    kw = dict(conf_thres=0.25, iou_thres=0.7)
    dev_rows = sarpost.postprocess_fused([x.to(cuda) for x in levels], spec, **kw)
    ctx = sarpost.HostContext(0)
    host_rows = ctx.postprocess(levels, spec, **kw)
    h2d, d2h = ctx.last_traffic()
    assert h2d > 0 and d2h > 0
    for a, b in zip(dev_rows, host_rows):
        assert torch.equal(a.cpu(), b)
    ctx.close()


def test_errors_and_api_contract(sarpost, cuda):
    y = sarpost.synth.decoded_prediction(1, 100, 2, 0, seed=0)
    with pytest.raises(RuntimeError):
        sarpost.non_max_suppression(y)  # CPU tensor: no fallback
    with pytest.raises(AssertionError):
        sarpost.non_max_suppression(y.to(cuda), conf_thres=1.5)
    with pytest.raises(AssertionError):
        sarpost.non_max_suppression(y.to(cuda), iou_thres=-0.1)
    with pytest.raises(NotImplementedError):
        sarpost.non_max_suppression(y.to(cuda), rotated=True)
    yc = y.to(cuda)
    before = yc.clone()
    out = sarpost.non_max_suppression((yc, None), conf_thres=0.1)  # tuple input (ops.py:219-220)
    assert torch.equal(yc, before), "input must not be modified"
    assert isinstance(out, list) and len(out) == 1 and out[0].shape[1] == 6
    out[0][:, :4] *= 2  # callers mutate outputs in place (jde/predict.py:49)
    empty = sarpost.non_max_suppression(yc, conf_thres=1.0)
    assert empty[0].shape == (0, 6)
    # end-to-end (B, N, 6) branch (ops.py:224-228)
    e2e = torch.rand(2, 300, 6, device=cuda)
    o = sarpost.non_max_suppression(e2e, conf_thres=0.5, max_det=10)
    assert all(t.shape[0] <= 10 and bool((t[:, 4] > 0.5).all()) for t in o)
    assert sarpost.ops.last_launch_count() >= 0


def test_sharded_tiles_then_merge_equals_single_device(sarpost, cuda):
    """cfg4 in miniature: 2 emulated ranks post-process disjoint tile ranges, their padded detections are
    concatenated (what the all-gather produces) and merged per frame; must equal one device doing it all."""
    strides = (8, 16, 32)
    shapes = sarpost.synth.level_shapes(320, strides)
    spec = sarpost.HeadSpec(nc=2, strides=strides, embed_dim=4, state_classes=0)
    origins = sarpost.dist.sahi_grid(900, 600, 320, 0.2).to(cuda)
    tpf = origins.shape[0]
    n_frames = 2
    levels = sarpost.synth.head_outputs(n_frames * tpf, shapes, 2, 4, 0, seed=55, blobs=3)
    kw = dict(conf_thres=0.25, iou_thres=0.7, max_det=50)
    full_out, full_cnt = sarpost.postprocess_fused([x.to(cuda) for x in levels], spec, return_padded=True, **kw)
    parts = []
    for r in range(2):
        lo, hi = sarpost.dist.shard_range(n_frames * tpf, r, 2)
        parts.append(sarpost.postprocess_fused([x[lo:hi].to(cuda) for x in levels], spec, return_padded=True, **kw))
    g_out = torch.cat([p[0] for p in parts])
    g_cnt = torch.cat([p[1] for p in parts])
    assert torch.equal(g_cnt, full_cnt)
    org = origins.repeat(n_frames, 1)
    a = sarpost.merge_tiles(full_out, full_cnt, org, tpf, iou_thres=0.5, max_det=100)
    b = sarpost.merge_tiles(g_out, g_cnt, org, tpf, iou_thres=0.5, max_det=100)
    assert len(a) == n_frames
    for x, y in zip(a, b):
        assert x.shape[0] > 0 and torch.equal(x, y)


def test_boxes_only_and_gather_extras(sarpost, cuda):
    """with_extras=False gives the same 6 columns; gather_extras reproduces the extras of the kept rows."""
    strides = (8, 16, 32)
    shapes = sarpost.synth.level_shapes(320, strides)
    spec = sarpost.HeadSpec(nc=1, strides=strides, embed_dim=32, state_classes=6)
    levels = [x.to(cuda) for x in sarpost.synth.head_outputs(3, shapes, 1, 32, 6, seed=8)]
    kw = dict(conf_thres=0.25, iou_thres=0.7, max_det=80)
    full, idx = sarpost.postprocess_fused(levels, spec, return_index=True, **kw)
    lean = sarpost.postprocess_fused(levels, spec, with_extras=False, **kw)
    for b in range(3):
        assert lean[b].shape[1] == 6 and torch.equal(lean[b], full[b][:, :6])
    img = torch.cat([torch.full((i.shape[0],), b, dtype=torch.int32, device=cuda) for b, i in enumerate(idx)])
    anc = torch.cat(idx) // spec.nc
    ex = sarpost.gather_extras(levels, spec, img, anc)
    assert torch.equal(ex, torch.cat([f[:, 6:] for f in full]))


def test_fused_scale_boxes_and_clip(sarpost, cuda):
    """§8f row 1: ops.scale_boxes + clip_boxes folded into the gather kernel, bit-exact to the torch ops."""
    strides = (8, 16, 32)
    img1 = (384, 640)
    shapes = sarpost.synth.level_shapes(img1, strides)
    spec = sarpost.HeadSpec(nc=3, strides=strides)
    levels = [x.to(cuda) for x in sarpost.synth.head_outputs(3, shapes, 3, 0, 0, seed=31)]
    orig = [(1080, 1920, 3), (3000, 4000, 3), (480, 640, 3)]
    kw = dict(conf_thres=0.25, iou_thres=0.7)
    plain = sarpost.postprocess_fused(levels, spec, **kw)
    scaled = sarpost.postprocess_fused(levels, spec, scale_to=(img1, orig), **kw)
    for r0, r1, s0 in zip(plain, scaled, orig):
        ref = r0.cpu().clone()
        ref[:, :4] = R.scale_boxes_ref(img1, ref[:, :4], s0)
        assert torch.equal(r1.cpu(), ref)
        assert float(r1[:, 2].max()) <= s0[1] and float(r1[:, 3].max()) <= s0[0] and float(r1[:, :4].min()) >= 0


@pytest.mark.parametrize("imgsz,strides,nc,ed,sc", [(640, (8, 16, 32), 1, 256, 6), ((96, 160), (8, 16, 32), 3, 0, 0),
                                                     ((88, 120), (8, 16, 32), 2, 8, 6)])
def test_fp16_level_tensors(sarpost, cuda, imgsz, strides, nc, ed, sc):
    """§8f row 4 (`half=True`): fp16 logits are upcast exactly and processed in fp32, so feeding the fp16
    tensors must give bit-identical rows to feeding their fp32 upcast; decode returns y in fp16."""
    shapes = sarpost.synth.level_shapes(imgsz, strides)
    spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
    lv16 = [x.to(cuda).half() for x in sarpost.synth.head_outputs(2, shapes, nc, ed, sc, seed=41)]
    lv32 = [x.float() for x in lv16]
    for kw in (dict(conf_thres=0.25, iou_thres=0.7), dict(conf_thres=0.001, iou_thres=0.7, multi_label=True)):
        a, ia = sarpost.postprocess_fused(lv16, spec, return_index=True, **kw)
        b, ib = sarpost.postprocess_fused(lv32, spec, return_index=True, **kw)
        for x, y, i, j in zip(a, b, ia, ib):
            assert x.dtype == torch.float32 and torch.equal(x, y) and torch.equal(i, j)
    y16 = sarpost.decode(lv16, spec)
    y32 = sarpost.decode(lv32, spec)
    assert y16.dtype == torch.float16 and torch.equal(y16, y32.half())
    # and the fp32 path on the upcast inputs is within tolerance of the oracle (same check as test_decode_matches_oracle)
    y_ref = R.decode_ref([x.cpu() for x in lv32], strides, nc, 16, ed, sc)
    st = torch.cat([torch.full((h * w,), float(s)) for (h, w), s in zip(shapes, strides)])
    err = (y32.cpu()[:, :4] - y_ref[:, :4]).abs()
    assert bool((err <= 1e-5 * y_ref[:, :4].abs() + 1e-5 * st).all())


@pytest.mark.parametrize("multi_label", [False, True])
def test_apriori_labels_save_hybrid(sarpost, cuda, multi_label):
    """ops.py:256-261: apriori labels are appended to each image's rows (prob 1.0, zero extras) before NMS."""
    g = torch.Generator().manual_seed(9)
    bs, na, nc, nm = 3, 3000, 4, 3
    y = sarpost.synth.decoded_prediction(bs, na, nc, nm, seed=77)
    labels = []
    for n in (5, 0, 150):
        cls = torch.randint(0, nc, (n, 1), generator=g).float()
        xy = torch.rand(n, 2, generator=g) * 600
        wh = 10 + torch.rand(n, 2, generator=g) * 80
        labels.append(torch.cat((cls, xy, wh), 1))
    kw = dict(conf_thres=0.3, iou_thres=0.6, nc=nc, multi_label=multi_label, classes=[0, 1, 3], max_det=400)
    rows, idx = sarpost.non_max_suppression(y.to(cuda), labels=[lb.to(cuda) for lb in labels], return_index=True, **kw)
    ref_rows, ref_idx = R.non_max_suppression_ref(y, labels=labels, return_index=True, **kw)
    for b in range(bs):
        assert torch.equal(rows[b].cpu(), ref_rows[b])
        a = idx[b].cpu().long() // nc
        src = ref_idx[b][:, 0]
        exp = torch.where(src >= 0, src, na + (-1 - src))  # oracle marks label row k as -1-k
        assert torch.equal(a, exp)
    assert any((r[:, 4] == 1.0).any() for r in rows)


def test_cuda_graph_capture_and_replay(sarpost, cuda):
    """The whole pipeline is stream-ordered (no host sync, no allocation inside the library), so a call can be
    captured into a CUDA graph and replayed on new data written into the same input buffers."""
    strides = (8, 16, 32)
    shapes = sarpost.synth.level_shapes(320, strides)
    spec = sarpost.HeadSpec(nc=2, strides=strides, embed_dim=16, state_classes=6)
    a = [x.to(cuda) for x in sarpost.synth.head_outputs(4, shapes, 2, 16, 6, seed=1)]
    b = [x.to(cuda) for x in sarpost.synth.head_outputs(4, shapes, 2, 16, 6, seed=2)]
    kw = dict(conf_thres=0.25, iou_thres=0.7, max_det=100)
    ref_a = sarpost.postprocess_fused(a, spec, return_padded=True, **kw)
    ref_b = sarpost.postprocess_fused(b, spec, return_padded=True, **kw)
    static_in = [x.clone() for x in a]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm-up on the capture stream (workspace allocation + prepare)
        sarpost.postprocess_fused(static_in, spec, return_padded=True, **kw)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out, counts = sarpost.postprocess_fused(static_in, spec, return_padded=True, **kw)
    for src, ref in ((a, ref_a), (b, ref_b), (a, ref_a)):
        for dst, s in zip(static_in, src):
            dst.copy_(s)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(counts, ref[1])
        for i, n in enumerate(counts.tolist()):
            assert torch.equal(out[i, :n], ref[0][i, :n])


def test_fused_gather_exchange_over_peer_memory(sarpost, cuda):
    """Multi-GPU: K5's peer stores (NVLink symmetric memory) must reproduce the NCCL all-gather bit for bit.
    Needs >= 2 GPUs (the 1-GPU box skips; `gpurun --gpus 2` runs it)."""
    import os, subprocess, sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}", "--master-addr",
           "127.0.0.1", "--master-port", "29577", os.path.join(root, "tools", "peer_gather_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "peer gather == nccl all-gather: True" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.parametrize("imgsz,strides,nc,bs,kw", [
    (320, (8, 16, 32), 80, 2, dict(conf_thres=0.001, iou_thres=0.7, multi_label=True)),       # COCO-like multi-label, 80 classes
    (160, (8, 16, 32), 200, 2, dict(conf_thres=0.05, iou_thres=0.6)),                         # 64+nc > 256 box rows -> LDG path
    (320, (8, 16, 32), 3, 2, dict(conf_thres=0.0, iou_thres=0.5, max_det=1)),                 # conf 0 (every anchor), max_det 1
    (160, (8, 16, 32), 2, 37, dict(conf_thres=0.25, iou_thres=0.7)),                          # 37 images: cluster of 4 CTAs
    (160, (8, 16, 32), 2, 75, dict(conf_thres=0.25, iou_thres=0.7)),                          # 75 images: cluster of 1
    (160, (8, 16, 32), 1, 160, dict(conf_thres=0.1, iou_thres=0.7, max_det=20)),              # more images than SMs
    (320, (8, 16, 32), 4, 2, dict(conf_thres=0.3, iou_thres=0.7, classes=[])),                # empty class filter keeps nothing
    (320, (8, 16, 32), 4, 2, dict(conf_thres=0.2, iou_thres=0.0)),                            # iou 0: any overlap suppresses
    (320, (8, 16, 32), 4, 2, dict(conf_thres=0.2, iou_thres=1.0)),                            # iou 1: nothing suppresses
    (320, (8, 16, 32), 4, 2, dict(conf_thres=0.2, iou_thres=0.7, max_wh=0.0)),                # offset 0 == agnostic
    (320, (8, 16, 32), 4, 2, dict(conf_thres=0.01, iou_thres=0.7, max_nms=7, max_det=300)),   # tiny max_nms
])
def test_fused_configuration_sweep(sarpost, cuda, imgsz, strides, nc, bs, kw):
    shapes = sarpost.synth.level_shapes(imgsz, strides)
    levels = [x.to(cuda) for x in sarpost.synth.head_outputs(bs, shapes, nc, 0, 0, seed=61, blobs=4)]
    spec = sarpost.HeadSpec(nc=nc, strides=strides)
    rows, idx = sarpost.postprocess_fused(levels, spec, return_index=True, **kw)
    y = sarpost.decode(levels, spec).cpu()
    ref_rows, ref_idx = R.non_max_suppression_ref(y, nc=nc, return_index=True, **kw)
    _assert_same(rows, idx, ref_rows, ref_idx, nc)


def test_match_predictions_gpu(sarpost, cuda):
    """§8f row 3: batched box_iou + match_predictions on the GPU vs the numpy restatement (tie-free inputs)."""
    from test_oracle import _match_case
    iouv = torch.linspace(0.5, 0.95, 10)
    cases = [_match_case(s, n_det=n, n_gt=m) for s, n, m in ((0, 120, 40), (1, 300, 7), (2, 5, 60), (3, 0, 10), (4, 50, 0))]
    max_det, max_gt = 300, 64
    dets = torch.zeros(len(cases), max_det, 8)
    gtb = torch.zeros(len(cases), max_gt, 4)
    gtc = torch.zeros(len(cases), max_gt)
    dn, gn = [], []
    for i, (d, g, c) in enumerate(cases):
        dets[i, : d.shape[0], :6] = d
        gtb[i, : g.shape[0]] = g
        gtc[i, : g.shape[0]] = c
        dn.append(d.shape[0]); gn.append(g.shape[0])
    correct, matched = sarpost.match_predictions(dets.to(cuda), torch.tensor(dn), gtb.to(cuda), gtc.to(cuda), torch.tensor(gn),
                                                 iouv.tolist(), tag_threshold_index=0)
    assert correct.shape == (len(cases), max_det, 10) and correct.dtype == torch.bool
    for i, (d, g, c) in enumerate(cases):
        n = d.shape[0]
        if n == 0 or g.shape[0] == 0:
            assert not bool(correct[i].any())
            continue
        iou = R.box_iou_ref(g, d[:, :4])
        ref, ref_m = R.match_predictions_ref(d[:, 5], c, iou, iouv, tag_thr=iouv[0].item())
        assert torch.equal(correct[i, :n].cpu(), ref), f"case {i}"
        assert not bool(correct[i, n:].any())
        assert torch.equal(matched[i, :n].cpu(), ref_m)
        assert ref.any()


def test_reentrant_from_two_threads_and_streams(sarpost, cuda):
    """SURVEY §8b threading: different predictor objects may run on different Python threads / streams.
    No global mutable state: concurrent calls must give exactly the single-threaded results."""
    import threading
    strides = (8, 16, 32)
    shapes = sarpost.synth.level_shapes(320, strides)
    spec = sarpost.HeadSpec(nc=2, strides=strides, embed_dim=8, state_classes=0)
    inputs = [[x.to(cuda) for x in sarpost.synth.head_outputs(3, shapes, 2, 8, 0, seed=70 + i)] for i in range(2)]
    kws = [dict(conf_thres=0.25, iou_thres=0.7), dict(conf_thres=0.01, iou_thres=0.5, multi_label=True, max_det=150)]
    expect = [sarpost.postprocess_fused(inputs[i], spec, **kws[i]) for i in range(2)]
    torch.cuda.synchronize()
    errors = []

    def worker(i):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(25):
                    rows = sarpost.postprocess_fused(inputs[i], spec, **kws[i])
                    for a, b in zip(rows, expect[i]):
                        if not torch.equal(a, b):
                            errors.append(f"thread {i}: mismatch")
                            return
            s.synchronize()
        except Exception as e:  # noqa: BLE001
            errors.append(f"thread {i}: {e!r}")

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_nms_decoded_fp16_prediction(sarpost, cuda):
    """`half=True` drop-in: a fp16 `y` is read directly (no fp32 copy of the whole tensor); results equal the
    fp32 path on the exactly upcast tensor, rows come back in the prediction's dtype like the reference."""
    y32 = sarpost.synth.decoded_prediction(2, 8400, 3, 5, seed=91)
    y16 = y32.half().to(cuda)
    kw = dict(conf_thres=0.25, iou_thres=0.7, nc=3)
    a, ia = sarpost.non_max_suppression(y16, return_index=True, **kw)
    b, ib = sarpost.non_max_suppression(y16.float(), return_index=True, **kw)
    for x, y, i, j in zip(a, b, ia, ib):
        assert x.dtype == torch.float16 and torch.equal(x, y.half()) and torch.equal(i, j)
    ref = R.non_max_suppression_ref(y16.float().cpu(), **kw)
    for x, r in zip(b, ref):
        assert torch.equal(x.cpu(), r)


def test_patch_fused_lazy_prediction(sarpost, cuda):
    """patch(fused=True): the unmodified call sequence `y = head._inference(x); ops.non_max_suppression(y, ...)`
    runs the fused kernels; touching `y` in any other way materialises it with the decode kernel."""
    import types

    class Detect:
        export = False
        reg_max = 16

        def __init__(self, nc, stride, embed_dim=0, state_classes=None):
            self.nc, self.stride, self.embed_dim, self.state_classes = nc, torch.tensor(stride), embed_dim, state_classes

        def _inference(self, x):
            raise AssertionError("reference decode must not run for CUDA inputs")

    class JDE(Detect):
        pass

    def ref_nms(*a, **k):
        raise AssertionError("reference NMS must not run for CUDA inputs")

    ops_mod = types.SimpleNamespace(non_max_suppression=ref_nms)
    head_mod = types.SimpleNamespace(Detect=Detect, JDE=JDE)
    strides = (8.0, 16.0, 32.0)
    shapes = sarpost.synth.level_shapes(320, (8, 16, 32))
    levels = [x.to(cuda) for x in sarpost.synth.head_outputs(2, shapes, 1, 16, 6, seed=12)]
    spec = sarpost.HeadSpec(nc=1, strides=strides, embed_dim=16, state_classes=6)
    want = sarpost.postprocess_fused(levels, spec, conf_thres=0.25, iou_thres=0.7, max_det=100)
    want_y = sarpost.decode(levels, spec)
    sarpost.patch(ops_mod, head_mod, fused=True)
    try:
        head = JDE(1, strides, 16, 6)
        y = head._inference(levels)
        assert isinstance(y, sarpost.plugin.LazyPrediction) and tuple(y.shape) == tuple(want_y.shape) and y.is_cuda
        # the predictor's call (models/yolo/jde/predict.py:31-39): positional conf/iou, keyword rest, tuple input
        rows = ops_mod.non_max_suppression((y, levels), 0.25, 0.7, agnostic=False, max_det=100, nc=1, classes=None)
        assert y._y is None, "y must not have been materialised"
        for a, b in zip(rows, want):
            assert torch.equal(a, b)
        # any other use materialises y transparently
        assert torch.equal(y[:, :4], want_y[:, :4]) and torch.equal(y + 0, want_y) and y._y is not None
        rows2 = ops_mod.non_max_suppression(y, 0.25, 0.7, max_det=100, nc=1)  # now the decoded-input kernels
        for a, b in zip(rows2, want):
            assert torch.equal(a, b)
    finally:
        sarpost.unpatch()


def test_borderline_iou_both_directions(sarpost, cuda):
    """Regression (found by tools/fuzz_parity.py): pairs whose IoU is within the approximate-quotient band of the
    threshold must be decided by the exact division in BOTH directions — e.g. identical boxes at iou_thres = 1.0
    (IoU == 1.0 is not > 1.0: nothing is suppressed), also when they meet in phase 1 (candidate vs kept list)."""
    n = 700  # > one sub-chunk, so later duplicates are tested against the kept list
    y = torch.zeros(1, 5, n)
    y[0, :4] = torch.tensor([100.0, 80.0, 37.3, 55.1])[:, None]
    y[0, 4] = torch.linspace(0.9, 0.3, n)
    for thr, expect in ((1.0, 300), (0.999999, 1)):
        kw = dict(conf_thres=0.25, iou_thres=thr)
        rows, idx, ref_rows, ref_idx = _nms_both(sarpost, y, cuda, **kw)
        _assert_same(rows, idx, ref_rows, ref_idx, 1)
        assert rows[0].shape[0] == expect


@pytest.mark.parametrize("spread_ulps,batch", [(20000, 1), (20000, 5), (60000, 40), (6, 2)])
@pytest.mark.parametrize("cluster", ["", "1", "8"])
def test_one_score_bucket_holds_thousands(sarpost, cuda, monkeypatch, spread_ulps, batch, cluster):
    """Saturated scores (clustered detections): thousands of DISTINCT scores inside one 0.4 %-wide score bucket.  The NMS
    kernel then zooms into the bucket (exact histogram over sub-buckets of 8 adjacent floats) instead of sorting it in
    global memory; with `spread_ulps` = 6 nearly everything shares one sub-bucket and the radix fallback has to finish.
    299 far-apart clusters of near-duplicates keep the walk going through all max_nms candidates."""
    if cluster:
        monkeypatch.setenv("SARPOST_NMS_CLUSTER", cluster)
    a, k = 9000, 120
    g = torch.Generator().manual_seed(spread_ulps + batch)
    cx = (torch.arange(k) % 12) * 60.0 + 40
    cy = (torch.arange(k) // 12) * 80.0 + 40
    y = torch.zeros(batch, 5, a)
    for b in range(batch):
        pick = torch.randint(0, k, (a,), generator=g)
        y[b, 0] = cx[pick] + torch.randn(a, generator=g) * 0.5
        y[b, 1] = cy[pick] + torch.randn(a, generator=g) * 0.5
        y[b, 2:4] = 30.0
        base = torch.tensor(0.99).view(torch.int32)  # bit pattern of 0.99: consecutive patterns = consecutive floats
        ulps = torch.randint(0, spread_ulps, (a,), generator=g, dtype=torch.int32)
        y[b, 4] = (base + ulps).view(torch.float32)
        y[b, 4, : a // 10] = torch.rand(a // 10, generator=g) * 0.9 + 0.05  # plus ordinary scores in the other buckets
    kw = dict(conf_thres=0.001, iou_thres=0.5, max_nms=8000)
    rows, idx, ref_rows, ref_idx = _nms_both(sarpost, y, cuda, **kw)
    _assert_same(rows, idx, ref_rows, ref_idx, 1)
    assert all(r.shape[0] <= k for r in rows) and rows[0].shape[0] > k // 2


def test_seeded_fuzz(sarpost, cuda):
    """150 random configurations (decoded / fused, fp32 / fp16, ties, clustered boxes, odd shapes, extreme
    thresholds) against the oracle; `python tools/fuzz_parity.py 2000 <seed>` runs more."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                               "tools", "fuzz_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(150, 7) == 0


@pytest.mark.parametrize("e,h,s,bsz,max_det", [(256, 128, 6, 16, 300), (128, 64, 6, 3, 100), (20, 10, 3, 2, 37), (64, 200, 17, 2, 50), (10, 5, 3, 2, 20), (36, 33, 64, 1, 17)])
@pytest.mark.parametrize("tiled", [False, True])
def test_state_head_kernel_vs_oracle(sarpost, cuda, monkeypatch, e, h, s, bsz, max_det, tiled):
    """§8f row 2: sarpost_state_head on padded rows vs the oracle MLP (head.py:189-190,247); fp32, atol 1e-5 on the
    probabilities.  Rows beyond counts[b] and all other columns stay untouched."""
    if tiled:  # the streaming kernel (any size); otherwise the resident kernel where it applies
        monkeypatch.setenv("SARPOST_STATE_TILED", "1")
    g = torch.Generator().manual_seed(e * 7 + h)
    w1, b1 = torch.randn(h, e, generator=g) * (2.0 / e ** 0.5), torch.randn(h, generator=g)
    w2, b2 = torch.randn(s, h, generator=g) * (2.0 / h ** 0.5), torch.randn(s, generator=g)
    row_len = 6 + e + s
    rows = torch.randn(bsz, max_det, row_len, generator=g)
    counts = torch.randint(0, max_det + 1, (bsz,), generator=g, dtype=torch.int32)
    counts[0] = max_det
    if bsz > 1:
        counts[1] = 0
    mlp = sarpost.StateMLP.from_tensors(w1, b1, w2, b2, device=cuda)
    out = sarpost.state_head(rows.to(cuda).clone(), counts.to(cuda), mlp).cpu()
    ref = R.state_head_ref(rows[..., 6: 6 + e], w1, b1, w2, b2)
    for b in range(bsz):
        n = int(counts[b])
        assert torch.allclose(out[b, :n, 6 + e:], ref[b, :n], rtol=0, atol=1e-5), float((out[b, :n, 6 + e:] - ref[b, :n]).abs().max())
        assert torch.equal(out[b, n:], rows[b, n:])
        assert torch.equal(out[b, :, : 6 + e], rows[b, :, : 6 + e])
    assert sarpost.ops.last_launch_count() == 1


def test_state_head_errors(sarpost, cuda):
    mlp = sarpost.StateMLP.from_tensors(torch.zeros(4, 8), torch.zeros(4), torch.zeros(2, 4), torch.zeros(2), device=cuda)
    rows = torch.zeros(1, 5, 6 + 8 + 2, device=cuda)
    cnt = torch.ones(1, dtype=torch.int32, device=cuda)
    with pytest.raises(sarpost.SarpostError, match="outside the row"):
        sarpost.state_head(rows, cnt, mlp, emb_col=6, state_col=15)
    with pytest.raises(sarpost.SarpostError, match="overlap"):
        sarpost.state_head(rows, cnt, mlp, emb_col=6, state_col=10)
    with pytest.raises(RuntimeError, match="only CUDA"):
        sarpost.state_head(rows.cpu(), cnt, mlp)
    with pytest.raises(ValueError, match="state MLP is"):
        spec = sarpost.HeadSpec(nc=1, strides=(8,), embed_dim=16, state_classes=2)
        sarpost.postprocess_fused([torch.zeros(1, 64 + 1 + 16, 4, 4, device=cuda)], spec, state_mlp=mlp)


def test_patch_defer_state_runs_the_mlp_on_kept_rows_only(sarpost, cuda):
    """patch(fused=True, defer_state=True): JDE.forward skips state_predictor on the anchor map and the patched NMS
    fills the state columns from the kept rows' embeddings; rows equal the undeferred path (state within 1e-5)."""
    import types

    calls = {"mlp": 0}

    class Counting(torch.nn.Sequential):
        def forward(self, x):
            calls["mlp"] += 1
            return super().forward(x)

    class JDE(torch.nn.Module):
        export, reg_max, end2end = False, 16, False

        def __init__(self, nc, stride, embed_dim, state_classes, ch):
            super().__init__()
            self.nc, self.stride, self.embed_dim, self.state_classes, self.nl = nc, torch.tensor(stride), embed_dim, state_classes, len(ch)
            self.cv2 = torch.nn.ModuleList(torch.nn.Conv2d(c, 64, 1) for c in ch)
            self.cv3 = torch.nn.ModuleList(torch.nn.Conv2d(c, nc, 1) for c in ch)
            self.cv4 = torch.nn.ModuleList(torch.nn.Conv2d(c, embed_dim, 1) for c in ch)
            self.state_predictor = Counting(torch.nn.Linear(embed_dim, embed_dim // 2), torch.nn.ReLU(), torch.nn.Dropout(0.1),
                                            torch.nn.Linear(embed_dim // 2, state_classes))

        def forward(self, x):  # the reference's structure (head.py:193-212)
            for i in range(self.nl):
                emb = self.cv4[i](x[i])
                b, c, h, w = emb.shape
                st = self.state_predictor(emb.view(b, c, -1).permute(0, 2, 1)).permute(0, 2, 1).view(b, self.state_classes, h, w)
                x[i] = torch.cat((self.cv2[i](x[i]), self.cv3[i](x[i]), emb, st), 1)
            return self._inference(x), x

        def _inference(self, x):
            raise AssertionError("reference decode must not run for CUDA inputs")

    Detect = type("Detect", (), {"_inference": lambda self, x: None})
    ops_mod = types.SimpleNamespace(non_max_suppression=lambda *a, **k: (_ for _ in ()).throw(AssertionError("reference NMS")))
    head_mod = types.SimpleNamespace(Detect=Detect, JDE=JDE)
    torch.manual_seed(3)
    head = JDE(1, (8.0, 16.0), 32, 6, (8, 12)).to(cuda).eval()
    feats = [torch.randn(2, 8, 20, 20, device=cuda), torch.randn(2, 12, 10, 10, device=cuda)]
    with torch.no_grad():
        sarpost.patch(ops_mod, head_mod, fused=True)
        try:
            y, _ = head([f.clone() for f in feats])
            want = ops_mod.non_max_suppression(y, 0.3, 0.7, max_det=50, nc=1)
            want_y = y + 0
        finally:
            sarpost.unpatch()
        assert calls["mlp"] == 2
        sarpost.patch(ops_mod, head_mod, fused=True, defer_state=True)
        try:
            y, x = head([f.clone() for f in feats])
            # defer_state implies the split layout: per level a (box, cls, emb) tuple, no state branch
            assert calls["mlp"] == 2 and [int(t.shape[1]) for t in x[0]] == [64, 1, 32] and tuple(y.shape) == tuple(want_y.shape)
            got = ops_mod.non_max_suppression(y, 0.3, 0.7, max_det=50, nc=1)
            assert calls["mlp"] == 2 and y._y is None
            assert sum(r.shape[0] for r in got) > 10
            for a, b in zip(got, want):
                assert a.shape == b.shape and torch.equal(a[:, :38], b[:, :38])
                assert torch.allclose(a[:, 38:], b[:, 38:], rtol=0, atol=1e-5)
            assert torch.allclose(y + 0, want_y, rtol=1e-5, atol=1e-5) and calls["mlp"] == 4   # materialising y runs the module's MLP
        finally:
            sarpost.unpatch()
    assert JDE.forward is not sarpost.plugin._jde_forward_split


def test_empty_batch_returns_empty_list(sarpost, cuda):
    """B = 0: the reference's per-image loop simply produces no outputs (ops.py:250)."""
    spec = sarpost.HeadSpec(nc=1, strides=(8, 16, 32), embed_dim=4, state_classes=0)
    levels = [torch.zeros(0, 69, h, w, device=cuda) for h, w in sarpost.synth.level_shapes(64, (8, 16, 32))]
    assert sarpost.postprocess_fused(levels, spec) == []
    out, counts = sarpost.postprocess_fused(levels, spec, return_padded=True, max_det=10)
    assert tuple(out.shape) == (0, 10, 10) and tuple(counts.shape) == (0,)
    assert sarpost.non_max_suppression(torch.zeros(0, 9, 50, device=cuda), nc=1) == []


def test_bench_batch_cfg3_against_oracle(sarpost, cuda):
    """The bench's own input (bench.py default workload: cfg3, JDE no=327, 16 images generated ON THE DEVICE with seed 3000,
    cls_mean -4) at full size: the 262-channel extras gather at 136 000 anchors and the deep max_nms cut are compared with
    the oracle — rows / indices / extras bit-exact on our decoded y for the first two images, and the raw-logit path inside
    the 1e-4 mismatch budget for the first image."""
    strides = (4, 8, 16, 32)
    shapes = sarpost.synth.level_shapes(1280, strides)
    spec = sarpost.HeadSpec(nc=1, strides=strides, embed_dim=256, state_classes=6)
    levels = sarpost.synth.head_outputs(16, shapes, 1, 256, 6, cls_mean=-4.0, seed=3000, device=cuda)  # as bench.py, rank 0
    kw = dict(conf_thres=0.001, iou_thres=0.7, max_det=300, max_nms=30000, multi_label=False)
    rows, idx = sarpost.postprocess_fused(levels, spec, return_index=True, **kw)
    assert len(rows) == 16 and all(r.shape == (300, 268) for r in rows)
    first = [x[:2].contiguous() for x in levels]
    y = sarpost.decode(first, spec).cpu()
    assert y.shape == (2, 267, 136000)
    ref_rows, ref_idx = R.non_max_suppression_ref(y, nc=1, return_index=True, **kw)
    _assert_same(rows[:2], idx[:2], ref_rows, ref_idx, 1)
    # raw logits -> oracle decode -> oracle NMS (image 0)
    lv0 = [x[:1].cpu() for x in levels]
    y_ref = R.decode_ref(lv0, strides, 1, 16, 256, 6)
    e_rows, e_idx = R.non_max_suppression_ref(y_ref, nc=1, return_index=True, **kw)
    ours = {int(k): r for k, r in zip(idx[0].cpu().tolist(), rows[0].cpu())}
    ref = {int(a): r for (a, c), r in zip(e_idx[0].tolist(), e_rows[0])}
    bad = sum(1 for k in set(ours) | set(ref)
              if k not in ours or k not in ref or not torch.allclose(ours[k], ref[k], rtol=1e-5, atol=1e-5 * 32))
    total = max(len(ours), len(ref))
    assert total == 300 and bad <= 1e-4 * total, f"{bad} mismatching detections out of {total}"


def _extreme_cases(S, sarpost):
    s4, s3 = (4, 8, 16, 32), (8, 16, 32)
    spec = sarpost.HeadSpec(nc=1, strides=s4, embed_dim=8, state_classes=0)
    spec80 = sarpost.HeadSpec(nc=80, strides=s3)
    return {
        # 544 000 anchors: more tiles than one tile-list round holds
        "p2_2560_544k_anchors": (lambda: S.head_outputs(2, S.level_shapes(2560, s4), 1, 8, 0, seed=1, cls_mean=-2.0), spec,
                                 dict(conf_thres=0.001, iou_thres=0.7)),
        "max_det_4096_max_nms_100000": (lambda: S.head_outputs(1, S.level_shapes(2560, s4), 1, 8, 0, seed=2, cls_mean=-2.0), spec,
                                        dict(conf_thres=0.001, iou_thres=0.7, max_det=4096, max_nms=100000)),
        "nc80_multi_label_672k_slots": (lambda: S.head_outputs(3, S.level_shapes(640, s3), 80, seed=3, cls_mean=-3.0), spec80,
                                        dict(conf_thres=0.001, iou_thres=0.7, multi_label=True)),
        "nc80_quantised_scores_ties": (lambda: [x.mul(2).round().div(2) for x in S.head_outputs(2, S.level_shapes(640, s3), 80, seed=4, cls_mean=-1.0)],
                                       spec80, dict(conf_thres=0.05, iou_thres=0.6, agnostic=True)),
        "batch_300_more_images_than_sms": (lambda: S.head_outputs(300, S.level_shapes(64, (16,)), 2, seed=5, cls_mean=0.0),
                                           sarpost.HeadSpec(nc=2, strides=(16,)), dict(conf_thres=0.25, iou_thres=0.5)),
        # every score equal: one giant score bucket -> the global radix-sort fallback
        "all_scores_equal_radix_fallback": (lambda: [torch.zeros(1, 65, h, w) for h, w in S.level_shapes(640, s3)],
                                            sarpost.HeadSpec(nc=1, strides=s3), dict(conf_thres=0.25, iou_thres=0.7)),
    }


@pytest.mark.parametrize("name", ["p2_2560_544k_anchors", "max_det_4096_max_nms_100000", "nc80_multi_label_672k_slots",
                                  "nc80_quantised_scores_ties", "batch_300_more_images_than_sms", "all_scores_equal_radix_fallback"])
def test_extreme_geometry(sarpost, cuda, name):
    """Edges of the supported geometry (formerly tools/extreme_parity.py): fused and decoded entry points bit-exact vs the oracle."""
    make, spec, kw = _extreme_cases(sarpost.synth, sarpost)[name]
    levels = [x.to(cuda) for x in make()]
    rows, idx = sarpost.postprocess_fused(levels, spec, return_index=True, **kw)
    y = sarpost.decode(levels, spec)
    ref_rows, ref_idx = R.non_max_suppression_ref(y.cpu(), nc=spec.nc, return_index=True, **kw)
    _assert_same(rows, idx, ref_rows, ref_idx, spec.nc)
    rows2 = sarpost.non_max_suppression(y, nc=spec.nc, **kw)
    for a, b in zip(rows2, ref_rows):
        assert torch.equal(a.cpu(), b)


@pytest.mark.parametrize("imgsz,strides,nc,ed,sc,bs", HEADS)
@pytest.mark.parametrize("emb_cl", [False, True])
def test_split_layout_equals_concatenated_layout(sarpost, cuda, imgsz, strides, nc, ed, sc, bs, emb_cl):
    """SARPOST_LAYOUT_SPLIT: the branch outputs handed over separately (no torch.cat, head.py:204-206), embedding NCHW or
    channels_last — rows, kept indices and extras must be bit-identical to the concatenated layout (TMA and LDG paths)."""
    shapes = sarpost.synth.level_shapes(imgsz, strides)
    cat = [x.to(cuda) for x in sarpost.synth.head_outputs(bs, shapes, nc, ed, sc, seed=13, blobs=3)]
    spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
    split = sarpost.split_levels(cat, spec, emb_channels_last=emb_cl)
    back = sarpost.cat_levels(split)
    assert all(torch.equal(a, b) for a, b in zip(back, cat))
    for kw in (dict(conf_thres=0.25, iou_thres=0.7), dict(conf_thres=0.001, iou_thres=0.7, multi_label=True, max_det=50)):
        a, ia = sarpost.postprocess_fused(cat, spec, return_index=True, **kw)
        b, ib = sarpost.postprocess_fused(split, spec, return_index=True, **kw)
        for x, y, i, j in zip(a, b, ia, ib):
            assert torch.equal(x, y) and torch.equal(i, j)
    if ed or sc:  # explicit extras gather from the split tensors
        rows, idx = sarpost.postprocess_fused(cat, spec, return_index=True, conf_thres=0.25, iou_thres=0.7, max_det=40)
        img = torch.cat([torch.full((i.shape[0],), b_, dtype=torch.int32, device=cuda) for b_, i in enumerate(idx)])
        ex = sarpost.gather_extras(split, spec, img, torch.cat(idx) // nc)
        assert torch.equal(ex, torch.cat([r[:, 6:] for r in rows]))
        lean = sarpost.postprocess_fused(split, spec, with_extras=False, conf_thres=0.25, iou_thres=0.7, max_det=40)
        assert all(torch.equal(l_, r[:, :6]) for l_, r in zip(lean, rows))
    # fp16 branches
    split16 = [tuple(None if t is None else t.half() for t in lv) for lv in split]
    cat16 = [x.half() for x in cat]
    a = sarpost.postprocess_fused(cat16, spec, conf_thres=0.25, iou_thres=0.7)
    b = sarpost.postprocess_fused(split16, spec, conf_thres=0.25, iou_thres=0.7)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_split_layout_deferred_state_and_errors(sarpost, cuda):
    strides = (8, 16, 32)
    shapes = sarpost.synth.level_shapes(160, strides)
    spec = sarpost.HeadSpec(nc=1, strides=strides, embed_dim=16, state_classes=6)
    cat = [x.to(cuda) for x in sarpost.synth.head_outputs(2, shapes, 1, 16, 6, seed=5)]
    g = torch.Generator().manual_seed(3)
    mlp = sarpost.StateMLP.from_tensors(torch.randn(8, 16, generator=g), torch.randn(8, generator=g), torch.randn(6, 8, generator=g),
                                        torch.randn(6, generator=g), device=cuda)
    no_state_cat = [x[:, :64 + 1 + 16].contiguous() for x in cat]
    no_state_split = [lv[:3] for lv in sarpost.split_levels(cat, spec)]
    kw = dict(conf_thres=0.25, iou_thres=0.7)
    a = sarpost.postprocess_fused(no_state_cat, spec, state_mlp=mlp, **kw)
    b = sarpost.postprocess_fused(no_state_split, spec, state_mlp=mlp, **kw)
    assert all(x.shape[1] == 6 + 16 + 6 and torch.equal(x, y) for x, y in zip(a, b))
    with pytest.raises(ValueError, match="no state branch"):
        sarpost.postprocess_fused(no_state_split, spec, **kw)
    with pytest.raises(ValueError):
        sarpost.postprocess_fused([lv[:1] + (lv[1][:, :, :-1],) + lv[2:] for lv in sarpost.split_levels(cat, spec)], spec, **kw)
    # y is defined on the concatenated layout: decode() accepts split levels by concatenating them first
    assert torch.equal(sarpost.decode(sarpost.split_levels(cat, spec), spec), sarpost.decode(cat, spec))


def test_match_from_iou_equals_reference_method(sarpost, cuda):
    """`patch(match=True)` boundary: BaseValidator.match_predictions(pred_classes, true_classes, iou) from a precomputed IoU
    matrix (engine/validator.py:222-262) and the JDE variant's tags (jde/val.py:683-736) vs the numpy oracle."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_oracle import _match_case
    iouv = torch.linspace(0.5, 0.95, 10)
    for seed, (nd, ng, nc) in enumerate([(300, 40, 3), (17, 200, 1), (1, 1, 2), (120, 0, 2), (0, 5, 2), (299, 299, 6)]):
        dets, gtb, gtc = _match_case(100 + seed, n_det=nd, n_gt=ng, nc=nc)
        iou = R.box_iou_ref(gtb, dets[:, :4]) if nd and ng else torch.zeros(ng, nd)
        got, matched = sarpost.match_from_iou(dets[:, 5].to(cuda), gtc.to(cuda), iou.to(cuda), iouv.to(cuda), tag_threshold_index=0)
        if nd and ng:
            ref, ref_m = R.match_predictions_ref(dets[:, 5], gtc, iou, iouv, tag_thr=iouv[0].item())
        else:
            ref, ref_m = torch.zeros(nd, 10, dtype=torch.bool), torch.full((nd,), -1, dtype=torch.int32)
        assert got.dtype == torch.bool and torch.equal(got.cpu(), ref)
        assert torch.equal(matched.cpu().long(), ref_m.long())
    # through the patched methods on stand-in validator classes (the real ones are exercised by tools/dropin_predict_check.py)
    import types

    class BaseValidator:
        def match_predictions(self, pred_classes, true_classes, iou, use_scipy=False):
            return "ref"

    class JDEValidator(BaseValidator):
        def match_predictions(self, pred_classes, true_classes, true_tags, iou, use_scipy=False):
            return "ref-jde"

    vm = types.SimpleNamespace(BaseValidator=BaseValidator, JDEValidator=JDEValidator)
    sarpost.patch(types.SimpleNamespace(non_max_suppression=lambda *a, **k: None), types.SimpleNamespace(Detect=type("Detect", (), {"_inference": None})),
                  match=True, ultralytics_validator=vm)
    try:
        dets, gtb, gtc = _match_case(7, n_det=150, n_gt=30, nc=2)
        iou = R.box_iou_ref(gtb, dets[:, :4])
        v = BaseValidator()
        v.iouv = iouv.to(cuda)
        assert v.match_predictions(dets[:, 5], gtc, iou) == "ref"                                  # CPU tensors: the reference's own method
        assert v.match_predictions(dets[:, 5].to(cuda), gtc.to(cuda), iou.to(cuda), use_scipy=True) == "ref"
        got = v.match_predictions(dets[:, 5].to(cuda), gtc.to(cuda), iou.to(cuda))
        ref, ref_m = R.match_predictions_ref(dets[:, 5], gtc, iou, iouv, tag_thr=iouv[0].item())
        assert torch.equal(got.cpu(), ref)
        j = JDEValidator()
        j.iouv, j.state_iou = iouv.to(cuda), 0.5
        tags = torch.arange(30).float() + 100
        c, t = j.match_predictions(dets[:, 5].to(cuda), gtc.to(cuda), tags.to(cuda), iou.to(cuda))
        assert torch.equal(c.cpu(), ref) and t.dtype == torch.int32
        want = torch.where(ref_m >= 0, tags[ref_m.clamp(min=0).long()].int(), torch.zeros_like(ref_m, dtype=torch.int))
        assert torch.equal(t.cpu(), want)
    finally:
        sarpost.unpatch()
    assert BaseValidator().match_predictions(None, None, None) == "ref" and JDEValidator().match_predictions(None, None, None, None) == "ref-jde"


def test_pipeline_matches_plain_calls(sarpost, cuda):
    """sarpost_pipeline_*: batches submitted back to back (decode of batch i+1 overlapping NMS + gather of batch i on a
    second, high-priority stream, rotating workspaces, dynamic tile scheduling) give the very rows of one plain
    `postprocess_fused` call per batch — also when geometry, batch size and thresholds change between submits."""
    strides = (8, 16, 32)
    spec = sarpost.HeadSpec(nc=2, strides=strides, embed_dim=8, state_classes=3)
    jobs = []
    for i, (imgsz, bs, kw) in enumerate([(320, 9, dict(conf_thres=0.25, iou_thres=0.7)), (320, 9, dict(conf_thres=0.25, iou_thres=0.7)),
                                         (256, 3, dict(conf_thres=0.05, iou_thres=0.5, max_det=40)), (320, 9, dict(conf_thres=0.25, iou_thres=0.7, multi_label=True)),
                                         (320, 12, dict(conf_thres=0.001, iou_thres=0.7, max_nms=500)), (320, 9, dict(conf_thres=0.25, iou_thres=0.7))] * 2):
        shapes = sarpost.synth.level_shapes(imgsz, strides)
        levels = [x.to(cuda) for x in sarpost.synth.head_outputs(bs, shapes, spec.nc, spec.embed_dim, spec.state_classes, seed=50 + i, cls_mean=-1.0)]
        jobs.append((levels, kw))
    for depth in (1, 2, 3):
        pl = sarpost.Pipeline(cuda, depth=depth)
        got = [pl.submit(levels, spec, return_index=True, **kw) for levels, kw in jobs]
        pl.wait()
        torch.cuda.current_stream().synchronize()
        for (levels, kw), (out, counts, kidx) in zip(jobs, got):
            want_out, want_counts, want_idx = sarpost.postprocess_fused(levels, spec, return_padded=True, return_index=True, **kw)
            assert torch.equal(counts, want_counts)
            for b, n in enumerate(want_counts.tolist()):
                assert torch.equal(out[b, :n], want_out[b, :n]) and torch.equal(kidx[b, :n], want_idx[b, :n])
        pl.close()


@pytest.mark.gpu
@pytest.mark.parametrize("imgsz,strides,nc,ed,sc,bs", HEADS)
@pytest.mark.parametrize("layout", ["cat", "split", "split_cl", "half"])
def test_plan_runs_equal_plain_calls(sarpost, cuda, imgsz, strides, nc, ed, sc, bs, layout):
    """sarpost_plan_*: a plan prepared from one set of level tensors and then run on OTHER tensors of the same geometry
    (new addresses -> tensor maps re-encoded, new values) gives the very rows of a plain `postprocess_fused` call, in every
    layout / dtype, with and without kept indices, caller-owned outputs, the results layout and fused scale_boxes."""
    spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
    shapes = sarpost.synth.level_shapes(imgsz, strides)
    kw = dict(conf_thres=0.05, iou_thres=0.6, multi_label=nc > 1, max_det=50)

    def make(seed):
        lv = [x.to(cuda) for x in sarpost.synth.head_outputs(bs, shapes, nc, ed, sc, seed=seed, cls_mean=-1.0)]
        if layout == "half":
            lv = [x.half() for x in lv]
        if layout.startswith("split"):
            lv = sarpost.split_levels(lv, spec, emb_channels_last=layout == "split_cl")
        return lv

    plan = sarpost.FusedPlan(make(1), spec, **kw)
    keep_alive = []
    for seed in (1, 2, 3, 2):
        lv = make(seed)
        keep_alive.append(lv)  # keep earlier inputs allocated so that every round sees new addresses
        want_out, want_counts, want_idx = sarpost.postprocess_fused(lv, spec, return_padded=True, return_index=True, **kw)
        out, counts, kidx = plan(lv, return_index=True)
        assert torch.equal(counts, want_counts)
        for b, n in enumerate(want_counts.tolist()):
            assert torch.equal(out[b, :n], want_out[b, :n]) and torch.equal(kidx[b, :n], want_idx[b, :n])
        # caller-owned outputs, no kept indices
        mine = (torch.full_like(want_out, -7.0), torch.zeros_like(want_counts))
        out2, counts2 = plan(lv, out=mine)
        assert out2 is mine[0] and torch.equal(counts2, want_counts)
        for b, n in enumerate(want_counts.tolist()):
            assert torch.equal(out2[b, :n], want_out[b, :n]) and bool((out2[b, n:] == -7.0).all())
    # fused scale_boxes + clip through the plan
    h0 = imgsz if isinstance(imgsz, int) else imgsz[0]
    w0 = imgsz if isinstance(imgsz, int) else imgsz[1]
    scale_to = ((h0, w0), [(h0 * 2 + 3 * i, w0 * 2 - 5 * i) for i in range(bs)])
    lv = make(5)
    want_out, want_counts = sarpost.postprocess_fused(lv, spec, return_padded=True, scale_to=scale_to, **kw)
    out, counts = plan(lv, scale_to=scale_to)
    assert torch.equal(counts, want_counts)
    for b, n in enumerate(want_counts.tolist()):
        assert torch.equal(out[b, :n], want_out[b, :n])
    plan.close()
    if ed:  # results layout (7-column boxes + contiguous embeddings)
        rplan = sarpost.FusedPlan(make(1), spec, results=True, **kw)
        lv = make(6)
        wb, we, wc = sarpost.postprocess_fused(lv, spec, return_padded=True, results=True, **kw)
        gb, ge, gc = rplan(lv)
        assert torch.equal(gc, wc)
        for b, n in enumerate(wc.tolist()):
            assert torch.equal(gb[b, :n], wb[b, :n]) and torch.equal(ge[b, :n], we[b, :n])
        rplan.close()


@pytest.mark.gpu
def test_plan_rejects_other_geometry(sarpost, cuda):
    strides = (8, 16, 32)
    spec = sarpost.HeadSpec(nc=2, strides=strides)
    mk = lambda imgsz, bs: [x.to(cuda) for x in sarpost.synth.head_outputs(bs, sarpost.synth.level_shapes(imgsz, strides), 2, 0, 0, seed=3)]  # noqa: E731
    plan = sarpost.FusedPlan(mk(320, 2), spec, conf_thres=0.25, iou_thres=0.7)
    for bad in (mk(320, 3), mk(256, 2), [x.half() for x in mk(320, 2)], sarpost.split_levels(mk(320, 2), spec),
                [x.permute(0, 1, 3, 2) for x in mk(320, 2)]):
        with pytest.raises(ValueError):
            plan(bad)
    with pytest.raises(RuntimeError):
        sarpost.FusedPlan([x.cpu() for x in mk(320, 2)], spec)
    out, counts = plan(mk(320, 2))  # still usable after the rejected calls
    want_out, want_counts = sarpost.postprocess_fused(mk(320, 2), spec, return_padded=True, conf_thres=0.25, iou_thres=0.7)
    assert torch.equal(counts, want_counts)
    plan.close()


@pytest.mark.gpu
def test_nms_cluster_parameter_does_not_change_results(sarpost, cuda):
    """`sarpost_nms_params_t.nms_cluster`: 1 / 2 / 4 / 8 CTAs per image give the rows of the automatic choice, through the
    general call, a plan and the merge; any other value is rejected."""
    strides = (8, 16, 32)
    spec = sarpost.HeadSpec(nc=3, strides=strides, embed_dim=8, state_classes=2)
    lv = [x.to(cuda) for x in sarpost.synth.head_outputs(3, sarpost.synth.level_shapes(320, strides), 3, 8, 2, seed=11, cls_mean=-1.0, blobs=4)]
    kw = dict(conf_thres=0.01, iou_thres=0.6, multi_label=True, max_det=120)
    want_out, want_counts, want_idx = sarpost.postprocess_fused(lv, spec, return_padded=True, return_index=True, **kw)
    for cl in (1, 2, 4, 8):
        out, counts, kidx = sarpost.postprocess_fused(lv, spec, return_padded=True, return_index=True, nms_cluster=cl, **kw)
        assert torch.equal(counts, want_counts)
        for b, n in enumerate(want_counts.tolist()):
            assert torch.equal(out[b, :n], want_out[b, :n]) and torch.equal(kidx[b, :n], want_idx[b, :n])
        plan = sarpost.FusedPlan(lv, spec, nms_cluster=cl, **kw)
        p_out, p_counts = plan(lv)
        assert torch.equal(p_counts, want_counts)
        for b, n in enumerate(want_counts.tolist()):
            assert torch.equal(p_out[b, :n], want_out[b, :n])
        plan.close()
    with pytest.raises(sarpost.SarpostError):
        sarpost.postprocess_fused(lv, spec, nms_cluster=3, **kw)
    # merge: two frames of 3 tiles each
    dets = torch.zeros(6, 50, 6, device=cuda)
    g = torch.Generator().manual_seed(5)
    xy = torch.rand(6, 50, 2, generator=g) * 200
    wh = torch.rand(6, 50, 2, generator=g) * 60 + 5
    dets[..., :2], dets[..., 2:4] = xy.to(cuda), (xy + wh).to(cuda)
    dets[..., 4] = torch.rand(6, 50, generator=g).to(cuda)
    cnt = torch.tensor([50, 20, 0, 50, 50, 7], dtype=torch.int32, device=cuda)
    org = torch.tensor([[0.0, 0.0], [100.0, 0.0], [0.0, 100.0]] * 2, device=cuda)
    want = sarpost.merge_tiles(dets, cnt, org, 3, iou_thres=0.5)
    for cl in (1, 2, 8):
        got = sarpost.merge_tiles(dets, cnt, org, 3, iou_thres=0.5, nms_cluster=cl)
        assert all(torch.equal(a, b) for a, b in zip(got, want))
