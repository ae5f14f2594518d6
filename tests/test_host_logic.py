"""CPU: host-side logic — API contract on CPU tensors, plugin patching, sharding, SAHI grid, gloo all-gather."""
import os
import subprocess
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpu_tensors_raise_no_fallback(sarpost):
    y = sarpost.synth.decoded_prediction(1, 64, 2, 0, seed=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sarpost.non_max_suppression(y)
    spec = sarpost.HeadSpec(nc=1, strides=(8,))
    with pytest.raises(RuntimeError):
        sarpost.decode([torch.zeros(1, 65, 4, 4)], spec)
    with pytest.raises(RuntimeError):
        sarpost.postprocess_fused([torch.zeros(1, 65, 4, 4)], spec)
    with pytest.raises(AssertionError, match="Invalid Confidence threshold"):
        sarpost.non_max_suppression(y, conf_thres=2.0)
    with pytest.raises(AssertionError, match="Invalid IoU"):
        sarpost.non_max_suppression(y, iou_thres=-1)


def test_headspec_from_module(sarpost):
    m = types.SimpleNamespace(nc=1, stride=torch.tensor([8.0, 16.0, 32.0]), reg_max=16, embed_dim=256, state_classes=6)
    s = sarpost.HeadSpec.from_module(m)
    assert (s.no, s.nm, s.strides) == (327, 262, (8.0, 16.0, 32.0))
    d = sarpost.HeadSpec.from_module(types.SimpleNamespace(nc=6, stride=[8, 16, 32], reg_max=16))
    assert (d.no, d.nm) == (70, 0)
    j = sarpost.HeadSpec.from_module(types.SimpleNamespace(nc=1, stride=[8], reg_max=16, embed_dim=128, state_classes=None))
    assert (j.no, j.nm) == (193, 128)


def test_level_shapes_and_anchor_counts(sarpost):
    s = sarpost.synth.level_shapes(640, (8, 16, 32))
    assert s == [(80, 80), (40, 40), (20, 20)] and sum(h * w for h, w in s) == 8400
    p2 = sarpost.synth.level_shapes(1280, (4, 8, 16, 32))
    assert sum(h * w for h, w in p2) == 136000
    from oracle import postprocess_ref as R
    pts, st = R.make_anchors_ref([(2, 3), (1, 2)], [8, 16])  # anchors are analytic inside the kernels; the oracle restates tal.py
    assert pts.tolist() == [[0.5, 0.5], [1.5, 0.5], [2.5, 0.5], [0.5, 1.5], [1.5, 1.5], [2.5, 1.5], [0.5, 0.5], [1.5, 0.5]]
    assert st.flatten().tolist() == [8.0] * 6 + [16.0] * 2


def test_plugin_patch_and_unpatch_forward_unaccelerated_calls(sarpost):
    calls = []

    def ref_nms(prediction, *a, **k):
        calls.append("nms")
        return ["ref"]

    class Detect:
        export = False
        reg_max = 16

        def _inference(self, x):
            calls.append("detect")
            return "ref-y"

    class JDE(Detect):
        def _inference(self, x):
            calls.append("jde")
            return "ref-y"

    ops_mod = types.SimpleNamespace(non_max_suppression=ref_nms)
    head_mod = types.SimpleNamespace(Detect=Detect, JDE=JDE)
    sarpost.patch(ops_mod, head_mod)
    try:
        assert sarpost.plugin.is_patched()
        assert ops_mod.non_max_suppression is not ref_nms
        # CPU tensors and rotated boxes go to the reference's own function
        y = torch.zeros(1, 6, 10)
        assert ops_mod.non_max_suppression(y) == ["ref"]
        assert ops_mod.non_max_suppression((y, None), 0.25, 0.45, None, False, False, (), 300, 0, 0.05, 30000, 7680, True, True) == ["ref"]
        assert Detect()._inference([torch.zeros(1, 65, 2, 2)]) == "ref-y"
        assert JDE()._inference([torch.zeros(1, 65, 2, 2)]) == "ref-y"
        assert calls == ["nms", "nms", "detect", "jde"]
    finally:
        sarpost.unpatch()
    assert ops_mod.non_max_suppression is ref_nms and not sarpost.plugin.is_patched()
    assert Detect._inference(Detect(), [torch.zeros(1)]) == "ref-y"


def test_shard_ranges_cover_everything_once(sarpost):
    for n in (0, 1, 7, 64, 512, 513):
        for w in (1, 2, 3, 8):
            r = [sarpost.dist.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    assert sarpost.dist.shard_frames(10, 48, 1, 4) == (3 * 48, 6 * 48)


def test_sahi_grid_cfg4(sarpost):
    g = sarpost.dist.sahi_grid(4000, 3000, 640, 0.2)
    assert g.shape == (48, 2)  # 8 x 6 tiles
    assert g[:, 0].max().item() == 4000 - 640 and g[:, 1].max().item() == 3000 - 640
    assert sarpost.dist.sahi_grid(600, 500).tolist() == [[0.0, 0.0]]


GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import sarpost
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=rank, world_size=world)
per, max_det, row = 3, 5, 6
out = torch.full((per, max_det, row), float(rank + 1))
counts = torch.tensor([rank + 1, 0, max_det], dtype=torch.int32)
g_out, g_cnt = sarpost.dist.allgather_detections(out, counts)
assert g_out.shape == (world * per, max_det, row) and g_cnt.tolist() == [1, 0, 5, 2, 0, 5]
assert bool((g_out[:per] == 1).all()) and bool((g_out[per:] == 2).all())
buf = sarpost.dist.GatherBuffer(per, max_det, row, "cpu")
buf.local_out.fill_(10.0 * (rank + 1)); buf.local_counts.copy_(counts)
o, c = buf.exchange()
assert bool((o[:per] == 10).all()) and bool((o[per:] == 20).all()) and c.tolist() == [1, 0, 5, 2, 0, 5]
lo, hi = sarpost.dist.shard_range(7, rank, world)
assert (lo, hi) == ((0, 4) if rank == 0 else (4, 7))
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_allgather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, port], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in o, o


def test_patch_on_live_reference_modules(sarpost):
    """With the real reference modules (build container only): patch() swaps the three attributes the call sites
    resolve at call time, CPU calls still reach the reference's own code (identical results), unpatch() restores."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    ops, _, head = ref_shim.load()
    orig_nms, orig_det, orig_jde = ops.non_max_suppression, head.Detect._inference, head.JDE._inference
    y = sarpost.synth.decoded_prediction(2, 500, 3, 2, seed=4)
    want = orig_nms(y.clone(), 0.25, 0.6, nc=3)
    shapes = sarpost.synth.level_shapes(64, (8, 16, 32))
    levels = sarpost.synth.head_outputs(1, shapes, 1, 16, 6, seed=5)
    m = head.JDE(nc=1, embed_dim=16, state_classes=6, ch=(64, 64, 64))
    m.stride = torch.tensor([8.0, 16.0, 32.0])
    m.eval()
    want_y = m._inference([x.clone() for x in levels])
    for fused in (False, True):
        sarpost.patch(fused=fused)
        try:
            assert ops.non_max_suppression is not orig_nms and head.JDE._inference is not orig_jde
            got = ops.non_max_suppression(y.clone(), 0.25, 0.6, nc=3)          # CPU tensor -> reference code path
            assert all(torch.equal(a, b) for a, b in zip(got, want))
            assert torch.equal(m._inference([x.clone() for x in levels]), want_y)  # CPU levels -> reference decode
        finally:
            sarpost.unpatch()
    assert ops.non_max_suppression is orig_nms and head.Detect._inference is orig_det and head.JDE._inference is orig_jde
    # defer_state also swaps JDE.forward; CPU features still run the reference's own forward (state_predictor on every anchor)
    orig_fwd = head.JDE.forward
    feats = [torch.randn(1, 64, h, w, generator=torch.Generator().manual_seed(i)) for i, (h, w) in enumerate(shapes)]
    with torch.no_grad():
        want_full, want_x = m([f.clone() for f in feats])
        sarpost.patch(fused=True, defer_state=True)
        try:
            assert head.JDE.forward is not orig_fwd
            got_full, got_x = m([f.clone() for f in feats])
            assert torch.equal(got_full, want_full) and all(torch.equal(a, b) for a, b in zip(got_x, want_x))
        finally:
            sarpost.unpatch()
    assert head.JDE.forward is orig_fwd
    with pytest.raises(ValueError):
        sarpost.patch(defer_state=True)


def test_only_plain_detect_and_jde_heads_take_the_fast_decode(sarpost):
    """`patch()` replaces `Detect._inference` on the base class, so every subclass inherits it.  OBB decodes
    with dist2rbox (head.py:303), end2end/v10 heads expect xyxy (head.py:147), Pose/Segment read the anchor cache only
    the reference `_inference` fills (head.py:354): all of those must keep the reference's own method.  Checked on the
    live reference classes (build container) — and, below, on stand-in classes so the rule is also covered on the GPU box."""
    from oracle import ref_shim
    P = sarpost.plugin
    if ref_shim.available():
        _, _, head = ref_shim.load()
        sarpost.patch()
        try:
            ch = (16, 16, 16)
            assert P._plain_head(head.Detect(nc=2, ch=ch), "detect")
            assert P._plain_head(head.JDE(nc=1, embed_dim=16, state_classes=6, ch=ch), "jde")
            assert P._plain_head(head.JDE(nc=1, embed_dim=16, ch=ch), "jde")
            for sub in (head.OBB(nc=2, ne=1, ch=ch), head.Pose(nc=1, kpt_shape=(17, 3), ch=ch), head.Segment(nc=2, nm=8, npr=16, ch=ch),
                        head.v10Detect(nc=2, ch=ch)):
                assert not P._plain_head(sub, "detect"), type(sub).__name__
                assert not P._plain_head(sub, "jde"), type(sub).__name__
            d = head.Detect(nc=2, ch=ch)
            d.export = True
            assert not P._plain_head(d, "detect")
            # patch -> (CPU call: reference decode, fills the anchor cache) -> unpatch -> reference call: same y, cache intact
            m = head.Detect(nc=2, ch=ch)
            m.stride = torch.tensor([8.0, 16.0, 32.0])
            m.eval()
            lv = [torch.randn(1, m.no, s, s) for s in (8, 4, 2)]
            y_patched = m._inference([x.clone() for x in lv])
        finally:
            sarpost.unpatch()
        assert m.anchors.shape[-1] == 8 * 8 + 4 * 4 + 2 * 2
        assert torch.equal(m._inference([x.clone() for x in lv]), y_patched)

    class Detect:
        export, end2end, reg_max = False, False, 16

        def decode_bboxes(self, b, a):
            return b

        def _inference(self, x):
            return "ref"

    class JDE(Detect):
        pass

    class OBB(Detect):
        def decode_bboxes(self, b, a):
            return "rotated"

    class Pose(Detect):
        pass

    class V10(Detect):
        end2end = True

    sarpost.patch(types.SimpleNamespace(non_max_suppression=lambda *a, **k: None), types.SimpleNamespace(Detect=Detect, JDE=JDE))
    try:
        assert P._plain_head(Detect(), "detect") and P._plain_head(JDE(), "jde")
        assert not P._plain_head(JDE(), "detect") and not P._plain_head(Detect(), "jde")
        for sub in (OBB(), Pose(), V10()):
            assert not P._plain_head(sub, "detect")
            assert sub._inference([torch.zeros(1, 65, 2, 2)]) == "ref"
    finally:
        sarpost.unpatch()


def test_nms_dispatch_forwards_what_the_library_rejects(sarpost):
    """argument ranges libsarpost rejects (max_det > 4096, nc > 2048, max_nms < 1, odd dtypes/ranks) must run the
    reference's function instead of raising inside an unmodified predictor.  `_nms_supported` is the gate (pure host logic;
    a meta-device tensor stands in for a CUDA one)."""
    P = sarpost.plugin
    y = torch.empty(2, 10, 50, device="meta")

    def sup(t, **kw):
        return P._nms_supported_fields(True, t.dim(), t.dtype, tuple(t.shape), kw)

    assert sup(y)
    assert sup(y, max_det=4096) and not sup(y, max_det=4097) and not sup(y, max_det=0)
    assert not sup(y, max_nms=0)
    assert not sup(y, rotated=True)
    assert not sup(torch.empty(2, 4000, 50, device="meta"))            # nc = 3996 > 2048
    assert sup(torch.empty(2, 4000, 50, device="meta"), nc=80)
    assert not sup(torch.empty(2, 10, 50, device="meta", dtype=torch.int32))
    assert not sup(torch.empty(10, 50, device="meta"))
    assert not P._nms_supported_fields(False, 3, torch.float32, (2, 10, 50), {})  # CPU tensor


def test_scale_params_equal_reference_host_arithmetic(sarpost):
    """ops.scale_params = the Python-float part of ops.scale_boxes (utils/ops.py:110-116): gain, rounded pads, clip bounds."""
    from oracle import postprocess_ref as R
    img1 = (640, 512)
    shapes = [(1080, 1920, 3), (333, 777), (640, 512), (17, 4000, 3)]
    prm = sarpost.ops.scale_params(img1, shapes, "cpu")
    assert prm.shape == (4, 5) and prm.dtype == torch.float32
    boxes = torch.tensor([[0.0, 0.0, 512.0, 640.0], [100.5, 200.25, 300.75, 400.125]])
    for row, s0 in zip(prm.tolist(), shapes):
        px, py, gain, w0, h0 = row
        got = boxes.clone()
        got[:, [0, 2]] = ((got[:, [0, 2]] - px) / gain).clamp(0, w0)   # what the gather kernel does, in torch fp32
        got[:, [1, 3]] = ((got[:, [1, 3]] - py) / gain).clamp(0, h0)
        assert torch.equal(got, R.scale_boxes_ref(img1, boxes, s0))


def test_state_mlp_validation_and_error_types(sarpost):
    w1, b1, w2, b2 = torch.zeros(4, 8), torch.zeros(4), torch.zeros(2, 4), torch.zeros(2)
    with pytest.raises(RuntimeError, match="only CUDA"):        # no CPU fallback for the deferred state head either
        sarpost.StateMLP.from_tensors(w1, b1, w2, b2)
    with pytest.raises(ValueError, match="2 Linear layers"):
        sarpost.StateMLP.from_module(torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.ReLU()))
    with pytest.raises(RuntimeError, match="only CUDA"):
        sarpost.state_head(torch.zeros(1, 3, 16), torch.zeros(1, dtype=torch.int32), None)
    assert issubclass(sarpost.SarpostError, RuntimeError)


def test_fused_dispatch_argument_mapping_on_cpu(sarpost):
    """The patched NMS takes the reference's positional order (ops.py:167-182): a CPU prediction with positional
    arguments must reach the ORIGINAL function unchanged, under both patch modes."""
    import types
    seen = {}

    def ref_nms(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, labels=(),
                max_det=300, nc=0, max_time_img=0.05, max_nms=30000, max_wh=7680, in_place=True, rotated=False):
        seen.update(conf_thres=conf_thres, iou_thres=iou_thres, classes=classes, agnostic=agnostic, max_det=max_det, nc=nc, rotated=rotated)
        return "ref"

    Detect = type("Detect", (), {"_inference": lambda self, x: "ref-decode"})
    for fused in (False, True):
        ops_mod = types.SimpleNamespace(non_max_suppression=ref_nms)
        head_mod = types.SimpleNamespace(Detect=Detect)
        sarpost.patch(ops_mod, head_mod, fused=fused)
        try:
            y = torch.zeros(1, 6, 10)
            assert ops_mod.non_max_suppression((y, None), 0.3, 0.6, [0], True, max_det=5, nc=2) == "ref"
            assert seen == dict(conf_thres=0.3, iou_thres=0.6, classes=[0], agnostic=True, max_det=5, nc=2, rotated=False)
            assert Detect()._inference([torch.zeros(1, 66, 2, 2)]) == "ref-decode"
        finally:
            sarpost.unpatch()
        assert ops_mod.non_max_suppression is ref_nms


def test_product_code_never_touches_the_oracle_or_a_cpu_fallback():
    """The oracle is test infrastructure: nothing under the package (Python or CUDA) may import, link or call it, nor
    torchvision's NMS, Triton or torch.compile; bench.py may use it only in its CPU legs."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "sar-yolo_b200")
    banned = re.compile(r"\b(import\s+oracle|from\s+oracle|oracle\.|torchvision|import\s+triton|torch\.compile)\b")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                src = open(os.path.join(dirpath, fn), encoding="utf-8").read()
                hit = [ln for ln in src.splitlines() if banned.search(ln) and not ln.lstrip().startswith(("#", "//", "*"))
                       and "no `torchvision" not in ln and "torchvision's" not in ln and "torchvision." not in ln.split("#")[-1]]
                hit = [ln for ln in hit if re.search(r"^\s*(import|from)\s+(oracle|torchvision|triton)|torch\.compile\(|oracle\.", ln)]
                assert not hit, f"{fn}: {hit[:2]}"
    bench = open(os.path.join(root, "bench.py"), encoding="utf-8").read()
    uses = [m.start() for m in re.finditer(r"from oracle|import oracle", bench)]
    assert uses, "bench.py's CPU legs use the oracle port"
    for at in uses:  # every use sits inside the baseline class both CPU legs and the reference_gpu leg go through
        owner = re.findall(r"\n(?:class|def) (\w+)[(:]", bench[:at])[-1]
        assert owner == "ReferencePath", owner
    # ... and the GPU arm's own step never goes near it: `ReferencePath` is only constructed in the three baseline legs
    for m in re.finditer(r"ReferencePath\(", bench):
        fn = re.findall(r"\ndef (\w+)\(", bench[:m.start()])[-1]
        assert fn in ("run_reference_arm", "leg_reference_gpu", "main"), fn


def test_level_signature_tracks_layout_shape_dtype_and_memory_format():
    """`_LevelSig` (the check a prepared plan / the pipeline's head cache runs instead of full validation): same geometry
    with new storage matches; another batch, dtype, layout, memory format or a missing branch does not."""
    import torch

    import sarpost
    from sarpost.ops import _LevelSig

    spec = sarpost.HeadSpec(nc=2, strides=(8, 16), embed_dim=4, state_classes=3)
    mk = lambda bs=2, dt=torch.float32: [torch.zeros(bs, spec.no, h, w, dtype=dt) for h, w in ((4, 6), (2, 3))]  # noqa: E731
    cat = mk()
    sig = _LevelSig(cat)
    assert sig.matches(mk()) and not sig.split
    assert not sig.matches(mk(bs=3)) and not sig.matches(mk(dt=torch.float16)) and not sig.matches(mk()[:1])
    assert not sig.matches([x.permute(0, 1, 3, 2) for x in mk()])            # same numel, other shape / strides
    assert not sig.matches([x[:, :, :, ::2] for x in mk()])
    split = sarpost.split_levels(cat, spec, emb_channels_last=True)
    assert not sig.matches(split)
    ssig = _LevelSig(split)
    assert ssig.split and ssig.matches(sarpost.split_levels(mk(), spec, emb_channels_last=True))
    assert not ssig.matches(sarpost.split_levels(mk(), spec, emb_channels_last=False))  # NCHW embedding where channels_last was prepared
    assert not ssig.matches([lv[:2] + (None, lv[3]) for lv in split])
    # addresses land in the right arrays of a head / io block
    io = sarpost._lib.PlanIO()
    ssig.fill(split, io)
    for i, lv in enumerate(split):
        assert (io.data[i], io.cls[i], io.emb[i], io.state[i]) == tuple(t.data_ptr() for t in lv)
    assert io.data[len(split)] is None
