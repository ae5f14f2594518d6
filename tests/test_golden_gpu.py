"""GPU: the CUDA path against the committed golden vectors (outputs of the live reference)."""
import pytest
import torch

from golden_util import canon, head_inputs, load, load_state_case, load_tail_case, names, nms_case_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", names("nms_"))
def test_cuda_nms_equals_reference_golden(sarpost, cuda, name):
    g = load(name)
    y, kw = nms_case_inputs(sarpost, g["meta"])
    if "labels" in kw:
        kw["labels"] = [lb.to(cuda) for lb in kw["labels"]]
    rows = sarpost.non_max_suppression(y.to(cuda), **kw)
    assert [r.shape[0] for r in rows] == g["counts"]
    for a, b in zip(rows, g["rows"]):
        # bit-exact kept rows, classes and extras; canon() only re-orders rows with exactly equal scores
        # (the reference's max_nms cut uses an unstable argsort, see golden_util.canon)
        assert torch.equal(canon(a), canon(b))
        assert bool((a[:-1, 4] >= a[1:, 4]).all())


@pytest.mark.parametrize("name", names("head_"))
def test_cuda_decode_and_fused_vs_reference_golden(sarpost, cuda, name):
    g = load(name)
    m = g["meta"]
    shapes, levels = head_inputs(sarpost, m)
    levels = [x.to(cuda) for x in levels]
    spec = sarpost.HeadSpec(nc=m["nc"], strides=tuple(m["strides"]), embed_dim=m["ed"], state_classes=m["sc"])
    y = sarpost.decode(levels, spec).cpu()
    assert tuple(y.shape) == g["y_shape"]
    ys, ref = y[:, :, :: m["sub"]], g["y_sub"]
    st = torch.cat([torch.full((h * w,), float(s)) for (h, w), s in zip(shapes, m["strides"])])[:: m["sub"]]
    err = (ys[:, :4] - ref[:, :4]).abs()
    assert bool((err <= 1e-5 * ref[:, :4].abs() + 1e-5 * st).all()), err.max().item()
    assert torch.allclose(ys[:, 4:], ref[:, 4:], rtol=1e-5, atol=1e-7)
    # end to end from raw logits: detections match the reference's within the decode tolerance;
    # an order/keep flip caused by a 1-ulp score difference would show up here (budget < 1e-4)
    rows = sarpost.postprocess_fused(levels, spec, **m["kw"])
    total = bad = 0
    for a, b in zip(rows, g["rows"]):
        a = a.cpu()
        total += max(a.shape[0], b.shape[0])
        if a.shape != b.shape:
            bad += abs(a.shape[0] - b.shape[0]) + 1
            continue
        ok = torch.isclose(canon(a), canon(b), rtol=1e-5, atol=1e-5 * max(m["strides"])).all(1)
        bad += int((~ok).sum())
    assert bad <= 1e-4 * total, f"{bad} of {total} detections differ from the reference"


def test_cuda_deferred_state_head_vs_reference_golden(sarpost, cuda):
    """§8f row 2: head WITHOUT the state channels + sarpost_state_head on the kept rows == the reference's rows, where
    state_predictor ran on every anchor inside JDE.forward (head.py:198-204).  Tolerance: 1e-5 on the state
    probabilities (fp32 summation order), the other columns as in the head goldens."""
    g = load_state_case()
    m = g["meta"]
    nc, ed, sc = m["nc"], m["ed"], m["sc"]
    spec = sarpost.HeadSpec(nc=nc, strides=tuple(m["strides"]), embed_dim=ed, state_classes=sc)
    mlp = sarpost.StateMLP.from_tensors(g["w1"], g["b1"], g["w2"], g["b2"], device=cuda)
    full = sarpost.postprocess_fused([x.to(cuda) for x in g["levels"]], spec, **m["kw"])
    deferred = sarpost.postprocess_fused([x[:, : 64 + nc + ed].contiguous().to(cuda) for x in g["levels"]], spec, state_mlp=mlp, **m["kw"])
    for a, d, b in zip(full, deferred, g["rows"]):
        assert a.shape == d.shape == b.shape
        assert torch.equal(a[:, : 6 + ed], d[:, : 6 + ed])                       # same kernels, same bits
        assert torch.allclose(d[:, 6 + ed:], a[:, 6 + ed:], rtol=0, atol=1e-5)   # deferred vs per-anchor (our sigmoid of the ref logits)
        assert torch.allclose(d.cpu()[:, : 6 + ed], b[:, : 6 + ed], rtol=1e-5, atol=1e-5 * max(m["strides"]))
        assert torch.allclose(d.cpu()[:, 6 + ed:], b[:, 6 + ed:], rtol=0, atol=1e-5)


def test_cuda_results_layout_vs_live_predictor_tail_golden(sarpost, cuda):
    """§8f row 1: `postprocess_fused(..., results=True, scale_to=...)` — boxes (n, 7) = x1,y1,x2,y2,state_id,conf,cls in
    original-image pixels + contiguous embeds written by the gather kernel — against what the live, unmodified
    JDEPredictor.postprocess (models/yolo/jde/predict.py:29-78) returned; from the raw levels (state from the levels'
    own state channels), with the deferred state head (`sarpost_state_ids`), and from the reference's decoded y."""
    st, tail = load_state_case(), load_tail_case()
    m, t = st["meta"], tail["meta"]["tail"]
    nc, ed, sc = m["nc"], m["ed"], m["sc"]
    spec = sarpost.HeadSpec(nc=nc, strides=tuple(m["strides"]), embed_dim=ed, state_classes=sc)
    kw = dict(m["kw"], scale_to=(tuple(t["img_shape"]), [tuple(s) for s in t["orig_shapes"]]), results=True)
    levels = [x.to(cuda) for x in st["levels"]]
    mlp = sarpost.StateMLP.from_tensors(st["w1"], st["b1"], st["w2"], st["b2"], device=cuda)
    full = sarpost.postprocess_fused(levels, spec, **kw)
    split = sarpost.postprocess_fused(sarpost.split_levels(levels, spec), spec, **kw)
    deferred = sarpost.postprocess_fused([x[:, : 64 + nc + ed].contiguous() for x in levels], spec, state_mlp=mlp, **kw)
    for got in (full, split, deferred):
        assert [int(b.shape[0]) for b in got[0]] == tail["counts"]
        for bx, em, rb, re in zip(got[0], got[1], tail["boxes"], tail["embeds"]):
            bx, em = bx.cpu(), em.cpu()
            assert tuple(bx.shape) == tuple(rb.shape) and em.is_contiguous()
            assert torch.allclose(bx[:, :4], rb[:, :4], rtol=1e-5, atol=1e-4)  # decode tolerance (1e-5 * stride) through the 1/gain scaling
            assert torch.equal(bx[:, 4], rb[:, 4]) and torch.equal(bx[:, 6], rb[:, 6])  # state id, class: exact
            assert torch.allclose(bx[:, 5], rb[:, 5], rtol=1e-6, atol=0) and torch.allclose(em, re, rtol=1e-6, atol=1e-7)
    for a, b in zip(full[0] + full[1], split[0] + split[1]):
        assert torch.equal(a, b)
    # the rows layout of the same call holds the same numbers (scale_boxes, embedding columns); ids = argmax of its states
    rows = sarpost.postprocess_fused(levels, spec, **dict(kw, results=False))
    for r, bx, em in zip(rows, full[0], full[1]):
        assert torch.equal(r[:, :4], bx[:, :4]) and torch.equal(r[:, 4:6], bx[:, 5:7]) and torch.equal(r[:, 6:6 + ed], em)
        assert torch.equal(r[:, 6 + ed:].argmax(1).float(), bx[:, 4])
    # decoded entry (the reference's own y): bit-exact boxes after scaling are the oracle's business; here columns + ids
    y = st["y"].to(cuda)
    bx_l, em_l = sarpost.non_max_suppression(y, nc=nc, results_state_cols=sc, **m["kw"])
    ref_rows = st["rows"]
    for bx, em, r in zip(bx_l, em_l, ref_rows):
        assert torch.equal(bx.cpu()[:, :4], r[:, :4]) and torch.equal(bx.cpu()[:, 5:7], r[:, 4:6])
        assert torch.equal(em.cpu(), r[:, 6:6 + ed]) and torch.equal(bx.cpu()[:, 4], r[:, 6 + ed:].argmax(1).float())
