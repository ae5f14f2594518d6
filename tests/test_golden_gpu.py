"""GPU: the CUDA path against the committed golden vectors (outputs of the live reference)."""
import pytest
import torch

from golden_util import canon, head_inputs, load, load_state_case, names, nms_case_inputs

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
