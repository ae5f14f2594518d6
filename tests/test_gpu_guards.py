"""GPU: out-of-bounds WRITE detection without compute-sanitizer (closed on this pool).  Every output / workspace buffer
handed to the C ABI is an interior slice of a larger allocation whose guard bands are filled with a pattern; after the
call the bands must be untouched and the results must equal the normal wrapper's."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 8192
PAT = 0xA5


class Guarded:
    def __init__(self, nbytes, dev):
        self.n = int(nbytes)
        self.buf = torch.full((self.n + 2 * GUARD,), PAT, dtype=torch.uint8, device=dev)
        self.ptr = self.buf.data_ptr() + GUARD

    def view(self, dtype, shape):
        return self.buf[GUARD:GUARD + self.n].view(dtype).view(shape)

    def intact(self):
        return bool((self.buf[:GUARD] == PAT).all()) and bool((self.buf[GUARD + self.n:] == PAT).all())


@pytest.mark.parametrize("imgsz,strides,nc,ed,sc,bs,kw,half", [
    (160, (8, 16, 32), 1, 16, 6, 3, dict(conf_thres=0.001, iou_thres=0.7), False),
    (320, (4, 8, 16, 32), 6, 0, 0, 2, dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=1000), False),
    ((88, 120), (8, 16, 32), 2, 8, 0, 5, dict(conf_thres=0.0, iou_thres=1.0, max_det=7, max_nms=50), False),   # odd levels: LDG path
    (96, (8, 16, 32), 3, 4, 6, 38, dict(conf_thres=0.25, iou_thres=0.45, agnostic=True), True),                # B > SMs/4: cluster 2
    (64, (16,), 17, 0, 0, 150, dict(conf_thres=0.05, iou_thres=0.5, multi_label=True, max_det=33), False),     # B > SMs: no cluster
])
def test_fused_writes_stay_inside_their_buffers(sarpost, cuda, imgsz, strides, nc, ed, sc, bs, kw, half):
    ops, lib = sarpost.ops, sarpost._lib.lib
    spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
    levels = [x.to(cuda) for x in sarpost.synth.head_outputs(bs, sarpost.synth.level_shapes(imgsz, strides), nc, ed, sc, seed=5, cls_mean=-1.0, blobs=2)]
    if half:
        levels = [x.half() for x in levels]
    want, want_idx = sarpost.postprocess_fused(levels, spec, return_index=True, **kw)
    full = dict(conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, max_det=300, max_nms=30000, max_wh=7680)
    full.update(kw)
    params, _keep = ops._make_params(**full)
    params.workspace_clean = 0
    head = ops._make_head(levels, spec)
    anchors = sum(int(x.shape[2]) * int(x.shape[3]) for x in levels)
    md, row = int(full["max_det"]), 6 + spec.nm
    ws_bytes = lib.sarpost_workspace_bytes(bs, anchors, nc, int(full["multi_label"]), md)
    ws, out, cnt, kid = Guarded(ws_bytes, cuda), Guarded(bs * md * row * 4, cuda), Guarded(bs * 4, cuda), Guarded(bs * md * 4, cuda)
    for rep in range(2):  # second call: the workspace now holds the previous call's leftovers
        rc = lib.sarpost_fused(C.byref(head), C.byref(params), out.ptr, cnt.ptr, kid.ptr, ws.ptr, ws_bytes,
                               torch.cuda.current_stream().cuda_stream)
        assert rc == 0, sarpost._lib.lib.sarpost_last_error()
        torch.cuda.synchronize()
        assert ws.intact() and out.intact() and cnt.intact() and kid.intact(), f"guard band overwritten (call {rep})"
        counts = cnt.view(torch.int32, (bs,)).tolist()
        rows = out.view(torch.float32, (bs, md, row))
        kidx = kid.view(torch.int32, (bs, md))
        assert counts == [r.shape[0] for r in want]
        for b in range(bs):
            assert torch.equal(rows[b, :counts[b]], want[b]) and torch.equal(kidx[b, :counts[b]], want_idx[b])
            assert bool((rows[b, counts[b]:].view(torch.uint8) == PAT).all()), "rows beyond counts[b] must stay unwritten"


def test_decoded_and_merge_writes_stay_inside_their_buffers(sarpost, cuda):
    ops, lib = sarpost.ops, sarpost._lib.lib
    stream = torch.cuda.current_stream().cuda_stream
    # decoded-input NMS with apriori labels
    bs, na, nc, nm, md = 3, 3001, 4, 5, 120
    y = sarpost.synth.decoded_prediction(bs, na, nc, nm, seed=9, clustered=True).to(cuda)
    kw = dict(conf_thres=0.1, iou_thres=0.5, multi_label=True, max_det=md)
    want = sarpost.non_max_suppression(y, nc=nc, **kw)
    full = dict(conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, max_det=300, max_nms=30000, max_wh=7680)
    full.update(kw)
    params, _keep = ops._make_params(**full)
    params.workspace_clean = 0
    ws_bytes = lib.sarpost_workspace_bytes(bs, na, nc, 1, md)
    ws, out, cnt = Guarded(ws_bytes, cuda), Guarded(bs * md * (6 + nm) * 4, cuda), Guarded(bs * 4, cuda)
    rc = lib.sarpost_nms_decoded(y.data_ptr(), bs, 4 + nc + nm, na, nc, C.byref(params), out.ptr, cnt.ptr, None, ws.ptr, ws_bytes, stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert ws.intact() and out.intact() and cnt.intact()
    counts = cnt.view(torch.int32, (bs,)).tolist()
    for b in range(bs):
        assert torch.equal(out.view(torch.float32, (bs, md, 6 + nm))[b, :counts[b]], want[b])
    # cross-tile merge
    tpf, d, nf, rl = 12, 50, 3, 9
    g = torch.Generator().manual_seed(3)
    dets = torch.rand(nf * tpf, d, rl, generator=g) * 100
    dets[..., 2:4] += dets[..., 0:2] + 5
    dets[..., 5] = torch.randint(0, 3, (nf * tpf, d), generator=g).float()
    dcnt = torch.randint(0, d + 1, (nf * tpf,), generator=g, dtype=torch.int32)
    org = (torch.rand(tpf, 2, generator=g) * 300).floor().repeat(nf, 1)
    want = sarpost.merge_tiles(dets.to(cuda), dcnt.to(cuda), org.to(cuda), tpf, iou_thres=0.5, max_det=md)
    params, _keep = ops._make_params(0.0, 0.5, None, False, False, md, 30000, 7680)
    params.workspace_clean = 0
    ws_bytes = lib.sarpost_merge_workspace_bytes(nf, tpf, d, md)
    ws, out, cnt = Guarded(ws_bytes, cuda), Guarded(nf * md * rl * 4, cuda), Guarded(nf * 4, cuda)
    dd, dc, oo = dets.to(cuda).contiguous(), dcnt.to(cuda), org.to(cuda).contiguous()
    rc = lib.sarpost_merge_tiles(dd.data_ptr(), dc.data_ptr(), oo.data_ptr(), nf, tpf, d, rl, C.byref(params), out.ptr, cnt.ptr, None,
                                 ws.ptr, ws_bytes, stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert ws.intact() and out.intact() and cnt.intact()
    counts = cnt.view(torch.int32, (nf,)).tolist()
    for f in range(nf):
        assert torch.equal(out.view(torch.float32, (nf, md, rl))[f, :counts[f]], want[f])


def test_decode_extras_match_state_writes_stay_inside_their_buffers(sarpost, cuda):
    ops, lib = sarpost.ops, sarpost._lib.lib
    stream = torch.cuda.current_stream().cuda_stream
    strides, nc, ed, sc, bs = (8, 16, 32), 2, 8, 6, 3
    spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
    shapes = sarpost.synth.level_shapes((88, 120), strides)
    anchors = sum(h * w for h, w in shapes)
    for half in (False, True):  # API-exact decode, y in the input dtype
        levels = [x.to(cuda) for x in sarpost.synth.head_outputs(bs, shapes, nc, ed, sc, seed=2)]
        if half:
            levels = [x.half() for x in levels]
        want = sarpost.decode(levels, spec)
        head = ops._make_head(levels, spec)
        esz = 2 if half else 4
        y = Guarded(bs * (4 + nc + ed + sc) * anchors * esz, cuda)
        assert lib.sarpost_decode(C.byref(head), y.ptr, stream) == 0
        torch.cuda.synchronize()
        assert y.intact() and torch.equal(y.view(want.dtype, tuple(want.shape)), want)
    # extras of explicit (image, anchor) pairs, including out-of-range ones (zero rows)
    ii = torch.tensor([0, 2, 1, 5, -1, 0], dtype=torch.int32, device=cuda)
    ai = torch.tensor([0, anchors - 1, 17, 3, 4, anchors], dtype=torch.int32, device=cuda)
    levels = [x.to(cuda) for x in sarpost.synth.head_outputs(bs, shapes, nc, ed, sc, seed=2)]
    want = sarpost.gather_extras(levels, spec, ii, ai)
    head = ops._make_head(levels, spec)
    ex = Guarded(6 * (ed + sc) * 4, cuda)
    assert lib.sarpost_gather_extras(C.byref(head), ii.data_ptr(), ai.data_ptr(), 6, ex.ptr, stream) == 0
    torch.cuda.synchronize()
    assert ex.intact() and torch.equal(ex.view(torch.float32, (6, ed + sc)), want)
    # validator matching
    from test_oracle import _match_case
    md, mg, nt = 40, 16, 10
    dets = torch.zeros(2, md, 7); gtb = torch.zeros(2, mg, 4); gtc = torch.zeros(2, mg)
    dn, gn = [], []
    for j in range(2):
        d_, g_, c_ = _match_case(20 + j, n_det=35 + j, n_gt=12 + j)
        dets[j, : d_.shape[0], :6] = d_; gtb[j, : g_.shape[0]] = g_; gtc[j, : g_.shape[0]] = c_
        dn.append(d_.shape[0]); gn.append(g_.shape[0])
    iouv = [0.5 + 0.05 * i for i in range(nt)]
    want_c, want_m = sarpost.match_predictions(dets.to(cuda), torch.tensor(dn), gtb.to(cuda), gtc.to(cuda), torch.tensor(gn), iouv, tag_threshold_index=0)
    cor, mat = Guarded(2 * md * nt, cuda), Guarded(2 * md * 4, cuda)
    dd, gb, gc = dets.to(cuda), gtb.to(cuda), gtc.to(cuda)
    dnt, gnt = torch.tensor(dn, dtype=torch.int32, device=cuda), torch.tensor(gn, dtype=torch.int32, device=cuda)
    thr = (C.c_float * nt)(*iouv)
    assert lib.sarpost_match_predictions(dd.data_ptr(), dnt.data_ptr(), 2, md, 7, gb.data_ptr(), gc.data_ptr(), gnt.data_ptr(), mg, thr, nt,
                                         cor.ptr, mat.ptr, 0, stream) == 0
    torch.cuda.synchronize()
    assert cor.intact() and mat.intact()
    assert torch.equal(cor.view(torch.uint8, (2, md, nt)).bool(), want_c) and torch.equal(mat.view(torch.int32, (2, md)), want_m)
    # deferred state head, both kernels: only the state columns of rows < counts[b] may change
    import os
    for e, h, s in ((32, 16, 6), (10, 5, 3)):
        g = torch.Generator().manual_seed(e)
        mlp = sarpost.StateMLP.from_tensors(torch.randn(h, e, generator=g), torch.randn(h, generator=g), torch.randn(s, h, generator=g),
                                            torch.randn(s, generator=g), device=cuda)
        md, rl = 21, 6 + e + s
        src = torch.randn(2, md, rl, generator=g).to(cuda)
        cnts = torch.tensor([md, 5], dtype=torch.int32, device=cuda)
        rows = Guarded(2 * md * rl * 4, cuda)
        rows.view(torch.float32, (2, md, rl)).copy_(src)
        assert lib.sarpost_state_head(rows.ptr, cnts.data_ptr(), 2, md, rl, 6, e, 6 + e, s, h, mlp.w1.data_ptr(), mlp.b1.data_ptr(),
                                      mlp.w2.data_ptr(), mlp.b2.data_ptr(), stream) == 0
        torch.cuda.synchronize()
        got = rows.view(torch.float32, (2, md, rl))
        assert rows.intact() and torch.equal(got[..., : 6 + e], src[..., : 6 + e]) and torch.equal(got[1, 5:], src[1, 5:])
        assert not torch.equal(got[0, :, 6 + e:], src[0, :, 6 + e:])
