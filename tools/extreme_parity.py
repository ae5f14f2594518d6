"""Parity at the edges of the supported geometry (one-off, oracle on the host cores takes a while):
   huge anchor counts (tile list > kTileListCap), nc=80 multi-label with the max_nms cut inside a dense bucket,
   max_det=4096 / max_nms=100000, batch > SM count.   python tools/extreme_parity.py"""
import sys, time
import torch
sys.path.insert(0, ".")
import sarpost
from oracle import postprocess_ref as R

dev = torch.device("cuda:0")


def check(name, levels, spec, **kw):
    t0 = time.time()
    rows, idx = sarpost.postprocess_fused([x.to(dev) for x in levels], spec, return_index=True, **kw)
    y = sarpost.decode([x.to(dev) for x in levels], spec).cpu()
    ref_rows, ref_idx = R.non_max_suppression_ref(y, nc=spec.nc, return_index=True, **kw)
    ok = all(torch.equal(a.cpu(), b) and torch.equal(i.cpu().long() // spec.nc, ri[:, 0]) and torch.equal(i.cpu().long() % spec.nc, ri[:, 1])
             for a, b, i, ri in zip(rows, ref_rows, idx, ref_idx))
    # and the decoded-input entry on the same y
    rows2 = sarpost.non_max_suppression(y.to(dev), nc=spec.nc, **kw)
    ok2 = all(torch.equal(a.cpu(), b) for a, b in zip(rows2, ref_rows))
    print(f"{name:44s} fused {'OK' if ok else 'MISMATCH'}  decoded {'OK' if ok2 else 'MISMATCH'}  kept {[r.shape[0] for r in ref_rows][:4]}  {time.time() - t0:.1f}s", flush=True)
    return ok and ok2


def main():
    S = sarpost.synth
    good = True
    strides = (4, 8, 16, 32)
    spec = sarpost.HeadSpec(nc=1, strides=strides, embed_dim=8, state_classes=0)
    good &= check("2560^2 P2, 544k anchors, conf .001", S.head_outputs(2, S.level_shapes(2560, strides), 1, 8, 0, seed=1, cls_mean=-2.0), spec,
                  conf_thres=0.001, iou_thres=0.7)
    good &= check("same, max_det 4096 / max_nms 100000", S.head_outputs(1, S.level_shapes(2560, strides), 1, 8, 0, seed=2, cls_mean=-2.0), spec,
                  conf_thres=0.001, iou_thres=0.7, max_det=4096, max_nms=100000)
    s3 = (8, 16, 32)
    spec80 = sarpost.HeadSpec(nc=80, strides=s3)
    good &= check("nc=80 multi_label 640, conf .001 (672k slots)", S.head_outputs(3, S.level_shapes(640, s3), 80, seed=3, cls_mean=-3.0), spec80,
                  conf_thres=0.001, iou_thres=0.7, multi_label=True)
    good &= check("nc=80 best-class, quantised scores (ties)", [x.mul(2).round().div(2) for x in S.head_outputs(2, S.level_shapes(640, s3), 80, seed=4, cls_mean=-1.0)],
                  spec80, conf_thres=0.05, iou_thres=0.6, agnostic=True)
    spec1 = sarpost.HeadSpec(nc=2, strides=(16,))
    good &= check("batch 300 (> 148 SMs), 64x64", S.head_outputs(300, S.level_shapes(64, (16,)), 2, seed=5, cls_mean=0.0), spec1,
                  conf_thres=0.25, iou_thres=0.5)
    good &= check("all scores equal (one giant bucket, radix path)", [torch.zeros(1, 65, h, w) for h, w in S.level_shapes(640, s3)],
                  sarpost.HeadSpec(nc=1, strides=s3), conf_thres=0.25, iou_thres=0.7)
    print("ALL OK" if good else "FAILURES")
    return 0 if good else 1


if __name__ == "__main__":
    sys.exit(main())
