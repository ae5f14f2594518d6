"""The UNMODIFIED reference's `YOLO(...).predict()` (JDE task) with and without `sarpost.patch()`, on a GPU.

    python tools/dropin_predict_check.py [--device 0] [--imgsz 640] [--conf 0.001]

Needs an importable reference: baseline/_ref (pip-installed copy, travels to the GPU box) or /root/reference (build
container).  What it does:
  * builds `YOLO("cfg/models/v13/yolov13-JDE.yaml", task="jde")` (random init) from the reference's own code;
  * a forward pre-hook feeds the JDE head seeded N(0,1) feature maps (the untrained trunk kills the signal)
    and the class bias is shifted so scores spread over (0.05, 0.9); a forward hook captures the raw per-level logits the
    head returns next to `y`;
  * runs `model.predict([ndarray, ...])` eight times (the last two modes also with `predictor=True`: JDEPredictor.postprocess
    replaced by one fused call whose gather kernel writes the 7-column boxes and the embeddings): reference as is (torch CUDA ops + torchvision CUDA NMS), then under
    `patch()` (API-exact decode + NMS kernels), `patch(fused=True)` (LazyPrediction -> fused kernels),
    `patch(fused=True, split=True)` with a channels_last and with an NCHW embedding branch (no torch.cat in the head) and
    `patch(fused=True, defer_state=True)` (state MLP on the kept rows only);
  * runs `BaseValidator.match_predictions` / `JDEValidator.match_predictions` of the reference with and without
    `patch(match=True)` on the same CUDA inputs;
  * checks every patched run against the CPU oracle evaluated on the logits captured IN THAT RUN (decode_ref + NMS +
    scale_boxes + state argmax + 7-column re-pack = models/yolo/jde/predict.py:29-78): rows matched in order, decode
    tolerance, mismatch budget 1e-4 of the detections; the deferred-state run is checked against the fused run;
  * reports the differences against the reference's own GPU result for information (torchvision's CUDA kernel is not
    bit-identical to its CPU kernel, SURVEY §7 hard part 1).
Prints one JSON object; exit code 0 = all checks passed.
"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="0")
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--conf", type=float, default=0.05)
    ap.add_argument("--iou", type=float, default=0.7)
    args = ap.parse_args()

    import numpy as np
    import torch

    from oracle import postprocess_ref as R
    from oracle import ref_shim

    if not ref_shim.available():
        print(json.dumps({"skipped": "no reference install (baseline/_ref) on this machine"}))
        return 0
    ref_shim._stub_third_party()
    os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp(prefix="sarpost_yolo_cfg_"))
    os.environ.setdefault("YOLO_OFFLINE", "1")
    sys.path.insert(0, ref_shim.REF_ROOT)
    from ultralytics import YOLO  # the reference's full stack

    import sarpost

    dev = f"cuda:{args.device}" if args.device != "cpu" else "cpu"
    cfg = os.path.join(ref_shim.REF_ROOT, "ultralytics", "cfg", "models", "v13", "yolov13-JDE.yaml")
    torch.manual_seed(0)
    model = YOLO(cfg, task="jde")
    model.model.person_states = {}
    head = model.model.model[-1]
    with torch.no_grad():
        for seq in head.cv3:
            seq[-1].bias.add_(6.5)  # bias_init puts scores at ~1e-3 (head.py:133-140); spread them so conf/NMS have work to do
            seq[-1].weight.mul_(8.0)
        for prm in head.state_predictor.parameters():
            prm.mul_(4.0)
    captured = {}

    def feed(mod, inp):  # same seeded features on every call with the same shapes
        g = torch.Generator().manual_seed(1234)
        return ([torch.randn(f.shape, generator=g).to(device=f.device, dtype=f.dtype) for f in inp[0]],)

    def grab(mod, inp, out):
        if isinstance(out, tuple) and isinstance(out[1], list):
            captured["split"] = bool(out[1]) and isinstance(out[1][0], tuple)
            captured["levels"] = [x.detach().float().cpu().clone() for x in sarpost.cat_levels(out[1])]

    head.register_forward_pre_hook(feed)
    head.register_forward_hook(grab)
    rng = np.random.default_rng(0)
    imgs = [(rng.random((480, 640, 3)) * 255).astype(np.uint8), (rng.random((360, 640, 3)) * 255).astype(np.uint8)]
    kw = dict(device=args.device, conf=args.conf, iou=args.iou, imgsz=args.imgsz, verbose=False, half=False)

    def run():
        res = model.predict(imgs, **kw)
        return [(r.boxes.data.detach().float().cpu(), r.embeds.data.detach().float().cpu()) for r in res], captured.pop("levels")

    def expected(levels):
        strides = [float(s) for s in head.stride]
        nc, ed, sc = int(head.nc), int(head.embed_dim), int(head.state_classes)
        full = levels
        if levels[0].shape[1] == 64 + nc + ed:  # deferred-state run: the head returned levels without state channels
            return None
        y = R.decode_ref(full, strides, nc, 16, ed, sc)
        rows = R.non_max_suppression_ref(y, conf_thres=args.conf, iou_thres=args.iou, nc=nc, max_det=300)
        img1 = (int(levels[0].shape[2] * strides[0]), int(levels[0].shape[3] * strides[0]))
        out = []
        for r, im in zip(rows, imgs):
            r = r.clone()
            r[:, :4] = R.scale_boxes_ref(img1, r[:, :4], im.shape)
            if r.shape[0]:
                sid = r[:, 6 + ed:6 + ed + sc].argmax(1, keepdim=True).float()
                boxes = torch.cat((r[:, :4], sid, r[:, 4:6]), 1)
            else:
                boxes = r[:, :6]
            out.append((boxes, r[:, 6:6 + ed]))
        return out

    def diff(a, b):
        """mismatching detections between two result lists (rows compared in order), and the total."""
        bad = tot = 0
        for (ba, ea), (bb, eb) in zip(a, b):
            n = max(ba.shape[0], bb.shape[0])
            tot += n
            m = min(ba.shape[0], bb.shape[0])
            bad += n - m
            if m and ba.shape[1] == bb.shape[1]:
                ok = torch.isclose(ba[:m], bb[:m], rtol=1e-5, atol=1e-5 * 32 * 4).all(1) & torch.isclose(ea[:m], eb[:m], rtol=1e-5, atol=1e-6).all(1)
                bad += int((~ok).sum())
            elif m:
                bad += m
        return bad, tot

    report = {"reference": ref_shim.source(), "device": dev, "runs": {}}
    ok = True
    ref_res, ref_levels = run()
    exp = expected(ref_levels)
    b, t = diff(ref_res, exp)
    report["runs"]["reference_unpatched"] = {"detections": [int(x[0].shape[0]) for x in ref_res], "vs_cpu_oracle": {"mismatch": b, "of": t},
                                             "box_columns": int(ref_res[0][0].shape[1])}
    fused_res = None
    for name, mode in (("patch", {}), ("patch_fused", dict(fused=True)), ("patch_fused_split", dict(fused=True, split=True)),
                       ("patch_fused_split_nchw_emb", dict(fused=True, split=True, emb_channels_last=False)),
                       ("patch_fused_predictor", dict(fused=True, split=True, predictor=True)),
                       ("patch_fused_defer_state", dict(fused=True, defer_state=True)),
                       ("patch_fused_defer_state_predictor", dict(fused=True, defer_state=True, predictor=True))):
        sarpost.patch(**mode)
        try:
            res, levels = run()
            was_split = captured.pop("split", False)
            launches = sarpost.ops.last_launch_count()
        finally:
            sarpost.unpatch()
        exp = expected(levels)
        entry = {"detections": [int(x[0].shape[0]) for x in res], "box_columns": int(res[0][0].shape[1]) if res else None,
                 "head_returned_split_levels": bool(was_split), "launches_of_last_library_call": int(launches)}
        if exp is not None:
            b, t = diff(res, exp)
            entry["vs_cpu_oracle"] = {"mismatch": b, "of": t}
            good = t > 0 and b <= 1e-4 * t
        else:  # deferred state: every column but the state-derived id must equal the fused run; ids equal unless the top-2 states tie
            b = tot = 0
            for (ba, ea), (bb, eb) in zip(res, fused_res):
                tot += max(ba.shape[0], bb.shape[0])
                same = (ba.shape == bb.shape and torch.equal(ba[:, [0, 1, 2, 3, 5, 6]], bb[:, [0, 1, 2, 3, 5, 6]])
                        and torch.allclose(ea, eb, rtol=1e-4, atol=1e-5))
                b += 0 if same else max(ba.shape[0], bb.shape[0])
                if same:
                    b += int((ba[:, 4] != bb[:, 4]).sum())
            entry["vs_fused_run"] = {"mismatch": b, "of": tot}
            entry["levels_channels"] = int(levels[0].shape[1])
            good = tot > 0 and b <= 1e-2 * tot
        b2, t2 = diff(res, ref_res)
        entry["vs_reference_gpu_run"] = {"mismatch": b2, "of": t2}
        entry["ok"] = bool(good)
        ok = ok and good
        if name == "patch_fused_split":  # same forward (channels_last cv4 branch) as the deferred-state run, state head on every anchor
            fused_res = res
        report["runs"][name] = entry
    # ---- validator side: BaseValidator.match_predictions / JDEValidator.match_predictions under patch(match=True) ----
    try:
        from ultralytics.engine.validator import BaseValidator
        from ultralytics.models.yolo.jde.val import JDEValidator
        from ultralytics.utils.metrics import box_iou

        g = torch.Generator().manual_seed(5)
        n_gt, n_det = 37, 300
        gxy = torch.rand(n_gt, 2, generator=g) * 500
        gt = torch.cat((gxy, gxy + 20 + torch.rand(n_gt, 2, generator=g) * 100), 1)
        src = torch.randint(0, n_gt, (n_det,), generator=g)
        det = gt[src] + torch.randn(n_det, 4, generator=g) * 8
        gcls = torch.randint(0, 3, (n_gt,), generator=g).float()
        dcls = torch.where(torch.rand(n_det, generator=g) < 0.8, gcls[src], torch.randint(0, 3, (n_det,), generator=g).float())
        tags = torch.randint(1, 90, (n_gt,), generator=g).float()
        d = torch.device(dev)
        bv, jv = BaseValidator.__new__(BaseValidator), JDEValidator.__new__(JDEValidator)
        bv.iouv = jv.iouv = torch.linspace(0.5, 0.95, 10).to(d)
        jv.state_iou = 0.5
        iou = box_iou(gt.to(d), det.to(d))
        want = bv.match_predictions(dcls.to(d), gcls.to(d), iou)
        want_j = jv.match_predictions(dcls.to(d), gcls.to(d), tags.to(d), iou)
        sarpost.patch(match=True)
        try:
            got = bv.match_predictions(dcls.to(d), gcls.to(d), iou)
            got_j = jv.match_predictions(dcls.to(d), gcls.to(d), tags.to(d), iou)
            patched = BaseValidator.match_predictions is not sarpost.plugin._SAVED["match"] and "jde_match" in sarpost.plugin._SAVED
        finally:
            sarpost.unpatch()
        vm_ok = (patched and torch.equal(got, want) and torch.equal(got_j[0], want_j[0]) and torch.equal(got_j[1].long(), want_j[1].long())
                 and got.device == want.device)
        report["validator_match"] = {"ok": bool(vm_ok), "true_positives_at_0.5": int(want[:, 0].sum()), "tags_assigned": int((want_j[1] != 0).sum())}
        ok = ok and (vm_ok or dev == "cpu")
    except Exception as e:  # noqa: BLE001
        report["validator_match"] = {"ok": False, "error": f"{type(e).__name__}: {e}"}
        ok = False
    report["ok"] = bool(ok)
    print(json.dumps(report))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
