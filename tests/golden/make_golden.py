"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

    python tests/golden/make_golden.py

Inputs are never stored: every case is regenerated from its seed with `sarpost.synth` (torch CPU
generator, deterministic for a given torch build); the files hold the reference's OUTPUTS —
`Detect/JDE._inference` (ultralytics/nn/modules/head.py:100-131, :214-249) and
`ops.non_max_suppression` (ultralytics/utils/ops.py:167-316) executed from /root/reference through
oracle/ref_shim.py — plus the case parameters.  tests/test_oracle.py pins the oracle to them on CPU,
tests/test_golden_gpu.py checks the CUDA path against them on the B200.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import sarpost  # noqa: E402
from oracle import ref_shim  # noqa: E402
from sarpost import synth  # noqa: E402

# name -> decoded-prediction NMS cases (input = synth.decoded_prediction(**gen))
NMS_CASES = {
    "nms_predict_nc1": dict(gen=dict(batch=2, anchors=8400, nc=1, nm=0, seed=11), kw=dict(conf_thres=0.25, iou_thres=0.7, nc=1)),
    "nms_predict_nc6": dict(gen=dict(batch=2, anchors=8400, nc=6, nm=0, seed=12), kw=dict(conf_thres=0.25, iou_thres=0.7, nc=6)),
    "nms_val_multilabel_nc6": dict(gen=dict(batch=2, anchors=8400, nc=6, nm=0, seed=13, score_pow=2.0),
                                   kw=dict(conf_thres=0.001, iou_thres=0.7, nc=6, multi_label=True)),
    "nms_agnostic_thr06_classes": dict(gen=dict(batch=2, anchors=5000, nc=5, nm=3, seed=14),
                                       kw=dict(conf_thres=0.05, iou_thres=0.6, nc=5, agnostic=True, classes=[0, 3])),
    "nms_clustered": dict(gen=dict(batch=2, anchors=12000, nc=3, nm=0, seed=15, clustered=True, score_pow=1.0),
                          kw=dict(conf_thres=0.05, iou_thres=0.7, nc=3)),
    "nms_topk_cut": dict(gen=dict(batch=1, anchors=70000, nc=1, nm=0, seed=16, score_pow=2.0),
                         kw=dict(conf_thres=0.001, iou_thres=0.7, nc=1)),
    "nms_jde_extras": dict(gen=dict(batch=1, anchors=8400, nc=1, nm=262, seed=17), kw=dict(conf_thres=0.25, iou_thres=0.7, nc=1, max_det=100)),
}

# raw-head cases (input = synth.head_outputs(...)); decode output stored subsampled every `sub` anchors
HEAD_CASES = {
    "head_jde_640": dict(imgsz=640, strides=[8, 16, 32], nc=1, ed=256, sc=6, batch=1, seed=21, sub=7,
                         kw=dict(conf_thres=0.25, iou_thres=0.7, max_det=100)),
    "head_detect6_640": dict(imgsz=640, strides=[8, 16, 32], nc=6, ed=0, sc=0, batch=2, seed=22, sub=5,
                             kw=dict(conf_thres=0.25, iou_thres=0.7)),
    "head_p2_jde_val": dict(imgsz=320, strides=[4, 8, 16, 32], nc=1, ed=16, sc=0, batch=2, seed=23, sub=3,
                            kw=dict(conf_thres=0.001, iou_thres=0.7)),
    "head_rect_odd": dict(imgsz=[88, 120], strides=[8, 16, 32], nc=2, ed=8, sc=6, batch=2, seed=24, sub=1,
                          kw=dict(conf_thres=0.1, iou_thres=0.7, multi_label=True)),
    "head_p2_1280_val": dict(imgsz=1280, strides=[4, 8, 16, 32], nc=1, ed=0, sc=0, batch=1, seed=25, sub=97,
                             kw=dict(conf_thres=0.001, iou_thres=0.7)),
}


def pack_rows(rows):
    n = np.array([r.shape[0] for r in rows], dtype=np.int64)
    cat = torch.cat(rows, 0).numpy() if len(rows) else np.zeros((0, 6), np.float32)
    return n, cat


def main():
    ref_shim.load()
    torch.set_num_threads(4)
    for name, case in NMS_CASES.items():
        y = synth.decoded_prediction(**case["gen"])
        rows = ref_shim.ref_nms(y, **case["kw"])
        n, cat = pack_rows(rows)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(case), counts=n, rows=cat)
        print(name, n.tolist())
    for name, c in HEAD_CASES.items():
        shapes = synth.level_shapes(c["imgsz"] if isinstance(c["imgsz"], int) else tuple(c["imgsz"]), c["strides"])
        levels = synth.head_outputs(c["batch"], shapes, c["nc"], c["ed"], c["sc"], seed=c["seed"])
        y = ref_shim.ref_decode(levels, c["strides"], c["nc"], c["ed"], c["sc"])
        rows = ref_shim.ref_nms(y, nc=c["nc"], **c["kw"])
        n, cat = pack_rows(rows)
        ysub = y[:, :, :: c["sub"]].contiguous().numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(c), counts=n, rows=cat, y_sub=ysub,
                            y_shape=np.array(y.shape))
        print(name, tuple(y.shape), n.tolist())
    # known-answer vectors probed on torchvision's CPU kernel through the reference
    ka = []
    for boxes, scores, thr in [
        ([[0, 0, 10, 10]] * 3, [0.5, 0.5, 0.5], 0.5),
        ([[0, 0, 6, 1], [0, 0, 3.6000001, 1]], [0.9, 0.8], 0.6),
        ([[0, 0, 6, 1], [0, 0, 3.6000001, 1]], [0.9, 0.8], 0.7),
        ([[5, 5, 5, 5], [5, 5, 5, 5]], [0.9, 0.8], 0.5),
    ]:
        b = torch.tensor(boxes, dtype=torch.float32)
        xywh = torch.stack(((b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]), 1)
        y = torch.cat((xywh, torch.tensor(scores)[:, None]), 1).t()[None].contiguous()
        rows = ref_shim.ref_nms(y, conf_thres=0.1, iou_thres=thr)
        ka.append(dict(boxes=boxes, scores=scores, iou_thres=thr, rows=rows[0].tolist()))
    json.dump(ka, open(os.path.join(HERE, "known_answers.json"), "w"), indent=1)
    print("known answers", [len(k["rows"]) for k in ka])


if __name__ == "__main__":
    main()
