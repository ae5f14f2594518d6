"""CPU restatement of the reference decode + non_max_suppression (TEST INFRASTRUCTURE).

See oracle/__init__.py for the rules.  Every function cites the reference lines it follows
(paths relative to /root/reference/ultralytics).  Arithmetic is torch CPU fp32 in the same
operation order as the reference so results are `torch.equal` to it (checked by
tests/test_oracle.py against golden vectors written by the live reference).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import build as _build

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = _build.build()
        lib = ctypes.CDLL(path)
        lib.oracle_nms_greedy.restype = ctypes.c_int64
        lib.oracle_nms_greedy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                          ctypes.c_int64, ctypes.c_void_p]
        _LIB = lib
    return _LIB


# ----------------------------------------------------------------------------------------------
# suppression (torchvision.ops.nms CPU semantics; third party, see nms_greedy.c)
# ----------------------------------------------------------------------------------------------
def nms_ref(boxes: torch.Tensor, scores: torch.Tensor, iou_thres: float, max_keep: int = 0) -> torch.Tensor:
    """Greedy NMS, kept indices (int64) in descending-score order.  C restatement (nms_greedy.c)."""
    b = np.ascontiguousarray(boxes.detach().cpu().to(torch.float32).numpy())
    s = np.ascontiguousarray(scores.detach().cpu().to(torch.float32).numpy())
    n = int(s.shape[0])
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = _lib().oracle_nms_greedy(b.ctypes.data, s.ctypes.data, n, float(iou_thres), int(max_keep), keep.ctypes.data)
    if k < 0:
        raise MemoryError("oracle_nms_greedy")
    return torch.from_numpy(keep[:k].copy())


def nms_torchvision(boxes: torch.Tensor, scores: torch.Tensor, iou_thres: float) -> torch.Tensor:
    """The third-party kernel the reference actually calls (utils/ops.py:296), CPU build."""
    import torchvision

    return torchvision.ops.nms(boxes.cpu().float(), scores.cpu().float(), iou_thres)


# ----------------------------------------------------------------------------------------------
# decode
# ----------------------------------------------------------------------------------------------
def make_anchors_ref(shapes: Sequence[Tuple[int, int]], strides: Sequence[float], offset: float = 0.5, device="cpu"):
    """utils/tal.py:366-378 — anchor centres (A,2) as (x,y) and stride column (A,1), levels concatenated."""
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sx = torch.arange(end=w, dtype=torch.float32, device=device) + offset
        sy = torch.arange(end=h, dtype=torch.float32, device=device) + offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=torch.float32, device=device))
    return torch.cat(pts), torch.cat(st)


def dfl_ref(box: torch.Tensor, reg_max: int = 16) -> torch.Tensor:
    """nn/modules/block.py:77-80 — (B, 4*reg_max, A) → (B, 4, A): softmax over the bins of each side,
    then the fixed 1x1 conv with weights arange(reg_max) (block.py:72-74) = expectation."""
    b, _, a = box.shape
    p = box.view(b, 4, reg_max, a).transpose(2, 1).softmax(1)  # (B, reg_max, 4, A)
    w = torch.arange(reg_max, dtype=torch.float32, device=box.device).view(1, reg_max, 1, 1)
    return torch.nn.functional.conv2d(p, w).view(b, 4, a)


def dist2bbox_ref(distance: torch.Tensor, anchor_points: torch.Tensor) -> torch.Tensor:
    """utils/tal.py:381-390 with xywh=True, dim=1."""
    lt, rb = distance.chunk(2, 1)
    x1y1 = anchor_points - lt
    x2y2 = anchor_points + rb
    c_xy = (x1y1 + x2y2) / 2
    wh = x2y2 - x1y1
    return torch.cat((c_xy, wh), 1)


def decode_ref(levels: Sequence[torch.Tensor], strides: Sequence[float], nc: int, reg_max: int = 16,
               embed_dim: int = 0, state_classes: int = 0, device="cpu") -> torch.Tensor:
    """nn/modules/head.py:100-131 (Detect._inference, embed_dim=state_classes=0) and
    :214-249 (JDE._inference).  levels[l]: (B, no, H_l, W_l) fp32, no = 4*reg_max+nc+embed_dim+state_classes.
    Returns y (B, 4+nc+embed_dim+state_classes, A).  `device`: where the torch ops run — "cpu" for every parity check;
    bench.py's `reference_gpu` leg passes a CUDA device to time the same op sequence as a `device=0` user runs it."""
    levels = [x.detach().to(device).float() for x in levels]
    bsz = levels[0].shape[0]
    no = 4 * reg_max + nc + embed_dim + state_classes
    x_cat = torch.cat([xi.reshape(bsz, no, -1) for xi in levels], 2)  # head.py:104 / :218
    anchors, stride_t = (t.transpose(0, 1) for t in
                         make_anchors_ref([tuple(x.shape[2:]) for x in levels], strides, device=device))  # head.py:106 / :220
    parts = x_cat.split([4 * reg_max, nc] + ([embed_dim] if embed_dim else []) + ([state_classes] if state_classes else []), 1)
    box, cls = parts[0], parts[1]
    dbox = dist2bbox_ref(dfl_ref(box, reg_max), anchors.unsqueeze(0)) * stride_t  # head.py:129 / :245
    out = [dbox, cls.sigmoid()]
    if embed_dim:
        out.append(parts[2])  # raw embedding, head.py:247
    if state_classes:
        out.append(parts[-1].sigmoid())  # head.py:247
    return torch.cat(out, 1)


# ----------------------------------------------------------------------------------------------
# non_max_suppression
# ----------------------------------------------------------------------------------------------
def xywh2xyxy_ref(x: torch.Tensor) -> torch.Tensor:
    """utils/ops.py:416-433."""
    y = torch.empty_like(x)
    xy = x[..., :2]
    wh = x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def non_max_suppression_ref(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                            classes: Optional[Sequence[int]] = None, agnostic: bool = False,
                            multi_label: bool = False, labels=(), max_det: int = 300, nc: int = 0,
                            max_nms: int = 30000, max_wh=7680, nms_fn=None, stable_topk: bool = True,
                            return_index: bool = False, device="cpu"):
    """utils/ops.py:167-316 (non-rotated, non end-to-end branch).

    Deviations, all deliberate and documented in SURVEY.md Appendix B:
      * the caller's tensor is not mutated (ops.py:243-244 writes xyxy back in place);
      * the wall-clock time limit (ops.py:312-314) is not applied;
      * `stable_topk=True` makes the max_nms cut stable (ops.py:286 uses an unstable argsort whose
        tie order is torch-version defined, SURVEY §7 hard part 2) — pass False to use the literal call.
    `return_index=True` additionally returns, per image, an int64 (n_i, 2) tensor of (anchor, class)
    identifying each output row in the input — the "kept-index set" the GPU path is checked against.
    `device`: where the torch ops run ("cpu" for every parity check; a CUDA device + a CUDA `nms_fn` only in bench.py's
    `reference_gpu` leg, which times this op sequence the way a `device=0` user of the reference runs it).
    """
    if nms_fn is None:  # early stop after max_det keeps == truncating the full result (ops.py:297)
        nms_fn = lambda b_, s_, t_: nms_ref(b_, s_, t_, max_keep=max_det)
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1  # ops.py:217-218
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]  # ops.py:219-220
    prediction = prediction.detach().to(device).float()
    dev = prediction.device
    cls_t = torch.tensor(list(classes), device=dev) if classes is not None else None
    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)  # ops.py:231
    nm = prediction.shape[1] - nc - 4
    mi = 4 + nc
    xc = prediction[:, 4:mi].amax(1) > conf_thres  # ops.py:234
    multi_label = multi_label and nc > 1  # ops.py:239
    prediction = prediction.transpose(-1, -2)  # (B, A, C)
    prediction = torch.cat((xywh2xyxy_ref(prediction[..., :4]), prediction[..., 4:]), dim=-1)  # ops.py:246
    output = [torch.zeros((0, 6 + nm), device=dev)] * bs
    index = [torch.zeros((0, 2), dtype=torch.int64, device=dev)] * bs
    for xi, x in enumerate(prediction):
        sel = xc[xi].nonzero().squeeze(1)
        x = x[sel]  # ops.py:253
        src_a = sel
        if labels and len(labels[xi]):  # ops.py:256-261 (save_hybrid)
            lb = labels[xi].to(dev)
            v = torch.zeros((len(lb), nc + nm + 4), device=dev)
            v[:, :4] = xywh2xyxy_ref(lb[:, 1:5])
            v[range(len(lb)), lb[:, 0].long() + 4] = 1.0
            x = torch.cat((x, v), 0)
            src_a = torch.cat((src_a, -1 - torch.arange(len(lb), device=dev)))
        if not x.shape[0]:
            continue
        box, cls, mask = x.split((4, nc, nm), 1)  # ops.py:268
        if multi_label:
            i, j = torch.where(cls > conf_thres)  # ops.py:271
            x = torch.cat((box[i], x[i, 4 + j, None], j[:, None].float(), mask[i]), 1)
            src = torch.stack((src_a[i], j), 1)
        else:
            conf, j = cls.max(1, keepdim=True)  # ops.py:274
            keep = conf.view(-1) > conf_thres
            x = torch.cat((box, conf, j.float(), mask), 1)[keep]
            src = torch.stack((src_a, j.view(-1)), 1)[keep]
        if cls_t is not None:
            keep = (x[:, 5:6] == cls_t).any(1)  # ops.py:279
            x, src = x[keep], src[keep]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:  # ops.py:285-286
            order = x[:, 4].argsort(descending=True, stable=True) if stable_topk else x[:, 4].argsort(descending=True)
            x, src = x[order[:max_nms]], src[order[:max_nms]]
        c = x[:, 5:6] * (0 if agnostic else max_wh)  # ops.py:289
        scores = x[:, 4]
        boxes = x[:, :4] + c  # ops.py:295
        i = nms_fn(boxes, scores, iou_thres)  # ops.py:296
        i = i[:max_det]  # ops.py:297
        output[xi] = x[i]  # ops.py:311
        index[xi] = src[i]
    return (output, index) if return_index else output


def scale_boxes_ref(img1_shape, boxes: torch.Tensor, img0_shape) -> torch.Tensor:
    """utils/ops.py:92-127 (ratio_pad=None, padding=True, xyxy) + clip_boxes :319-338, on a copy."""
    boxes = boxes.clone()
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    boxes[..., 0] -= pad[0]
    boxes[..., 1] -= pad[1]
    boxes[..., 2] -= pad[0]
    boxes[..., 3] -= pad[1]
    boxes[..., :4] /= gain
    boxes[..., 0] = boxes[..., 0].clamp(0, img0_shape[1])
    boxes[..., 1] = boxes[..., 1].clamp(0, img0_shape[0])
    boxes[..., 2] = boxes[..., 2].clamp(0, img0_shape[1])
    boxes[..., 3] = boxes[..., 3].clamp(0, img0_shape[0])
    return boxes


def jde_results_ref(rows: torch.Tensor, img1_shape, img0_shape, embed_dim: int, state_classes: int):
    """The per-image tail of JDEPredictor.postprocess (models/yolo/jde/predict.py:48-77) on one image's NMS rows
    `(n, 6 + embed_dim + state_classes)`: boxes scaled to the original image (:49), then — with a state head and at
    least one row — `boxes (n, 7)` = x1,y1,x2,y2,state_id,conf,cls with state_id = argmax over the state columns
    (:61-64); otherwise the plain 6 columns (:66-69, :73-75).  Returns `(boxes, embeds)` as `Results` receives them."""
    pred = rows.clone()
    pred[:, :4] = scale_boxes_ref(img1_shape, pred[:, :4], img0_shape)
    if not state_classes:
        return pred[:, :6], pred[:, 6:]
    box6, emb, states = pred[:, :6], pred[:, 6:6 + embed_dim], pred[:, 6 + embed_dim:6 + embed_dim + state_classes]
    if len(states) == 0:
        return box6, emb
    ids = states.argmax(dim=1).unsqueeze(1)
    return torch.cat([box6[:, :4], ids, box6[:, 4:]], dim=1), emb


def box_iou_ref(box1: torch.Tensor, box2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """utils/metrics.py:55-75."""
    (a1, a2), (b1, b2) = box1.float().unsqueeze(1).chunk(2, 2), box2.float().unsqueeze(0).chunk(2, 2)
    inter = (torch.min(a2, b2) - torch.max(a1, b1)).clamp_(0).prod(2)
    return inter / ((a2 - a1).prod(2) + (b2 - b1).prod(2) - inter + eps)


def match_predictions_ref(pred_classes: torch.Tensor, true_classes: torch.Tensor, iou: torch.Tensor, iouv: torch.Tensor,
                          tag_thr: Optional[float] = None):
    """engine/validator.py:222-262 (use_scipy=False), numpy like the reference; with `tag_thr` also returns the label
    index matched at that threshold per detection or -1 (the pairs jde/val.py:731-735 iterates over)."""
    correct = np.zeros((pred_classes.shape[0], iouv.shape[0])).astype(bool)
    matched = np.full((pred_classes.shape[0],), -1, dtype=np.int32)
    correct_class = true_classes[:, None] == pred_classes
    iou = (iou * correct_class).cpu().numpy()
    for i, threshold in enumerate(iouv.cpu().tolist()):
        matches = np.nonzero(iou >= threshold)
        matches = np.array(matches).T
        if matches.shape[0]:
            if matches.shape[0] > 1:
                matches = matches[iou[matches[:, 0], matches[:, 1]].argsort()[::-1]]
                matches = matches[np.unique(matches[:, 1], return_index=True)[1]]
                matches = matches[np.unique(matches[:, 0], return_index=True)[1]]
            correct[matches[:, 1].astype(int), i] = True
            if tag_thr is not None and threshold == tag_thr:
                for gt_idx, pred_idx in matches:
                    matched[pred_idx] = gt_idx
    c = torch.tensor(correct, dtype=torch.bool)
    return (c, torch.from_numpy(matched)) if tag_thr is not None else c


def state_head_ref(emb: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """nn/modules/head.py:189-190 (`state_predictor` = Linear, ReLU, Dropout — identity in eval —, Linear), applied per
    anchor at :198-204, sigmoid at :247.  `emb (..., E)` -> `(..., S)` probabilities, fp32 on the CPU."""
    h = torch.relu(torch.nn.functional.linear(emb.float(), w1.float(), b1.float()))
    return torch.nn.functional.linear(h, w2.float(), b2.float()).sigmoid()
