"""Host-side mirror of the reference's post-processing interface, backed by libsarpost.so.

Same names, argument meaning and error behaviour as the reference so it drops in:
  * non_max_suppression   ultralytics/utils/ops.py:167-316
  * decode                ultralytics/nn/modules/head.py:100-131 (Detect._inference), :214-249 (JDE._inference)
  * postprocess_fused     decode + non_max_suppression in one pass from the raw level logits
  * merge_tiles           cross-tile merge for sliced inference (no reference counterpart)
  * postprocess_host      the fused path for HOST tensors (H2D / D2H included) — the `e2e` entry

PyTorch is used for device memory, streams and the final list-of-views only.  CUDA tensors are
required: there is no CPU fallback (`RuntimeError`).
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import Head, NmsParams, lib

__all__ = ["HeadSpec", "FusedPlan", "StateMLP", "state_head", "non_max_suppression", "decode", "postprocess_fused", "split_levels", "cat_levels", "merge_tiles", "gather_extras", "match_predictions", "match_from_iou", "postprocess_host",
           "HostContext", "Pipeline", "last_launch_count", "stage_timing", "stage_times"]


@dataclass(frozen=True)
class HeadSpec:
    """The attributes read from `model.model[-1]` (head.py:37-41, :180-188)."""
    nc: int
    strides: Tuple[float, ...]
    reg_max: int = 16
    embed_dim: int = 0        # JDE raw embedding channels (head.py:180)
    state_classes: int = 0    # JDE state logits, sigmoid on output (head.py:181,247)

    @property
    def no(self) -> int:
        return 4 * self.reg_max + self.nc + self.embed_dim + self.state_classes

    @property
    def nm(self) -> int:
        return self.embed_dim + self.state_classes

    @staticmethod
    def from_module(m) -> "HeadSpec":
        """Build from a reference Detect/JDE module instance."""
        return HeadSpec(nc=int(m.nc), strides=tuple(float(s) for s in m.stride), reg_max=int(m.reg_max),
                        embed_dim=int(getattr(m, "embed_dim", 0) or 0),
                        state_classes=int(getattr(m, "state_classes", 0) or 0))


@dataclass(frozen=True)
class StateMLP:
    """Weights of `JDE.state_predictor` (head.py:189-190): Linear(E, H) -> ReLU -> Dropout (identity in eval) ->
    Linear(H, S), as contiguous fp32 CUDA tensors in `nn.Linear` layout: w1 (H, E), b1 (H), w2 (S, H), b2 (S)."""
    w1: torch.Tensor
    b1: torch.Tensor
    w2: torch.Tensor
    b2: torch.Tensor

    def __post_init__(self):
        h, e = self.w1.shape
        s, h2 = self.w2.shape
        if h2 != h or tuple(self.b1.shape) != (h,) or tuple(self.b2.shape) != (s,):
            raise ValueError(f"sarpost: state MLP shapes {tuple(self.w1.shape)}, {tuple(self.b1.shape)}, "
                             f"{tuple(self.w2.shape)}, {tuple(self.b2.shape)} are not Linear(E,H) -> Linear(H,S)")
        for t in (self.w1, self.b1, self.w2, self.b2):
            _require_cuda(t, "state MLP weight")
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("sarpost: state MLP weights must be contiguous float32")

    @property
    def embed_dim(self) -> int:
        return int(self.w1.shape[1])

    @property
    def hidden(self) -> int:
        return int(self.w1.shape[0])

    @property
    def n_state(self) -> int:
        return int(self.w2.shape[0])

    @staticmethod
    def from_tensors(w1, b1, w2, b2, device=None) -> "StateMLP":
        dev = device if device is not None else w1.device
        return StateMLP(*(t.detach().to(device=dev, dtype=torch.float32).contiguous() for t in (w1, b1, w2, b2)))

    @staticmethod
    def from_module(m, device=None) -> "StateMLP":
        """From a reference JDE module (or its `state_predictor`): the two `nn.Linear` layers of the Sequential."""
        seq = getattr(m, "state_predictor", m)
        lin = [layer for layer in seq if isinstance(layer, torch.nn.Linear)]
        if len(lin) != 2:
            raise ValueError(f"sarpost: expected 2 Linear layers in state_predictor, found {len(lin)}")
        return StateMLP.from_tensors(lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, device)


def state_head(rows: torch.Tensor, counts: torch.Tensor, mlp: StateMLP, emb_col: int = 6, state_col: Optional[int] = None) -> torch.Tensor:
    """Deferred JDE state head (head.py:189-190,198-204,247) on padded rows, IN PLACE:
    `rows[b, r, state_col:state_col+S] = sigmoid(state_predictor(rows[b, r, emb_col:emb_col+E]))` for `r < counts[b]`.
    `rows (B, max_det, row_len)` fp32 contiguous CUDA, `counts (B,)` int32.  Returns `rows`."""
    _require_cuda(rows, "rows")
    if rows.dtype != torch.float32 or not rows.is_contiguous() or rows.dim() != 3:
        raise ValueError("sarpost: rows must be a contiguous float32 (B, max_det, row_len) tensor")
    dev = rows.device
    counts = counts.to(device=dev, dtype=torch.int32).contiguous()
    bsz, max_det, row_len = (int(v) for v in rows.shape)
    if counts.shape != (bsz,):
        raise ValueError("sarpost: counts must have one entry per image")
    if mlp.w1.device != dev:
        raise RuntimeError(f"sarpost: state MLP weights are on {mlp.w1.device}, rows on {dev}")
    sc = emb_col + mlp.embed_dim if state_col is None else int(state_col)
    with torch.cuda.device(dev):
        _lib.check(lib.sarpost_state_head(rows.data_ptr(), counts.data_ptr(), bsz, max_det, row_len, int(emb_col), mlp.embed_dim,
                                          sc, mlp.n_state, mlp.hidden, mlp.w1.data_ptr(), mlp.b1.data_ptr(), mlp.w2.data_ptr(),
                                          mlp.b2.data_ptr(), _stream_ptr(dev)))
    return rows


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"sarpost: {what} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"sarpost: {what} is on {t.device}; only CUDA tensors are supported "
                           "(no CPU fallback — the reference path is the CPU implementation)")


def _stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def scale_params(img1_shape, img0_shapes, device) -> torch.Tensor:
    """Per-image `(pad_x, pad_y, gain, w0, h0)` of `ops.scale_boxes(img1_shape, boxes, img0_shape)` with
    `ratio_pad=None, padding=True` (utils/ops.py:110-116) — the host part of that function, evaluated exactly
    like the reference (Python floats, `round`); the arithmetic on the boxes runs in the gather kernel."""
    rows = []
    for s0 in img0_shapes:
        gain = min(img1_shape[0] / s0[0], img1_shape[1] / s0[1])  # gain  = old / new
        pad = (round((img1_shape[1] - s0[1] * gain) / 2 - 0.1), round((img1_shape[0] - s0[0] * gain) / 2 - 0.1))
        rows.append([float(pad[0]), float(pad[1]), float(gain), float(s0[1]), float(s0[0])])
    return torch.tensor(rows, dtype=torch.float32).to(device, non_blocking=True)


def _make_params(conf_thres, iou_thres, classes, agnostic, multi_label, max_det, max_nms, max_wh, rescale=None, nms_cluster=0):
    p = NmsParams()
    p.nms_cluster = int(nms_cluster)
    p.rescale = rescale.data_ptr() if rescale is not None else None
    p.conf_thres = float(conf_thres)
    p.iou_thres = float(iou_thres)
    p.agnostic = int(bool(agnostic))
    p.multi_label = int(bool(multi_label))
    p.max_det = int(max_det)
    p.max_nms = int(max_nms)
    p.max_wh = float(max_wh)
    p.workspace_clean = 1  # every device entry point below uses a persistent, prepared _Workspace
    keep = None
    if classes is not None:
        cl = [int(c) for c in (classes.tolist() if isinstance(classes, torch.Tensor) else classes)]
        keep = (C.c_int32 * max(len(cl), 1))(*cl)
        p.classes = C.cast(keep, C.POINTER(C.c_int32))
        p.n_classes = len(cl)
    else:
        p.classes = None
        p.n_classes = 0
    return p, keep


def _is_split(levels) -> bool:
    """Split layout (SARPOST_LAYOUT_SPLIT): every level is a tuple `(box, cls[, emb[, state]])` of branch outputs."""
    return len(levels) > 0 and isinstance(levels[0], (tuple, list))


def _box_of(level) -> torch.Tensor:
    return level[0] if isinstance(level, (tuple, list)) else level


def _emb_channels_last(e: torch.Tensor) -> bool:
    """`e (B, E, H, W)` stored as (B, H, W, E) — what a channels_last convolution writes — and not also plain NCHW."""
    return (not e.is_contiguous()) and e.is_contiguous(memory_format=torch.channels_last)


def _make_head(levels, spec: HeadSpec, host: bool = False, with_extras: bool = True) -> Head:
    if len(levels) != len(spec.strides):
        raise ValueError(f"sarpost: {len(levels)} level tensors but {len(spec.strides)} strides")
    if len(levels) > _lib.MAX_LEVELS:
        raise ValueError(f"sarpost: at most {_lib.MAX_LEVELS} levels")
    split = _is_split(levels)
    first = _box_of(levels[0])
    h = Head()
    h.nl = len(levels)
    h.batch = int(first.shape[0])
    h.no = spec.no
    h.nc = spec.nc
    h.reg_max = spec.reg_max
    h.n_extra_raw = spec.embed_dim if with_extras else 0
    h.n_extra_sigmoid = spec.state_classes if with_extras else 0
    h.dtype = 1 if first.dtype == torch.float16 else 0
    h.layout = 1 if split else 0

    def check(x, i, channels, what, channels_last_ok=False):
        if x.dim() != 4 or x.shape[0] != h.batch or x.shape[1] != channels:
            raise ValueError(f"sarpost: level {i} {what} has shape {tuple(x.shape)}, expected (B={h.batch}, {channels}, H, W)")
        ok_layout = x.is_contiguous() or (channels_last_ok and x.is_contiguous(memory_format=torch.channels_last))
        if x.dtype != first.dtype or x.dtype not in (torch.float32, torch.float16) or not ok_layout:
            raise ValueError("sarpost: level tensors must be contiguous and all float32 or all float16")
        if host == x.is_cuda:
            raise RuntimeError(f"sarpost: level {i} is on {x.device}, expected {'host' if host else 'CUDA'} memory")

    if split:
        if host:
            raise ValueError("sarpost: the host entry takes concatenated level tensors")
        embs = [lv[2] for lv in levels if len(lv) > 2 and lv[2] is not None]
        h.emb_channels_last = int(bool(h.n_extra_raw) and len(embs) == len(levels) and all(_emb_channels_last(e) for e in embs))
    for i, lv in enumerate(levels):
        x = _box_of(lv)
        check(x, i, 4 * spec.reg_max if split else spec.no, "box branch" if split else "tensor")
        h.h[i] = int(x.shape[2])
        h.w[i] = int(x.shape[3])
        h.stride[i] = float(spec.strides[i])
        h.data[i] = x.data_ptr()
        if split:
            hw = tuple(x.shape[2:])
            check(lv[1], i, spec.nc, "class branch")
            h.cls[i] = lv[1].data_ptr()
            if h.n_extra_raw:
                if len(lv) < 3 or lv[2] is None:
                    raise ValueError(f"sarpost: level {i} has no embedding branch but the head has embed_dim {spec.embed_dim}")
                check(lv[2], i, spec.embed_dim, "embedding branch", channels_last_ok=True)
                if bool(h.emb_channels_last) != _emb_channels_last(lv[2]):
                    raise ValueError("sarpost: embedding branches must all be channels_last or all NCHW (see _prep_levels)")
                h.emb[i] = lv[2].data_ptr()
            if h.n_extra_sigmoid:
                if len(lv) < 4 or lv[3] is None:
                    raise ValueError(f"sarpost: level {i} has no state branch but the head has state_classes {spec.state_classes}")
                check(lv[3], i, spec.state_classes, "state branch")
                h.state[i] = lv[3].data_ptr()
            for t in lv[1:]:
                if t is not None and tuple(t.shape[2:]) != hw:
                    raise ValueError(f"sarpost: level {i}: branch tensors disagree on H, W")
    return h


def _prep_levels(levels) -> list:
    """Device / dtype / contiguity normalisation.  fp16 (`half=True` pipelines) is read as it is; anything else but fp32 is
    upcast.  Split levels: every branch made NCHW-contiguous, except the embedding branch which may stay channels_last
    (and is kept so only when every level's embedding is)."""
    split = _is_split(levels)
    flat = [t for lv in levels for t in (lv if split else (lv,)) if t is not None]
    half = all(x.dtype == torch.float16 for x in flat)  # `half=True` pipelines: fp16 logits are read as they are

    def norm(x, keep_cl=False):
        _require_cuda(x, "level tensor")
        if not half and x.dtype != torch.float32:
            x = x.float()
        if keep_cl and _emb_channels_last(x):
            return x
        return x.contiguous()

    if not split:
        return [norm(x) for x in levels]
    embs = [lv[2] for lv in levels if len(lv) > 2 and lv[2] is not None]
    keep_cl = len(embs) == len(levels) and all(_emb_channels_last(e) for e in embs)
    out = []
    for lv in levels:
        lv = tuple(lv)
        out.append(tuple(None if t is None else norm(t, keep_cl=(j == 2 and keep_cl)) for j, t in enumerate(lv)))
    return out


def split_levels(levels: Sequence[torch.Tensor], spec: HeadSpec, emb_channels_last: bool = True):
    """Concatenated level tensors `(B, no, H, W)` -> the split layout `[(box, cls, emb, state), ...]` holding the same values
    (what a head that skips `torch.cat` hands over; used by tests and bench.py to feed both layouts the same numbers)."""
    out = []
    c0, c1 = 4 * spec.reg_max, 4 * spec.reg_max + spec.nc
    for x in levels:
        emb = state = None
        if spec.embed_dim:
            emb = x[:, c1:c1 + spec.embed_dim]
            emb = emb.contiguous(memory_format=torch.channels_last) if emb_channels_last else emb.contiguous()
        if spec.state_classes:
            state = x[:, c1 + spec.embed_dim:c1 + spec.embed_dim + spec.state_classes].contiguous()
        out.append((x[:, :c0].contiguous(), x[:, c0:c1].contiguous(), emb, state))
    return out


def cat_levels(levels) -> list:
    """Inverse of `split_levels`: the reference's concatenated layout (head.py:204-206)."""
    if not _is_split(levels):
        return list(levels)
    return [torch.cat([t.contiguous() for t in lv if t is not None], 1) for lv in levels]


def _split(out: torch.Tensor, counts: torch.Tensor) -> List[torch.Tensor]:
    # one D2H of B ints (the reference syncs several times per image, ops.py:253,275)
    n = counts.tolist()
    return [out[b, : n[b]] for b in range(out.shape[0])]


_WS_CACHE = {}  # (device index, stream, nbytes, batch) -> persistent workspace whose histogram head is clean
_WS_CACHE_MAX = 32  # e.g. 4 steps in flight x (per-tile + merge workspace) x 2 streams each, without evicting a live entry


class _Workspace:
    """Workspace for one call.  Persistent per (device, stream, size, batch): the library leaves the score
    histogram at the head of the buffer zeroed after every successful call, so only the first use pays a
    memset (`sarpost_workspace_prepare`).  Stream order makes reuse on the same stream safe; a failed call
    drops the entry."""

    def __init__(self, nbytes: int, batch: int, device: torch.device):
        self.key = (device.index, _stream_ptr(device), int(nbytes), int(batch))
        self.nbytes = int(nbytes)
        ent = _WS_CACHE.pop(self.key, None)
        if ent is None:
            ent = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
            _lib.check(lib.sarpost_workspace_prepare(ent.data_ptr(), self.nbytes, int(batch), _stream_ptr(device)))
        self.buf = ent

    def ptr(self) -> int:
        return self.buf.data_ptr()

    def release(self) -> None:
        """Call after a successful launch: the buffer goes back to the cache (most recently used last)."""
        _WS_CACHE[self.key] = self.buf
        while len(_WS_CACHE) > _WS_CACHE_MAX:
            _WS_CACHE.pop(next(iter(_WS_CACHE)))


def clear_workspace_cache() -> None:
    _WS_CACHE.clear()


def non_max_suppression(
    prediction,
    conf_thres=0.25,
    iou_thres=0.45,
    classes=None,
    agnostic=False,
    multi_label=False,
    labels=(),
    max_det=300,
    nc=0,
    max_time_img=0.05,
    max_nms=30000,
    max_wh=7680,
    in_place=True,
    rotated=False,
    return_index=False,
    results_state_cols=None,
):
    """Drop-in for `ultralytics.utils.ops.non_max_suppression` (utils/ops.py:167-316) on CUDA tensors.

    Same arguments and result: a list of length batch with one `(n_i, 6 + nm)` tensor per image, columns
    `x1, y1, x2, y2, confidence, class, extras...`, rows in descending confidence, `n_i <= max_det`.
    Kept sets are bit-exact to the reference (torchvision CPU NMS semantics).  Documented deviations
    (SURVEY.md Appendix B.10): the input is never modified (`in_place` is accepted and ignored), the
    wall-clock limit `max_time_img` is accepted and ignored (nothing here can time out), score ties at
    the `max_nms` cut resolve to the lower anchor index (the reference's order there is torch-version
    defined).  `rotated=True` raises NotImplementedError.  Apriori `labels` (save_hybrid) are supported; with
    `return_index` a label row reports anchor index `A + label_row`.
    `return_index=True` (extension) also returns per image the int32 `anchor*nc + class` of each row.
    `results_state_cols=S` (extension): the last S of the `nm` extras columns are state probabilities;
    returns `(boxes, embeds)` per-image lists in the layout `JDEPredictor.postprocess` builds (models/yolo/jde/predict.py:52-66):
    boxes `(n, 7)` = x1,y1,x2,y2,argmax(state),conf,cls (6 columns when S == 0) and the contiguous `(n, nm - S)` embeddings.
    """
    # the drop-in keeps the reference's assertion texts (ops.py:217-218): callers and tests match on them
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]  # a validator hands over (y, raw_levels); only y is post-processed (ops.py:219-220)
    _require_cuda(prediction, "prediction")
    if rotated:
        raise NotImplementedError("sarpost: rotated=True (OBB probiou NMS, ops.py:146-164) is outside the accelerated path")
    has_labels = bool(labels) and any(len(lb) for lb in labels)

    if prediction.shape[-1] == 6:
        # NMS-free heads (v10-style) already emit (B, N, 6) rows x1,y1,x2,y2,conf,cls: the reference only thresholds,
        # truncates and applies the class filter (ops.py:224-228).  No kernel involved: a handful of tiny torch ops.
        allow = None if classes is None else torch.as_tensor(classes, device=prediction.device)
        picked = []
        for rows in prediction:
            rows = rows[rows[:, 4] > conf_thres][:max_det]
            if allow is not None:
                rows = rows[(rows[:, 5:6] == allow).any(1)]
            picked.append(rows)
        return picked

    in_dtype = prediction.dtype
    # fp32 and fp16 (`half=True` models) are read as they are; anything else is upcast first
    pred = prediction if in_dtype in (torch.float32, torch.float16) else prediction.float()
    pred = pred.contiguous()
    bs, ch, na = (int(s) for s in pred.shape)
    nc = int(nc) or (ch - 4)  # nc=0 means "every channel after the box is a class" (ops.py:231)
    nm = ch - nc - 4  # trailing per-anchor payload carried through NMS (mask coefficients, JDE embedding + state)
    dev = pred.device
    if bs == 0 or na == 0:
        empty = [torch.zeros((0, 6 + nm), device=dev)] * bs
        if results_state_cols is not None:
            s_ = int(results_state_cols)
            empty = ([torch.zeros((0, 7 if s_ else 6), device=dev)] * bs, [torch.zeros((0, nm - s_), device=dev)] * bs)
            return empty + ([torch.zeros((0,), dtype=torch.int32, device=dev)] * bs,) if return_index else empty
        return (empty, [torch.zeros((0,), dtype=torch.int32, device=dev)] * bs) if return_index else empty

    params, _keep = _make_params(conf_thres, iou_thres, classes, agnostic, multi_label, max_det, max_nms, max_wh)
    params.prediction_dtype = 1 if pred.dtype == torch.float16 else 0
    n_lab = 0
    if has_labels:  # apriori labels for autolabelling (ops.py:256-261): padded (B, L, 5) cls,x,y,w,h + counts
        if len(labels) != bs:
            raise ValueError(f"sarpost: {len(labels)} label tensors for a batch of {bs}")
        lbs = [torch.as_tensor(lb, dtype=torch.float32).reshape(-1, 5) for lb in labels]
        n_lab = max(int(lb.shape[0]) for lb in lbs)
        lab = torch.zeros((bs, n_lab, 5), dtype=torch.float32)
        for i, lb in enumerate(lbs):
            lab[i, : lb.shape[0]] = lb.detach().cpu()
        lab = lab.to(dev)
        lab_cnt = torch.tensor([int(lb.shape[0]) for lb in lbs], dtype=torch.int32).to(dev)
        params.labels, params.label_counts, params.max_labels = lab.data_ptr(), lab_cnt.data_ptr(), n_lab
    with torch.cuda.device(dev):
        ws_bytes = lib.sarpost_workspace_bytes(bs, na + n_lab, nc, int(bool(multi_label)), int(max_det))
        if ws_bytes < 0:
            _lib.check(int(ws_bytes))
        ws = _Workspace(ws_bytes, bs, dev)
        res = results_state_cols is not None
        if res:
            n_sig = int(results_state_cols)
            if not 0 <= n_sig <= nm:
                raise ValueError(f"sarpost: results_state_cols {n_sig} outside [0, nm={nm}]")
            res_boxes = torch.empty((bs, int(max_det), 7), dtype=torch.float32, device=dev)
            res_embeds = torch.empty((bs, int(max_det), nm - n_sig), dtype=torch.float32, device=dev)
            params.res_boxes, params.res_state_cols = res_boxes.data_ptr(), n_sig
            params.res_embeds = res_embeds.data_ptr() if nm - n_sig else None
        out = None if res else torch.empty((bs, int(max_det), 6 + nm), dtype=torch.float32, device=dev)
        counts = torch.empty((bs,), dtype=torch.int32, device=dev)
        kidx = torch.empty((bs, int(max_det)), dtype=torch.int32, device=dev) if return_index else None
        _lib.check(lib.sarpost_nms_decoded(pred.data_ptr(), bs, ch, na, nc, C.byref(params), out.data_ptr() if out is not None else None,
                                           counts.data_ptr(), kidx.data_ptr() if return_index else None,
                                           ws.ptr(), ws_bytes, _stream_ptr(dev)))
        ws.release()
    if res:
        if n_sig == 0:
            res_boxes = torch.cat((res_boxes[..., :4], res_boxes[..., 5:]), -1)
        if in_dtype != torch.float32:
            res_boxes, res_embeds = res_boxes.to(in_dtype), res_embeds.to(in_dtype)
        n = counts.tolist()
        bx, em = [res_boxes[b, : n[b]] for b in range(bs)], [res_embeds[b, : n[b]] for b in range(bs)]
        return (bx, em, [kidx[b, : n[b]] for b in range(bs)]) if return_index else (bx, em)
    if in_dtype != torch.float32:
        out = out.to(in_dtype)
    rows = _split(out, counts)
    if return_index:
        n = [r.shape[0] for r in rows]
        return rows, [kidx[b, : n[b]] for b in range(bs)]
    return rows


def decode(levels: Sequence[torch.Tensor], spec: HeadSpec) -> torch.Tensor:
    """`Detect._inference` / `JDE._inference` (head.py:100-131, :214-249): raw level logits ->
    `y (B, 4 + nc + embed_dim + state_classes, A)` with xywh boxes in pixels, class probabilities, raw
    embedding and sigmoid state."""
    levels = _prep_levels(cat_levels(levels))  # y is defined on the concatenated layout (head.py:218)
    head = _make_head(levels, spec)
    dev = levels[0].device
    anchors = sum(int(x.shape[2]) * int(x.shape[3]) for x in levels)
    with torch.cuda.device(dev):
        y = torch.empty((head.batch, 4 + spec.nc + spec.nm, anchors), dtype=levels[0].dtype, device=dev)
        _lib.check(lib.sarpost_decode(C.byref(head), y.data_ptr(), _stream_ptr(dev)))
    return y


def postprocess_fused(levels: Sequence[torch.Tensor], spec: HeadSpec, conf_thres=0.25, iou_thres=0.45, classes=None,
                      agnostic=False, multi_label=False, max_det=300, max_nms=30000, max_wh=7680,
                      return_index=False, return_padded=False, with_extras=True, scale_to=None, peer_out=None,
                      state_mlp: Optional[StateMLP] = None, out=None, nms_stats: Optional[torch.Tensor] = None,
                      results: bool = False, nms_cluster: int = 0):
    """decode + non_max_suppression in one pass (never materialises y; the extras channels are read only
    for the kept rows).  `levels`: the reference's concatenated `(B, no, H_l, W_l)` tensors, or the split layout —
    per level a tuple `(box, cls[, emb[, state]])` of the branch outputs before `torch.cat` (head.py:204-206), the
    embedding optionally channels_last (see `split_levels`, include/sarpost.h SARPOST_LAYOUT_SPLIT).  Result as `non_max_suppression`; `return_padded=True` returns the raw
    `(out (B, max_det, 6+nm), counts (B,) int32[, kept_index])` device tensors without any host sync.
    `with_extras=False` returns 6-column rows even for a JDE head (use `gather_extras` later for the rows that
    survive a subsequent stage such as the cross-tile merge).
    `scale_to=(img1_shape, [img0_shape, ...])` additionally applies `ops.scale_boxes` + `clip_boxes` per image
    (the loop of models/yolo/jde/predict.py:48-49) inside the gather kernel: `img1_shape` = (h, w) of the network
    input, one (h, w[, c]) per original image.
    `peer_out=dist.PeerGatherBuffer` (multi-GPU): the gather kernel stores this rank's rows and counts into every
    rank's buffer over NVLink peer memory — the all-gather is fused into the kernel; returns `(rows, counts)` views
    of the full buffers, valid after `peer_out.barrier()`.
    `state_mlp=StateMLP` (deferred JDE state head): `levels` come from a head that SKIPPED
    `state_predictor` (channels = box, cls, embedding only; `spec.state_classes` still names S); the MLP + sigmoid
    run on the kept rows' embeddings after the gather and fill the last S columns — same row layout as the reference.
    `out=(rows, counts)`: write into caller-owned contiguous `(B, max_det, row_len)` fp32 / `(B,)` int32 CUDA tensors (e.g. a
    slice of a larger gather buffer) instead of allocating; implies the padded return form.
    `nms_stats`: optional `(B, 4)` int64 CUDA tensor receiving the NMS kernel's instrumentation counters
    (`sarpost_nms_params_t.stats`).
    `results=True`: the gather kernel writes what `JDEPredictor.postprocess` assembles per image with
    split / argmax / cat (models/yolo/jde/predict.py:52-66) — `boxes (B, max_det, 7)` = x1,y1,x2,y2,state_id,conf,cls
    (6 columns x1,y1,x2,y2,conf,cls when the head has no state classes, :73-75) and a contiguous `embeds (B, max_det, E)` —
    instead of the `(6+nm)`-column rows.  Returns per-image lists `(boxes, embeds)`, or with `return_padded=True` the
    device tensors `(boxes, embeds, counts)`.  Combines with `scale_to` (boxes come out in original-image pixels,
    :49) and `state_mlp` (the deferred MLP then only produces the state id).
    `nms_cluster` (`sarpost_nms_params_t.nms_cluster`): CTAs the NMS kernel spends per image, 0 = automatic; pass 1 when
    several calls are kept in flight on different streams."""
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if len(levels) == 0:
        raise ValueError("sarpost: no level tensors")
    _require_cuda(_box_of(levels[0]), "level tensor")
    if int(_box_of(levels[0]).shape[0]) == 0 and peer_out is None:  # empty batch: the reference returns an empty list (ops.py:250)
        dev0, cols = _box_of(levels[0]).device, 6 + (spec.nm if with_extras else 0)
        if results:
            cols = 7 if (spec.state_classes > 0 or state_mlp is not None) else 6
            if return_padded:
                e = (torch.zeros((0, int(max_det), cols), device=dev0), torch.zeros((0, int(max_det), spec.embed_dim), device=dev0),
                     torch.zeros((0,), dtype=torch.int32, device=dev0))
                return e + (torch.zeros((0, int(max_det)), dtype=torch.int32, device=dev0),) if return_index else e
            return ([], [], []) if return_index else ([], [])
        if return_padded:
            e = (torch.zeros((0, int(max_det), cols), device=dev0), torch.zeros((0,), dtype=torch.int32, device=dev0))
            return e + (torch.zeros((0, int(max_det)), dtype=torch.int32, device=dev0),) if return_index else e
        return ([], []) if return_index else []
    if results and (peer_out is not None or out is not None or not with_extras):
        raise ValueError("sarpost: results=True does not combine with peer_out=, out= or with_extras=False")
    tail = 0
    if state_mlp is not None:
        if not with_extras or peer_out is not None:
            raise ValueError("sarpost: state_mlp needs with_extras=True and does not combine with peer_out")
        if (state_mlp.embed_dim, state_mlp.n_state) != (spec.embed_dim, spec.state_classes):
            raise ValueError(f"sarpost: state MLP is {state_mlp.embed_dim}->{state_mlp.n_state}, head has embed_dim "
                             f"{spec.embed_dim}, state_classes {spec.state_classes}")
        tail = 0 if results else spec.state_classes
        spec = HeadSpec(nc=spec.nc, strides=spec.strides, reg_max=spec.reg_max, embed_dim=spec.embed_dim, state_classes=0)
    levels, head = _head_for(levels, spec, with_extras)
    nm = spec.nm if with_extras else 0
    dev = _box_of(levels[0]).device
    anchors = sum(int(_box_of(x).shape[2]) * int(_box_of(x).shape[3]) for x in levels)
    bs = head.batch
    rescale = None
    if scale_to is not None:
        img1_shape, img0_shapes = scale_to
        if len(img0_shapes) != bs:
            raise ValueError(f"sarpost: scale_to has {len(img0_shapes)} original shapes for a batch of {bs}")
        rescale = scale_params(img1_shape, img0_shapes, dev)
    params, _keep = _make_params(conf_thres, iou_thres, classes, agnostic, multi_label, max_det, max_nms, max_wh, rescale, nms_cluster)
    params.out_tail_cols = tail
    if nms_stats is not None:
        if nms_stats.dtype != torch.int64 or tuple(nms_stats.shape) != (bs, 4) or not nms_stats.is_contiguous() or nms_stats.device != dev:
            raise ValueError("sarpost: nms_stats must be a contiguous (B, 4) int64 tensor on the levels' device")
        nms_stats.zero_()  # the kernel accumulates into it
        params.stats = nms_stats.data_ptr()
    if peer_out is not None:
        _bind_peer(params, peer_out, bs, max_det, 6 + nm)
    with torch.cuda.device(dev):
        ws_bytes = lib.sarpost_workspace_bytes(bs, anchors, spec.nc, int(bool(multi_label)), int(max_det))
        if ws_bytes < 0:
            _lib.check(int(ws_bytes))
        ws = _Workspace(ws_bytes, bs, dev)
        res_boxes = res_embeds = None
        if results:
            res_boxes = torch.empty((bs, int(max_det), 7), dtype=torch.float32, device=dev)
            res_embeds = torch.empty((bs, int(max_det), spec.embed_dim), dtype=torch.float32, device=dev)
            params.res_boxes = res_boxes.data_ptr()
            params.res_embeds = res_embeds.data_ptr() if spec.embed_dim else None
        if out is not None:
            out, counts = out
            if peer_out is not None:
                raise ValueError("sarpost: out= and peer_out= are exclusive")
            for t, shp, dt in ((out, (bs, int(max_det), 6 + nm + tail), torch.float32), (counts, (bs,), torch.int32)):
                if tuple(t.shape) != shp or t.dtype != dt or not t.is_contiguous() or t.device != dev:
                    raise ValueError(f"sarpost: out= tensors must be contiguous {shp} {dt} on {dev}")
            return_padded = True
        else:
            out = torch.empty((bs, int(max_det), 6 + nm + tail), dtype=torch.float32, device=dev) if (peer_out is None and not results) else None
            counts = torch.empty((bs,), dtype=torch.int32, device=dev)
        want_idx = return_index
        kidx = torch.empty((bs, int(max_det)), dtype=torch.int32, device=dev) if want_idx else None
        _lib.check(lib.sarpost_fused(C.byref(head), C.byref(params), out.data_ptr() if out is not None else None,
                                     counts.data_ptr(), kidx.data_ptr() if want_idx else None, ws.ptr(), ws_bytes,
                                     _stream_ptr(dev)))
        ws.release()
        if state_mlp is not None and results:
            _lib.check(lib.sarpost_state_ids(res_embeds.data_ptr(), counts.data_ptr(), res_boxes.data_ptr(), bs, int(max_det),
                                             state_mlp.embed_dim, state_mlp.n_state, state_mlp.hidden, state_mlp.w1.data_ptr(),
                                             state_mlp.b1.data_ptr(), state_mlp.w2.data_ptr(), state_mlp.b2.data_ptr(), _stream_ptr(dev)))
        elif state_mlp is not None:
            state_head(out, counts, state_mlp, emb_col=6, state_col=6 + nm)
    if results:
        has_state = spec.state_classes > 0 or state_mlp is not None
        if not has_state:  # no state id to show: the reference hands Boxes the plain 6 columns (predict.py:73-75)
            res_boxes = torch.cat((res_boxes[..., :4], res_boxes[..., 5:]), -1)
        if return_padded:
            return (res_boxes, res_embeds, counts, kidx) if want_idx else (res_boxes, res_embeds, counts)
        n = counts.tolist()
        bx, em = [res_boxes[b, : n[b]] for b in range(bs)], [res_embeds[b, : n[b]] for b in range(bs)]
        return (bx, em, [kidx[b, : n[b]] for b in range(bs)]) if want_idx else (bx, em)
    if peer_out is not None:
        return (peer_out.rows, peer_out.counts, kidx) if want_idx else (peer_out.rows, peer_out.counts)
    if return_padded:
        return (out, counts, kidx) if want_idx else (out, counts)
    rows = _split(out, counts)
    if return_index:
        n = [r.shape[0] for r in rows]
        return rows, [kidx[b, : n[b]] for b in range(bs)]
    return rows


def gather_extras(levels: Sequence[torch.Tensor], spec: HeadSpec, image_index: torch.Tensor,
                  anchor_index: torch.Tensor) -> torch.Tensor:
    """Extras (raw embedding + sigmoid state, head.py:247) of explicit `(image, anchor)` pairs -> `(n, nm)`."""
    levels = _prep_levels(levels)
    head = _make_head(levels, spec)
    dev = _box_of(levels[0]).device
    ii = image_index.to(device=dev, dtype=torch.int32).contiguous()
    ai = anchor_index.to(device=dev, dtype=torch.int32).contiguous()
    if ii.shape != ai.shape or ii.dim() != 1:
        raise ValueError("sarpost: image_index and anchor_index must be 1-D tensors of equal length")
    out = torch.empty((ii.shape[0], spec.nm), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sarpost_gather_extras(C.byref(head), ii.data_ptr(), ai.data_ptr(), int(ii.shape[0]), out.data_ptr(),
                                             _stream_ptr(dev)))
    return out


def _bind_peer(params, peer_out, bs: int, max_det: int, row_len: int) -> None:
    """Point the gather kernel's output at every rank's `dist.PeerGatherBuffer` (fused gather + exchange)."""
    if (peer_out.per, peer_out.max_det, peer_out.row_len) != (bs, int(max_det), row_len):
        raise ValueError("sarpost: peer_out buffer geometry does not match this call")
    rp, cp = peer_out.peer_ptrs()
    params.n_peers = len(rp)
    params.peer_slot_offset = peer_out.slot_offset
    for q in range(len(rp)):
        params.peer_out[q] = rp[q]
        params.peer_counts[q] = cp[q]


def merge_tiles(dets: torch.Tensor, det_counts: torch.Tensor, origins: torch.Tensor, tiles_per_frame: int,
                iou_thres=0.45, agnostic=False, max_det=300, max_nms=30000, max_wh=7680, return_index=False,
                return_padded=False, peer_out=None, nms_cluster: int = 0):
    """Cross-tile merge of sliced inference: `dets (T, D, row_len)` padded per-tile detections
    (x1,y1,x2,y2,conf,cls,extras...), `det_counts (T,)` int32, `origins (T, 2)` tile offsets in frame
    pixels; T = n_frames * tiles_per_frame with the tiles of a frame contiguous.  Per frame: shift by
    the tile origin, then the same class-offset NMS as ops.py:289-297.  Returns a list of per-frame
    `(n_f, row_len)` tensors.  `peer_out=dist.PeerGatherBuffer` (frames sharded over ranks): the merged rows and counts of
    this rank's frames are stored into every rank's buffer by the gather kernel; returns the full `(rows, counts)` buffers,
    valid after `peer_out.barrier()`."""
    _require_cuda(dets, "dets")
    dets = dets.float().contiguous()
    det_counts = det_counts.to(device=dets.device, dtype=torch.int32).contiguous()
    origins = origins.to(device=dets.device, dtype=torch.float32).contiguous()
    t, d, row_len = (int(s) for s in dets.shape)
    if t % tiles_per_frame:
        raise ValueError("sarpost: number of tiles is not a multiple of tiles_per_frame")
    nf = t // tiles_per_frame
    dev = dets.device
    params, _keep = _make_params(0.0, iou_thres, None, agnostic, False, max_det, max_nms, max_wh, None, nms_cluster)
    if peer_out is not None:
        _bind_peer(params, peer_out, nf, max_det, row_len)
    with torch.cuda.device(dev):
        ws_bytes = lib.sarpost_merge_workspace_bytes(nf, tiles_per_frame, d, int(max_det))
        if ws_bytes < 0:
            _lib.check(int(ws_bytes))
        ws = _Workspace(ws_bytes, nf, dev)
        out = torch.empty((nf, int(max_det), row_len), dtype=torch.float32, device=dev) if peer_out is None else None
        counts = torch.empty((nf,), dtype=torch.int32, device=dev)
        kidx = torch.empty((nf, int(max_det)), dtype=torch.int32, device=dev) if return_index else None
        _lib.check(lib.sarpost_merge_tiles(dets.data_ptr(), det_counts.data_ptr(), origins.data_ptr(), nf,
                                           tiles_per_frame, d, row_len, C.byref(params), out.data_ptr() if out is not None else None,
                                           counts.data_ptr(), kidx.data_ptr() if return_index else None,
                                           ws.ptr(), ws_bytes, _stream_ptr(dev)))
        ws.release()
    if peer_out is not None:
        return (peer_out.rows, peer_out.counts, kidx) if return_index else (peer_out.rows, peer_out.counts)
    if return_padded:
        return (out, counts, kidx) if return_index else (out, counts)
    rows = _split(out, counts)
    if return_index:
        n = [r.shape[0] for r in rows]
        return rows, [kidx[b, : n[b]] for b in range(nf)]
    return rows


def match_predictions(dets: torch.Tensor, det_counts: torch.Tensor, gt_boxes: torch.Tensor, gt_cls: torch.Tensor,
                      gt_counts: torch.Tensor, iouv: Sequence[float] = tuple(0.5 + 0.05 * i for i in range(10)),
                      tag_threshold_index: Optional[int] = None):
    """Batched validator matching on the GPU: `box_iou` (utils/metrics.py:55-75) + `match_predictions`
    (engine/validator.py:222-262, `use_scipy=False`) for every image of the batch in one launch, no host copies.
      dets (B, max_det, row_len>=6) padded detections, det_counts (B,), gt_boxes (B, max_gt, 4) xyxy,
      gt_cls (B, max_gt), gt_counts (B,)
    Returns `correct (B, max_det, len(iouv))` bool; with `tag_threshold_index` also `matched_gt (B, max_det)` int32 —
    the label index each detection matched at that threshold or -1 (JDE: `true_tags[matched_gt]`, jde/val.py:731-735)."""
    _require_cuda(dets, "dets")
    dev = dets.device
    dets = dets.float().contiguous()
    bsz, max_det, row_len = (int(v) for v in dets.shape)
    gt_boxes = gt_boxes.to(device=dev, dtype=torch.float32).contiguous()
    max_gt = int(gt_boxes.shape[1])
    if max_gt == 0:
        correct = torch.zeros((bsz, max_det, len(iouv)), dtype=torch.bool, device=dev)
        return (correct, torch.full((bsz, max_det), -1, dtype=torch.int32, device=dev)) if tag_threshold_index is not None else correct
    gt_cls = gt_cls.to(device=dev, dtype=torch.float32).contiguous()
    det_counts = det_counts.to(device=dev, dtype=torch.int32).contiguous()
    gt_counts = gt_counts.to(device=dev, dtype=torch.int32).contiguous()
    thr = (C.c_float * len(iouv))(*[float(v) for v in iouv])
    correct = torch.empty((bsz, max_det, len(iouv)), dtype=torch.uint8, device=dev)
    matched = torch.empty((bsz, max_det), dtype=torch.int32, device=dev) if tag_threshold_index is not None else None
    with torch.cuda.device(dev):
        _lib.check(lib.sarpost_match_predictions(dets.data_ptr(), det_counts.data_ptr(), bsz, max_det, row_len,
                                                 gt_boxes.data_ptr(), gt_cls.data_ptr(), gt_counts.data_ptr(), max_gt, thr,
                                                 len(iouv), correct.data_ptr(), matched.data_ptr() if matched is not None else None,
                                                 int(tag_threshold_index) if tag_threshold_index is not None else -1,
                                                 _stream_ptr(dev)))
    correct = correct.bool()
    return (correct, matched) if matched is not None else correct


def match_from_iou(pred_classes: torch.Tensor, true_classes: torch.Tensor, iou: torch.Tensor, iouv,
                   tag_threshold_index: Optional[int] = None):
    """`BaseValidator.match_predictions(pred_classes, true_classes, iou)` (engine/validator.py:222-262, `use_scipy=False`) on
    the GPU for one image: `iou (n_gt, n_det)` as the validator computed it, `iouv` the thresholds (tensor or sequence).
    Returns the `(n_det, len(iouv))` bool tensor on the device; with `tag_threshold_index` also `matched_gt (n_det,)` int32
    (label index matched at that threshold, else -1 — JDE tags, jde/val.py:731-735)."""
    _require_cuda(iou, "iou")
    dev = iou.device
    thr_list = [float(v) for v in (iouv.detach().cpu().tolist() if isinstance(iouv, torch.Tensor) else iouv)]
    n_det = int(pred_classes.shape[0])
    n_gt = int(true_classes.shape[0])
    correct = torch.zeros((n_det, len(thr_list)), dtype=torch.uint8, device=dev)
    matched = torch.full((n_det,), -1, dtype=torch.int32, device=dev) if tag_threshold_index is not None else None
    if n_det and n_gt:
        iou = iou.to(torch.float32)
        if iou.dim() != 2 or tuple(iou.shape) != (n_gt, n_det):
            raise ValueError(f"sarpost: iou has shape {tuple(iou.shape)}, expected ({n_gt}, {n_det})")
        if iou.stride(1) != 1:
            iou = iou.contiguous()
        pc = pred_classes.to(device=dev, dtype=torch.float32).contiguous()
        tc = true_classes.to(device=dev, dtype=torch.float32).contiguous()
        thr = (C.c_float * len(thr_list))(*thr_list)
        with torch.cuda.device(dev):
            _lib.check(lib.sarpost_match_from_iou(iou.data_ptr(), n_gt, n_det, int(iou.stride(0)), pc.data_ptr(), tc.data_ptr(), thr,
                                                  len(thr_list), correct.data_ptr(), matched.data_ptr() if matched is not None else None,
                                                  int(tag_threshold_index) if tag_threshold_index is not None else -1, _stream_ptr(dev)))
    correct = correct.bool()
    return (correct, matched) if matched is not None else correct


class _LevelSig:
    """What a set of level tensors has to reproduce for a prepared `sarpost_head_t` to stay valid with new addresses only:
    layout, and per tensor shape / dtype / memory format / device."""

    def __init__(self, levels):
        self.split = _is_split(levels)
        flat = [t for lv in levels for t in lv] if self.split else list(levels)
        self.device = _box_of(levels[0]).device
        self.sig = [None if t is None else (t.shape, t.dtype, _emb_channels_last(t)) for t in flat]

    def matches(self, levels) -> bool:
        if _is_split(levels) != self.split:
            return False
        flat = [t for lv in levels for t in lv] if self.split else levels
        sig, dev = self.sig, self.device
        if len(flat) != len(sig):
            return False
        for t, want in zip(flat, sig):
            if want is None:
                if t is not None:
                    return False
                continue
            if t is None or t.shape != want[0] or t.dtype != want[1] or t.device != dev:
                return False
            if not (t.is_contiguous(memory_format=torch.channels_last) if want[2] else t.is_contiguous()):
                return False
        return True

    def fill(self, levels, io) -> None:
        """Addresses into `io.data / cls / emb / state` (a sarpost_head_t or sarpost_plan_io_t)."""
        if self.split:
            for i, lv in enumerate(levels):
                io.data[i] = lv[0].data_ptr()
                io.cls[i] = lv[1].data_ptr()
                io.emb[i] = lv[2].data_ptr() if len(lv) > 2 and lv[2] is not None else None
                io.state[i] = lv[3].data_ptr() if len(lv) > 3 and lv[3] is not None else None
        else:
            for i, x in enumerate(levels):
                io.data[i] = x.data_ptr()


_TLS = threading.local()  # per-thread: the cached head block is updated in place right before the library reads it


def _head_for(levels, spec: HeadSpec, with_extras: bool):
    """`(levels, sarpost_head_t)` for a device call.  A loop that hands over the same geometry every time (a predictor, a
    validator) pays the full normalisation + validation once: while layout / shapes / dtype / memory format / device match
    the previous call of this thread, the validated block is reused with the new addresses."""
    ent = getattr(_TLS, "head", None)
    if ent is not None and ent[1] == spec and ent[2] == with_extras and ent[0].matches(levels):
        ent[0].fill(levels, ent[3])
        return levels, ent[3]
    levels = _prep_levels(levels)
    head = _make_head(levels, spec, with_extras=with_extras)
    if head.batch > 0:
        _TLS.head = (_LevelSig(levels), spec, with_extras, head)
    return levels, head


class FusedPlan:
    """`postprocess_fused` for a serving loop (`sarpost_plan_*`): geometry, thresholds, workspace, launch configuration and
    the TMA tensor maps are prepared once from a first set of level tensors; a call then only hands over the addresses
    of tensors with the SAME shapes, dtype, layout and device — three kernel launches, no validation beyond a shape /
    dtype / contiguity comparison, no allocation when `out=` is given.  Rows are bit-identical to `postprocess_fused`.

        plan = sarpost.FusedPlan(levels, spec, conf_thres=0.25, iou_thres=0.7)
        for levels in stream_of_batches:
            out, counts = plan(levels)            # (B, max_det, 6+nm) padded rows, (B,) int32 — no host sync

    One plan belongs to one CUDA stream at a time (its workspace is reused in stream order)."""

    def __init__(self, levels, spec: HeadSpec, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                 max_det=300, max_nms=30000, max_wh=7680, with_extras=True, results: bool = False, nms_cluster: int = 0):
        assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
        assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
        levels = _prep_levels(levels)
        head = _make_head(levels, spec, with_extras=with_extras)
        if head.batch == 0:
            raise ValueError("sarpost: a plan needs a non-empty batch")
        self.spec, self.with_extras, self.results = spec, with_extras, bool(results)
        self.split = _is_split(levels)
        self.device = _box_of(levels[0]).device
        self.batch, self.max_det = head.batch, int(max_det)
        self.nm = spec.nm if with_extras else 0
        if results and not with_extras:
            raise ValueError("sarpost: results=True needs with_extras=True")
        self._sig = _LevelSig(levels)
        params, _keep = _make_params(conf_thres, iou_thres, classes, agnostic, multi_label, max_det, max_nms, max_wh, None, nms_cluster)
        anchors = sum(int(_box_of(x).shape[2]) * int(_box_of(x).shape[3]) for x in levels)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            ws_bytes = lib.sarpost_workspace_bytes(self.batch, anchors, spec.nc, int(bool(multi_label)), self.max_det)
            if ws_bytes < 0:
                _lib.check(int(ws_bytes))
            self._ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=self.device)
            _lib.check(lib.sarpost_workspace_prepare(self._ws.data_ptr(), int(ws_bytes), self.batch, _stream_ptr(self.device)))
            _lib.check(lib.sarpost_plan_create(C.byref(head), C.byref(params), self._ws.data_ptr(), int(ws_bytes), C.byref(self._h)))
        self._io = _lib.PlanIO()
        self._nl = len(levels)
        self._dev_index = self.device.index

    def close(self):
        if getattr(self, "_h", None):
            lib.sarpost_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, levels, out=None, return_index: bool = False, scale_to=None, nms_stats: Optional[torch.Tensor] = None):
        """`levels` as at creation (same shapes / dtype / strides / device; anything else raises).  Returns the padded device
        tensors `(out, counts[, kept_index])`, or with `results=True` `(boxes (B, max_det, 7), embeds (B, max_det, E), counts
        [, kept_index])`.  `out=` caller-owned tensors of exactly those shapes (the same tuple without kept_index)."""
        if not self._sig.matches(levels):
            raise ValueError("sarpost: level tensors differ from the ones the plan was created for (layout / shape / dtype / "
                             "memory format); a plan takes tensors as `postprocess_fused` would pass them on unchanged")
        io, dev, bs, md = self._io, self.device, self.batch, self.max_det
        self._sig.fill(levels, io)
        cur = torch.cuda.current_device()
        if cur != self._dev_index:
            torch.cuda.set_device(self._dev_index)
        try:
            rescale = None
            if scale_to is not None:
                img1_shape, img0_shapes = scale_to
                if len(img0_shapes) != bs:
                    raise ValueError(f"sarpost: scale_to has {len(img0_shapes)} original shapes for a batch of {bs}")
                rescale = scale_params(img1_shape, img0_shapes, dev)
            io.rescale = rescale.data_ptr() if rescale is not None else None
            if nms_stats is not None:
                if nms_stats.dtype != torch.int64 or tuple(nms_stats.shape) != (bs, 4) or not nms_stats.is_contiguous() or nms_stats.device != dev:
                    raise ValueError("sarpost: nms_stats must be a contiguous (B, 4) int64 tensor on the levels' device")
                nms_stats.zero_()  # the kernel accumulates into it
                io.stats = nms_stats.data_ptr()
            else:
                io.stats = None
            if out is not None:
                want = ((bs, md, 7), (bs, md, self.spec.embed_dim), (bs,)) if self.results else ((bs, md, 6 + self.nm), (bs,))
                if len(out) != len(want):
                    raise ValueError(f"sarpost: out= must hold {len(want)} tensors")
                for t, shp in zip(out, want):
                    dt = torch.int32 if len(shp) == 1 else torch.float32
                    if tuple(t.shape) != shp or t.dtype != dt or not t.is_contiguous() or t.device != dev:
                        raise ValueError(f"sarpost: out= tensors must be contiguous {shp} {dt} on {dev}")
            if self.results:
                if out is not None:
                    boxes, embeds, counts = out
                else:
                    boxes = torch.empty((bs, md, 7), dtype=torch.float32, device=dev)
                    embeds = torch.empty((bs, md, self.spec.embed_dim), dtype=torch.float32, device=dev)
                    counts = torch.empty((bs,), dtype=torch.int32, device=dev)
                io.out = None
                io.res_boxes = boxes.data_ptr()
                io.res_embeds = embeds.data_ptr() if self.spec.embed_dim else None
            else:
                if out is not None:
                    rows, counts = out
                else:
                    rows = torch.empty((bs, md, 6 + self.nm), dtype=torch.float32, device=dev)
                    counts = torch.empty((bs,), dtype=torch.int32, device=dev)
                io.out = rows.data_ptr()
                io.res_boxes = io.res_embeds = None
            io.counts = counts.data_ptr()
            kidx = torch.empty((bs, md), dtype=torch.int32, device=dev) if return_index else None
            io.kept_index = kidx.data_ptr() if return_index else None
            rc = lib.sarpost_plan_run(self._h, C.byref(io), torch.cuda.current_stream().cuda_stream)
            if rc != 0:
                _lib.check(rc)
            if self.results:
                if self.spec.state_classes == 0:  # no state id to show: the plain 6 columns (predict.py:73-75)
                    boxes = torch.cat((boxes[..., :4], boxes[..., 5:]), -1)
                ret = (boxes, embeds, counts)
            else:
                ret = (rows, counts)
        finally:
            if cur != self._dev_index:
                torch.cuda.set_device(cur)
        return ret + (kidx,) if return_index else ret


class Pipeline:
    """Throughput mode for a stream of batches in device memory (`sarpost_pipeline_*`): `submit()` enqueues the fused
    decode + NMS of one batch on the pipeline's own streams and returns the padded output tensors immediately; the
    decode kernels of successive batches run back to back on one stream, the NMS + gather of batch i on a high-priority
    stream of their own under the decode kernel of a later batch (`depth` = batches whose NMS + gather may be pending or
    running behind the decode stream).  `wait()` makes the current stream wait for everything
    submitted so far — the outputs may be read (on the current stream) after it, without any host synchronisation.

        pl = sarpost.Pipeline(device)
        for levels in batches:
            results.append(pl.submit(levels, spec, conf_thres=0.001, iou_thres=0.7))   # (out, counts)
        pl.wait()
    """

    def __init__(self, device=None, depth: int = 3):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("sarpost: Pipeline needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self._h = C.c_void_p()
        self._held = []  # inputs / outputs / parameter blocks of batches in flight: alive until wait()
        self._head_cache = None  # (_LevelSig, spec, with_extras, Head) of the last submit
        _lib.check(lib.sarpost_pipeline_create(self.device.index, int(depth), C.byref(self._h)))

    def close(self):
        if self._h:
            lib.sarpost_pipeline_destroy(self._h)
            self._h = C.c_void_p()
        self._held.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, levels, spec: HeadSpec, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
               max_det=300, max_nms=30000, max_wh=7680, return_index=False, with_extras=True, scale_to=None, out=None):
        """Arguments as `postprocess_fused`.  Returns `(out (B, max_det, 6+nm), counts (B,) int32[, kept_index])` device
        tensors that are being written asynchronously: call `wait()` before using them.  `out=(rows, counts)`: caller-owned
        output tensors (a long-running loop should rotate a few preallocated sets: tensors handed out here stay referenced
        until `wait()`, so nothing is recycled by the allocator in between)."""
        assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
        assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
        # steady state of a serving loop: the same geometry as the previous submit -> the validated head block is reused
        # with new addresses only
        ent = self._head_cache
        if ent is not None and ent[1] == spec and ent[2] == with_extras and ent[0].matches(levels):
            head = ent[3]
            ent[0].fill(levels, head)
            dev = ent[0].device
        else:
            levels = _prep_levels(levels)
            head = _make_head(levels, spec, with_extras=with_extras)
            dev = _box_of(levels[0]).device
            if dev != self.device:
                raise RuntimeError(f"sarpost: level tensors are on {dev}, the pipeline on {self.device}")
            if head.batch == 0:
                raise ValueError("sarpost: empty batch")
            self._head_cache = (_LevelSig(levels), spec, with_extras, head)
        nm = spec.nm if with_extras else 0
        rescale = None
        if scale_to is not None:
            img1_shape, img0_shapes = scale_to
            if len(img0_shapes) != head.batch:
                raise ValueError(f"sarpost: scale_to has {len(img0_shapes)} original shapes for a batch of {head.batch}")
            rescale = scale_params(img1_shape, img0_shapes, dev)
        params, keep = _make_params(conf_thres, iou_thres, classes, agnostic, multi_label, max_det, max_nms, max_wh, rescale)
        with torch.cuda.device(dev):
            if out is not None:
                out, counts = out
                for t, shp, dt in ((out, (head.batch, int(max_det), 6 + nm), torch.float32), (counts, (head.batch,), torch.int32)):
                    if tuple(t.shape) != shp or t.dtype != dt or not t.is_contiguous() or t.device != dev:
                        raise ValueError(f"sarpost: out= tensors must be contiguous {shp} {dt} on {dev}")
            else:
                out = torch.empty((head.batch, int(max_det), 6 + nm), dtype=torch.float32, device=dev)
                counts = torch.empty((head.batch,), dtype=torch.int32, device=dev)
            kidx = torch.empty((head.batch, int(max_det)), dtype=torch.int32, device=dev) if return_index else None
            _lib.check(lib.sarpost_pipeline_submit(self._h, C.byref(head), C.byref(params), out.data_ptr(), counts.data_ptr(),
                                                   kidx.data_ptr() if return_index else None, _stream_ptr(dev)))
        self._held.append((levels, rescale, out, counts, kidx, keep))
        return (out, counts, kidx) if return_index else (out, counts)

    def wait(self) -> None:
        """The current stream waits (device side) for every submitted batch; held references are released."""
        with torch.cuda.device(self.device):
            _lib.check(lib.sarpost_pipeline_wait(self._h, _stream_ptr(self.device)))
        # the tensors were allocated on the current stream and that stream now waits for their last use: freeing is safe
        self._held.clear()


class HostContext:
    """Owns the device buffers, pinned staging and stream of the HOST-buffer entry point
    (`sarpost_host_ctx_*`).  One per caller thread / device."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        self.device = int(device)
        _lib.check(lib.sarpost_host_ctx_create(self.device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib.sarpost_host_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_traffic(self) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        _lib.check(lib.sarpost_host_ctx_last_traffic(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def postprocess(self, levels: Sequence[torch.Tensor], spec: HeadSpec, conf_thres=0.25, iou_thres=0.45,
                    classes=None, agnostic=False, multi_label=False, max_det=300, max_nms=30000, max_wh=7680,
                    out: Optional[torch.Tensor] = None, return_index=False):
        """Fused post-processing of HOST level tensors; returns a list of per-image CPU tensors."""
        assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
        head = _make_head(levels, spec, host=True)
        bs = head.batch
        params, _keep = _make_params(conf_thres, iou_thres, classes, agnostic, multi_label, max_det, max_nms, max_wh)
        if out is None:
            out = torch.empty((bs, int(max_det), 6 + spec.nm), dtype=torch.float32)
        counts = torch.empty((bs,), dtype=torch.int32)
        kidx = torch.empty((bs, int(max_det)), dtype=torch.int32) if return_index else None
        _lib.check(lib.sarpost_fused_host(self._h, C.byref(head), C.byref(params), out.data_ptr(), counts.data_ptr(),
                                          kidx.data_ptr() if return_index else None))
        n = counts.tolist()
        rows = [out[b, : n[b]] for b in range(bs)]
        if return_index:
            return rows, [kidx[b, : n[b]] for b in range(bs)]
        return rows


_HOST_CTX = {}


def postprocess_host(levels: Sequence[torch.Tensor], spec: HeadSpec, device: int = 0, **kw):
    """Convenience wrapper: fused post-processing of CPU level tensors on GPU `device` (cached context)."""
    ctx = _HOST_CTX.get(device)
    if ctx is None:
        ctx = _HOST_CTX[device] = HostContext(device)
    return ctx.postprocess(levels, spec, **kw)


def last_launch_count() -> int:
    """Kernels launched by the last libsarpost call on this thread."""
    return int(lib.sarpost_last_launch_count())


def stage_timing(enabled, accumulate: bool = False) -> None:
    """Per-stage CUDA events inside the library.  `accumulate=True`: `stage_times()` returns the mean over every
    call since (no host sync per call) — used by bench.py to time K1 inside the timed region itself."""
    _lib.check(lib.sarpost_set_stage_timing(2 if (enabled and accumulate) else int(bool(enabled))))


def stage_times() -> Tuple[float, float, float, float]:
    """Milliseconds of (K1 candidates incl. histogram memset, K2-K4 select+sort+NMS, K5 gather, whole call)
    of the last timed call."""
    buf = (C.c_float * 4)()
    _lib.check(lib.sarpost_stage_times(buf))
    return tuple(float(v) for v in buf)
