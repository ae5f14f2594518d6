"""Install / remove the accelerated path under the reference's call sites.

Every reference call site resolves `ops.non_max_suppression` through the module attribute at call
time (models/yolo/jde/predict.py:31, jde/val.py:648, detect/predict.py:25, detect/val.py:94, ...) and
the heads call `self._inference(x)` (nn/modules/head.py:73, :211), so replacing those attributes
is enough for `model.predict()` / `model.val()`.  Calls this package does not
accelerate are forwarded to the ORIGINAL reference function that was saved at patch time — never to a
re-implementation of ours: CPU tensors, rotated boxes, export mode, head subclasses with their own box
decode (OBB `dist2rbox`, end2end/v10 heads that expect xyxy, Pose/Segment which read the anchor cache the
reference `_inference` fills), and argument ranges the library rejects (`max_det > 4096`, `nc > 2048`, ...).
"""
from __future__ import annotations

import importlib

import torch

from . import _lib
from . import ops as _ops

_SAVED = {}

_NMS_ARG_NAMES = ("conf_thres", "iou_thres", "classes", "agnostic", "multi_label", "labels", "max_det", "nc", "max_time_img",
                  "max_nms", "max_wh", "in_place", "rotated")


def _nms_kwargs(args, kwargs) -> dict:
    """Positional + keyword arguments of `ops.non_max_suppression` (utils/ops.py:167-182) by name."""
    kw = dict(zip(_NMS_ARG_NAMES, args))
    kw.update(kwargs)
    return kw


def _nms_supported_fields(is_cuda: bool, ndim: int, dtype, shape, kw) -> bool:
    if not is_cuda or kw.get("rotated", False) or ndim != 3 or not dtype.is_floating_point:
        return False
    if shape[-1] == 6:  # end-to-end layout: plain thresholding, handled by ops.non_max_suppression in torch
        return True
    try:
        max_det, max_nms = int(kw.get("max_det", 300)), int(kw.get("max_nms", 30000))
        nc = int(kw.get("nc", 0) or 0) or int(shape[1]) - 4
    except (TypeError, ValueError):
        return False
    return 1 <= max_det <= 4096 and max_nms >= 1 and 1 <= nc <= _lib.MAX_CLASSES and int(shape[1]) >= 4 + nc


def _nms_supported(pred, kw) -> bool:
    """True when libsarpost covers this call; everything else runs the reference's own function."""
    if not isinstance(pred, torch.Tensor):
        return False
    return _nms_supported_fields(pred.is_cuda, pred.dim(), pred.dtype, tuple(pred.shape), kw)


def _nms_dispatch(prediction, *args, **kwargs):
    pred = prediction[0] if isinstance(prediction, (list, tuple)) else prediction
    if not _nms_supported(pred, _nms_kwargs(args, kwargs)):
        return _SAVED["nms"](prediction, *args, **kwargs)  # the reference's own code path, untouched
    return _ops.non_max_suppression(prediction, *args, **kwargs)


def _plain_head(self, kind: str) -> bool:
    """Only the exact `Detect` / `JDE` classes with the stock xywh `decode_bboxes` take the accelerated decode.
    Subclasses (OBB, Pose, Segment, v10Detect, ...) override the decode or read `self.anchors/self.strides`, which
    only the reference `_inference` maintains (head.py:105-107, :303, :354)."""
    head_mod = _SAVED.get("head_mod")
    base = getattr(head_mod, "Detect", None)
    want = getattr(head_mod, "JDE", None) if kind == "jde" else base
    if want is None or type(self) is not want:
        return False
    if getattr(self, "end2end", False) or getattr(self, "export", False) or getattr(self, "reg_max", 16) != 16:
        return False
    dec, base_dec = getattr(type(self), "decode_bboxes", None), getattr(base, "decode_bboxes", None)
    return dec is base_dec


def _levels_ok(self, x) -> bool:
    return (len(x) > 0 and all(isinstance(xi, torch.Tensor) and xi.is_cuda and xi.dim() == 4 for xi in x)
            and all(int(xi.shape[1]) == int(getattr(self, "no", xi.shape[1])) for xi in x))


def _make_inference(kind: str):
    """`Detect._inference(self, x)` (head.py:100-131) / `JDE._inference` (:214-249) replacement returning the real
    `y = cat(dbox, cls.sigmoid()[, emb, state.sigmoid()])` from the decode kernel.  The reference's anchor cache
    (`self.shape/anchors/strides`, head.py:105-107) is left untouched so a later call of the ORIGINAL method (after
    `unpatch()`, or for a CPU input) recomputes it as usual."""

    def _inference(self, x):
        if not _plain_head(self, kind) or not _levels_ok(self, x):
            return _SAVED[kind](self, x)
        return _ops.decode(x, _ops.HeadSpec.from_module(self))

    return _inference


def patch(ultralytics_ops=None, ultralytics_head=None, decode: bool = True, fused: bool = False,
          defer_state: bool = False, split: bool = False, emb_channels_last: bool = True, match: bool = False,
          ultralytics_validator=None, predictor: bool = False, ultralytics_jde_predict=None) -> None:
    """Monkey-patch the reference.  Modules default to the importable `ultralytics` package.

    `fused=False`: API-exact — `_inference` returns the real `y` (decode kernel), `non_max_suppression` runs the
    decoded-input kernels.  `fused=True`: `_inference` returns a `LazyPrediction` handle and `non_max_suppression`
    runs the single-pass fused kernels from the raw logits (the unmodified predictor / validator get the fused
    speed); `y` is materialised only if something else touches it.
    `split=True` (with `fused=True`): `Detect.forward` / `JDE.forward` in eval mode hand the convolution branch outputs
    to the kernels as they are instead of concatenating them per level (head.py:204-206 writes all `no` channels only for
    `_inference` to split them again, :232-235) — with `emb_channels_last` the JDE embedding branch (`cv4`) runs in
    channels_last memory format, so the gather kernel reads one contiguous `embed_dim*4`-byte run per kept row.  The second
    element of the head's return value then holds per-level `(box, cls[, emb[, state]])` tuples instead of `(B, no, H, W)`
    tensors, so keep it off while a validator computes a loss from `preds[1]`.
    `defer_state=True` (implies `split`): `JDE.forward` in eval mode also skips `state_predictor` on the
    `(B, A, embed_dim)` embedding map (head.py:198-204) and the patched NMS evaluates that MLP only on the kept rows
    (`sarpost_state_head`).
    `match=True`: `BaseValidator.match_predictions` (engine/validator.py:222-262) runs on the GPU
    (`sarpost_match_predictions`) for CUDA inputs with `use_scipy=False`.
    `predictor=True` (with `fused=True`): `JDEPredictor.postprocess` (models/yolo/jde/predict.py:29-78) —
    NMS, then per image `scale_boxes`, split of the row, `argmax` over the states and `cat` into the 7-column boxes — becomes
    one fused call whose gather kernel writes the boxes (already in original-image pixels, state id in column 4) and the
    contiguous embeddings that `Results` takes."""
    if _SAVED:
        return
    if (defer_state or split or predictor) and not fused:
        raise ValueError("sarpost: split=True / defer_state=True / predictor=True need fused=True")
    split = split or defer_state
    ops_mod = ultralytics_ops or importlib.import_module("ultralytics.utils.ops")
    _SAVED["ops_mod"] = ops_mod
    _SAVED["nms"] = ops_mod.non_max_suppression
    _SAVED["opts"] = dict(defer_state=bool(defer_state), emb_channels_last=bool(emb_channels_last))
    ops_mod.non_max_suppression = _nms_dispatch_fused if fused else _nms_dispatch
    if decode or fused:
        make = _make_lazy_inference if fused else _make_inference
        head_mod = ultralytics_head or importlib.import_module("ultralytics.nn.modules.head")
        _SAVED["head_mod"] = head_mod
        _SAVED["detect"] = head_mod.Detect._inference
        head_mod.Detect._inference = make("detect")
        if split and hasattr(head_mod.Detect, "forward"):
            _SAVED["detect_forward"] = head_mod.Detect.forward
            head_mod.Detect.forward = _detect_forward_split
        if hasattr(head_mod, "JDE"):
            _SAVED["jde"] = head_mod.JDE._inference
            head_mod.JDE._inference = make("jde")
            if split and hasattr(head_mod.JDE, "forward"):
                _SAVED["jde_forward"] = head_mod.JDE.forward
                head_mod.JDE.forward = _jde_forward_split
    if predictor:
        pred_mod = ultralytics_jde_predict or importlib.import_module("ultralytics.models.yolo.jde.predict")
        _SAVED["pred_mod"] = pred_mod
        _SAVED["jde_post"] = pred_mod.JDEPredictor.postprocess
        pred_mod.JDEPredictor.postprocess = _jde_predictor_postprocess
    if match:
        val_mod = ultralytics_validator or importlib.import_module("ultralytics.engine.validator")
        _SAVED["val_mod"] = val_mod
        _SAVED["match"] = val_mod.BaseValidator.match_predictions
        val_mod.BaseValidator.match_predictions = _match_predictions_dispatch
        jde_val = getattr(val_mod, "JDEValidator", None)  # tests hand the class over on the same namespace
        if jde_val is None and ultralytics_validator is None:
            try:
                jde_val = importlib.import_module("ultralytics.models.yolo.jde.val").JDEValidator
            except Exception:  # noqa: BLE001  (light installs without the model zoo)
                jde_val = None
        if jde_val is not None and "match_predictions" in vars(jde_val):
            _SAVED["jde_val"] = jde_val
            _SAVED["jde_match"] = jde_val.match_predictions
            jde_val.match_predictions = _jde_match_predictions_dispatch


def unpatch() -> None:
    if not _SAVED:
        return
    _SAVED["ops_mod"].non_max_suppression = _SAVED["nms"]
    head_mod = _SAVED.get("head_mod")
    if head_mod is not None:
        head_mod.Detect._inference = _SAVED["detect"]
        if "detect_forward" in _SAVED:
            head_mod.Detect.forward = _SAVED["detect_forward"]
        if "jde" in _SAVED:
            head_mod.JDE._inference = _SAVED["jde"]
        if "jde_forward" in _SAVED:
            head_mod.JDE.forward = _SAVED["jde_forward"]
    if "jde_post" in _SAVED:
        _SAVED["pred_mod"].JDEPredictor.postprocess = _SAVED["jde_post"]
    if "match" in _SAVED:
        _SAVED["val_mod"].BaseValidator.match_predictions = _SAVED["match"]
    if "jde_match" in _SAVED:
        _SAVED["jde_val"].match_predictions = _SAVED["jde_match"]
    _SAVED.clear()


def is_patched() -> bool:
    return bool(_SAVED)


def _match_predictions_dispatch(self, pred_classes, true_classes, iou, use_scipy=False):
    """`BaseValidator.match_predictions(pred_classes, true_classes, iou)` (engine/validator.py:222-262) on the GPU: the
    reference moves `iou` to the host and loops over the 10 thresholds in numpy (one D2H per image).  `iou (n_gt, n_det)`
    arrives already computed by `box_iou`; the kernel redoes the class masking, the per-threshold greedy unique matching
    and returns the `(n_det, n_thr)` bool tensor on `pred_classes.device`.  scipy matching and CPU inputs keep the
    reference's own method."""
    if use_scipy or not (isinstance(iou, torch.Tensor) and iou.is_cuda and pred_classes.is_cuda):
        return _SAVED["match"](self, pred_classes, true_classes, iou, use_scipy)
    return _ops.match_from_iou(pred_classes, true_classes, iou, _iouv_list(self))


def _iouv_list(validator):
    """`self.iouv` (a device tensor, engine/validator.py / detect/val.py:34) as Python floats, read back once per validator."""
    cached = getattr(validator, "_sarpost_iouv", None)
    if cached is None or cached[0] is not validator.iouv:
        cached = (validator.iouv, [float(v) for v in validator.iouv.detach().cpu().tolist()])
        validator._sarpost_iouv = cached
    return cached[1]


def _jde_match_predictions_dispatch(self, pred_classes, true_classes, true_tags, iou, use_scipy=False):
    """`JDEValidator.match_predictions(pred_classes, true_classes, true_tags, iou)` (models/yolo/jde/val.py:683-736): the
    same matching plus, for the pairs matched at `threshold == self.state_iou`, the label's tag per detection (0 where
    unmatched — the reference's `[False] * n` turned into an int tensor)."""
    if use_scipy or not (isinstance(iou, torch.Tensor) and iou.is_cuda and pred_classes.is_cuda):
        return _SAVED["jde_match"](self, pred_classes, true_classes, true_tags, iou, use_scipy)
    thr = _iouv_list(self)
    tag_idx = next((i for i, t in enumerate(thr) if t == getattr(self, "state_iou", None)), None)
    n = int(pred_classes.shape[0])
    tags = torch.zeros((n,), dtype=torch.int, device=pred_classes.device)
    if tag_idx is None or true_tags.dim() == 0 or true_tags.numel() == 0 or n == 0:
        return _ops.match_from_iou(pred_classes, true_classes, iou, thr), tags
    correct, matched = _ops.match_from_iou(pred_classes, true_classes, iou, thr, tag_threshold_index=tag_idx)
    picked = true_tags.to(pred_classes.device)[matched.clamp(min=0).long()].to(torch.int)
    return correct, torch.where(matched >= 0, picked, tags)


def _jde_predictor_postprocess(self, preds, img, orig_imgs):
    """`JDEPredictor.postprocess(preds, img, orig_imgs)` (models/yolo/jde/predict.py:29-78) as ONE fused call: decode + NMS
    from the raw logits, `ops.scale_boxes` to the original image (:49), the state `argmax` and the 7-column re-pack
    `[xyxy, state_id, conf, cls]` (:61-64) and the contiguous embeddings all come out of the gather kernel; what is left
    here is building the `Results` objects.  Anything the fused path does not cover runs the reference's own method."""
    pred = preds[0] if isinstance(preds, (list, tuple)) else preds
    model = getattr(self, "model", None)
    names = getattr(model, "names", None)
    a = self.args
    ok = (isinstance(pred, LazyPrediction) and pred._y is None and names is not None and len(names) == pred._spec.nc
          and pred._spec.embed_dim > 0 and 1 <= int(a.max_det) <= 4096)
    if not ok:
        return _SAVED["jde_post"](self, preds, img, orig_imgs)
    if not isinstance(orig_imgs, list):  # a tensor source: the reference converts it to a list of HWC arrays (:41-42)
        orig_imgs = _SAVED["ops_mod"].convert_torch2numpy_batch(orig_imgs)
    boxes, embeds = _ops.postprocess_fused(pred._levels, pred._spec, conf_thres=a.conf, iou_thres=a.iou, classes=a.classes,
                                           agnostic=a.agnostic_nms, max_det=a.max_det, state_mlp=pred.state_mlp(), results=True,
                                           scale_to=(tuple(img.shape[2:]), [tuple(o.shape) for o in orig_imgs]))
    Results = importlib.import_module(_SAVED["pred_mod"].__name__.split(".models.")[0] + ".engine.results").Results \
        if not hasattr(_SAVED["pred_mod"], "Results") else _SAVED["pred_mod"].Results
    out = []
    for bx, em, orig_img, img_path in zip(boxes, embeds, orig_imgs, self.batch[0]):
        if pred.dtype != torch.float32:
            bx, em = bx.to(pred.dtype), em.to(pred.dtype)  # the reference's rows carry the prediction's dtype
        if bx.shape[0] == 0 and bx.shape[1] == 7:
            bx = bx[:, :6]  # no detections: the reference skips the state column (:66-69)
        out.append(Results(orig_img, path=img_path, names=names, person_states=getattr(model, "person_states", None), boxes=bx, embeds=em))
    return out


def fused_postprocess(preds, head_module, img_shape=None, orig_shapes=None, **nms_kwargs):
    """For custom predictors/validators (`Model.predict(predictor=...)`, engine/model.py:505,552):
    `preds` is what the PyTorch model returns, `[y, x_levels]` (head.py:212, nn/autobackend.py:700-704);
    the raw per-level logits `preds[1]` go straight through the fused kernels and `y` is ignored."""
    levels = preds[1]
    if isinstance(levels, dict):  # end2end heads return {"one2many":..., "one2one":...}
        raise NotImplementedError("sarpost: end2end heads are not on the accelerated path")
    spec = _ops.HeadSpec.from_module(head_module)
    nms_kwargs.pop("nc", None)
    if img_shape is not None and orig_shapes is not None:  # fold predict.py:49 (scale_boxes + clip) into the gather
        nms_kwargs["scale_to"] = (tuple(img_shape), [tuple(s) for s in orig_shapes])
    return _ops.postprocess_fused(levels, spec, **nms_kwargs)


# ---------------------------------------------------------------------------------------------------------------
# fused drop-in: patch(fused=True)
# ---------------------------------------------------------------------------------------------------------------
class LazyPrediction(torch.Tensor):
    """What the patched `_inference` returns under `patch(fused=True)`: a tensor-shaped handle on the raw level
    logits.  The patched `ops.non_max_suppression` recognises it and runs the fused kernels straight from the
    logits, so the `(B, 4+nc+nm, A)` tensor `y` (2.3 GB at 1280² P2, batch 16) is never written or read.  Any other
    use (a torch op, indexing, `.cpu()`, printing) transparently materialises `y` with the decode kernel first
    (`__torch_dispatch__`), so code that really needs `y` keeps working — it just pays for it."""

    @staticmethod
    def __new__(cls, levels, spec, state_module=None):
        first = _ops._box_of(levels[0])  # `levels`: concatenated tensors, or split-layout tuples (box, cls[, emb[, state]])
        b = int(first.shape[0])
        a = sum(int(_ops._box_of(x).shape[2]) * int(_ops._box_of(x).shape[3]) for x in levels)
        r = torch.Tensor._make_wrapper_subclass(cls, (b, 4 + spec.nc + spec.nm, a), dtype=first.dtype,
                                                device=first.device, requires_grad=False)
        r._levels, r._spec, r._y = list(levels), spec, None
        # deferred state head: `levels` carry no state channels; `state_module` is the JDE head that owns state_predictor
        r._state_module = state_module
        return r

    def state_mlp(self):
        return None if self._state_module is None else _ops.StateMLP.from_module(self._state_module, self.device)

    def materialize(self) -> torch.Tensor:
        if self._y is None:
            levels = _ops.cat_levels(self._levels)  # y is defined on the concatenated layout (head.py:218)
            if self._state_module is not None:  # somebody wants the whole y: run the module's own MLP on every anchor
                e0 = 4 * self._spec.reg_max + self._spec.nc
                full = []
                with torch.no_grad():
                    for x in levels:
                        b, _, h, w = x.shape
                        emb = x[:, e0: e0 + self._spec.embed_dim].flatten(2).permute(0, 2, 1)
                        st = self._state_module.state_predictor(emb.to(next(self._state_module.state_predictor.parameters()).dtype))
                        full.append(torch.cat((x, st.to(x.dtype).permute(0, 2, 1).reshape(b, -1, h, w)), 1))
                levels = full
            self._y = _ops.decode(levels, self._spec)
        return self._y

    def __repr__(self):  # noqa: D105
        return f"LazyPrediction(shape={tuple(self.shape)}, device={self.device}, materialized={self._y is not None})"

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        def unwrap(v):
            if isinstance(v, LazyPrediction):
                return v.materialize()
            if isinstance(v, (list, tuple)):
                return type(v)(unwrap(u) for u in v)
            return v

        return func(*unwrap(args), **{k: unwrap(v) for k, v in (kwargs or {}).items()})


def _make_lazy_inference(kind: str):
    def _inference(self, x):
        if not _plain_head(self, kind) or not _levels_ok(self, x):
            return _SAVED[kind](self, x)
        return LazyPrediction([_as_float(xi) for xi in x], _ops.HeadSpec.from_module(self))

    return _inference


def _as_float(t):
    return t if t.dtype in (torch.float32, torch.float16) else t.float()


def _detect_forward_split(self, x):
    """Detect.forward (head.py:64-74) in eval mode without the per-level `torch.cat` (:70): the cv2 / cv3 outputs go to the
    kernels as they are (split layout)."""
    if (self.training or not _plain_head(self, "detect") or not all(isinstance(xi, torch.Tensor) and xi.is_cuda for xi in x)):
        return _SAVED["detect_forward"](self, x)
    levels = [(_as_float(self.cv2[i](x[i])), _as_float(self.cv3[i](x[i]))) for i in range(self.nl)]
    return LazyPrediction(levels, _ops.HeadSpec.from_module(self)), levels


def _jde_forward_split(self, x):
    """JDE.forward (head.py:193-212) in eval mode without the per-level `torch.cat` (:204-206): box, class, embedding
    [and state] branch outputs are handed over separately.  With `emb_channels_last` the cv4 branch runs in
    channels_last memory format (its input feature map is converted once), so the embedding lands as (B, H, W, E).
    With `defer_state` the per-anchor state_predictor (:198-204) is skipped and left to the NMS (kept rows only)."""
    opts = _SAVED.get("opts", {})
    if (self.training or not _plain_head(self, "jde") or not all(isinstance(xi, torch.Tensor) and xi.is_cuda for xi in x)):
        return _SAVED["jde_forward"](self, x)
    defer = opts.get("defer_state", False) and self.state_classes is not None
    levels = []
    for i in range(self.nl):
        box, cls = self.cv2[i](x[i]), self.cv3[i](x[i])
        xin = x[i].contiguous(memory_format=torch.channels_last) if opts.get("emb_channels_last", True) else x[i]
        emb = self.cv4[i](xin)
        state = None
        if self.state_classes is not None and not defer:
            b, c, h, w = emb.shape
            flat = emb.permute(0, 2, 3, 1).reshape(b, h * w, c)  # (B, H*W, E): a view when emb is channels_last
            state = self.state_predictor(flat).permute(0, 2, 1).reshape(b, self.state_classes, h, w)
        levels.append(tuple(None if t is None else _as_float(t) for t in (box, cls, emb, state)))
    if defer:
        levels = [lv[:3] for lv in levels]
    return LazyPrediction(levels, _ops.HeadSpec.from_module(self), state_module=self if defer else None), levels


def _nms_dispatch_fused(prediction, *args, **kwargs):
    pred = prediction[0] if isinstance(prediction, (list, tuple)) else prediction
    if isinstance(pred, LazyPrediction):
        kw = _nms_kwargs(args, kwargs)
        has_labels = bool(kw.get("labels")) and any(len(lb) for lb in kw["labels"])
        nc = int(kw.get("nc") or 0)
        if (not kw.get("rotated") and not has_labels and nc in (0, pred._spec.nc) and pred._y is None
                and 1 <= int(kw.get("max_det", 300)) <= 4096 and int(kw.get("max_nms", 30000)) >= 1):
            fused_kw = {k: kw[k] for k in ("conf_thres", "iou_thres", "classes", "agnostic", "multi_label", "max_det",
                                           "max_nms", "max_wh") if k in kw}
            rows = _ops.postprocess_fused(pred._levels, pred._spec, state_mlp=pred.state_mlp(), **fused_kw)
            if pred.dtype != torch.float32:
                rows = [r.to(pred.dtype) for r in rows]  # the reference returns rows in the prediction's dtype
            return rows
        prediction = pred.materialize()
    return _nms_dispatch(prediction, *args, **kwargs)
