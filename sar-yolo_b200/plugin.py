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
          defer_state: bool = False) -> None:
    """Monkey-patch the reference.  Modules default to the importable `ultralytics` package.

    `fused=False`: API-exact — `_inference` returns the real `y` (decode kernel), `non_max_suppression` runs the
    decoded-input kernels.  `fused=True`: `_inference` returns a `LazyPrediction` handle and `non_max_suppression`
    runs the single-pass fused kernels from the raw logits (the unmodified predictor / validator get the fused
    speed); `y` is materialised only if something else touches it.
    `defer_state=True` (with `fused=True`): `JDE.forward` in eval mode skips `state_predictor` on the
    `(B, A, embed_dim)` embedding map (head.py:198-204) and the patched NMS evaluates that MLP only on the kept rows
    (`sarpost_state_head`); the returned level list then has no state channels, so keep it off while a validator
    computes a loss from `preds[1]`."""
    if _SAVED:
        return
    if defer_state and not fused:
        raise ValueError("sarpost: defer_state=True needs fused=True")
    ops_mod = ultralytics_ops or importlib.import_module("ultralytics.utils.ops")
    _SAVED["ops_mod"] = ops_mod
    _SAVED["nms"] = ops_mod.non_max_suppression
    ops_mod.non_max_suppression = _nms_dispatch_fused if fused else _nms_dispatch
    if decode or fused:
        make = _make_lazy_inference if fused else _make_inference
        head_mod = ultralytics_head or importlib.import_module("ultralytics.nn.modules.head")
        _SAVED["head_mod"] = head_mod
        _SAVED["detect"] = head_mod.Detect._inference
        head_mod.Detect._inference = make("detect")
        if hasattr(head_mod, "JDE"):
            _SAVED["jde"] = head_mod.JDE._inference
            head_mod.JDE._inference = make("jde")
            if defer_state:
                _SAVED["jde_forward"] = head_mod.JDE.forward
                head_mod.JDE.forward = _jde_forward_deferred


def unpatch() -> None:
    if not _SAVED:
        return
    _SAVED["ops_mod"].non_max_suppression = _SAVED["nms"]
    head_mod = _SAVED.get("head_mod")
    if head_mod is not None:
        head_mod.Detect._inference = _SAVED["detect"]
        if "jde" in _SAVED:
            head_mod.JDE._inference = _SAVED["jde"]
        if "jde_forward" in _SAVED:
            head_mod.JDE.forward = _SAVED["jde_forward"]
    _SAVED.clear()


def is_patched() -> bool:
    return bool(_SAVED)


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
        b = int(levels[0].shape[0])
        a = sum(int(x.shape[2]) * int(x.shape[3]) for x in levels)
        r = torch.Tensor._make_wrapper_subclass(cls, (b, 4 + spec.nc + spec.nm, a), dtype=levels[0].dtype,
                                                device=levels[0].device, requires_grad=False)
        r._levels, r._spec, r._y = list(levels), spec, None
        # deferred state head: `levels` carry no state channels; `state_module` is the JDE head that owns state_predictor
        r._state_module = state_module
        return r

    def state_mlp(self):
        return None if self._state_module is None else _ops.StateMLP.from_module(self._state_module, self.device)

    def materialize(self) -> torch.Tensor:
        if self._y is None:
            levels = self._levels
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
        return LazyPrediction([xi if xi.dtype in (torch.float32, torch.float16) else xi.float() for xi in x],
                              _ops.HeadSpec.from_module(self))

    return _inference


def _jde_forward_deferred(self, x):
    """JDE.forward (head.py:193-212) in eval mode without the per-anchor state_predictor (:198-204): the three
    convolution branches are concatenated as in the stateless branch (:206) and the state MLP is left to the NMS."""
    if (self.training or self.state_classes is None or not _plain_head(self, "jde")
            or not all(isinstance(xi, torch.Tensor) and xi.is_cuda for xi in x)):
        return _SAVED["jde_forward"](self, x)
    for i in range(self.nl):
        x[i] = torch.cat((self.cv2[i](x[i]), self.cv3[i](x[i]), self.cv4[i](x[i])), 1)
    levels = [xi if xi.dtype in (torch.float32, torch.float16) else xi.float() for xi in x]
    return LazyPrediction(levels, _ops.HeadSpec.from_module(self), state_module=self), x


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
