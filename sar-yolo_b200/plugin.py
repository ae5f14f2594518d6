"""Install / remove the accelerated path under the reference's call sites.

Every reference call site resolves `ops.non_max_suppression` through the module attribute at call
time (models/yolo/jde/predict.py:31, jde/val.py:648, detect/predict.py:25, detect/val.py:94, ...) and
the heads call `self._inference(x)` (nn/modules/head.py:73, :211), so replacing those three attributes
is enough for `model.predict()` / `model.val()`.  Calls this package does not
accelerate (CPU tensors, rotated boxes, export mode) are forwarded to the ORIGINAL
reference function that was saved at patch time — never to a re-implementation of ours.
"""
from __future__ import annotations

import importlib

from . import head as _head
from . import ops as _ops

_SAVED = {}


def _nms_dispatch(prediction, *args, **kwargs):
    orig = _SAVED["nms"]
    pred = prediction[0] if isinstance(prediction, (list, tuple)) else prediction
    rotated = kwargs.get("rotated", args[12] if len(args) > 12 else False)
    if (not getattr(pred, "is_cuda", False)) or rotated:
        return orig(prediction, *args, **kwargs)  # the reference's own code path, untouched
    return _ops.non_max_suppression(prediction, *args, **kwargs)


def _make_inference(kind: str):
    fast = _head.jde_inference if kind == "jde" else _head.detect_inference

    def _inference(self, x):
        if getattr(self, "export", False) or not x[0].is_cuda or getattr(self, "reg_max", 16) != 16:
            return _SAVED[kind](self, x)
        return fast(self, x)

    return _inference


def patch(ultralytics_ops=None, ultralytics_head=None, decode: bool = True) -> None:
    """Monkey-patch the reference.  Modules default to the importable `ultralytics` package."""
    if _SAVED:
        return
    ops_mod = ultralytics_ops or importlib.import_module("ultralytics.utils.ops")
    _SAVED["ops_mod"] = ops_mod
    _SAVED["nms"] = ops_mod.non_max_suppression
    ops_mod.non_max_suppression = _nms_dispatch
    if decode:
        head_mod = ultralytics_head or importlib.import_module("ultralytics.nn.modules.head")
        _SAVED["head_mod"] = head_mod
        _SAVED["detect"] = head_mod.Detect._inference
        head_mod.Detect._inference = _make_inference("detect")
        if hasattr(head_mod, "JDE"):
            _SAVED["jde"] = head_mod.JDE._inference
            head_mod.JDE._inference = _make_inference("jde")


def unpatch() -> None:
    if not _SAVED:
        return
    _SAVED["ops_mod"].non_max_suppression = _SAVED["nms"]
    head_mod = _SAVED.get("head_mod")
    if head_mod is not None:
        head_mod.Detect._inference = _SAVED["detect"]
        if "jde" in _SAVED:
            head_mod.JDE._inference = _SAVED["jde"]
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
