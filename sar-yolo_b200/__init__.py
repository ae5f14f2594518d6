"""sarpost — B200-native (sm_100a) detection post-processing for SAR-YOLO.

One hot path of HaoqianSong/SAR-YOLO rebuilt from scratch behind the reference's own interface:
Detect/JDE head decode (ultralytics/nn/modules/head.py:100-131, :214-249) and
`ops.non_max_suppression` (ultralytics/utils/ops.py:167-316).  Hand-written CUDA kernels live in
`csrc/` and are reached through the C ABI in `include/sarpost.h` (ctypes, `_lib.py`); this package is
the Python host side that mirrors the reference's function signatures.

Importing the package loads `libsarpost.so` and raises ImportError if it has not been built —
there is no CPU / PyTorch fallback.  (The directory is named `sar-yolo_b200`; `import sarpost` via
the alias module at the repo root, or `importlib.import_module("sar-yolo_b200")`.)
"""
from . import synth  # noqa: F401  (pure torch, no native code)
from . import _lib  # noqa: F401  (raises if libsarpost.so is missing)
from . import ops, plugin, dist  # noqa: F401
from .ops import (HeadSpec, HostContext, Pipeline, FusedPlan, StateMLP, state_head, decode, gather_extras, match_predictions, match_from_iou, split_levels, cat_levels, merge_tiles, non_max_suppression, postprocess_fused,  # noqa: F401
                  postprocess_host)
from .plugin import patch, unpatch  # noqa: F401
from ._lib import SarpostError  # noqa: F401

__version__ = "0.1.0"
