"""ctypes binding of libsarpost.so (C ABI declared in include/sarpost.h).

The library is the product: if it is missing or does not load, importing this module raises —
there is no Python/torch fallback for any entry point.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SARPOST_LIB_PATH") or os.path.join(HERE, "libsarpost.so")  # env override: A/B builds

MAX_LEVELS = 8
MAX_CLASSES = 2048

OK, EINVAL, ECUDA, EWORKSPACE, EUNSUPPORTED = 0, -1, -2, -3, -4


class Head(C.Structure):
    """sarpost_head_t"""
    _fields_ = [
        ("nl", C.c_int32), ("batch", C.c_int32), ("no", C.c_int32), ("nc", C.c_int32), ("reg_max", C.c_int32),
        ("n_extra_raw", C.c_int32), ("n_extra_sigmoid", C.c_int32), ("dtype", C.c_int32),
        ("h", C.c_int32 * MAX_LEVELS), ("w", C.c_int32 * MAX_LEVELS), ("stride", C.c_float * MAX_LEVELS),
        ("data", C.c_void_p * MAX_LEVELS),
        ("layout", C.c_int32), ("emb_channels_last", C.c_int32),
        ("cls", C.c_void_p * MAX_LEVELS), ("emb", C.c_void_p * MAX_LEVELS), ("state", C.c_void_p * MAX_LEVELS),
    ]


class NmsParams(C.Structure):
    """sarpost_nms_params_t"""
    _fields_ = [
        ("conf_thres", C.c_float), ("iou_thres", C.c_double), ("agnostic", C.c_int32), ("multi_label", C.c_int32),
        ("max_det", C.c_int32), ("max_nms", C.c_int32), ("max_wh", C.c_float),
        ("classes", C.POINTER(C.c_int32)), ("n_classes", C.c_int32), ("labels", C.c_void_p),
        ("label_counts", C.c_void_p), ("max_labels", C.c_int32), ("rescale", C.c_void_p),
        ("peer_out", C.c_void_p * 8), ("peer_counts", C.c_void_p * 8), ("n_peers", C.c_int32),
        ("peer_slot_offset", C.c_int32), ("prediction_dtype", C.c_int32), ("workspace_clean", C.c_int32), ("out_tail_cols", C.c_int32),
        ("stats", C.c_void_p),
        ("res_boxes", C.c_void_p), ("res_embeds", C.c_void_p), ("res_state_cols", C.c_int32), ("nms_cluster", C.c_int32),
    ]


class PlanIO(C.Structure):
    """sarpost_plan_io_t"""
    _fields_ = [
        ("data", C.c_void_p * MAX_LEVELS), ("cls", C.c_void_p * MAX_LEVELS), ("emb", C.c_void_p * MAX_LEVELS),
        ("state", C.c_void_p * MAX_LEVELS),
        ("out", C.c_void_p), ("counts", C.c_void_p), ("kept_index", C.c_void_p), ("rescale", C.c_void_p), ("stats", C.c_void_p),
        ("res_boxes", C.c_void_p), ("res_embeds", C.c_void_p),
    ]


class SarpostError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsarpost error {code}: {msg}")
        self.code = code


# name -> (restype, argtypes); every symbol include/sarpost.h declares
SYMBOLS = {
    "sarpost_last_error": (C.c_char_p, []),
    "sarpost_version": (C.c_int32, []),
    "sarpost_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "sarpost_merge_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "sarpost_workspace_clean_bytes": (C.c_int64, [C.c_int32]),
    "sarpost_workspace_prepare": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "sarpost_decode": (C.c_int32, [C.POINTER(Head), C.c_void_p, C.c_void_p]),
    "sarpost_nms_decoded": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.POINTER(NmsParams),
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "sarpost_fused": (C.c_int32, [C.POINTER(Head), C.POINTER(NmsParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int64, C.c_void_p]),
    "sarpost_plan_create": (C.c_int32, [C.POINTER(Head), C.POINTER(NmsParams), C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "sarpost_plan_run": (C.c_int32, [C.c_void_p, C.POINTER(PlanIO), C.c_void_p]),
    "sarpost_plan_destroy": (None, [C.c_void_p]),
    "sarpost_gather_extras": (C.c_int32, [C.POINTER(Head), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "sarpost_merge_tiles": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.POINTER(NmsParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_void_p]),
    "sarpost_match_predictions": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.c_int32, C.c_void_p, C.c_void_p,
                                              C.c_int32, C.c_void_p]),
    "sarpost_match_from_iou": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_float),
                                           C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "sarpost_state_head": (C.c_int32, [C.c_void_p, C.c_void_p] + [C.c_int32] * 8 + [C.c_void_p] * 5),
    "sarpost_state_ids": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p] * 5),
    "sarpost_abi_sizes": (C.c_int32, [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "sarpost_host_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(C.c_void_p)]),
    "sarpost_host_ctx_destroy": (None, [C.c_void_p]),
    "sarpost_fused_host": (C.c_int32, [C.c_void_p, C.POINTER(Head), C.POINTER(NmsParams), C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "sarpost_host_ctx_last_traffic": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "sarpost_pipeline_create": (C.c_int32, [C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "sarpost_pipeline_destroy": (None, [C.c_void_p]),
    "sarpost_pipeline_submit": (C.c_int32, [C.c_void_p, C.POINTER(Head), C.POINTER(NmsParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sarpost_pipeline_wait": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "sarpost_last_launch_count": (C.c_int32, []),
    "sarpost_set_stage_timing": (C.c_int32, [C.c_int32]),
    "sarpost_stage_times": (C.c_int32, [C.POINTER(C.c_float)]),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(nvcc -gencode arch=compute_100a,code=sm_100a). sarpost has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def _check_abi() -> None:
    """The ctypes mirrors above must be the structs the library was compiled with (a stale .so would read garbage)."""
    hb, pb = C.c_int32(), C.c_int32()
    lib.sarpost_abi_sizes(C.byref(hb), C.byref(pb))
    if (hb.value, pb.value) != (C.sizeof(Head), C.sizeof(NmsParams)):
        raise ImportError(f"{LIB_PATH} was built from a different include/sarpost.h: sarpost_head_t {hb.value} B vs binding "
                          f"{C.sizeof(Head)} B, sarpost_nms_params_t {pb.value} B vs binding {C.sizeof(NmsParams)} B — rebuild it")


_check_abi()


def last_error() -> str:
    return (lib.sarpost_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise SarpostError(int(rc), last_error())
