"""Import the LIVE, UNMODIFIED reference for golden generation, parity checks and the baseline legs of bench.py
(TEST / BASELINE INFRASTRUCTURE — never imported by the product package).

Where the reference comes from, first hit wins:
  1. $SARPOST_REFERENCE
  2. /root/reference            (read-only source tree; exists in the build container only)
  3. <repo>/baseline/_ref       (the reference pip-installed with
         python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>
     — git-ignored, but it travels to the GPU box with the gpurun snapshot, so the `-m gpu` tests and bench.py can run
     the reference's own code there.  Nothing reads /root/reference at run time on the GPU box.)
Recipe = SURVEY.md Appendix A: stub the three missing third-party imports, pre-seed bare `ultralytics` /
`ultralytics.nn` namespace modules so the heavy package __init__ files are skipped, then import the hot-path modules.
Nothing under the reference tree is modified or copied.
"""
import contextlib
import os
import sys
import tempfile
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    for cand in (os.environ.get("SARPOST_REFERENCE"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "ultralytics")):
            return cand
    return os.environ.get("SARPOST_REFERENCE") or "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "ultralytics"))


def source() -> str:
    """Which copy of the reference is in use: 'tree' (/root/reference or $SARPOST_REFERENCE) or 'baseline/_ref'."""
    return "baseline/_ref" if os.path.abspath(REF_ROOT) == os.path.join(_REPO, "baseline", "_ref") else "tree"


def _stub_third_party():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            mpl.__version__ = "3.8.0"
            mpl.__path__ = []
            mpl.use = lambda *a, **k: None
            plt = types.ModuleType("matplotlib.pyplot")
            plt.get_backend = lambda: "agg"
            plt.switch_backend = lambda *a, **k: None
            plt.close = lambda *a, **k: None
            plt.rc = lambda *a, **k: None
            plt.rc_context = lambda *a, **k: contextlib.nullcontext()
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if "thop" not in sys.modules:
        thop = types.ModuleType("thop")

        def _profile(*a, **k):
            raise RuntimeError("thop stub")

        thop.profile = _profile
        sys.modules["thop"] = thop
    if "pytorch_metric_learning" not in sys.modules:
        pml = types.ModuleType("pytorch_metric_learning")
        pml.__path__ = []
        for s in ("miners", "distances", "losses", "reducers"):
            m = types.ModuleType("pytorch_metric_learning." + s)
            setattr(pml, s, m)
            sys.modules["pytorch_metric_learning." + s] = m
        sys.modules["pytorch_metric_learning"] = pml


_CACHE = None


def load():
    """Returns (ops, tal, head) modules of the live reference."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    _stub_third_party()
    os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp(prefix="sarpost_yolo_cfg_"))
    os.environ.setdefault("YOLO_OFFLINE", "1")
    root = os.path.join(REF_ROOT, "ultralytics")
    if "ultralytics" not in sys.modules:
        u = types.ModuleType("ultralytics")
        u.__path__ = [root]
        u.__version__ = "8.3.63"
        sys.modules["ultralytics"] = u
        n = types.ModuleType("ultralytics.nn")
        n.__path__ = [os.path.join(root, "nn")]
        sys.modules["ultralytics.nn"] = n
    import ultralytics.nn.modules.head as head
    import ultralytics.utils.ops as ops
    import ultralytics.utils.tal as tal

    _CACHE = (ops, tal, head)
    return _CACHE


def ref_decode(levels, strides, nc, embed_dim=0, state_classes=0):
    """Run the reference's own Detect/JDE._inference (head.py:100-131 / :214-249) on raw level logits."""
    import torch

    _, _, head = load()
    ch = tuple(64 for _ in levels)
    with torch.no_grad():
        if embed_dim:
            m = head.JDE(nc=nc, embed_dim=embed_dim, state_classes=(state_classes or None), ch=ch)
        else:
            m = head.Detect(nc=nc, ch=ch)
        m.stride = torch.tensor([float(s) for s in strides])
        m.eval()
        return m._inference([x.clone() for x in levels])


def ref_jde_forward(feats, strides, nc, embed_dim, state_classes, seed=0):
    """Run the reference's whole JDE head (head.py:174-249: cv2/cv3/cv4 convolutions, state_predictor on every
    anchor, _inference) in eval mode on backbone features `feats`, with seeded random weights.
    Returns (y, x_levels, module) — x_levels are the raw per-level head outputs the module returns next to y."""
    import torch

    _, _, head = load()
    torch.manual_seed(seed)
    with torch.no_grad():
        m = head.JDE(nc=nc, embed_dim=embed_dim, state_classes=state_classes, ch=tuple(int(f.shape[1]) for f in feats))
        for prm in m.state_predictor.parameters():  # default init is tiny; spread the logits so the sigmoid is exercised
            prm.mul_(4.0)
        m.stride = torch.tensor([float(s) for s in strides])
        m.eval()
        y, x = m([f.clone() for f in feats])
    return y, x, m


def ref_nms(prediction, **kw):
    """Run the reference's own ops.non_max_suppression (utils/ops.py:167-316) on a clone of `prediction`."""
    ops, _, _ = load()
    kw.setdefault("max_time_img", 1e9)  # never hit the soft time limit (ops.py:312-314)
    return ops.non_max_suppression(prediction.clone(), **kw)


def ref_jde_predictor_postprocess(y, head_module, img_shape, orig_shapes, conf, iou, max_det=300, classes=None, agnostic=False,
                                  names=None):
    """Run the reference's own, unmodified `JDEPredictor.postprocess` (models/yolo/jde/predict.py:29-78) on a decoded
    prediction `y`: NMS, `scale_boxes`, state argmax, 7-column re-pack, `Results`.  The method only reads
    `self.args`, `self.model.names / person_states / model.model[-1]` and `self.batch[0]`, so a namespace stands in for
    the predictor object.  `orig_shapes`: (h, w) per image (zero images of that size are handed over as `orig_imgs`).
    Returns per image `(boxes.data, embeds.data)`."""
    import types as _t

    import numpy as np

    load()
    import ultralytics.models.yolo.jde.predict as P

    names = names or {i: str(i) for i in range(int(head_module.nc))}
    fake = _t.SimpleNamespace(
        args=_t.SimpleNamespace(conf=conf, iou=iou, agnostic_nms=agnostic, max_det=max_det, classes=classes),
        model=_t.SimpleNamespace(names=names, person_states={0: "s"}, model=_t.SimpleNamespace(model=[head_module])),
        batch=[[f"img{i}.jpg" for i in range(len(orig_shapes))]])
    import torch

    img = torch.zeros((len(orig_shapes), 3, int(img_shape[0]), int(img_shape[1])))
    orig = [np.zeros((int(h), int(w), 3), dtype=np.uint8) for h, w in orig_shapes]
    res = P.JDEPredictor.postprocess(fake, [y.clone()], img, orig)
    return [(r.boxes.data, r.embeds.data) for r in res]
