"""Loading of tests/golden/*.npz (written by tests/golden/make_golden.py from the live reference)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    counts = z["counts"].tolist()
    rows = torch.from_numpy(z["rows"])
    split = list(torch.split(rows, counts)) if counts else []
    out = dict(meta=meta, counts=counts, rows=split)
    if "y_sub" in z.files:
        out["y_sub"] = torch.from_numpy(z["y_sub"])
        out["y_shape"] = tuple(int(v) for v in z["y_shape"])
    return out


def head_inputs(sarpost, meta):
    imgsz = meta["imgsz"] if isinstance(meta["imgsz"], int) else tuple(meta["imgsz"])
    shapes = sarpost.synth.level_shapes(imgsz, meta["strides"])
    levels = sarpost.synth.head_outputs(meta["batch"], shapes, meta["nc"], meta["ed"], meta["sc"], seed=meta["seed"])
    return shapes, levels


def canon(rows: torch.Tensor) -> torch.Tensor:
    """Canonical order for comparing detections whose equal-score rows may be permuted: descending
    score, ties ordered by the remaining columns.  The reference cuts to max_nms with an UNSTABLE argsort
    (ops.py:286), so among exactly equal scores its row order is torch-build defined; this repo resolves
    ties to the lower source index (SURVEY.md §7 hard part 2).  Kept SETS must still agree."""
    r = rows.detach().cpu().numpy()
    if r.shape[0] == 0:
        return rows.detach().cpu()
    keys = [r[:, c] for c in range(r.shape[1] - 1, -1, -1) if c != 4] + [-r[:, 4]]
    return torch.from_numpy(r[np.lexsort(keys)])


def synth_labels(seed, nc, counts=(5, 0, 150)):
    """Same generator as tests/golden/make_golden.py::synth_labels."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in counts:
        cls = torch.randint(0, nc, (n, 1), generator=g).float()
        xy = torch.rand(n, 2, generator=g) * 600
        wh = 10 + torch.rand(n, 2, generator=g) * 80
        out.append(torch.cat((cls, xy, wh), 1))
    return out


def nms_case_inputs(sarpost, meta):
    y = sarpost.synth.decoded_prediction(**meta["gen"])
    kw = dict(meta["kw"])
    if "labels_seed" in meta:
        kw["labels"] = synth_labels(meta["labels_seed"], meta["gen"]["nc"])
    return y, kw


def load_state_case():
    """tests/golden/state_jde_mlp.npz: the live reference's whole JDE head (state_predictor on every anchor)."""
    z = np.load(os.path.join(GOLDEN_DIR, "state_jde_mlp.npz"))
    meta = json.loads(str(z["meta"]))
    t = {k: torch.from_numpy(z[k]) for k in ("w1", "b1", "w2", "b2", "y", "rows")}
    t["levels"] = [torch.from_numpy(z[f"level{i}"]) for i in range(len(meta["strides"]))]
    t["rows"] = list(torch.split(t["rows"], z["counts"].tolist()))
    t["meta"] = meta
    return t


def load_tail_case():
    """tests/golden/predict_tail_jde.npz: boxes (n, 7) / embeds (n, E) per image from the live, unmodified
    JDEPredictor.postprocess (models/yolo/jde/predict.py:29-78) on the state case's prediction."""
    z = np.load(os.path.join(GOLDEN_DIR, "predict_tail_jde.npz"))
    meta = json.loads(str(z["meta"]))
    counts = z["counts"].tolist()
    return dict(meta=meta, counts=counts, boxes=list(torch.split(torch.from_numpy(z["boxes"]), counts)),
                embeds=list(torch.split(torch.from_numpy(z["embeds"]), counts)))
