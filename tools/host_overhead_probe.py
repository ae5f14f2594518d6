"""Host time per call (time.perf_counter around N enqueue-only calls on an idle device, N small enough that the launch
queue never fills): what the CPU spends per step in the general entry point, a prepared plan and the pipeline.
nsys is not in the image; this is the host-span evidence for the per-call overhead."""
import sys
import time

import torch

sys.path.insert(0, "/root/repo")
import sarpost  # noqa: E402
from sarpost import synth  # noqa: E402
from bench import WORKLOADS  # noqa: E402

dev = torch.device("cuda:0")
N = 100
for wl in sys.argv[1:] or ["cfg1", "cfg5", "cfg3"]:
    imgsz, strides, nc, ed, sc, bs, kw, cls_mean, desc = WORKLOADS[wl]
    spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
    cat = synth.head_outputs(bs, synth.level_shapes(imgsz, strides), nc, ed, sc, cls_mean=cls_mean, seed=3000, device=dev)
    for layout, levels in (("cat", cat), ("split", sarpost.split_levels(cat, spec, emb_channels_last=True))):
        plan = sarpost.FusedPlan(levels, spec, **kw)
        outs = (torch.empty((bs, kw["max_det"], 6 + spec.nm), device=dev), torch.empty((bs,), dtype=torch.int32, device=dev))
        pl = sarpost.Pipeline(dev, depth=2)
        calls = {
            "general (postprocess_fused)": lambda: sarpost.postprocess_fused(levels, spec, return_padded=True, **kw),
            "plan (FusedPlan, out=)": lambda: plan(levels, out=outs),
            "plan (FusedPlan, allocating)": lambda: plan(levels),
            "pipeline.submit(out=)": lambda: pl.submit(levels, spec, out=outs, **kw),
        }
        for name, fn in calls.items():
            for _ in range(5):
                fn()
            pl.wait()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                t0 = time.perf_counter()
                for _ in range(N):
                    fn()
                best = min(best, (time.perf_counter() - t0) / N)
                pl.wait()
                torch.cuda.synchronize()
            print(f"{wl:5s} {layout:5s} {name:32s} {best * 1e6:7.1f} us/call (host, enqueue only)")
        pl.close()
        plan.close()
