"""Where does the time go inside the software pipeline?  Per-kernel CUDA-event durations (library stage timing, recorded
on each batch's own stream, so a kernel that waits for SMs or shares them shows up as a longer stage) + the step time."""
import sys, torch
sys.path.insert(0, ".")
import sarpost
from bench import WORKLOADS
imgsz, strides, nc, ed, sc, bs, kw, cls_mean, desc = WORKLOADS["cfg3"]
dev = torch.device("cuda:0")
spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
shapes = sarpost.synth.level_shapes(imgsz, strides)
sets = [sarpost.synth.head_outputs(bs, shapes, nc, ed, sc, cls_mean=cls_mean, seed=3000 + i, device=dev) for i in range(2)]
if len(sys.argv) > 2 and sys.argv[2] == "split":
    sets = [sarpost.split_levels(s, spec) for s in sets]
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
outs = [(torch.empty((bs, 300, 6 + spec.nm), device=dev), torch.empty((bs,), dtype=torch.int32, device=dev)) for _ in range(4)]
pl = sarpost.Pipeline(dev, depth=depth)
for i in range(6): pl.submit(sets[i % 2], spec, out=outs[i % 4], **kw)
pl.wait(); torch.cuda.synchronize()
N = 200
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(N): pl.submit(sets[i % 2], spec, out=outs[i % 4], **kw)
pl.wait(); e1.record(); torch.cuda.synchronize()
print(f"depth {depth}: {e0.elapsed_time(e1) / N * 1e3:.1f} us/step untimed")
sarpost.ops.stage_timing(True, accumulate=True)
e0.record()
for i in range(N): pl.submit(sets[i % 2], spec, out=outs[i % 4], **kw)
pl.wait(); e1.record(); torch.cuda.synchronize()
st = sarpost.ops.stage_times(); sarpost.ops.stage_timing(False)
print(f"depth {depth}: {e0.elapsed_time(e1) / N * 1e3:.1f} us/step with stage events; K1 {st[0]*1e3:.1f}  K4 {st[1]*1e3:.1f}  K5 {st[2]*1e3:.1f}  call {st[3]*1e3:.1f} us")
