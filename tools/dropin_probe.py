"""Time the API-exact drop-in pair: decode (Detect/JDE._inference) then non_max_suppression(y)."""
import sys, torch; sys.path.insert(0, ".")
import sarpost
from sarpost import synth
from bench import WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
imgsz, strides, nc, ed, sc, bs, kw, cls_mean, desc = WORKLOADS[wl]
if len(sys.argv) > 2: bs = int(sys.argv[2])
dev = torch.device("cuda:0")
spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
lv = synth.head_outputs(bs, synth.level_shapes(imgsz, strides), nc, ed, sc, cls_mean=cls_mean, seed=1, device=dev)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
y = sarpost.decode(lv, spec)
A = y.shape[2]
td = t(lambda: sarpost.decode(lv, spec))
bytes_d = bs * A * (spec.no + 4 + nc + spec.nm) * 4
tn = t(lambda: sarpost.non_max_suppression(y, nc=nc, **kw))
tf = t(lambda: sarpost.postprocess_fused(lv, spec, **kw))
print(f"{wl} B={bs}: decode {td*1e3:.0f} us ({bytes_d/td/1e6:.0f} GB/s), nms(y) {tn*1e3:.0f} us, decode+nms {1e3*(td+tn):.0f} us = {bs/(td+tn)*1e3:.0f} img/s; fused {tf*1e3:.0f} us = {bs/tf*1e3:.0f} img/s")
