"""K1 on fp16 vs fp32 level tensors (cfg3 shape): mean stage times over N calls (library stage events, accumulate mode).
   python tools/fp16_probe.py [half|f32] [N]"""
import sys, torch; sys.path.insert(0, ".")
import sarpost
from sarpost import synth
dev = torch.device("cuda:0")
strides = (4, 8, 16, 32); spec = sarpost.HeadSpec(nc=1, strides=strides, embed_dim=256, state_classes=6)
half = len(sys.argv) < 2 or sys.argv[1] == "half"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
lv = synth.head_outputs(16, synth.level_shapes(1280, strides), 1, 256, 6, seed=1, device=dev)
if half: lv = [x.half() for x in lv]
kw = dict(conf_thres=0.001, iou_thres=0.7)
for _ in range(20): sarpost.postprocess_fused(lv, spec, return_padded=True, **kw)
torch.cuda.synchronize()
sarpost.ops.stage_timing(True, accumulate=True)
for _ in range(n): sarpost.postprocess_fused(lv, spec, return_padded=True, **kw)
t = list(sarpost.ops.stage_times())
sarpost.ops.stage_timing(False)
esz = 2 if half else 4
print("half" if half else "f32", "stages ms:", [round(x, 4) for x in t], "K1 GB/s:", round((16 * 136000 * 65 * esz + 16 * 126000 * 24) / t[0] / 1e6))
