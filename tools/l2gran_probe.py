import sys, ctypes, torch; sys.path.insert(0, ".")
import sarpost
from sarpost import synth
dev = torch.device("cuda:0"); torch.cuda.init(); torch.zeros(1, device=dev)
rt = ctypes.CDLL("libcudart.so.12") if len(sys.argv) < 3 else ctypes.CDLL(sys.argv[2])
gran = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cudaLimitMaxL2FetchGranularity = 0x05
v = ctypes.c_size_t()
print("get rc", rt.cudaDeviceGetLimit(ctypes.byref(v), cudaLimitMaxL2FetchGranularity), "value", v.value)
if gran:
    print("set rc", rt.cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, ctypes.c_size_t(gran)))
    rt.cudaDeviceGetLimit(ctypes.byref(v), cudaLimitMaxL2FetchGranularity); print("now", v.value)
strides = (4, 8, 16, 32); spec = sarpost.HeadSpec(nc=1, strides=strides, embed_dim=256, state_classes=6)
lv = synth.head_outputs(16, synth.level_shapes(1280, strides), 1, 256, 6, seed=1, device=dev)
kw = dict(conf_thres=0.001, iou_thres=0.7)
for _ in range(5): sarpost.postprocess_fused(lv, spec, return_padded=True, **kw)
sarpost.ops.stage_timing(True)
t = [0, 0, 0, 0]
for _ in range(20):
    sarpost.postprocess_fused(lv, spec, return_padded=True, **kw)
    for i, x in enumerate(sarpost.ops.stage_times()): t[i] += x / 20
print("stages ms", [round(x, 4) for x in t])
