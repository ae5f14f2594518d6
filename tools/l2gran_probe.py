"""Does cudaLimitMaxL2FetchGranularity change the DRAM traffic / time of the strided extras gather (k5_gather)?
   python tools/l2gran_probe.py [granularity 0|32|64|128] [workload cfg3|cfg2]
   (run under `ncu --metrics dram__bytes_read.sum -k regex:k5_gather` to see the traffic)"""
import sys, ctypes, torch; sys.path.insert(0, ".")
gran = int(sys.argv[1]) if len(sys.argv) > 1 else 0
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
rt = ctypes.CDLL("libcudart.so.12")
cudaLimitMaxL2FetchGranularity = 0x05
v = ctypes.c_size_t()
if gran:  # before the context does any work
    print("set rc", rt.cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, ctypes.c_size_t(gran)))
import sarpost
from sarpost import synth
from bench import WORKLOADS
dev = torch.device("cuda:0"); torch.cuda.init(); torch.zeros(1, device=dev)
print("get rc", rt.cudaDeviceGetLimit(ctypes.byref(v), cudaLimitMaxL2FetchGranularity), "value", v.value)
imgsz, strides, nc, ed, sc, bs, kw, cls_mean, desc = WORKLOADS[wl]
spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
lv = synth.head_outputs(bs, synth.level_shapes(imgsz, strides), nc, ed, sc, cls_mean=cls_mean, seed=1, device=dev)
for _ in range(5): sarpost.postprocess_fused(lv, spec, return_padded=True, **kw)
sarpost.ops.stage_timing(True)
t = [0, 0, 0, 0]
for _ in range(20):
    sarpost.postprocess_fused(lv, spec, return_padded=True, **kw)
    for i, x in enumerate(sarpost.ops.stage_times()): t[i] += x / 20
print(wl, "gran", gran, "stages ms", [round(x, 4) for x in t])
