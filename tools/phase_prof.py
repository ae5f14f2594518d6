import sys, ctypes as C, torch
sys.path.insert(0,'/root/repo')
import sarpost
from sarpost import synth
from bench import WORKLOADS
wl=sys.argv[1] if len(sys.argv)>1 else 'cfg3'
imgsz, strides, nc, ed, sc, bs, kw, cls_mean, desc = WORKLOADS[wl]
dev=torch.device('cuda:0')
spec=sarpost.HeadSpec(nc=nc,strides=strides,embed_dim=ed,state_classes=sc)
blobs=int(sys.argv[2]) if len(sys.argv)>2 else 0
levels=synth.head_outputs(bs, synth.level_shapes(imgsz,strides), nc, ed, sc, cls_mean=cls_mean, seed=3000, device=dev, blobs=blobs)
lib=sarpost._lib.lib
f=lib.sarpost_debug_phase_detail; f.restype=C.c_int32; f.argtypes=[C.c_void_p,C.c_int32]
for _ in range(3): sarpost.postprocess_fused(levels,spec,return_padded=True,**kw)
buf=(C.c_ulonglong*48)(); f(buf,1)
N=20
for _ in range(N): sarpost.postprocess_fused(levels,spec,return_padded=True,**kw)
f(buf,1)
names={0:'prologue(hist scan)',1:'collect',2:'share phase1',3:'deliver+barrier1',4:'sort',5:'radix fallback',6:'zoom histogram',9:'tail+replicate+barrier2',8:'publish',7:'fused gather',
       10:'ps: load',13:'ps: pair round',11:'ps: sweep (warp 0)',14:'ps: append+barriers'}
tot=sum(buf[i] for i in range(16))
print(wl, 'blobs', blobs)
for i,n in names.items(): print(f"{n:26s} {buf[i]/N:10.0f} cyc  {buf[i]/tot*100:5.1f}%   visits/launch {buf[16+i]/N:6.2f}  mean {buf[i]/max(buf[16+i],1):8.0f}  max {buf[32+i]:8d}")
print("total cyc/launch", tot/N)
