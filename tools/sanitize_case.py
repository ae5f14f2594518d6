"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck): every kernel, tiny shapes."""
import sys, torch; sys.path.insert(0, ".")
import sarpost
from sarpost import synth
dev = torch.device("cuda:0")
strides = (8, 16, 32)
for (imgsz, nc, ed, sc, kw) in [
    (160, 1, 8, 6, dict(conf_thres=0.25, iou_thres=0.7)),
    ((88, 120), 3, 0, 0, dict(conf_thres=0.001, iou_thres=0.7, multi_label=True, max_nms=500)),
]:
    shapes = synth.level_shapes(imgsz, strides)
    spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
    lv = [x.to(dev) for x in synth.head_outputs(2, shapes, nc, ed, sc, seed=5, blobs=3)]
    rows = sarpost.postprocess_fused(lv, spec, **kw)
    y = sarpost.decode(lv, spec)
    rows2 = sarpost.non_max_suppression(y, nc=nc, **kw)
    rows3 = sarpost.postprocess_fused([x.half() for x in lv], spec, **kw)
    print([r.shape[0] for r in rows], [r.shape[0] for r in rows2], [r.shape[0] for r in rows3])
# ties -> oversized bucket fallback (radix sort in global memory)
y = synth.decoded_prediction(1, 3000, 1, 0, seed=5); y[:, 4] = 0.5
print(sarpost.non_max_suppression(y.to(dev), conf_thres=0.1, iou_thres=0.7)[0].shape)
# merge
org = sarpost.dist.sahi_grid(900, 600, 320, 0.2).to(dev)
d = torch.rand(org.shape[0], 20, 7, device=dev) * 100; d[..., 2:4] += d[..., 0:2]; d[..., 5] = 0
print(sarpost.merge_tiles(d, torch.full((org.shape[0],), 20, dtype=torch.int32, device=dev), org, org.shape[0], iou_thres=0.5)[0].shape)
torch.cuda.synchronize()
print("done")
