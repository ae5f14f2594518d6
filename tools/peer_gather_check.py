"""torchrun --nproc-per-node N tools/peer_gather_check.py — fused gather+exchange (K5 peer stores) vs NCCL all-gather."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sarpost
from sarpost import synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
strides = (8, 16, 32); spec = sarpost.HeadSpec(nc=2, strides=strides, embed_dim=8, state_classes=6)
per = 6
lv = [x.to(dev) for x in synth.head_outputs(per, synth.level_shapes(320, strides), 2, 8, 6, seed=100 + rank)]
kw = dict(conf_thres=0.25, iou_thres=0.7, max_det=60)
ok = True
for with_extras in (True, False):
    row_len = 6 + (spec.nm if with_extras else 0)
    out, counts = sarpost.postprocess_fused(lv, spec, return_padded=True, with_extras=with_extras, **kw)
    ref_rows, ref_cnt = sarpost.dist.allgather_detections(out, counts)
    peer = sarpost.dist.PeerGatherBuffer(per, kw["max_det"], row_len, dev)
    for it in range(3):
        rows, cnt = sarpost.postprocess_fused(lv, spec, with_extras=with_extras, peer_out=peer.next(), **kw)
        peer.barrier()
        torch.cuda.synchronize()
        same = torch.equal(cnt, ref_cnt)
        for i, n in enumerate(ref_cnt.tolist()):
            same = same and torch.equal(rows[i, :n], ref_rows[i, :n])
        ok = ok and same
t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("peer gather == nccl all-gather:", bool(t.item()))
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
