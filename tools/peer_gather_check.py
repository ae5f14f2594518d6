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
# rows beyond counts stay untouched in every rank's buffer (sentinel), odd max_det (row-by-row path), low-confidence run
# with partly filled blocks of rows
for max_det, conf in ((61, 0.25), (60, 0.6), (8, 0.25), (300, 0.05)):
    kw2 = dict(conf_thres=conf, iou_thres=0.7, max_det=max_det)
    out, counts = sarpost.postprocess_fused(lv, spec, return_padded=True, with_extras=False, **kw2)
    ref_rows, ref_cnt = sarpost.dist.allgather_detections(out, counts)
    peer = sarpost.dist.PeerGatherBuffer(per, max_det, 6, dev)
    for r_, c_, _, _ in peer._bufs:
        r_.fill_(-7.0)
    dist.barrier()
    torch.cuda.synchronize()
    rows, cnt = sarpost.postprocess_fused(lv, spec, with_extras=False, peer_out=peer.next(), **kw2)
    peer.barrier()
    torch.cuda.synchronize()
    same = torch.equal(cnt, ref_cnt)
    for i, n in enumerate(ref_cnt.tolist()):
        same = same and torch.equal(rows[i, :n], ref_rows[i, :n]) and bool((rows[i, n:] == -7.0).all())
    ok = ok and same
    dist.barrier()
# cross-tile merge with the merged frames exchanged by the merge's own gather kernel (frames sharded over ranks)
g = torch.Generator().manual_seed(77 + rank)
nf, tpf, d = 2, 3, 40
dets = torch.zeros(nf * tpf, d, 6)
xy = torch.rand(nf * tpf, d, 2, generator=g) * 150
dets[..., :2], dets[..., 2:4] = xy, xy + torch.rand(nf * tpf, d, 2, generator=g) * 50 + 4
dets[..., 4] = torch.rand(nf * tpf, d, generator=g)
dets[..., 5] = torch.randint(0, 2, (nf * tpf, d), generator=g).float()
dets = dets.to(dev)
dcnt = torch.tensor([40, 13, 0, 40, 1, 27], dtype=torch.int32, device=dev)
org = torch.tensor([[0.0, 0.0], [90.0, 0.0], [0.0, 90.0]] * nf, device=dev)
for max_det in (50, 33):
    m_out, m_cnt = sarpost.merge_tiles(dets, dcnt, org, tpf, iou_thres=0.5, max_det=max_det, return_padded=True)
    ref_rows, ref_cnt = sarpost.dist.allgather_detections(m_out, m_cnt)
    peer = sarpost.dist.PeerGatherBuffer(nf, max_det, 6, dev)
    for r_, c_, _, _ in peer._bufs:
        r_.fill_(-7.0)
    dist.barrier()
    torch.cuda.synchronize()
    rows, cnt = sarpost.merge_tiles(dets, dcnt, org, tpf, iou_thres=0.5, max_det=max_det, peer_out=peer.next())
    peer.barrier()
    torch.cuda.synchronize()
    same = torch.equal(cnt, ref_cnt)
    for i, n in enumerate(ref_cnt.tolist()):
        same = same and torch.equal(rows[i, :n], ref_rows[i, :n]) and bool((rows[i, n:] == -7.0).all())
    ok = ok and same
    dist.barrier()
t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("peer gather == nccl all-gather:", bool(t.item()))
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
