"""Randomised parity fuzz: CUDA non_max_suppression / fused path vs the oracle on many small random configurations.
   python tools/fuzz_parity.py [n_cases] [seed]"""
import random, sys, time
import torch
sys.path.insert(0, ".")
import sarpost
from oracle import postprocess_ref as R



def run(n_cases: int, seed: int) -> int:
  rng = random.Random(seed)
  dev = torch.device("cuda:0")
  bad = 0
  stats = {}
  for case in range(n_cases):
      mode = rng.choice(["decoded", "decoded", "fused", "fused", "merge", "labels", "host", "match", "deep"])
      nc = rng.choice([1, 1, 2, 3, 6, 9, 17])
      kw = dict(conf_thres=rng.choice([0.0, 0.001, 0.05, 0.25, 0.5, 0.9]), iou_thres=rng.choice([0.0, 0.3, 0.45, 0.6, 0.7, 0.95, 1.0]),
                agnostic=rng.random() < 0.3, multi_label=rng.random() < 0.4, max_det=rng.choice([1, 7, 100, 300, 1000]),
                max_nms=rng.choice([5, 300, 2000, 30000]), max_wh=rng.choice([7680, 7680, 0, 100.5]))
      if rng.random() < 0.25:
          kw["classes"] = rng.sample(range(nc), k=rng.randint(0, nc))
      cs = rng.randint(0, 10 ** 6)
      try:
          if mode == "match":
              sys.path.insert(0, "tests")
              from test_oracle import _match_case
              nb = rng.choice([1, 4])
              md, mg = rng.choice([10, 300]), rng.choice([1, 64, 200])
              iouv = torch.linspace(0.5, 0.95, 10)
              cases = [_match_case(cs + j, n_det=rng.randint(0, md), n_gt=rng.randint(0, mg), nc=nc) for j in range(nb)]
              dets = torch.zeros(nb, md, 7); gtb = torch.zeros(nb, mg, 4); gtc = torch.zeros(nb, mg)
              for j, (d_, g_, c_) in enumerate(cases):
                  dets[j, : d_.shape[0], :6] = d_; gtb[j, : g_.shape[0]] = g_; gtc[j, : g_.shape[0]] = c_
              ti = rng.choice([0, 5])
              correct, matched = sarpost.match_predictions(dets.to(dev), torch.tensor([c[0].shape[0] for c in cases]), gtb.to(dev), gtc.to(dev),
                                                           torch.tensor([c[1].shape[0] for c in cases]), iouv.tolist(), tag_threshold_index=ti)
              rows, idx, ref_rows, ref_idx = [], [], [], []
              for j, (d_, g_, c_) in enumerate(cases):
                  n_ = d_.shape[0]
                  if n_ == 0 or g_.shape[0] == 0:
                      rc_, rm_ = torch.zeros(n_, 10), torch.full((n_,), -1)
                  else:
                      rc_, rm_ = R.match_predictions_ref(d_[:, 5], c_, R.box_iou_ref(g_, d_[:, :4]), iouv, tag_thr=iouv[ti].item())
                      rc_, rm_ = rc_.float(), rm_.long()
                  rows.append(correct[j, :n_].float()); idx.append(matched[j, :n_].long() + 1)
                  ref_rows.append(rc_); ref_idx.append(torch.stack((rm_ + 1, torch.zeros_like(rm_)), 1))
                  ok_pad = not bool(correct[j, n_:].any())
                  if not ok_pad:
                      raise RuntimeError("padding rows marked correct")
              nc = 1
          elif mode == "merge":
              tpf, d, nf, rl = rng.choice([1, 3, 12, 48]), rng.choice([1, 20, 300]), rng.choice([1, 2, 5]), rng.choice([6, 9])
              g = torch.Generator().manual_seed(cs)
              dets = torch.zeros(nf * tpf, d, rl)
              cnt = torch.randint(0, d + 1, (nf * tpf,), generator=g, dtype=torch.int32)
              xy = torch.rand(nf * tpf, d, 2, generator=g) * 300
              dets[..., 0:2] = xy
              dets[..., 2:4] = xy + 10 + torch.rand(nf * tpf, d, 2, generator=g) * 100
              sc_ = torch.rand(nf * tpf, d, generator=g)
              if rng.random() < 0.3:
                  sc_ = (sc_ * 8).floor() / 8 + 0.05
              dets[..., 4] = sc_
              dets[..., 5] = torch.randint(0, nc, (nf * tpf, d), generator=g).float()
              if rl > 6:
                  dets[..., 6:] = torch.randn(nf * tpf, d, rl - 6, generator=g)
              org = (torch.rand(tpf, 2, generator=g) * 400).floor().repeat(nf, 1)
              mk = dict(iou_thres=kw["iou_thres"], agnostic=kw["agnostic"], max_det=kw["max_det"], max_nms=kw["max_nms"], max_wh=kw["max_wh"])
              rows, idx = sarpost.merge_tiles(dets.to(dev), cnt.to(dev), org.to(dev), tpf, return_index=True, **mk)
              ref_rows, ref_idx = [], []
              for f in range(nf):
                  cand, src = [], []
                  for t in range(tpf):
                      k = f * tpf + t
                      n_ = int(cnt[k])
                      r_ = dets[k, :n_].clone()
                      r_[:, 0] += org[k, 0]; r_[:, 2] += org[k, 0]; r_[:, 1] += org[k, 1]; r_[:, 3] += org[k, 1]
                      cand.append(r_); src.append(torch.arange(n_) + t * d)
                  x = torch.cat(cand); srcc = torch.cat(src)
                  if x.shape[0] > mk["max_nms"]:
                      o = x[:, 4].argsort(descending=True, stable=True)[: mk["max_nms"]]
                      x, srcc = x[o], srcc[o]
                  c_ = x[:, 5:6] * (0 if mk["agnostic"] else mk["max_wh"])
                  keep = R.nms_ref(x[:, :4] + c_, x[:, 4], mk["iou_thres"])[: mk["max_det"]]
                  ref_rows.append(x[keep]); ref_idx.append(torch.stack((srcc[keep], torch.zeros_like(srcc[keep])), 1))
              nc = 1  # kept_index of the merge is tile*dets_per_tile + row
          elif mode == "labels":
              bs, na, nm = rng.choice([1, 3]), rng.choice([129, 1000, 4000]), rng.choice([0, 3])
              y = sarpost.synth.decoded_prediction(bs, na, nc, nm, seed=cs, clustered=rng.random() < 0.5)
              g = torch.Generator().manual_seed(cs + 1)
              labels = []
              for _ in range(bs):
                  n_ = rng.choice([0, 1, 40, 200])
                  labels.append(torch.cat((torch.randint(0, nc, (n_, 1), generator=g).float(), torch.rand(n_, 2, generator=g) * 600,
                                           5 + torch.rand(n_, 2, generator=g) * 90), 1))
              rows, idx = sarpost.non_max_suppression(y.to(dev), nc=nc, labels=[lb.to(dev) for lb in labels], return_index=True, **kw)
              ref_rows, ri_ = R.non_max_suppression_ref(y, nc=nc, labels=labels, return_index=True, **kw)
              ref_idx = [torch.stack((torch.where(r[:, 0] >= 0, r[:, 0], na + (-1 - r[:, 0])), r[:, 1]), 1) for r in ri_]
          elif mode == "host":
              strides = rng.choice([(8, 16, 32), (16,)])
              imgsz = rng.choice([64, 160])
              bs, ed, sc = rng.choice([1, 5]), rng.choice([0, 4]), rng.choice([0, 6])
              shapes = sarpost.synth.level_shapes(imgsz, strides)
              lv = sarpost.synth.head_outputs(bs, shapes, nc, ed, sc, seed=cs, cls_mean=rng.choice([-4.0, 1.0]))
              spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
              lvd = lv
              rows, idx = sarpost.postprocess_host(lv, spec, return_index=True, **kw)
              y = sarpost.decode([x.to(dev) for x in lv], spec).cpu()
              ref_rows, ref_idx = R.non_max_suppression_ref(y, nc=nc, return_index=True, **kw)
          elif mode == "deep":
              # suppression-heavy inputs that make the NMS kernel walk many chunks: few far-apart clusters of near-duplicate boxes
              # (kept list saturates below max_det, everything else is suppressed), optionally crossing the max_nms rank cut
              bs, na = rng.choice([1, 2, 3]), rng.choice([3000, 12000, 40000, 140000])
              k = rng.choice([3, 40, 299, 1000])
              g = torch.Generator().manual_seed(cs)
              y = torch.zeros(bs, 4 + nc, na)
              cx = (torch.arange(k) % 40) * 60.0 + 40
              cy = (torch.arange(k) // 40) * 80.0 + 40
              pick = torch.randint(0, k, (bs, na), generator=g)
              jit = rng.choice([0.5, 4.0, 12.0])
              y[:, 0] = cx[pick] + torch.randn(bs, na, generator=g) * jit
              y[:, 1] = cy[pick] + torch.randn(bs, na, generator=g) * jit
              y[:, 2:4] = 30.0
              y[:, 4:4 + nc] = torch.rand(bs, nc, na, generator=g) * 0.9 + 0.05
              if rng.random() < 0.3:
                  y[:, 4:4 + nc] = (y[:, 4:4 + nc] * 64).floor() / 64 + 1 / 128
              kw["conf_thres"] = rng.choice([0.001, 0.3])
              kw["iou_thres"] = rng.choice([0.3, 0.5, 0.7])
              kw["max_nms"] = rng.choice([300, 2000, 30000, 30000])
              kw["max_det"] = rng.choice([7, 100, 300, 300, 1000])
              kw.pop("classes", None)
              rows, idx = sarpost.non_max_suppression(y.to(dev), nc=nc, return_index=True, **kw)
              ref_rows, ref_idx = R.non_max_suppression_ref(y, nc=nc, return_index=True, **kw)
          elif mode == "decoded":
              bs, na, nm = rng.choice([1, 2, 5, 40]), rng.choice([1, 31, 128, 129, 1000, 5000, 20000]), rng.choice([0, 0, 3, 40])
              y = sarpost.synth.decoded_prediction(bs, na, nc, nm, seed=cs, clustered=rng.random() < 0.5, score_pow=rng.choice([1.0, 2.0, 4.0]))
              if rng.random() < 0.3:  # quantised scores: heavy ties
                  q = rng.choice([4, 16, 256])
                  y[:, 4:4 + nc] = (y[:, 4:4 + nc] * q).floor() / q + 0.5 / q
              rows, idx = sarpost.non_max_suppression(y.to(dev), nc=nc, return_index=True, **kw)
              ref_rows, ref_idx = R.non_max_suppression_ref(y, nc=nc, return_index=True, **kw)
          else:
              strides = rng.choice([(8, 16, 32), (4, 8, 16, 32), (16,)])
              imgsz = rng.choice([64, 96, (88, 120), 160, 320])
              bs, ed, sc = rng.choice([1, 3, 38]), rng.choice([0, 4]), rng.choice([0, 6])
              shapes = sarpost.synth.level_shapes(imgsz, strides)
              lv = sarpost.synth.head_outputs(bs, shapes, nc, ed, sc, seed=cs, cls_mean=rng.choice([-4.0, -1.0, 1.0]), blobs=rng.choice([0, 3]))
              spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
              lvd = [x.to(dev) for x in lv]
              if rng.random() < 0.3:
                  lvd = [x.half() for x in lvd]
              scale_to = None
              if rng.random() < 0.3:
                  ih, iw = (imgsz, imgsz) if isinstance(imgsz, int) else imgsz
                  scale_to = ((ih, iw), [(rng.randint(20, 900), rng.randint(20, 900), 3) for _ in range(bs)])
              mlp = None
              if sc and ed and rng.random() < 0.4:  # deferred state head: levels without the state channels + MLP on kept rows
                  g = torch.Generator().manual_seed(cs + 7)
                  hd = rng.choice([2, 8, 40])
                  wts = (torch.randn(hd, ed, generator=g), torch.randn(hd, generator=g), torch.randn(sc, hd, generator=g), torch.randn(sc, generator=g))
                  mlp = sarpost.StateMLP.from_tensors(*wts, device=dev)
                  lvd = [x[:, : 64 + nc + ed].contiguous() for x in lvd]
              lv_cat = lvd
              via = "plain"
              if mlp is None:  # the same call through the other host-side forms: split layout, a prepared plan, the batch pipeline
                  if rng.random() < 0.4:
                      lvd = sarpost.split_levels(lvd, spec, emb_channels_last=rng.random() < 0.5)
                  via = rng.choice(["plain", "plain", "plan", "pipeline"])
              if via == "plan":  # prepared from OTHER tensors of the same geometry, so every address is new at run time
                  warm = [tuple(None if t is None else t.clone(memory_format=torch.preserve_format) for t in x) if isinstance(x, tuple) else x.clone() for x in lvd]
                  plan = sarpost.FusedPlan(warm, spec, **kw)
                  out_, cnt_, kidx_ = plan(lvd, return_index=True, scale_to=scale_to)
                  n_ = cnt_.tolist()
                  rows, idx = [out_[b_, :n_[b_]] for b_ in range(bs)], [kidx_[b_, :n_[b_]] for b_ in range(bs)]
                  plan.close()
              elif via == "pipeline":
                  pl = sarpost.Pipeline(dev, depth=rng.choice([1, 2]))
                  pl.submit(lvd, spec, **kw)  # a batch in front, so the measured one runs in the chained steady state
                  out_, cnt_, kidx_ = pl.submit(lvd, spec, return_index=True, scale_to=scale_to, **kw)
                  pl.wait()
                  n_ = cnt_.tolist()
                  rows, idx = [out_[b_, :n_[b_]] for b_ in range(bs)], [kidx_[b_, :n_[b_]] for b_ in range(bs)]
                  pl.close()
              else:
                  rows, idx = sarpost.postprocess_fused(lvd, spec, return_index=True, scale_to=scale_to, state_mlp=mlp, **kw)
              lvd = lv_cat
              if mlp is not None:
                  spec_ns = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=0)
                  y = sarpost.decode([x.float() for x in lvd], spec_ns).cpu()
                  ref_rows, ref_idx = R.non_max_suppression_ref(y, nc=nc, return_index=True, **kw)
                  for b_ in range(len(rows)):  # state columns: tolerance (fp32 summation order), then compared as equal
                      st_ref = R.state_head_ref(ref_rows[b_][:, 6:6 + ed], *wts)
                      got = rows[b_].cpu()
                      if got.shape[0] == st_ref.shape[0] and torch.allclose(got[:, 6 + ed:], st_ref, rtol=0, atol=2e-5):
                          ref_rows[b_] = torch.cat((ref_rows[b_], got[:, 6 + ed:]), 1)
                      else:
                          ref_rows[b_] = torch.cat((ref_rows[b_], st_ref), 1)
              else:
                  y = sarpost.decode([x.float() for x in lvd], spec).cpu()
                  ref_rows, ref_idx = R.non_max_suppression_ref(y, nc=nc, return_index=True, **kw)
              if scale_to is not None:
                  for r_, o_ in zip(ref_rows, scale_to[1]):
                      r_[:, :4] = R.scale_boxes_ref(scale_to[0], r_[:, :4], o_)
          ok = len(rows) == len(ref_rows)
          st = stats.setdefault(mode, [0, 0])
          st[0] += 1
          st[1] += sum(int(r.shape[0]) for r in ref_rows)
          for r, i, rr, ri in zip(rows, idx, ref_rows, ref_idx):
              ok = ok and tuple(r.shape) == tuple(rr.shape) and torch.equal(r.cpu(), rr) and torch.equal(i.cpu().long() // nc, ri[:, 0]) \
                   and torch.equal(i.cpu().long() % nc, ri[:, 1])
      except Exception as e:  # noqa: BLE001
          ok = False
          print("EXC", repr(e))
      if not ok:
          bad += 1
          extra = {}
          if mode in ("decoded", "labels"):
              extra = dict(bs=bs, na=na, nm=nm)
          elif mode == "deep":
              extra = dict(bs=bs, na=na, k=k, jit=jit)
          elif mode in ("fused", "host"):
              extra = dict(strides=strides, imgsz=imgsz, bs=bs, ed=ed, sc=sc, half=lvd[0].dtype == torch.float16)
          print("MISMATCH case", case, mode, "nc", nc, kw, "seed", cs, extra)
          try:
              for b_, (r, rr, i, ri) in enumerate(zip(rows, ref_rows, idx, ref_idx)):
                  if tuple(r.shape) != tuple(rr.shape):
                      print("  image", b_, "shape", tuple(r.shape), "vs", tuple(rr.shape))
                      continue
                  d = (r.cpu() != rr).any(1).nonzero().flatten()
                  if d.numel():
                      k = int(d[0])
                      print("  image", b_, "first differing row", k, "ours", r[k, :6].tolist(), int(i[k]), "ref", rr[k, :6].tolist(), ri[k].tolist())
          except Exception as e2:  # noqa: BLE001
              print("  (detail failed)", repr(e2))
  print("cases / oracle rows compared per mode:", {k: tuple(v) for k, v in sorted(stats.items())})
  return bad


if __name__ == "__main__":
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t0 = time.time()
    bad = run(n_cases, seed)
    print(f"{n_cases} cases, {bad} mismatches, {time.time() - t0:.1f} s")
    sys.exit(1 if bad else 0)
