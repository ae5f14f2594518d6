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
  for case in range(n_cases):
      mode = rng.choice(["decoded", "decoded", "fused"])
      nc = rng.choice([1, 1, 2, 3, 6, 9, 17])
      kw = dict(conf_thres=rng.choice([0.0, 0.001, 0.05, 0.25, 0.5, 0.9]), iou_thres=rng.choice([0.0, 0.3, 0.45, 0.6, 0.7, 0.95, 1.0]),
                agnostic=rng.random() < 0.3, multi_label=rng.random() < 0.4, max_det=rng.choice([1, 7, 100, 300, 1000]),
                max_nms=rng.choice([5, 300, 2000, 30000]), max_wh=rng.choice([7680, 7680, 0, 100.5]))
      if rng.random() < 0.25:
          kw["classes"] = rng.sample(range(nc), k=rng.randint(0, nc))
      cs = rng.randint(0, 10 ** 6)
      try:
          if mode == "decoded":
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
              rows, idx = sarpost.postprocess_fused(lvd, spec, return_index=True, **kw)
              y = sarpost.decode([x.float() for x in lvd], spec).cpu()
              ref_rows, ref_idx = R.non_max_suppression_ref(y, nc=nc, return_index=True, **kw)
          ok = len(rows) == len(ref_rows)
          for r, i, rr, ri in zip(rows, idx, ref_rows, ref_idx):
              ok = ok and tuple(r.shape) == tuple(rr.shape) and torch.equal(r.cpu(), rr) and torch.equal(i.cpu().long() // nc, ri[:, 0]) \
                   and torch.equal(i.cpu().long() % nc, ri[:, 1])
      except Exception as e:  # noqa: BLE001
          ok = False
          print("EXC", repr(e))
      if not ok:
          bad += 1
          extra = dict(bs=bs, na=na, nm=nm) if mode == "decoded" else dict(strides=strides, imgsz=imgsz, bs=bs, ed=ed, sc=sc,
                                                                            half=lvd[0].dtype == torch.float16)
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
  return bad


if __name__ == "__main__":
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t0 = time.time()
    bad = run(n_cases, seed)
    print(f"{n_cases} cases, {bad} mismatches, {time.time() - t0:.1f} s")
    sys.exit(1 if bad else 0)
