"""GPU validator matching vs the reference's per-image numpy loop (oracle port) — §8f row 3 measurement."""
import sys, time, torch; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import sarpost
from oracle import postprocess_ref as R
from test_oracle import _match_case
dev = torch.device("cuda:0")
B, max_det, max_gt = 64, 300, 64
cases = [_match_case(s, n_det=300, n_gt=50) for s in range(B)]
dets = torch.zeros(B, max_det, 6); gtb = torch.zeros(B, max_gt, 4); gtc = torch.zeros(B, max_gt)
for i, (d, g, c) in enumerate(cases):
    dets[i, :300] = d; gtb[i, :50] = g; gtc[i, :50] = c
dn = torch.full((B,), 300, dtype=torch.int32); gn = torch.full((B,), 50, dtype=torch.int32)
args = (dets.to(dev), dn.to(dev), gtb.to(dev), gtc.to(dev), gn.to(dev))
for _ in range(3): sarpost.match_predictions(*args)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): sarpost.match_predictions(*args)
e1.record(); torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / 50
iouv = torch.linspace(0.5, 0.95, 10)
t0 = time.perf_counter()
for d, g, c in cases:
    R.match_predictions_ref(d[:, 5], c, R.box_iou_ref(g, d[:, :4]), iouv)
cpu_ms = (time.perf_counter() - t0) * 1e3
print(f"match_predictions B={B} x 300 dets x 50 labels x 10 thresholds: GPU {gpu_ms*1e3:.1f} us/batch ({B/gpu_ms*1e3:.0f} img/s), "
      f"CPU numpy port {cpu_ms:.1f} ms/batch ({B/cpu_ms*1e3:.0f} img/s)")
