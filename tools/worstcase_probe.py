"""Adversarial inputs for the NMS kernel: time and check against the oracle."""
import sys, time, torch; sys.path.insert(0, ".")
import sarpost
from oracle import postprocess_ref as R
dev = torch.device("cuda:0")
def run(name, y, **kw):
    yd = y.to(dev)
    for _ in range(2): rows = sarpost.non_max_suppression(yd, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): rows = sarpost.non_max_suppression(yd, **kw)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    t0 = time.perf_counter(); ref = R.non_max_suppression_ref(y, **kw); cpu = time.perf_counter() - t0
    ok = all(torch.equal(a.cpu(), b) for a, b in zip(rows, ref))
    print(f"{name:46s} kept {[r.shape[0] for r in rows]} gpu {ms:8.3f} ms  oracle(early-stop) {cpu*1e3:8.1f} ms  equal={ok}")
g = torch.Generator().manual_seed(0)
A = 136000
# 1. every box identical: 1 kept, everything else suppressed by it
y = torch.zeros(1, 5, A); y[0, :4] = torch.tensor([300., 300., 80., 80.])[:, None]; y[0, 4] = torch.rand(A, generator=g) * 0.9 + 0.05
run("identical boxes, 136k candidates", y, conf_thres=0.001, iou_thres=0.7)
# 2. 299 far-apart clusters, each with hundreds of near-duplicates: kept saturates at 299, all 30000 walked
k = 299
cx = (torch.arange(k) % 20) * 60.0 + 40; cy = (torch.arange(k) // 20) * 80.0 + 40
pick = torch.randint(0, k, (A,), generator=g)
y = torch.zeros(1, 5, A); y[0, 0] = cx[pick] + torch.randn(A, generator=g) * 0.5; y[0, 1] = cy[pick] + torch.randn(A, generator=g) * 0.5
y[0, 2:4] = 30.0; y[0, 4] = torch.rand(A, generator=g) * 0.9 + 0.05
run("299 clusters of duplicates (walks all 30000)", y, conf_thres=0.001, iou_thres=0.5)
# 3. all scores equal: one oversized bucket -> global radix fallback over 136k, then kept list
y = sarpost.synth.decoded_prediction(1, A, 1, 0, seed=3); y[:, 4] = 0.37
run("all scores equal (radix fallback, 136k)", y, conf_thres=0.001, iou_thres=0.7)
# 4. heavy overlap everywhere (cfg1 literal fixture: every box ~ same size, centres on the grid)
y = sarpost.synth.decoded_prediction(1, 8400, 1, 0, seed=4); y[0, 2:4] = 120.0; y[0, 4] = 0.45 + torch.rand(8400, generator=g) * 0.1
run("8400 big overlapping boxes (cfg1 literal-like)", y, conf_thres=0.25, iou_thres=0.7)
