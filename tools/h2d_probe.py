import sys, time, torch
sys.path.insert(0, '/root/repo')
import sarpost
from sarpost import synth
dev = torch.device('cuda:0')
B, no, hw = 16, 327, 320 * 320
x = torch.empty((B, no, hw), dtype=torch.float32, pin_memory=True).normal_()
d_full = torch.empty_like(x, device=dev)
d_part = torch.empty((B, 65, hw), dtype=torch.float32, device=dev)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
s = t(lambda: d_full.copy_(x, non_blocking=True))
print(f"contiguous H2D {x.numel()*4/1e6:.0f} MB: {s*1e3:.2f} ms = {x.numel()*4/s/1e9:.1f} GB/s")
s = t(lambda: d_part.copy_(x[:, :65], non_blocking=True))
print(f"strided (2D) H2D {d_part.numel()*4/1e6:.0f} MB: {s*1e3:.2f} ms = {d_part.numel()*4/s/1e9:.1f} GB/s")
xp = torch.empty((B, 65, hw), dtype=torch.float32, pin_memory=True)
s = t(lambda: d_part.copy_(xp, non_blocking=True))
print(f"contiguous H2D {xp.numel()*4/1e6:.0f} MB: {s*1e3:.2f} ms = {xp.numel()*4/s/1e9:.1f} GB/s")
h = torch.empty((16, 300, 268), dtype=torch.float32, pin_memory=True); dd = torch.empty_like(h, device=dev)
s = t(lambda: h.copy_(dd, non_blocking=True))
print(f"D2H {h.numel()*4/1e6:.1f} MB: {s*1e3:.3f} ms")
