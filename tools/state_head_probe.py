"""Deferred JDE state head on the cfg3 shape: kernel time for B*max_det kept rows vs the reference's
per-anchor evaluation (torch ops of head.py:198-204 on the GPU) over all B*A anchors.
   python tools/state_head_probe.py"""
import sys, time
import torch
sys.path.insert(0, ".")
import sarpost

dev = torch.device("cuda:0")
B, A, E, H, S, MD = 16, 136000, 256, 128, 6, 300
g = torch.Generator().manual_seed(0)
lin1, lin2 = torch.nn.Linear(E, H), torch.nn.Linear(H, S)
mlp = sarpost.StateMLP.from_tensors(lin1.weight, lin1.bias, lin2.weight, lin2.bias, device=dev)
rows = torch.randn(B, MD, 6 + E + S, device=dev)
counts = torch.full((B,), MD, dtype=torch.int32, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


t_call = timeit(lambda: sarpost.state_head(rows, counts, mlp), 200)
# device time without the Python/ctypes call overhead: 20 calls captured in one CUDA graph
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    sarpost.state_head(rows, counts, mlp)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(20):
            sarpost.state_head(rows, counts, mlp)
t_k = timeit(graph.replay, 20) / 20
print(f"sarpost_state_head  {B}x{MD} rows, {E}->{H}->{S}: {t_k:.1f} us on the device ({t_call:.1f} us per eager Python call)")
# the reference's formulation on the same GPU: every anchor of one level-0-sized map (B, E, A) -> permute -> MLP -> permute
seq = torch.nn.Sequential(lin1, torch.nn.ReLU(), torch.nn.Dropout(0.1), lin2).to(dev).eval()
emb = torch.randn(B, E, A, device=dev)
with torch.no_grad():
    t_ref = timeit(lambda: seq(emb.permute(0, 2, 1)).permute(0, 2, 1).contiguous(), 5)
print(f"reference formulation (torch, all {B}x{A} anchors, same GPU): {t_ref:.0f} us  -> x{t_ref / t_k:.0f}")
