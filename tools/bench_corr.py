"""Micro-benchmark of the correlation kernels alone (development tool): CUDA events, inputs > L2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sizes = [(80, 80), (40, 40), (20, 20)]
dev = "cuda"
torch.manual_seed(0)
NB = 4  # rotate input sets so q does not sit in L2
qs = [[torch.randn(B, h, w, 128, device=dev).permute(0, 3, 1, 2) for h, w in sizes] for _ in range(NB)]
taps = [torch.randn(C, 7, 128, device=dev) * 0.3 for _ in sizes]
w3 = torch.randn(128, 256, device=dev) * 0.05
b3 = torch.randn(128, device=dev) * 0.1
M = sum(h * w for h, w in sizes)
bytes_alg = 1024 * M * B * C


def timeit(fn, iters=20, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


old = lambda i: [ops.correlate(q, t, w3, b3) for q, t in zip(qs[i % NB], taps)]
new = lambda i: ops.correlate_levels(qs[i % NB], taps, w3, b3)
a = old(0)
b = new(0)
torch.cuda.synchronize()
for x, y in zip(a, b):
    d = (x - y).abs().max().item()
    print("max |old-new| =", d, " max |old| =", x.abs().max().item())
t_old = timeit(old)
t_new = timeit(new)
print(f"B={B} C={C}: old {t_old*1e3:.1f} us ({bytes_alg/t_old/1e6:.0f} GB/s)   new {t_new*1e3:.1f} us ({bytes_alg/t_new/1e6:.0f} GB/s, "
      f"{bytes_alg/t_new/1e6/6546.2*100:.1f}% of 6546 GB/s)")
