"""Run only the tensor-core correlation a few times (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops
B, C = 64, 1
sizes = [(80, 80), (40, 40), (20, 20)]
qs = [torch.randn(B, h, w, 128, device="cuda").permute(0, 3, 1, 2) for h, w in sizes]
taps = [torch.randn(C, 7, 128, device="cuda") * 0.3 for _ in sizes]
w3 = torch.randn(128, 256, device="cuda") * 0.05
b3 = torch.randn(128, device="cuda") * 0.1
for _ in range(3):
    ops.correlate_levels(qs, taps, w3, b3)
torch.cuda.synchronize()
print("ok")
