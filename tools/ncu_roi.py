import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops
B, C, cap, n = 64, 1, 320, 256
dev = "cuda"
torch.manual_seed(0)
P = B * C
feats = [torch.randn(B, h, w, 128, device=dev).permute(0, 3, 1, 2) for h, w in ((80, 80), (40, 40), (20, 20))]
ctr = torch.rand(P, cap, 2, device=dev) * 500 + 70
wh = torch.rand(P, cap, 2, device=dev) * 100 + 60
rois = torch.cat((ctr - wh / 2, ctr + wh / 2), -1).contiguous()
counts = torch.full((P,), n, dtype=torch.int32, device=dev)
for _ in range(3):
    ops.roi_align(feats, (8, 16, 32), rois, counts, C, 8, tiled=True)
torch.cuda.synchronize()
print("ok")
