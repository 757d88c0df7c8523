import sys, os
sys.path.insert(0, "/root/repo")
import torch
from faster_orefsdet_b200 import ops
B, C, cap, n = 64, 1, 320, 256
dev = "cuda"
P = B * C
pooled = torch.randn(P, 3, 256, 128, 32, device=dev)
ctr = torch.rand(P, cap, 2, device=dev) * 500 + 70
wh = torch.rand(P, cap, 2, device=dev) * 100 + 60
rois = torch.cat((ctr - wh / 2, ctr + wh / 2), -1).contiguous()
counts = torch.full((P,), n, dtype=torch.int32, device=dev)
w_fold = ops.relation_pack(torch.randn(128, 8192, device=dev) * 0.01)
bias = torch.randn(C, 128, device=dev) * 0.1
w_out = torch.randn(6, 128, device=dev) * 0.05
b_out = torch.zeros(6, device=dev)
for _ in range(2):
    db, ds, lg, dl = ops.relation_head(pooled, w_fold, bias, w_out, b_out, rois, counts, C, (10., 10., 5., 5.), want_raw=True)
torch.cuda.synchronize()
d = dl.view(-1)[1024:1024 + 256].view(torch.int64).cpu().view(32, 4)
prev = None
for r in d.tolist():
    print(r[0] - (prev if prev else r[0]), r[1], r[2], r[3])
    prev = r[0]
