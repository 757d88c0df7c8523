"""Phase cycle counts of roi_align_kernel<8> (first 64 ROIs of problem 0), needs a -DFOD_ROI_PROF build."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops, _lib
B, C, cap, n = 64, 1, 320, 256
dev = "cuda"
torch.manual_seed(0)
P = B * C
feats = [torch.randn(B, h, w, 128, device=dev).permute(0, 3, 1, 2) for h, w in ((80, 80), (40, 40), (20, 20))]
ctr = torch.rand(P, cap, 2, device=dev) * 500 + 70
wh = torch.rand(P, cap, 2, device=dev) * 100 + 60
rois = torch.cat((ctr - wh / 2, ctr + wh / 2), -1).contiguous()
counts = torch.full((P,), n, dtype=torch.int32, device=dev)
L = _lib.lib()
for _ in range(3):
    ops.roi_align(feats, (8, 16, 32), rois, counts, C, 8, tiled=True)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 8)()
L.fod_roi_prof(buf, 1)
iters = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.roi_align(feats, (8, 16, 32), rois, counts, C, 8, tiled=True)
e1.record()
torch.cuda.synchronize()
L.fod_roi_prof(buf, 0)
v = [x / iters / 64 for x in buf]
print(f"roi_align {e0.elapsed_time(e1)/iters*1e3:.1f} us; per CTA cycles (cumulative): blob load {v[1]:.0f}, accumulate {v[2]:.0f}, stores {v[3]:.0f}")
