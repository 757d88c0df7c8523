"""Micro-benchmark of roi_align + relation_head alone (development tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cap, n = 320, 256
dev = "cuda"
torch.manual_seed(0)
P = B * C
feats = [torch.randn(B, h, w, 128, device=dev).permute(0, 3, 1, 2) for h, w in ((80, 80), (40, 40), (20, 20))]
ctr = torch.rand(P, cap, 2, device=dev) * 500 + 70
wh = torch.rand(P, cap, 2, device=dev) * 100 + 60
rois = torch.cat((ctr - wh / 2, ctr + wh / 2), -1).contiguous()
counts = torch.full((P,), n, dtype=torch.int32, device=dev)
w_fold = ops.relation_pack(torch.randn(128, 8192, device=dev) * 0.01)
bias = torch.randn(C, 128, device=dev) * 0.1
w_out = torch.randn(6, 128, device=dev) * 0.05
b_out = torch.zeros(6, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


pooled = ops.roi_align(feats, (8, 16, 32), rois, counts, C, 8, tiled=True)
t_roi = timeit(lambda: ops.roi_align(feats, (8, 16, 32), rois, counts, C, 8, tiled=True))
xa = torch.cat([ops.absmax(f.contiguous(memory_format=torch.channels_last)) for f in feats])
t_rel = timeit(lambda: ops.relation_head(pooled, w_fold, bias, w_out, b_out, rois, counts, C, (10., 10., 5., 5.), x_amax=xa))
flops = 2.0 * P * n * 8192 * 128
print(f"B={B} C={C}: roi_align {t_roi*1e3:.1f} us   relation_head {t_rel*1e3:.1f} us ({flops/t_rel/1e9:.1f} fp32-equivalent TFLOP/s, "
      f"pooled read {P*n*32768/t_rel/1e6:.0f} GB/s)")
