"""Micro-benchmark of decode_topk + nms_proposals + final_detect (development tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
torch.manual_seed(0)
n, cap = 2400, 3000
ctr = torch.rand(P, cap, 2, device=dev) * 600 + 20
wh = torch.rand(P, cap, 2, device=dev) * 60 + 100
boxes = torch.cat((ctr - wh / 2, ctr + wh / 2), -1).contiguous()
scores = torch.rand(P, cap, device=dev)
count = torch.full((P,), n, dtype=torch.int32, device=dev)
status = ops.new_status(dev)


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


keep, ob, os_, oc = ops.nms_proposals(boxes, scores, count, 0.6, 256, 320, status)
t = timeit(lambda: ops.nms_proposals(boxes, scores, count, 0.6, 256, 320, status))
print(f"P={P} n={n}: nms_proposals {t*1e3:.1f} us  kept avg {oc.float().mean().item():.0f}")
