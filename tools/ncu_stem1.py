"""stem1_u8 launches for an ncu capture, plus a pure-write reference (fill of the same 1.68 GB) timed with CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops
x = (torch.rand(64, 3, 640, 640, device="cuda") * 255).to(torch.uint8)
w = torch.randn(64, 3, 3, 3, device="cuda") * 0.1
b = torch.zeros(64, device="cuda")
mean, std = [103.53, 116.28, 123.675], [1.0, 1.0, 1.0]
for _ in range(3):
    y = ops.stem1_u8(x, mean, std, w, b)
torch.cuda.synchronize()
if len(sys.argv) > 1:
    def t(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    z = torch.empty_like(y)
    print("fill 1.68 GB: %.3f ms" % t(lambda: y.fill_(1.0)))
    print("copy 1.68 GB -> 1.68 GB: %.3f ms" % t(lambda: z.copy_(y)))
    print("relu_ in place 1.68 GB: %.3f ms" % t(lambda: y.relu_()))
print("ok")
