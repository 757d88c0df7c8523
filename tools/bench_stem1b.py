import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops
import bench
w = torch.randn(64, 3, 3, 3, device="cuda") * 0.1
b = torch.zeros(64, device="cuda")
mean, std = [103.53, 116.28, 123.675], [1.0, 1.0, 1.0]
for name, x in (("random", (torch.rand(64, 3, 640, 640, device="cuda") * 255).to(torch.uint8)),
                ("ore", torch.stack(bench._images(64, 1000)).cuda()),
                ("constant", torch.full((64, 3, 640, 640), 37, dtype=torch.uint8, device="cuda"))):
    ts = []
    for _ in range(6):
        torch.cuda.synchronize(); time.sleep(0.05)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.stem1_u8(x, mean, std, w, b); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(name, "isolated launches (ms):", [round(t, 3) for t in ts])
