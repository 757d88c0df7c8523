"""Kernel-time breakdown of one detector step (torch.profiler / CUPTI).  Dev tool, not part of the product."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from faster_orefsdet_b200 import synth
from faster_orefsdet_b200.modeling import build_model

if os.environ.get("CUDNN_BENCH") == "1":
    torch.backends.cudnn.benchmark = True
B = int(os.environ.get("B", "64"))
model = build_model(bench._cfg("cuda:0")).eval()
model.load_state_dict(synth.state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}))
model.set_prototypes(synth.prototypes([1], 25, 7))
x8 = torch.stack(bench._images(B, 1000)).cuda()
sizes = [(640, 640)] * B


def step():
    return model.head(model.features_from_uint8(x8), sizes, sizes)


with torch.no_grad():
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("ms/step", e0.elapsed_time(e1) / 5)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("total device us", tot)
for k, t, c in rows[:40]:
    print(f"{t:10.1f} us {c:4d}x {100 * t / tot:5.1f}%  {k[:110]}")
