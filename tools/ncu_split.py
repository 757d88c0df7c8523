"""stem_1 -> stem_2 at the bench shape (batch 64 x 640x640 uint8) with the fp32 hand-off and with the split hand-off, for
    ncu --set full --clock-control none -k regex:conv_tc_kernel -s 4 -c 2 -o gpurun_out/r3_split python tools/ncu_split.py
(launches: 2 warm-up pairs, then fp32-fed stem_2, then pre-split-fed stem_2; also prints CUDA-event times)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops, synth
dev = "cuda"
n, h, w = 64, 640, 640
x = torch.stack([synth.ore_image(h, w, 1000 + i % 8) for i in range(n)]).to(dev)
mean, std = [103.53, 116.28, 123.675], [57.375, 57.12, 58.395]
w1 = synth.tensor((64, 3, 3, 3), 82, -0.3, 0.3)
b1 = synth.tensor((64,), 83, -1.0, 1.0).to(dev)
w2 = (synth.tensor((64, 64, 3, 3), 84, -0.08, 0.08)).to(dev)
b2 = synth.tensor((64,), 85, -0.5, 0.5).to(dev)
w32 = torch.cat((w1.permute(0, 2, 3, 1).reshape(64, 27), torch.zeros(64, 5)), 1).reshape(64, 32, 1, 1).contiguous()
pk1, pk2 = ops.conv2d_pack(w32.to(dev)), ops.conv2d_pack(w2)
xmax = torch.tensor([max(abs(0.0 - m), abs(255.0 - m)) / sd for m, sd in zip(mean, std)], dtype=torch.float64)
bound = (((w1.double().abs().sum((2, 3)) * xmax.view(1, 3)).sum(1) + b1.cpu().double().abs()).max() * 1.001).float().reshape(1).to(dev)
out = torch.empty((n, h // 2, w // 2, 64), device=dev).permute(0, 3, 1, 2)


def fp32():
    a1 = torch.zeros(n, device=dev)
    y1 = ops.stem1_u8_tc(x, mean, std, pk1, b1, y_amax=a1)
    return ops.conv2d_nhwc(y1, pk2, b2, 64, 3, True, out=out, x_amax=a1.view(1, n))


def split():
    y1 = ops.stem1_u8_tc(x, mean, std, pk1, b1, y_bound=bound)
    return ops.conv2d_nhwc(y1, pk2, b2, 64, 3, True, out=out, x_amax=bound, x_presplit=True)


for _ in range(2):
    fp32(); split()
torch.cuda.synchronize()
ya = fp32().clone()
yb = split().clone()
torch.cuda.synchronize()
print("max rel diff", float((ya - yb).abs().max() / ya.abs().max()))
if len(sys.argv) > 1:
    for name, fn in (("fp32 hand-off", fp32), ("split hand-off", split)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"stem_1 + stem_2, {name}: {e0.elapsed_time(e1) / 10:.3f} ms")
