"""stem1_u8 alone vs right after a burst of tensor-core convolutions (development tool: power / clock context)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops
x = (torch.rand(64, 3, 640, 640, device="cuda") * 255).to(torch.uint8)
w = torch.randn(64, 3, 3, 3, device="cuda") * 0.1
b = torch.zeros(64, device="cuda")
mean, std = [103.53, 116.28, 123.675], [1.0, 1.0, 1.0]
xc = torch.randn(64, 160, 160, 128, device="cuda").permute(0, 3, 1, 2)
pk = ops.conv2d_pack(torch.randn(64, 128, 3, 3, device="cuda") * 0.05)
ax = ops.absmax(xc)


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


print("stem1 alone, back to back: %.3f ms" % t(lambda: ops.stem1_u8(x, mean, std, w, b)))
print("conv 128->64 3x3 @160^2 alone: %.3f ms" % t(lambda: ops.conv2d_nhwc(xc, pk, None, 64, 3, True, x_amax=ax)))


def mixed():
    for _ in range(8):
        ops.conv2d_nhwc(xc, pk, None, 64, 3, True, x_amax=ax)
    ops.stem1_u8(x, mean, std, w, b)


tm = t(mixed)
print("8 convs + stem1: %.3f ms -> stem1 share if convs unchanged: %.3f ms" % (tm, tm - 8 * t(lambda: ops.conv2d_nhwc(xc, pk, None, 64, 3, True, x_amax=ax))))

import subprocess, time
for name, fn in (("stem1", lambda: ops.stem1_u8(x, mean, std, w, b)), ("conv", lambda: ops.conv2d_nhwc(xc, pk, None, 64, 3, True, x_amax=ax))):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu", "--format=csv,noheader,nounits", "-lms", "200"],
                         stdout=subprocess.PIPE, text=True)
    t0 = time.time()
    n = 0
    while time.time() - t0 < 2.5:
        for _ in range(50):
            fn()
        torch.cuda.synchronize()
        n += 50
    dt = time.time() - t0
    p.terminate()
    out = p.stdout.read().strip().splitlines()
    print(f"{name}: {dt / n * 1e3:.3f} ms per launch over {dt:.1f} s; nvidia-smi samples (sm MHz, W, power cap, hw slowdown, sw thermal, temp):")
    print("   " + " | ".join(out[2:10]))
