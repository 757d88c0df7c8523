"""stem1_u8 alone vs right after a burst of tensor-core convolutions (development tool: power / clock context)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops
x = (torch.rand(64, 3, 640, 640, device="cuda") * 255).to(torch.uint8)
w = torch.randn(64, 3, 3, 3, device="cuda") * 0.1
b = torch.zeros(64, device="cuda")
mean, std = [103.53, 116.28, 123.675], [1.0, 1.0, 1.0]
std2 = [1.0, 57.4, 58.4]
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

# the same layer through the tensor-core stem kernel (csrc/stem1_tc.cu)
w32 = torch.cat((w.permute(0, 2, 3, 1).reshape(64, 27), torch.zeros(64, 5, device="cuda")), 1).reshape(64, 32, 1, 1).contiguous()
pk1 = ops.conv2d_pack(w32)
print("stem1 on tensor cores, back to back: %.3f ms" % t(lambda: ops.stem1_u8_tc(x, mean, std, pk1, b)))
ya, yb = ops.stem1_u8(x, mean, std, w, b), ops.stem1_u8_tc(x, mean, std, pk1, b)
print("max |fma - tc| = %.3e of %.3e" % (float((ya - yb).abs().max()), float(ya.abs().max())))
xs4 = [(torch.rand(64, 3, 640, 640, device="cuda") * 255).to(torch.uint8) for _ in range(4)]
am = ops.new_amax("cuda", 64)
k = [0]
def rot():
    k[0] += 1
    ops.stem1_u8_tc(xs4[k[0] % 4], mean, std, pk1, b, y_amax=am)
print("stem1 on tensor cores, rotating inputs + per-image bounds: %.3f ms" % t(rot))
def rot_fma():
    k[0] += 1
    ops.stem1_u8(xs4[k[0] % 4], mean, std, w, b, y_amax=am)
print("stem1 FMA kernel, rotating inputs + per-image bounds: %.3f ms" % t(rot_fma))


def mixed_tc():
    for _ in range(8):
        ops.conv2d_nhwc(xc, pk, None, 64, 3, True, x_amax=ax)
    rot()


tc8 = t(lambda: [ops.conv2d_nhwc(xc, pk, None, 64, 3, True, x_amax=ax) for _ in range(8)])
tm = t(mixed_tc)
print("8 convs + tensor-core stem1: %.3f ms -> stem1 share if convs unchanged (%.3f): %.3f ms" % (tm, tc8, tm - tc8))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for _ in range(10):
    for _ in range(8):
        ops.conv2d_nhwc(xc, pk, None, 64, 3, True, x_amax=ax)
    e0.record()
    rot()
    e1.record()
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print("tensor-core stem1 right after 8 convs, by its own events: %.3f ms" % (tot / 10))
print("stem1 on tensor cores, non-unit pixel_std (table look-up path): %.3f ms" % t(lambda: ops.stem1_u8_tc(x, mean, std2, pk1, b)))
