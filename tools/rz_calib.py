"""Calibration of the tensor core's accumulator rounding (development tool).  Same-sign operands, one partial sum
over the whole K loop, no compensation: the mean relative error against float64 grows linearly with the number of
accumulated MMAs n as -kappa * n / 2.   FOD_CONV_PART_CHUNKS=100000 FOD_CONV_RZ_KAPPA=0 python tools/rz_calib.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from faster_orefsdet_b200 import ops, synth
for cin in (32, 64, 128, 256, 384, 512):
    x = synth.tensor((1, cin, 24, 32), 21, 0.5, 1.0)
    w = synth.tensor((64, cin, 3, 3), 22, 0.5, 1.0) / (9.0 * cin)
    ref = F.conv2d(x.double(), w.double(), padding=1)
    y = ops.conv2d_nhwc(x.cuda().contiguous(memory_format=torch.channels_last), ops.conv2d_pack(w.cuda()), None, 64, 3)
    rel = ((y.cpu().double() - ref) / ref)[:, :, 2:-2, 2:-2]
    n = 9 * (cin // 32) * 6
    print(f"Cin {cin:4d}: {n:5d} MMAs  mean rel {float(rel.mean()):+.3e}  std {float(rel.std()):.2e}  -> kappa = {-2 * float(rel.mean()) / n:.3e}")
# mixed signs: the error should have no systematic sign
x = synth.tensor((1, 256, 24, 32), 31, -1.0, 1.0)
w = synth.tensor((64, 256, 3, 3), 32, -1.0, 1.0) / 48.0
ref = F.conv2d(x.double(), w.double(), padding=1)
y = ops.conv2d_nhwc(x.cuda().contiguous(memory_format=torch.channels_last), ops.conv2d_pack(w.cuda()), None, 64, 3)
err = (y.cpu().double() - ref) / ref.abs().max()
print(f"mixed signs: mean of err*sign(ref) / max {float((err * ref.sign()).mean()):+.3e}   rms {float(err.pow(2).mean().sqrt()):.2e}")
