"""One convolution layer for an ncu capture of conv_tc_kernel:  python tools/ncu_conv.py H W Cin Cout k [stride]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops
a = [int(v) for v in sys.argv[1:]] or [320, 320, 64, 64, 3]
H, W, cin, cout, k = a[:5]
st = a[5] if len(a) > 5 else 1
x = torch.randn(64, H, W, cin, device="cuda").permute(0, 3, 1, 2)
packed = ops.conv2d_pack(torch.randn(cout, cin, k, k, device="cuda") * 0.05)
b = torch.randn(cout, device="cuda")
ax = ops.absmax(x)
for _ in range(3):
    ops.conv2d_nhwc(x, packed, b, cout, k, True, stride=st, x_amax=ax)
torch.cuda.synchronize()
print("ok")
