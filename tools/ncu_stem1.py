"""One launch pattern for ncu: stem_1 on the tensor cores (csrc/stem1_tc.cu) at batch 64, 640x640 (development tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops
xs = [(torch.rand(64, 3, 640, 640, device="cuda") * 255).to(torch.uint8) for _ in range(2)]
w = torch.randn(64, 3, 3, 3, device="cuda") * 0.1
w32 = torch.cat((w.permute(0, 2, 3, 1).reshape(64, 27), torch.zeros(64, 5, device="cuda")), 1).reshape(64, 32, 1, 1).contiguous()
pk = ops.conv2d_pack(w32)
b = torch.zeros(64, device="cuda")
am = ops.new_amax("cuda", 64)
for i in range(4):
    ops.stem1_u8_tc(xs[i % 2], [103.53, 116.28, 123.675], [1.0, 1.0, 1.0], pk, b, y_amax=am)
torch.cuda.synchronize()
print("ok")
