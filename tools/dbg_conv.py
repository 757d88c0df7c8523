"""Per-chunk clock64 stamps of the conv_tc pipeline roles (needs build/libfod_dbg.so built with -DFOD_DBG; development tool).
   FOD_B200_LIB_DEV=build/libfod_dbg.so python tools/dbg_conv.py [H W Cin Cout k]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops, _lib
a = [int(v) for v in sys.argv[1:]] or [80, 80, 128, 128, 3]
H, W, cin, cout, k = a
B, dev = 64, "cuda"
x = torch.randn(B, H, W, cin, device=dev).permute(0, 3, 1, 2)
w = torch.randn(cout, cin, k, k, device=dev) * 0.05
packed = ops.conv2d_pack(w)
N = 64
dbg = torch.zeros(4 * N * 4, dtype=torch.int64, device=dev)
L = _lib.lib()
L.fod_conv2d_debug.argtypes = [ctypes.c_void_p]
ax = ops.absmax(x)
for _ in range(2):
    ops.conv2d_nhwc(x, packed, None, cout, k, True, x_amax=ax)
L.fod_conv2d_debug(ctypes.c_void_p(dbg.data_ptr()))
ops.conv2d_nhwc(x, packed, None, cout, k, True, x_amax=ax)
torch.cuda.synchronize()
d = dbg.cpu().view(4, N, 4)
t0 = int(d[1, 0, 0])
names = ["Bprod: start, st_free ok", "MMA: poll start, flag seen, issued+commit", "watch: start, acc_empty ok, b_full ok, ready ok",
         "conv: start, lds+split done, st_free ok, arrived"]
for role in range(4):
    print(names[role])
    for g in range(0, 40):
        print(f"  g={400 + g:4d} " + " ".join(f"{int(v) - t0:8d}" if v else "       -" for v in d[role, g]))
