"""timeline of CTA 0 of one conv GEMM: python tools/gemm_trace.py n cin cout hw"""
import os, sys, ctypes
os.environ["MVLDM_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
from helpers import *  # noqa
from mvldm_b200 import _lib
n, cin, cout, hw = (int(a) for a in sys.argv[1:5])
x = nhwc_bf16(torch.randn(n, cin, hw, hw)).cuda()
wp = pack_conv_weight(torch.randn(cout, cin, 3, 3) / (3 * cin ** 0.5)).cuda()
b = torch.randn(cout).cuda()
out = None
for _ in range(4):
    out = run_gemm(0, [conv_seg(x)], n, hw, hw, wp, bias=b, out=out)
buf = (ctypes.c_int64 * 16)()
_lib.check(_lib.load().mvldm_debug_gemm_trace(buf, 16))
t = np.array(buf)[:11]; t = t - t[0]
names = ["entry", "prologue done", "first TMA issued", "last TMA issued", "first tile landed", "all MMAs issued", "accumulator ready", "epilogue done", "exit", "splits arrived", "slice reduced"]
print(f"conv n={n} {cin}->{cout} @{hw}: CTA 0 timeline (cycles from kernel entry)")
for nm, v in zip(names, t):
    print(f"  {nm:20s} {int(v):8d}  ({v/1.9e3:6.2f} us)")
