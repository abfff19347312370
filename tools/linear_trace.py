"""per-item timeline of CTA 0 of a linear GEMM (persistent CTA, several tiles): is the epilogue of item i hidden behind the
main loop of item i+1?   python tools/linear_trace.py n_img side cin cout [mode]"""
import os, sys, ctypes
os.environ["MVLDM_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
from helpers import *  # noqa
from mvldm_b200 import _lib
n, side, cin, cout = (int(a) for a in sys.argv[1:5])
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0
x = torch.randn(n, side, side, cin).to(torch.bfloat16).cuda()
w = (torch.randn(cout, cin) / cin ** 0.5).to(torch.bfloat16).cuda()
b = torch.randn(cout).cuda()
out = None
for _ in range(4):
    out = run_gemm(0, [conv_seg(x, taps=1)], n, side, side, w, bias=b, mode=mode, out=out)
buf = (ctypes.c_int64 * 64)()
_lib.check(_lib.load().mvldm_debug_gemm_trace(buf, 64))
t = np.array(buf); t0 = t[0]
print(f"linear M={n*side*side} N={cout} K={cin} mode={mode}: CTA 0, us from kernel entry (1.9 GHz)")
print(f"  prologue done {(t[1]-t0)/1.9e3:.2f}  first TMA {(t[2]-t0)/1.9e3:.2f}  first tile landed {(t[4]-t0)/1.9e3:.2f}  exit {(t[8]-t0)/1.9e3:.2f}")
for i in range(12):
    a, e, mm = t[16 + 4 * i], t[17 + 4 * i], t[18 + 4 * i]
    if a <= t0:
        break
    print(f"  item {i}: MMAs issued {(mm-t0)/1.9e3:6.2f}  accumulator ready {(a-t0)/1.9e3:6.2f}  epilogue done {(e-t0)/1.9e3:6.2f}  (epilogue {(e-a)/1.9e3:.2f} us)")
