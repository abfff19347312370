"""run one linear (1x1) GEMM shape repeatedly (for ncu / timing): python tools/prof_linear.py n_img side cin cout [mode] [res]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import *  # noqa
n, side, cin, cout = (int(a) for a in sys.argv[1:5])
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0
res = int(sys.argv[6]) if len(sys.argv) > 6 else 0
x = torch.randn(n, side, side, cin).to(torch.bfloat16).cuda()
w = (torch.randn(cout, cin) / cin ** 0.5).to(torch.bfloat16).cuda()
b = torch.randn(cout).cuda()
r = torch.randn(n * side * side, cout).to(torch.bfloat16).cuda() if res else None
out = None
for _ in range(5):
    out = run_gemm(0, [conv_seg(x, taps=1)], n, side, side, w, bias=b, residual=r, mode=mode, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run_gemm(0, [conv_seg(x, taps=1)], n, side, side, w, bias=b, residual=r, mode=mode, out=out)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
M = n * side * side
print(f"linear M={M} N={cout} K={cin} mode={mode}: {us:.1f} us (stream launches)  {2 * M * cout * cin / us * 1e-6:.0f} TF/s")
