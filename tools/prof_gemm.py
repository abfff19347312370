"""run one implicit-GEMM shape repeatedly (for ncu): python tools/prof_gemm.py n cin cout hw [reps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import *  # noqa
n, cin, cout, hw = (int(a) for a in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
x = nhwc_bf16(torch.randn(n, cin, hw, hw)).cuda()
wp = pack_conv_weight(torch.randn(cout, cin, 3, 3) / (3 * cin ** 0.5)).cuda()
b = torch.randn(cout).cuda()
out = None
for _ in range(reps):
    out = run_gemm(0, [conv_seg(x)], n, hw, hw, wp, bias=b, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run_gemm(0, [conv_seg(x)], n, hw, hw, wp, bias=b, out=out)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"conv n={n} {cin}->{cout} @{hw}: {us:.1f} us  {2*n*hw*hw*cout*9*cin/us*1e-6:.1f} TF/s")
