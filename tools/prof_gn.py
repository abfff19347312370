"""run one GroupNorm shape repeatedly: python tools/prof_gn.py n hw c0 c1"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import *  # noqa
from mvldm_b200 import _lib
n, hw, c0, c1 = (int(a) for a in sys.argv[1:5])
lib = _lib.load()
x0 = torch.randn(n, hw, c0).to(torch.bfloat16).cuda()
x1 = torch.randn(n, hw, c1).to(torch.bfloat16).cuda() if c1 else None
C = c0 + c1
g, b = torch.randn(C).cuda(), torch.randn(C).cuda()
out = torch.empty(n, hw, C, dtype=torch.bfloat16, device="cuda")
scratch = torch.empty(n * 32 * 2 * 64, device="cuda")
def run():
    _lib.check(lib.mvldm_op_groupnorm(stream_ptr(), x0.data_ptr(), c0, x1.data_ptr() if c1 else None, c1, n, hw, 32, 1e-5,
                                      g.data_ptr(), b.data_ptr(), 1, out.data_ptr(), scratch.data_ptr()))
for _ in range(5): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"groupnorm n={n} hw={hw} C={c0}+{c1}: {us:.1f} us  {4*n*hw*C/us*1e-3:.0f} GB/s (read+write)")
