"""GroupNorm per-launch time inside a CUDA graph (a dependent chain of 40 launches per shape, as in the forward):
python tools/gn_graph_bench.py   (MVLDM_GN_REG=0 selects the previous slab/cluster kernel)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import *  # noqa
from mvldm_b200 import _lib

lib = _lib.load()
SHAPES = [(8, 1024, 320, 0, 10), (8, 256, 640, 0, 8), (8, 64, 1280, 0, 8), (8, 16, 1280, 0, 12), (8, 1024, 320, 320, 2),
          (8, 16, 1280, 1280, 3), (8, 64, 1280, 1280, 2), (8, 1024, 640, 320, 1), (8, 256, 1280, 640, 1), (8, 256, 640, 320, 1),
          (8, 256, 1280, 0, 1), (8, 256, 320, 0, 1), (8, 64, 640, 0, 1), (8, 64, 1280, 640, 1), (64, 1024, 320, 0, 0), (64, 1024, 320, 320, 0), (64, 256, 640, 0, 0), (64, 256, 1280, 640, 0), (64, 64, 1280, 0, 0),
          (64, 16, 1280, 0, 0), (64, 1024, 640, 320, 0)]
tot = 0.0
for (n, hw, c0, c1, count) in SHAPES:
    C = c0 + c1
    a0 = torch.randn(n, hw, c0).to(torch.bfloat16).cuda()
    a1 = torch.randn(n, hw, c1).to(torch.bfloat16).cuda() if c1 else None
    g, b = torch.randn(C).cuda(), torch.randn(C).cuda()
    outs = [torch.empty(n, hw, C, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    scratch = torch.empty(n * 32 * 2 * 64, device="cuda")

    def chain(k):
        for i in range(k):
            _lib.check(lib.mvldm_op_groupnorm(stream_ptr(), a0.data_ptr(), c0, a1.data_ptr() if c1 else None, c1, n, hw, 32,
                                              1e-5, g.data_ptr(), b.data_ptr(), 1, outs[i & 1].data_ptr(), scratch.data_ptr()))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        chain(3)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            chain(40)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    tot += us * count
    print(f"groupnorm n={n} hw={hw:5d} C={c0}+{c1}: {us:6.2f} us/launch (x{count} per forward)  {4 * n * hw * C / us * 1e-3:6.0f} GB/s")
print(f"sum over one 1x8 forward: {tot:.0f} us")
