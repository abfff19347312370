"""forward time (CUDA-graph replay) vs number of views / scenes: separates fixed (launch, weight streaming) from per-token cost"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvldm_b200 as mv
from mvldm_b200 import synthetic
m = mv.MultiViewUNet(mv.default_cfg(), 11, 4)
synthetic.randomise_weights(m, 0)
m = m.cuda().eval()
for (B, V) in [(1, 1), (1, 2), (1, 4), (1, 8), (2, 8), (4, 8), (8, 8)]:
    x = torch.randn(B, V, 11, 32, 32, device="cuda")
    t = torch.full((B, V), 500, device="cuda", dtype=torch.long)
    for _ in range(3):
        m(x, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        m(x, t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gf = B * (137.9 * V + 3.069 * V * V)
    print(f"B={B} V={V}: {ms:7.3f} ms/forward  launches {m.last_launch_count()}  {gf/ms:7.1f} TF/s  {ms/B:6.3f} ms/scene", flush=True)
