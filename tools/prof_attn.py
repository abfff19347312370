"""run one attention shape repeatedly (for ncu / timing): python tools/prof_attn.py B N d [reps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import *  # noqa
from mvldm_b200 import _lib
B, N, d = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
heads, dpad = 8, (d + 63) // 64 * 64
qkv = torch.randn(B * N, 3, heads, dpad)
qkv[:, 2, :, d] = 1.0
qkv = qkv.reshape(B * N, -1).to(torch.bfloat16).cuda()
out = torch.empty(B * N, heads * dpad, dtype=torch.bfloat16, device="cuda")
lib = _lib.load()
for _ in range(reps):
    _lib.check(lib.mvldm_op_attention(stream_ptr(), 0, qkv.data_ptr(), out.data_ptr(), B, N, heads, d, dpad))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    _lib.check(lib.mvldm_op_attention(stream_ptr(), 0, qkv.data_ptr(), out.data_ptr(), B, N, heads, d, dpad))
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
tiles = B * heads * ((N + 127) // 128) ** 2
print(f"attn B={B} N={N} d={d}: {us:.1f} us  {4*B*N*N*heads*d/us*1e-6:.1f} TF/s (un-padded)  {us*1e3*1.9/ (tiles/148):.0f} cycles per 128x128 tile per SM")
