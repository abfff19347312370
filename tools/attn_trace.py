"""timeline of one attention CTA: MVLDM_ATTN_TRACE=1 python tools/attn_trace.py B N d"""
import os, sys, ctypes
os.environ["MVLDM_ATTN_TRACE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
from helpers import *  # noqa
from mvldm_b200 import _lib
B, N, d = (int(a) for a in sys.argv[1:4])
heads, dpad = 8, (d + 63) // 64 * 64
qkv = torch.randn(B * N, 3, heads, dpad)
qkv[:, 2, :, d] = 1.0
qkv = qkv.reshape(B * N, -1).to(torch.bfloat16).cuda()
out = torch.empty(B * N, heads * dpad, dtype=torch.bfloat16, device="cuda")
lib = _lib.load()
for _ in range(3):
    _lib.check(lib.mvldm_op_attention(stream_ptr(), 0, qkv.data_ptr(), out.data_ptr(), B, N, heads, d, dpad))
buf = (ctypes.c_int64 * (8 * 512))()
_lib.check(lib.mvldm_debug_attn_trace(buf, 8 * 512))
t = np.array(buf).reshape(8, 512)
T = min(512, (N + (64 if dpad == 192 else 128) - 1) // (64 if dpad == 192 else 128))
base = t[0, 0]
print("tile: waitS gotS handP | gotP pvIssued qkIssued   (cycles since first stamp)")
for j in list(range(min(T, 6))) + list(range(max(6, T - 3), T)):
    print(j, *(int(t[s, j] - base) for s in range(6)))
if T > 8:
    a, b = 8, T - 2
    per = (t[2, b] - t[2, a]) / (b - a)
    print(f"steady state: {per:.0f} cycles per tile per CTA; softmax busy {np.mean(t[2, a:b] - t[1, a:b]):.0f}, waiting for S {np.mean(t[1, a:b] - t[0, a:b]):.0f}; "
          f"MMA: wait for P {np.mean(t[3, a+1:b] - t[5, a:b-1]):.0f}, issue PV {np.mean(t[4, a:b] - t[3, a:b]):.0f}, issue QK {np.mean(t[5, a:b] - t[4, a:b]):.0f}")
