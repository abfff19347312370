"""Cost of an op boundary inside the fused sequence kernel: launches of N empty ops (N - 1 grid barriers), CUDA events.
MVLDM_SEQ_FLAGS selects the fences (1 tensormap acquire, 2 async-proxy fence, 4 weight prefetch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvldm_b200 import _lib

lib = _lib.load()
s = torch.cuda.current_stream().cuda_stream
for n in (1, 2, 11, 101, 401):
    for _ in range(3):
        _lib.check(lib.mvldm_debug_seq_empty_ops(s, n))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        _lib.check(lib.mvldm_debug_seq_empty_ops(s, n))
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"flags={os.environ.get('MVLDM_SEQ_FLAGS', '7')} ops={n:4d}: {us:8.1f} us per launch, {us / n:6.2f} us per op")
