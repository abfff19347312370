"""Barrier timeline of every fused sequence launch of one forward: for each op, when the first / median / last CTA finished its
work (relative to the release of the previous barrier) and how long the barrier took after the last arrival.

    python tools/seq_trace.py [views] > gpurun_out/seq_trace.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mvldm_b200 as mv
from mvldm_b200 import _lib, synthetic

V = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = mv.MultiViewUNet(mv.default_cfg(), 11, 4, use_cuda_graph=False)
synthetic.randomise_weights(m, 0)
m = m.cuda().eval()
x = torch.randn(1, V, 11, 32, 32, device="cuda")
t = torch.tensor([[0, 0] + [500] * (V - 2)], device="cuda")
for _ in range(3):
    m(x, t)
m.set_profiling(True)
m(x, t)
prof = m.profile()
m.set_profiling(False)
grid = torch.cuda.get_device_properties(0).multi_processor_count
buf = torch.zeros(8 << 20, dtype=torch.int64, device="cuda")
lib = _lib.load()
torch.cuda.synchronize()
torch.cuda._sleep(int(60e6))
lib.mvldm_debug_seq_trace(buf.data_ptr())
m(x, t)
torch.cuda.synchronize()
lib.mvldm_debug_seq_trace(None)
h = buf.cpu().numpy()
ops = [o for o in prof["ops"] if not o["cat"].startswith("attention")]
launches = [l for l in prof["launches"] if l["kind"] == "seq"]
off, oi = 0, 0
print(f"{'op':50s} {'first':>7s} {'median':>7s} {'last':>7s} {'t0lag':>6s} {'release':>7s}   (us; arrival of CTAs after the previous release; barrier cost after the last arrival)")
for l in launches:
    n = l["ops"]
    rows = h[off:off + 2 * n * grid * 16].reshape(2 * n, grid, 16)
    off += 2 * n * grid * 16
    prev_release = None
    names = ops[oi:oi + n]
    oi += n
    for b in range(2 * n):
        r = rows[b]
        if r[:, 2].max() == 0:
            break
        opi = int(r[0, 3])
        entry, done, rel = r[:, 0], r[:, 1], r[:, 2]
        base = prev_release if prev_release is not None else entry.min()
        a = np.sort(done - base) / 1e3
        name = f"{names[opi]['cat']} {names[opi]['what']}"
        print(f"{name[:50]:50s} {a[0]:7.2f} {np.median(a):7.2f} {a[-1]:7.2f} {(done - entry).max() / 1e3:6.2f} {(rel.min() - done.max()) / 1e3:7.2f}"
              f"   last CTA {int(np.argmax(done))}", end="")
        lc = int(np.argmax(done))
        print(f"  polls {int(np.median(r[:, 5]))}", end="")
        if names[opi]["cat"].startswith("gemm") and r[lc, 9] > 0:
            e = r[lc]
            b0 = base
            seq = [e[6], e[7], e[8], e[9], e[13], e[10], e[11], e[14], e[15], e[12], e[4]]
            print("  last CTA: tma0 %.2f tile0 %.2f wait_acc %.2f acc %.2f chunk0 %.2f stored %.2f splits_in %.2f cp_issued %.2f staged %.2f reduced %.2f c1wait %.2f" % tuple(x / 1965.0 if x > 0 else -1 for x in seq), end="")
        print()
        prev_release = rel.min()
    print("-- launch end --")
