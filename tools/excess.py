"""Where the B=1 step loses time: per-op device time of one forward (1 scene x V views) against each op's own
roofline floor max(FLOPs / sustained bf16 peak, algorithmic bytes / HBM peak).  Uses the in-library CUDA-event
profiler (eager launches behind a spin kernel, so kernels run back to back)."""
import collections
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mvldm_b200 as mv  # noqa: E402
from profile_step import profile_forward  # noqa: E402

PEAK_TF, PEAK_GB = 1395.1, 6453.7


def floor_us(o):
    t = o["gflop"] / PEAK_TF * 1e3   # GFLOP / (TFLOP/s) = ms; x 1e3 -> us
    b = 0.0
    m = re.match(r"M(\d+) N(\d+) K(\d+)", o["what"])
    if m:
        M, N, K = map(int, m.groups())
        ka = K / 9 if o["cat"] == "gemm_conv3x3" else K
        b = 2.0 * (N * K + M * N + M * ka)
    m = re.match(r"tokens(\d+) C(\d+)", o["what"])
    if m:
        T, C = map(int, m.groups())
        b = 4.0 * T * C
    return max(t, b / (PEAK_GB * 1e9) * 1e6)


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    m = mv.MultiViewUNet(mv.default_cfg(), 11, 4).cuda().eval()
    for k, p in m.named_parameters():
        if k.endswith("proj_out.weight"):
            torch.nn.init.normal_(p, std=0.02)
    m.mark_dirty()
    x = torch.randn(B, V, 11, 32, 32, device="cuda")
    t = torch.full((B, V), 500, dtype=torch.int64, device="cuda")
    for _ in range(3):
        m(x, t)
    p = profile_forward(m, x, t, reps=4)
    agg = collections.OrderedDict()
    for o in p["ops"]:
        a = agg.setdefault((o["cat"], o["what"]), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += o["us"]
        a[2] += floor_us(o)
    tot = sum(a[1] for a in agg.values())
    print(f"B={B} V={V}: {len(p['ops'])} ops, {tot:.0f} us summed, floor {sum(a[2] for a in agg.values()):.0f} us")
    cats = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for (cat, _), a in agg.items():
        for i in range(3):
            cats[cat][i] += a[i]
    for cat, a in sorted(cats.items(), key=lambda kv: -kv[1][1]):
        print(f"  {cat:20s} n={a[0]:3d} {a[1]:7.1f} us  floor {a[2]:7.1f}  excess {a[1] - a[2]:7.1f}")
    print("per shape, by excess:")
    for (cat, what), a in sorted(agg.items(), key=lambda kv: -(kv[1][1] - kv[1][2])):
        print(f"  {cat:18s} {what:28s} n={a[0]:2d} {a[1]:7.1f} us ({a[1] / a[0]:6.1f} each)  floor {a[2] / a[0]:6.1f} each  excess {a[1] - a[2]:7.1f}")
    json.dump(p, open(f"gpurun_out/excess_b{B}_v{V}.json", "w"))


if __name__ == "__main__":
    main()
