"""Variant B (SD-2.1 topology) quick check on the GPU box: steady-state forward time at 1 scene x 8 views and the per-op profile."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mvldm_b200 as mv  # noqa: E402


def main():
    cfg = mv.default_cfg()
    cfg.pretrained_from = "stabilityai/stable-diffusion-2-1"
    m = mv.MultiViewUNet(cfg, 11, 4).cuda().eval()
    for k, p in m.named_parameters():
        if k.endswith("proj_out.weight"):
            torch.nn.init.normal_(p, std=0.02)
    m.mark_dirty()
    out = {}
    for (B, V) in [(1, 8), (8, 8)]:
        x = torch.randn(B, V, 11, 32, 32, device="cuda")
        t = torch.full((B, V), 500, dtype=torch.int64, device="cuda")
        for _ in range(5):
            m(x, t)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            m(x, t)
        e1.record()
        torch.cuda.synchronize()
        out[f"B{B}V{V}_ms"] = e0.elapsed_time(e1) / 20
        out[f"B{B}V{V}_launches"] = m.last_launch_count()
    m.set_profiling(True)
    x = torch.randn(1, 8, 11, 32, 32, device="cuda")
    t = torch.full((1, 8), 500, dtype=torch.int64, device="cuda")
    m(x, t); m(x, t)
    out["profile_B1V8"] = m.profile()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
