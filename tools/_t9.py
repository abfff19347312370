import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, torch.nn.functional as F
from helpers import conv_seg, rel_err, run_gemm
torch.manual_seed(8)
for (n, hw, K, N, mode) in [(4, 32, 128, 320, 0), (4, 32, 320, 320, 4)]:
    M = n * hw * hw
    a = torch.randn(M, K).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N).cuda()
    ref = a.float() @ w.float().t() + b
    out = torch.full((M, N), 777.0, device="cuda", dtype=torch.float32 if mode == 4 else torch.bfloat16)
    run_gemm(3, [conv_seg(a.view(n, hw, hw, K), 1, 1)], n, hw, hw, w, bias=b, mode=mode, out=out)
    o = out.float()
    unwritten = (o == 777.0)
    good = (o - ref).abs() < 0.05 * ref.abs().max()
    print(f"K{K} mode{mode}: unwritten {int(unwritten.sum())} good {int(good.sum())} nan {int(torch.isnan(o).sum())} of {o.numel()}")
    print(" row0 good cols:", good[0].nonzero().flatten()[:24].tolist())
    print(" row0 unwritten cols:", unwritten[0].nonzero().flatten()[:24].tolist())
    print(" rows with any good:", good.any(1).sum().item(), "row 200 good cols", good[200].nonzero().flatten()[:12].tolist())
