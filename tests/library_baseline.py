"""The "library" GPU baseline of SURVEY.md §8d: the oracle (= the reference's arithmetic, restated) run by eager PyTorch
on the same B200 under bf16 autocast, i.e. cuDNN convolutions, cuBLAS linears and torch's softmax on materialised
N x N scores - what the reference itself would execute on this GPU.  Not a test (run as a script); it lives under tests/
because it executes the oracle.   python tests/library_baseline.py [--cfg]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mvldm_oracle as O  # noqa: E402

use_cfg = "--cfg" in sys.argv
cfg = O.OracleCfg()
sd = {k: v.cuda() for k, v in O.init_weights(cfg, 0).items()}
ctx, x_t, extr, intr = O.synthetic_scene(1, 2, 6)
sched = O.DDIMOracle()
sched.set_timesteps(25)
rays = O.raymap(extr, intr, 32, 32).cuda()
cin = torch.cat([ctx, torch.zeros(1, 2, 1, 32, 32)], 2).cuda()
mask = torch.ones(1, 6, 1, 32, 32).cuda()
x_t = x_t.cuda()
ts = [int(t) for t in sched.timesteps]
t_ctx = torch.zeros(1, 2, dtype=torch.int64, device="cuda")


def step(x, t):
    """DiffusionWrapper.step (diffusion_wrapper.py:413-453) with every tensor on the GPU"""
    t_tgt = torch.full((1, 6), t, dtype=torch.int64, device="cuda")
    tgt = torch.cat([sched.scale_model_input(x, t), mask], 2)
    inputs = torch.cat([torch.cat([cin, tgt], 1), rays], 2)
    pred = O.unet_forward(sd, inputs, torch.cat([t_ctx, t_tgt], 1), cfg)[:, 2:]
    if use_cfg:
        pred_u = O.unet_forward(sd, torch.cat([tgt, rays[:, 2:]], 2), t_tgt, cfg)
        pred = pred_u + 3.0 * (pred - pred_u)
    return sched.step(pred.float(), t, x)


for mode, ctxmgr in (("bf16 autocast", lambda: torch.autocast("cuda", dtype=torch.bfloat16)),
                     ("fp32", lambda: torch.autocast("cuda", enabled=False))):
    with torch.no_grad(), ctxmgr():
        x = x_t
        for i in range(3):
            x = step(x, ts[i])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 10
        for i in range(n):
            x = step(x, ts[(3 + i) % 25])
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
    print(f"eager PyTorch oracle on {torch.cuda.get_device_name(0)}, {mode}, cfg={use_cfg}: {dt * 1e3:.2f} ms/step = {1 / dt:.1f} steps/s "
          f"(peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)")
