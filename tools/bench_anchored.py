"""BASELINE.json config 3: anchored sampling of an 80-frame trajectory (1 context view, 4 anchors, 25 chunks of 3 targets),
25 DDIM steps, latent space.  Batched (all chunk calls of the second phase as one DDIM run: `mvldm_b200.sample_anchored`)
against the reference's order of execution (one `sample` call per chunk).   python tools/bench_anchored.py [--cfg]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvldm_b200 as mv
from mvldm_b200 import synthetic

use_cfg = "--cfg" in sys.argv
T, STEPS = 79, 25                      # 1 context + 79 targets = 80 frames
m = mv.MultiViewUNet(mv.default_cfg(), 11, 4).cuda().eval()
synthetic.randomise_weights(m, 0)
m = m.cuda()
sched = mv.DDIMScheduler(clip_sample=False)
path = mv.DenoisingPath(m, sched, use_cfg=use_cfg, cfg_scale=3.0)
path.set_timesteps(STEPS)
extr, intr = synthetic.cameras(1, 1 + T)
extr, intr = extr.cuda(), intr.cuda()
g = torch.Generator(device="cuda").manual_seed(3)
ctx = torch.randn(1, 1, 4, 32, 32, device="cuda", generator=g)
noise = torch.randn(1, T, 4, 32, 32, device="cuda", generator=g)


def batched():
    return mv.sample_anchored(path, ctx, extr[:, :1], intr[:, :1], extr[:, 1:], intr[:, 1:], noise, 4)


def sequential(plan, lat):
    out = lat.clone()
    for a, tg in plan.chunks:
        ai = plan.anchors.index(a)
        c = torch.cat([ctx, lat[:, plan.anchors][:, ai:ai + 1]], dim=1)
        e = torch.cat([extr[:, :1], extr[:, 1 + a:2 + a], extr[:, [1 + t for t in tg]]], dim=1)
        e = torch.linalg.inv(e[:, 1:2]) @ e
        k = torch.cat([intr[:, :1], intr[:, 1 + a:2 + a], intr[:, [1 + t for t in tg]]], dim=1)
        out[:, tg] = path.sample(c, noise[:, tg], e, k)
    return out


lat, done, plan = batched()
torch.cuda.synchronize()
t0 = time.perf_counter(); lat, done, plan = batched(); torch.cuda.synchronize(); tb = time.perf_counter() - t0
sequential(plan, lat); torch.cuda.synchronize()
t0 = time.perf_counter()
anchors_only = path.sample(ctx, noise[:, plan.anchors], torch.cat([extr[:, :1], extr[:, [1 + a for a in plan.anchors]]], 1),
                           torch.cat([intr[:, :1], intr[:, [1 + a for a in plan.anchors]]], 1))
sequential(plan, lat); torch.cuda.synchronize(); ts = time.perf_counter() - t0
n = int(done.sum())
print(f"anchored 80-frame trajectory, {STEPS} DDIM steps, cfg={use_cfg}: {plan.num_sample_calls} sample() calls in the reference's order; "
      f"{n} target frames generated")
print(f"  batched chunks (this package): {tb * 1e3:8.1f} ms per trajectory  = {n / tb:7.1f} frames/s")
print(f"  one sample() per chunk       : {ts * 1e3:8.1f} ms per trajectory  = {n / ts:7.1f} frames/s")
