"""Generate tests/golden/*.npz by executing the REFERENCE's own python.  TEST INFRASTRUCTURE ONLY.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py [--skip-trajectory]

What runs unchanged from /root/reference: ``src/model/denoiser/mvunet.py`` (``MultiViewUNet``),
``src/model/denoiser/mvdream/attention.py`` (``SpatialTransformer3D``), ``src/geometry/projection.py``
(``get_world_rays``, ``sample_image_grid``), ``src/misc/camera_utils.py``.  ``diffusers`` is the shim in
``oracle/diffusers_shim``.  ``DiffusionWrapper.step/sample`` cannot be imported
(hydra / lightning / moviepy absent): their 40 lines are restated in ``oracle/mvldm_oracle.py`` and driven
here with the reference module as the denoiser.

Every golden is produced by the reference module; the script also asserts that the oracle restatement
(`oracle/mvldm_oracle.py`) agrees with it, which is what pins the oracle.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from einops import rearrange  # noqa: E402
from src.geometry.projection import get_world_rays, sample_image_grid  # noqa: E402
from src.misc.camera_utils import absolute_to_relative_camera  # noqa: E402
from src.model.denoiser.mvdream.attention import SpatialTransformer3DCfg  # noqa: E402
from src.model.denoiser.mvunet import MultiViewUNet, MultiViewUNetCfg, UNet2DModelCfg  # noqa: E402

from oracle import mvldm_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def build_reference(sd):
    ucfg = UNet2DModelCfg("unet", ["DownBlock2D"] * 4, "UNetMidBlock2D", ["UpBlock2D"] * 4, False,
                          [320, 640, 1280, 1280])
    mcfg = MultiViewUNetCfg("mv_unet", ucfg, SpatialTransformer3DCfg("spatial_transformer_3d", num_heads=8),
                            use_ray_encoding=False)
    m = MultiViewUNet(mcfg, 11, 4)
    m.load_state_dict(sd, strict=True)   # strict: proves the oracle's key names/shapes == the module's
    return m.eval()


def reference_rays(extr, intr, h, w, plucker):
    """diffusion_wrapper.py:169-190,301-322 driven with the reference's projection.py."""
    xy, _ = sample_image_grid((h, w))
    o, d = get_world_rays(rearrange(xy, "h w xy -> (h w) xy"),
                          rearrange(extr, "b v i j -> b v () i j"),
                          rearrange(intr, "b v i j -> b v () i j"))
    if plucker:
        o = torch.cross(o, d, dim=-1)
    return rearrange(torch.cat([o, d], dim=-1), "b v (h w) c -> b v c h w", h=h, w=w)


def stats(t):
    f = t.flatten().double()
    idx = torch.linspace(0, f.numel() - 1, 64).long()
    return np.concatenate([[f.mean().item(), f.std().item(), f.abs().max().item()], f[idx].numpy()])


def golden_ray_encoding():
    """G8: `use_ray_encoding: true` (config/main.yaml:28-33, 10 origin / 8 direction octaves): DiffusionWrapper.ray_encode
    (diffusion_wrapper.py:301-322) with the reference's own projection.py AND its own PositionalEncoding class."""
    from src.model.encodings.positional_encoding import PositionalEncoding
    extr8, intr8 = O.synthetic_cameras(1, 4)
    h, w = 16, 24                                        # small fixture; non-square on purpose
    xy, _ = sample_image_grid((h, w))
    o, d = get_world_rays(rearrange(xy, "h w xy -> (h w) xy"), rearrange(extr8, "b v i j -> b v () i j"),
                          rearrange(intr8, "b v i j -> b v () i j"))
    out = {"extr": extr8.numpy(), "intr": intr8.numpy()}
    for fo, fd in ((10, 8), (4, 0)):
        oe = PositionalEncoding(fo)(o) if fo > 0 else o
        de = PositionalEncoding(fd)(d) if fd > 0 else d
        r_ref = rearrange(torch.cat([oe, de], dim=-1), "b v (h w) c -> b v c h w", h=h, w=w)
        r_ora = O.raymap(extr8, intr8, h, w, False, fo, fd)
        err = (r_ref - r_ora).abs().max().item()
        print(f"ray encoding octaves ({fo},{fd}): {tuple(r_ref.shape)}  max|ref-oracle| = {err:.3e}")
        assert r_ref.shape == r_ora.shape and err < 1e-4
        out[f"rays_{fo}_{fd}"] = r_ref.numpy()
    # srt_ray_encoding: true - the reference's own SRT RayEncoder on "(b v) (h w) c" inputs (diffusion_wrapper.py:311-315)
    from src.model.srt.layers import RayEncoder
    enc = RayEncoder(pos_octaves=6, ray_octaves=5)
    r = enc(rearrange(o, "b v n c -> (b v) n c"), rearrange(d, "b v n c -> (b v) n c"))
    r_ref = rearrange(r, "(b v) (h w) c -> b v c h w", b=1, h=h, w=w)
    r_ora = O.raymap(extr8, intr8, h, w, False, 6, 5, srt=True)
    err = (r_ref - r_ora).abs().max().item()
    print(f"SRT ray encoder (6,5): {tuple(r_ref.shape)}  max|ref-oracle| = {err:.3e}")
    assert r_ref.shape == r_ora.shape and err < 1e-4
    out["rays_srt_6_5"] = r_ref.numpy()
    np.savez_compressed(os.path.join(GOLD, "g8_ray_encoding.npz"), **out)


def golden_standard():
    """G7: the reference's DEFAULT multi-view block (config/model/denoiser/mv_unet.yaml:5 -> standard_attention.yaml):
    `StandardTransformer` at the 9 multi-view positions, V = 4 (2 context + 2 target), fp32 CPU.  The reference's
    Attention pins SDPA to the EFFICIENT_ATTENTION backend (transformer/attention.py:94), which has no CPU
    implementation; the context manager is made a no-op for this call only (same math: softmax(q k^T / sqrt(d)) v;
    the reference source is untouched)."""
    import contextlib
    import src.model.transformer.attention as ref_attn
    from src.model.denoiser.standard.transformer import CrossAttentionCfg
    cfg_s = O.OracleCfg(mv_block="standard")
    sd_s = O.init_weights(cfg_s, seed=0)
    ucfg = UNet2DModelCfg("unet", ["DownBlock2D"] * 4, "UNetMidBlock2D", ["UpBlock2D"] * 4, False, [320, 640, 1280, 1280])
    mcfg = MultiViewUNetCfg("mv_unet", ucfg, CrossAttentionCfg("standard", num_heads=8, d_mlp_multiplier=1),
                            use_ray_encoding=False)
    ref_s = MultiViewUNet(mcfg, 11, 4)
    ref_s.load_state_dict(sd_s, strict=True)      # strict: the oracle's / library's key names == the module's
    ref_s.eval()
    ctx, x_T, extr, intr = O.synthetic_scene(1, 2, 2)
    rays = reference_rays(extr, intr, 32, 32, False)
    inp, _ = O.build_inputs(x_T, torch.cat([ctx, torch.zeros(1, 2, 1, 32, 32)], 2), rays, torch.ones(1, 2, 1, 32, 32))
    ts = torch.tensor([[0, 0, 500, 500]])
    orig = ref_attn.sdpa_kernel
    ref_attn.sdpa_kernel = lambda *a, **k: contextlib.nullcontext()
    try:
        t0 = time.time()
        y_ref = ref_s.forward(inp, ts)
    finally:
        ref_attn.sdpa_kernel = orig
    taps = {}
    y_ora = O.unet_forward(sd_s, inp, ts, cfg_s, taps)
    err = (y_ref - y_ora).abs().max().item()
    print(f"g7_forward_standard_v4: ref {time.time() - t0:.1f}s  max|ref-oracle| = {err:.3e}  std {y_ref.std().item():.4f}")
    assert err < 1e-4 * y_ref.abs().max().item() + 1e-5
    out = {"inputs": inp.numpy(), "timesteps": ts.numpy(), "eps": y_ref.numpy()}
    for k, v in taps.items():
        out["tap/" + k] = stats(v)
    np.savez_compressed(os.path.join(GOLD, "g7_forward_standard_v4.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-trajectory", action="store_true")
    ap.add_argument("--skip-variant-b", action="store_true")
    ap.add_argument("--only-standard", action="store_true", help="(re)generate only g7 (StandardTransformer blocks)")
    ap.add_argument("--only-ray-encoding", action="store_true", help="(re)generate only g8 (positionally encoded ray maps)")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_grad_enabled(False)
    golden_ray_encoding()
    if args.only_ray_encoding:
        return
    golden_standard()
    if args.only_standard:
        return
    cfg = O.OracleCfg()
    sd = O.init_weights(cfg, seed=0)
    ref = build_reference(sd)

    # ---- G4: DDIM tables -------------------------------------------------------------------
    sched = O.DDIMOracle()
    out = {"alphas_cumprod": sched.alphas_cumprod.numpy()}
    for n in (25, 50, 70):
        sched.set_timesteps(n)
        out[f"timesteps_{n}"] = sched.timesteps.numpy()
        out[f"coef_{n}"] = np.array([sched.coefficients(int(t)) for t in sched.timesteps], dtype=np.float64)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 3, 4, 8, 8, generator=g)
    e = torch.randn(1, 3, 4, 8, 8, generator=g)
    sched.set_timesteps(25)
    out["step_x"], out["step_eps"] = x.numpy(), e.numpy()
    out["step_out_960"] = sched.step(e, 960, x).numpy()
    out["step_out_0"] = sched.step(e, 0, x).numpy()
    np.savez_compressed(os.path.join(GOLD, "g4_ddim.npz"), **out)

    # ---- G5: ray maps vs the reference's projection.py ------------------------------------------
    extr8, intr8 = O.synthetic_cameras(1, 8)
    # (the oracle's synthetic_cameras already applies absolute_to_relative; check that too)
    raw = torch.eye(4).expand(1, 8, 4, 4).clone()
    assert torch.allclose(absolute_to_relative_camera(raw, 0), O.absolute_to_relative(raw, 0))
    out = {"extr": extr8.numpy(), "intr": intr8.numpy()}
    for pl in (False, True):
        r_ref = reference_rays(extr8, intr8, 32, 32, pl)
        r_ora = O.raymap(extr8, intr8, 32, 32, pl)
        err = (r_ref - r_ora).abs().max().item()
        print(f"raymap plucker={pl}: max|ref-oracle| = {err:.3e}")
        assert err < 1e-5
        out[f"rays_plucker{int(pl)}"] = r_ref.numpy()
    np.savez_compressed(os.path.join(GOLD, "g5_rays.npz"), **out)

    # ---- G1 / G2: single forwards ----------------------------------------------------------
    for name, (v_c, v_t, t) in {"g1_forward_v4": (2, 2, 500), "g2_forward_v8": (2, 6, 500)}.items():
        ctx, x_T, extr, intr = O.synthetic_scene(1, v_c, v_t)
        rays = reference_rays(extr, intr, 32, 32, False)
        cin = torch.cat([ctx, torch.zeros(1, v_c, 1, 32, 32)], 2)
        inp, _ = O.build_inputs(x_T, cin, rays, torch.ones(1, v_t, 1, 32, 32))
        ts = torch.tensor([[0] * v_c + [t] * v_t])
        t0 = time.time()
        y_ref = ref.forward(inp, ts)
        t_ref = time.time() - t0
        taps = {}
        y_ora = O.unet_forward(sd, inp, ts, cfg, taps)
        err = (y_ref - y_ora).abs().max().item()
        print(f"{name}: ref {t_ref:.1f}s  max|ref-oracle| = {err:.3e}  std {y_ref.std().item():.4f}")
        assert err < 1e-4 * y_ref.abs().max().item() + 1e-5
        out = {"inputs": inp.numpy(), "timesteps": ts.numpy(), "eps": y_ref.numpy()}
        for k, v in taps.items():
            out["tap/" + k] = stats(v)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)

    # ---- G6: Variant B (SD-2.1 topology, random init): per-view Transformer2DModels + the quirks of mvunet.py ----
    # mvunet.py hard-codes `.cuda()` on the zero context (lines 125-128,155-157); there is no GPU here, so Tensor.cuda
    # is made a no-op for this call only (the reference source itself is untouched).
    if not args.skip_variant_b:
        cfg_b = O.OracleCfg(variant_b=True)
        sd_b = O.init_weights(cfg_b, seed=0)
        ucfg = UNet2DModelCfg("unet", ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"], "UNetMidBlock2DCrossAttn",
                              ["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3, False, [320, 640, 1280, 1280])
        mcfg = MultiViewUNetCfg("mv_unet", ucfg, SpatialTransformer3DCfg("spatial_transformer_3d", num_heads=8),
                                use_ray_encoding=False, pretrained_from="stabilityai/stable-diffusion-2-1")
        ref_b = MultiViewUNet(mcfg, 11, 4)
        ref_b.load_state_dict(sd_b, strict=True)
        ref_b.eval()
        ctx, x_T, extr, intr = O.synthetic_scene(1, 2, 2)
        rays = reference_rays(extr, intr, 32, 32, False)
        inp, _ = O.build_inputs(x_T, torch.cat([ctx, torch.zeros(1, 2, 1, 32, 32)], 2), rays, torch.ones(1, 2, 1, 32, 32))
        ts = torch.tensor([[0, 0, 500, 500]])
        orig_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            t0 = time.time()
            y_ref = ref_b.forward(inp, ts)
        finally:
            torch.Tensor.cuda = orig_cuda
        y_ora = O.unet_forward(sd_b, inp, ts, cfg_b)
        err = (y_ref - y_ora).abs().max().item()
        print(f"g6_forward_variant_b_v4: ref {time.time() - t0:.1f}s  max|ref-oracle| = {err:.3e}  std {y_ref.std().item():.4f}")
        assert err < 1e-4 * y_ref.abs().max().item() + 1e-5
        np.savez_compressed(os.path.join(GOLD, "g6_forward_variant_b_v4.npz"), inputs=inp.numpy(), timesteps=ts.numpy(),
                            eps=y_ref.numpy())
        del ref_b, sd_b

    if args.skip_trajectory:
        return
    # ---- G3: 25-step DDIM trajectories, reference module as the denoiser -------------------------
    ctx, x_T, extr, intr = O.synthetic_scene(1, 2, 6)
    fwd = lambda lat, t: ref.forward(lat, t)  # noqa: E731
    for use_cfg in (False, True):
        rec = []
        t0 = time.time()
        x0 = O.sample(sd, cfg, ctx, x_T, extr, intr, 25, use_cfg, 3.0, False, forward=fwd, record=rec)
        print(f"trajectory cfg={use_cfg}: {time.time() - t0:.0f}s  final std {x0.std().item():.4f}")
        out = {"context_latents": ctx.numpy(), "x_T": x_T.numpy(), "extr": extr.numpy(), "intr": intr.numpy(),
               "x_0": x0.numpy(), "timesteps": np.array([r[0] for r in rec]),
               "x_stats": np.stack([stats(r[1]) for r in rec]), "eps_stats": np.stack([stats(r[2]) for r in rec])}
        for i in (0, 4, 9, 14, 19, 24):
            out[f"x_after_step{i}"] = rec[i][1].numpy()
            out[f"eps_step{i}"] = rec[i][2].numpy()
        np.savez_compressed(os.path.join(GOLD, f"g3_traj25_cfg{int(use_cfg)}.npz"), **out)


if __name__ == "__main__":
    main()
