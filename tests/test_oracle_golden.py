"""CPU: the oracle restatement against the golden vectors produced by the reference's own python
(oracle/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import torch

from helpers import GOLD, rel_err
from oracle import mvldm_oracle as O


def test_param_table_matches_golden_module_count(oracle_cfg):
    shapes = O.param_shapes(oracle_cfg)
    assert len(shapes) == 494
    assert sum(int(np.prod(s)) for s, _ in shapes.values()) == 764_089_924     # 764.09 M


def test_forward_v4_matches_reference(oracle_cfg, oracle_weights):
    g = np.load(os.path.join(GOLD, "g1_forward_v4.npz"))
    taps = {}
    with torch.no_grad():
        y = O.unet_forward(oracle_weights, torch.tensor(g["inputs"]), torch.tensor(g["timesteps"]), oracle_cfg, taps)
    assert y.shape == (1, 4, 4, 32, 32)
    assert rel_err(y, torch.tensor(g["eps"])) < 1e-4            # fp32 vs the reference module's fp32
    for k, v in taps.items():                                    # per-block regression fingerprints
        ref = g["tap/" + k]
        f = v.flatten().double()
        got = np.array([f.mean().item(), f.std().item(), f.abs().max().item()])
        np.testing.assert_allclose(got, ref[:3], rtol=2e-4, atol=1e-5, err_msg=k)


def test_ddim_tables():
    g = np.load(os.path.join(GOLD, "g4_ddim.npz"))
    s = O.DDIMOracle()
    np.testing.assert_array_equal(s.alphas_cumprod.numpy(), g["alphas_cumprod"])
    for n in (25, 50, 70):
        s.set_timesteps(n)
        np.testing.assert_array_equal(s.timesteps.numpy(), g[f"timesteps_{n}"])
    s.set_timesteps(25)
    assert s.timesteps[0] == 960 and s.timesteps[-1] == 0
    x, e = torch.tensor(g["step_x"]), torch.tensor(g["step_eps"])
    np.testing.assert_array_equal(s.step(e, 960, x).numpy(), g["step_out_960"])
    np.testing.assert_array_equal(s.step(e, 0, x).numpy(), g["step_out_0"])
    # last step lands exactly on x0 (alpha_prev = 1): prev = (x - sqrt(1-a) e)/sqrt(a)
    a = s.alphas_cumprod[0]
    np.testing.assert_allclose(g["step_out_0"], ((x - (1 - a) ** 0.5 * e) / a ** 0.5).numpy(), rtol=1e-6)


def test_raymap_matches_reference_projection():
    g = np.load(os.path.join(GOLD, "g5_rays.npz"))
    extr, intr = torch.tensor(g["extr"]), torch.tensor(g["intr"])
    for pl in (0, 1):
        r = O.raymap(extr, intr, 32, 32, bool(pl))
        np.testing.assert_allclose(r.numpy(), g[f"rays_plucker{pl}"], atol=1e-6)
    d = O.raymap(extr, intr, 32, 32)[:, :, 3:]
    np.testing.assert_allclose(d.norm(dim=2).numpy(), 1.0, atol=1e-5)   # unit directions


def test_trajectory_first_step_matches_reference(oracle_cfg, oracle_weights):
    g = np.load(os.path.join(GOLD, "g3_traj25_cfg0.npz"))
    ctx, x_T = torch.tensor(g["context_latents"]), torch.tensor(g["x_T"])
    extr, intr = torch.tensor(g["extr"]), torch.tensor(g["intr"])
    sched = O.DDIMOracle()
    sched.set_timesteps(25)
    rays = O.raymap(extr, intr, 32, 32)
    cin = torch.cat([ctx, torch.zeros(1, 2, 1, 32, 32)], 2)
    with torch.no_grad():
        x1, eps = O.ddim_step(oracle_weights, oracle_cfg, sched, x_T, 960, cin, rays, torch.ones(1, 6, 1, 32, 32))
    assert rel_err(eps, torch.tensor(g["eps_step0"])) < 1e-4
    assert rel_err(x1, torch.tensor(g["x_after_step0"])) < 1e-4


def test_variant_b_matches_reference():
    """Variant B (SD-2.1 topology, `pretrained_from` set): mvunet.py:118-131,150-160 — per-view Transformer2DModel after
    each resnet of down blocks 0-2 and between the two mid resnets, cross-attending to one zero token; up-block
    attentions exist in the state dict (920 keys) but never run."""
    cfg = O.OracleCfg(variant_b=True)
    shapes = O.param_shapes(cfg)
    assert len(shapes) == 920
    assert sum(k.startswith("unet.up_blocks.") and ".attentions." in k for k in shapes) > 0
    g = np.load(os.path.join(GOLD, "g6_forward_variant_b_v4.npz"))
    sd = O.init_weights(cfg, seed=0)
    with torch.no_grad():
        y = O.unet_forward(sd, torch.tensor(g["inputs"]), torch.tensor(g["timesteps"]), cfg)
    assert rel_err(y, torch.tensor(g["eps"])) < 1e-4


def test_standard_transformer_matches_reference():
    """multi_view_attention "standard" (mv_unet.yaml's default): the oracle's restatement of StandardTransformer
    (standard/transformer.py:45-136, transformer/{transformer,attention,feed_forward,pre_norm}.py) against golden g7, the
    reference module's own output (oracle/make_golden.py golden_standard)."""
    cfg = O.OracleCfg(mv_block="standard")
    shapes = O.param_shapes(cfg)
    assert "cross_attn_blocks_mid.0.transformer.layers.0.0.fn.to_qkv.weight" in shapes
    assert shapes["cross_attn_blocks_encoder.1.transformer.layers.0.1.fn.net.3.weight"][0] == (640, 640)
    g = np.load(os.path.join(GOLD, "g7_forward_standard_v4.npz"))
    sd = O.init_weights(cfg, seed=0)
    taps = {}
    with torch.no_grad():
        y = O.unet_forward(sd, torch.tensor(g["inputs"]), torch.tensor(g["timesteps"]), cfg, taps)
    assert rel_err(y, torch.tensor(g["eps"])) < 1e-4
    assert sum(k.endswith(".attn") for k in taps) == 9


def test_ray_encoding_matches_reference():
    """use_ray_encoding: true: the oracle's positional encoding of the ray maps against golden g8 (reference projection.py +
    reference PositionalEncoding, oracle/make_golden.py golden_ray_encoding)"""
    g = np.load(os.path.join(GOLD, "g8_ray_encoding.npz"))
    extr, intr = torch.tensor(g["extr"]), torch.tensor(g["intr"])
    for fo, fd in ((10, 8), (4, 0)):
        ref = torch.tensor(g[f"rays_{fo}_{fd}"])
        got = O.raymap(extr, intr, ref.shape[-2], ref.shape[-1], False, fo, fd)
        assert got.shape == ref.shape and (got - ref).abs().max() < 1e-5
    ref = torch.tensor(g["rays_srt_6_5"])
    got = O.raymap(extr, intr, ref.shape[-2], ref.shape[-1], False, 6, 5, srt=True)
    assert got.shape == ref.shape and (got - ref).abs().max() < 1e-5
