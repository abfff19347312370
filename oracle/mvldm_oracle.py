"""CPU oracle for the MV-LDM denoising hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``mvldm_b200``) never does.

This file is a functional (state-dict in, tensors out) fp32 restatement in plain PyTorch-CPU of

* ``MultiViewUNet.forward``            reference ``src/model/denoiser/mvunet.py:90-208``
* ``SpatialTransformer3D`` & friends    reference ``src/model/denoiser/mvdream/attention.py:60-100,156-205,257-286,357-439``
* ``DiffusionWrapper.step`` / ``sample`` / ``ray_encode`` / ``generate_image_rays``
                                        reference ``src/model/diffusion_wrapper.py:169-190,301-322,413-490``
* ``get_world_rays`` / ``sample_image_grid``   reference ``src/geometry/projection.py:74-138``
* ``absolute_to_relative_camera``       reference ``src/misc/camera_utils.py:7-27``
* the pieces of the un-vendored dependency ``diffusers==0.27.2`` (reference ``requirements.txt:8``)
  that the path touches: ``UNet2DConditionModel`` (DownBlock2D / UNetMidBlock2D / UpBlock2D,
  ``ResnetBlock2D``, ``Downsample2D``, ``Upsample2D``, ``Timesteps``, ``TimestepEmbedding``) and
  ``DDIMScheduler`` (``set_timesteps``, ``step``, ``add_noise``).  Their published arithmetic is
  restated from SURVEY.md Appendix A; call sites: ``mvunet.py:54-63,107-113,121,147,150,177,200,203-205``,
  ``scheduler/__init__.py:37``, ``diffusion_wrapper.py:198,370,417,451,474``.

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned against the
reference's OWN python (``mvunet.py`` + ``mvdream/attention.py`` executed unchanged on a minimal
``diffusers`` shim, ``projection.py`` / ``camera_utils.py`` imported unchanged) by ``oracle/make_golden.py``,
which also writes ``tests/golden/*.npz``.  The diffusers arithmetic itself (Appendix A) cannot be diffed
against the real package in this sandbox (not installed, no network): for that part parity is
"restated from the published algorithm", shared by shim and oracle.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration (Variant A of SURVEY.md §0: pretrained_from=None, DownBlock2D x4 / UpBlock2D x4)
# --------------------------------------------------------------------------------------
@dataclass
class OracleCfg:
    in_channels: int = 11          # 4 latent + 1 mask + 3 origin/moment + 3 direction (diffusion_wrapper.py:98-129)
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)   # config/model/denoiser/mv_unet.yaml:10
    layers_per_block: int = 2      # diffusers default
    norm_groups: int = 32
    num_heads: int = 8             # spatial_transformer_3d.yaml:5
    temb_dim_mult: int = 4         # time_embed_dim = 4 * block_out_channels[0]
    max_attn_res: int = 32         # mvunet.py:137,190  (h<=32 and w<=32)
    # Variant B: SD-2.1 topology = per-view Transformer2DModel after every resnet of down blocks
    # 0-2 and inside the mid block (attention_head_dim = heads, head dim 64, cross_attention_dim 1024); the up-block
    # transformers exist in the state dict but are never executed (mvunet.py:178)
    variant_b: bool = False
    t2d_heads: Tuple[int, ...] = (5, 10, 20, 20)
    cross_attention_dim: int = 1024
    # multi_view_attention.name (src/model/denoiser/attention.py:8-27): "spatial_transformer_3d" (released experiment) or
    # "standard" (mv_unet.yaml's default): StandardTransformer, standard/transformer.py:45-136
    mv_block: str = "spatial_transformer_3d"
    mv_num_layers: int = 1
    mv_d_mlp_multiplier: int = 1       # multi_view_attention/standard_attention.yaml:8

    @property
    def temb_dim(self) -> int:
        return self.block_out_channels[0] * self.temb_dim_mult


# --------------------------------------------------------------------------------------
# deterministic weights.  Keys are exactly the reference module's state_dict keys
# (SURVEY.md §3.3 / Appendix A), so the same dict loads (strict) into the reference's MultiViewUNet.
# --------------------------------------------------------------------------------------
def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2**63 - 1))
    return g


def _uniform(seed, key, shape, bound):
    return (torch.rand(shape, generator=_gen(seed, key), dtype=torch.float32) * 2 - 1) * bound


def _normal(seed, key, shape, mean, std):
    return torch.randn(shape, generator=_gen(seed, key), dtype=torch.float32) * std + mean


def param_shapes(cfg: OracleCfg) -> "Dict[str, Tuple[Tuple[int, ...], str]]":
    """key -> (shape, kind) for every parameter of the Variant-A MultiViewUNet, in a fixed order."""
    P: Dict[str, Tuple[Tuple[int, ...], str]] = {}
    boc = cfg.block_out_channels
    T = cfg.temb_dim

    def conv(k, cout, cin, ks):
        P[k + ".weight"] = ((cout, cin, ks, ks), "w")
        P[k + ".bias"] = ((cout,), "b")

    def lin(k, cout, cin, bias=True):
        P[k + ".weight"] = ((cout, cin), "w")
        if bias:
            P[k + ".bias"] = ((cout,), "b")

    def norm(k, c):
        P[k + ".weight"] = ((c,), "g")
        P[k + ".bias"] = ((c,), "beta")

    def resnet(k, cin, cout):
        norm(k + ".norm1", cin)
        conv(k + ".conv1", cout, cin, 3)
        lin(k + ".time_emb_proj", cout, T)
        norm(k + ".norm2", cout)
        conv(k + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(k + ".conv_shortcut", cout, cin, 1)

    def mvblock(k, c):
        if cfg.mv_block == "standard":   # Transformer(c, depth, heads, c // heads, c * mult): transformer/transformer.py:24-47
            for j in range(cfg.mv_num_layers):
                lk = f"{k}.transformer.layers.{j}"
                norm(lk + ".0.norm", c)
                lin(lk + ".0.fn.to_qkv", 3 * c, c, bias=False)
                lin(lk + ".0.fn.to_out.0", c, c)
                norm(lk + ".1.norm", c)
                lin(lk + ".1.fn.net.0", c * cfg.mv_d_mlp_multiplier, c)
                lin(lk + ".1.fn.net.3", c, c * cfg.mv_d_mlp_multiplier)
            return
        norm(k + ".norm", c)
        conv(k + ".proj_in", c, c, 1)
        tb = k + ".transformer_blocks.0"
        for a in ("attn1", "attn2"):
            lin(f"{tb}.{a}.to_q", c, c, bias=False)
            lin(f"{tb}.{a}.to_k", c, c, bias=False)
            lin(f"{tb}.{a}.to_v", c, c, bias=False)
            lin(f"{tb}.{a}.to_out.0", c, c)
        lin(f"{tb}.ff.net.0.proj", 8 * c, c)
        lin(f"{tb}.ff.net.2", c, 4 * c)
        for n in ("norm1", "norm2", "norm3"):
            norm(f"{tb}.{n}", c)
        conv(k + ".proj_out", c, c, 1)

    def t2d(k, c):
        norm(k + ".norm", c)
        lin(k + ".proj_in", c, c)
        tb = k + ".transformer_blocks.0"
        for n in ("norm1", "norm2", "norm3"):
            norm(f"{tb}.{n}", c)
        for a, kv in (("attn1", c), ("attn2", cfg.cross_attention_dim)):
            lin(f"{tb}.{a}.to_q", c, c, bias=False)
            lin(f"{tb}.{a}.to_k", c, kv, bias=False)
            lin(f"{tb}.{a}.to_v", c, kv, bias=False)
            lin(f"{tb}.{a}.to_out.0", c, c)
        lin(f"{tb}.ff.net.0.proj", 8 * c, c)
        lin(f"{tb}.ff.net.2", c, 4 * c)
        lin(k + ".proj_out", c, c)

    conv("unet.conv_in", boc[0], cfg.in_channels, 3)
    lin("unet.time_embedding.linear_1", T, boc[0])
    lin("unet.time_embedding.linear_2", T, T)
    # down
    cout = boc[0]
    for l, c in enumerate(boc):
        cin, cout = cout, c
        for i in range(cfg.layers_per_block):
            resnet(f"unet.down_blocks.{l}.resnets.{i}", cin if i == 0 else cout, cout)
            if cfg.variant_b and l != len(boc) - 1:
                t2d(f"unet.down_blocks.{l}.attentions.{i}", cout)
        if l != len(boc) - 1:
            conv(f"unet.down_blocks.{l}.downsamplers.0.conv", cout, cout, 3)
    resnet("unet.mid_block.resnets.0", boc[-1], boc[-1])
    if cfg.variant_b:
        t2d("unet.mid_block.attentions.0", boc[-1])
        resnet("unet.mid_block.resnets.1", boc[-1], boc[-1])
    # up
    rev = list(reversed(boc))
    out_c = rev[0]
    for l in range(len(boc)):
        prev, out_c = out_c, rev[l]
        in_c = rev[min(l + 1, len(boc) - 1)]
        for i in range(cfg.layers_per_block + 1):
            skip = in_c if i == cfg.layers_per_block else out_c
            rin = prev if i == 0 else out_c
            resnet(f"unet.up_blocks.{l}.resnets.{i}", rin + skip, out_c)
            if cfg.variant_b and l != 0:
                t2d(f"unet.up_blocks.{l}.attentions.{i}", out_c)   # present in the checkpoint, never executed
        if l != len(boc) - 1:
            conv(f"unet.up_blocks.{l}.upsamplers.0.conv", out_c, out_c, 3)
    norm("unet.conv_norm_out", boc[0])
    conv("unet.conv_out", cfg.out_channels, boc[0], 3)
    for l, c in enumerate(boc):
        mvblock(f"cross_attn_blocks_encoder.{l}", c)
    mvblock("cross_attn_blocks_mid.0", boc[-1])
    for l, c in enumerate(rev):
        mvblock(f"cross_attn_blocks_decoder.{l}", c)
    return P


def init_weights(cfg: OracleCfg, seed: int = 0) -> Dict[str, Tensor]:
    """Random-init state dict.  Conv/Linear ~ U(+-1/sqrt(fan_in)) (torch default scale); the zero-init
    ``proj_out`` of every multi-view block (mvdream/attention.py:406-411) is re-randomised like any other
    conv so the attention branch is exercised; norm affine params are perturbed around
    (1, 0) so that gamma/beta handling is tested.  Each tensor has its own generator keyed on
    (seed, name): the result does not depend on construction order or platform."""
    sd: Dict[str, Tensor] = {}
    shapes = param_shapes(cfg)
    for k, (shape, kind) in shapes.items():
        if kind == "w":
            fan_in = int(math.prod(shape[1:]))
            sd[k] = _uniform(seed, k, shape, 1.0 / math.sqrt(fan_in))
        elif kind == "b":
            wshape = shapes[k[: -len("bias")] + "weight"][0]
            fan_in = int(math.prod(wshape[1:]))
            sd[k] = _uniform(seed, k, shape, 1.0 / math.sqrt(fan_in))
        elif kind == "g":
            sd[k] = _normal(seed, k, shape, 1.0, 0.1)
        else:
            sd[k] = _normal(seed, k, shape, 0.0, 0.1)
    return sd


# --------------------------------------------------------------------------------------
# diffusers pieces (Appendix A)
# --------------------------------------------------------------------------------------
def timestep_sinusoid(t: Tensor, dim: int) -> Tensor:
    """diffusers ``Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0)``; mvunet.py:107."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    arg = t.to(torch.float32)[:, None] * freqs[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


def time_embedding(sd, t_flat: Tensor, cfg: OracleCfg) -> Tensor:
    """``TimestepEmbedding``: linear_1 -> SiLU -> linear_2; mvunet.py:108."""
    e = timestep_sinusoid(t_flat, cfg.block_out_channels[0])
    e = F.linear(e, sd["unet.time_embedding.linear_1.weight"], sd["unet.time_embedding.linear_1.bias"])
    e = F.silu(e)
    return F.linear(e, sd["unet.time_embedding.linear_2.weight"], sd["unet.time_embedding.linear_2.bias"])


def resnet_block(sd, k: str, x: Tensor, emb: Tensor, groups: int) -> Tensor:
    """``ResnetBlock2D.forward`` (eps=1e-5, silu, dropout 0, output_scale_factor 1)."""
    h = F.group_norm(x, groups, sd[k + ".norm1.weight"], sd[k + ".norm1.bias"], eps=1e-5)
    h = F.conv2d(F.silu(h), sd[k + ".conv1.weight"], sd[k + ".conv1.bias"], padding=1)
    t = F.linear(F.silu(emb), sd[k + ".time_emb_proj.weight"], sd[k + ".time_emb_proj.bias"])
    h = h + t[:, :, None, None]
    h = F.group_norm(h, groups, sd[k + ".norm2.weight"], sd[k + ".norm2.bias"], eps=1e-5)
    h = F.conv2d(F.silu(h), sd[k + ".conv2.weight"], sd[k + ".conv2.bias"], padding=1)
    if (k + ".conv_shortcut.weight") in sd:
        x = F.conv2d(x, sd[k + ".conv_shortcut.weight"], sd[k + ".conv_shortcut.bias"])
    return x + h


# --------------------------------------------------------------------------------------
# multi-view block (mvdream/attention.py)
# --------------------------------------------------------------------------------------
def _attention(sd, k: str, x: Tensor, heads: int) -> Tensor:
    """``CrossAttention.forward`` with context=None (self-attention), mvdream/attention.py:174-205:
    bias-free q/k/v, fp32 scores * d^-0.5, softmax, PV, to_out.0 (with bias)."""
    b, n, c = x.shape
    d = c // heads
    q = F.linear(x, sd[k + ".to_q.weight"])
    kk = F.linear(x, sd[k + ".to_k.weight"])
    v = F.linear(x, sd[k + ".to_v.weight"])

    def split(t):
        return t.reshape(b, n, heads, d).permute(0, 2, 1, 3).reshape(b * heads, n, d)

    q, kk, v = split(q), split(kk), split(v)
    out = torch.empty_like(q)
    scale = d ** -0.5
    # chunk queries so the N x N score matrix never exceeds ~256 MB (the reference materialises it whole)
    step = max(1, min(n, (64 * 1024 * 1024) // max(n, 1)))
    for bh in range(b * heads):
        for s in range(0, n, step):
            sim = (q[bh, s:s + step].float() @ kk[bh].float().t()) * scale
            out[bh, s:s + step] = sim.softmax(dim=-1) @ v[bh]
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, c)
    return F.linear(out, sd[k + ".to_out.0.weight"], sd[k + ".to_out.0.bias"])


def standard_block(sd, k: str, x: Tensor, b: int, v: int, heads: int, layers: int,
                   taps: Optional[dict] = None) -> Tensor:
    """``StandardTransformer.forward`` (standard/transformer.py:96-136; downscale 1, pos_enc off) around
    ``Transformer.forward`` (transformer/transformer.py:68-72): per layer ``x = attn(LN(x)) + x; x = ff(LN(x)) + x`` with
    ``Attention`` (transformer/attention.py:59-99: fused bias-free to_qkv, chunk(3), softmax(q k^T / sqrt(d)) v over ALL
    (view, pixel) tokens of the scene, to_out Linear) and ``FeedForward`` (feed_forward.py:28-39: Linear, exact GELU,
    Linear).  ``x`` is [(b v), c, h, w]."""
    bv, c, h, w = x.shape
    d = c // heads
    y = x.reshape(b, v, c, h, w).permute(0, 1, 3, 4, 2).reshape(b, v * h * w, c)     # "b v c h w -> b (v h w) c"
    for j in range(layers):
        lk = f"{k}.transformer.layers.{j}"
        z = F.layer_norm(y, (c,), sd[lk + ".0.norm.weight"], sd[lk + ".0.norm.bias"], eps=1e-5)
        q, kk, vv = F.linear(z, sd[lk + ".0.fn.to_qkv.weight"]).chunk(3, dim=-1)
        sp = lambda t: t.reshape(b, -1, heads, d).permute(0, 2, 1, 3)                # noqa: E731  "b n (h d) -> b h n d"
        att = torch.softmax(sp(q) @ sp(kk).transpose(-1, -2) * d ** -0.5, dim=-1) @ sp(vv)
        att = att.permute(0, 2, 1, 3).reshape(b, -1, c)
        y = F.linear(att, sd[lk + ".0.fn.to_out.0.weight"], sd[lk + ".0.fn.to_out.0.bias"]) + y
        if taps is not None and j == 0:
            taps[k + ".attn"] = y.reshape(bv, h * w, c)
        z = F.layer_norm(y, (c,), sd[lk + ".1.norm.weight"], sd[lk + ".1.norm.bias"], eps=1e-5)
        z = F.gelu(F.linear(z, sd[lk + ".1.fn.net.0.weight"], sd[lk + ".1.fn.net.0.bias"]))
        y = F.linear(z, sd[lk + ".1.fn.net.3.weight"], sd[lk + ".1.fn.net.3.bias"]) + y
    return y.reshape(b, v, h, w, c).permute(0, 1, 4, 2, 3).reshape(bv, c, h, w)


def _mv_block_3d(sd, k: str, x: Tensor, b: int, v: int, heads: int, groups: int,
                 taps: Optional[dict] = None) -> Tensor:
    """``SpatialTransformer3D.forward`` (mvdream/attention.py:416-439) with one
    ``BasicTransformerBlock3D`` (:362-368).  ``x`` is [(b v), c, h, w]."""
    bv, c, h, w = x.shape
    x_in = x
    y = F.group_norm(x, groups, sd[k + ".norm.weight"], sd[k + ".norm.bias"], eps=1e-6)   # Normalize(): :99-100
    y = F.conv2d(y, sd[k + ".proj_in.weight"], sd[k + ".proj_in.bias"])
    y = y.permute(0, 2, 3, 1).reshape(bv, h * w, c)
    tb = k + ".transformer_blocks.0"

    def ln(t, name):
        return F.layer_norm(t, (c,), sd[f"{tb}.{name}.weight"], sd[f"{tb}.{name}.bias"], eps=1e-5)

    # joint attention over all views' tokens: "(b f) l c -> b (f l) c"
    y = y.reshape(b, v * h * w, c)
    y = _attention(sd, tb + ".attn1", ln(y, "norm1"), heads) + y
    if taps is not None:
        taps[k + ".attn1"] = y.reshape(bv, h * w, c)
    # per-view attention (context=None -> self attention over one view's tokens)
    y = y.reshape(bv, h * w, c)
    y = _attention(sd, tb + ".attn2", ln(y, "norm2"), heads) + y
    if taps is not None:
        taps[k + ".attn2"] = y
    # GEGLU feed-forward: proj -> chunk(x, gate) -> x * gelu(gate) -> linear  (:60-87)
    z = F.linear(ln(y, "norm3"), sd[tb + ".ff.net.0.proj.weight"], sd[tb + ".ff.net.0.proj.bias"])
    a, gate = z.chunk(2, dim=-1)
    z = F.linear(a * F.gelu(gate), sd[tb + ".ff.net.2.weight"], sd[tb + ".ff.net.2.bias"])
    y = z + y
    y = y.reshape(bv, h, w, c).permute(0, 3, 1, 2)
    y = F.conv2d(y, sd[k + ".proj_out.weight"], sd[k + ".proj_out.bias"])
    return y + x_in


def transformer2d(sd, k: str, x: Tensor, heads: int, groups: int, cross_dim: int) -> Tensor:
    """diffusers ``Transformer2DModel`` (use_linear_projection=True, one ``BasicTransformerBlock``) as the reference
    calls it at mvunet.py:131-134,158 with ``encoder_hidden_states = zeros(b*v, 1, 1024)``:
    GN(eps 1e-6) -> Linear -> [LN -> self-attn (head dim 64) -> +; LN -> cross-attn to the zero token -> +;
    LN -> GEGLU FF -> +] -> Linear -> + input."""
    bv, c, h, w = x.shape
    y = F.group_norm(x, groups, sd[k + ".norm.weight"], sd[k + ".norm.bias"], eps=1e-6)
    y = y.permute(0, 2, 3, 1).reshape(bv, h * w, c)
    y = F.linear(y, sd[k + ".proj_in.weight"], sd[k + ".proj_in.bias"])
    tb = k + ".transformer_blocks.0"

    def ln(t, name):
        return F.layer_norm(t, (c,), sd[f"{tb}.{name}.weight"], sd[f"{tb}.{name}.bias"], eps=1e-5)

    y = _attention(sd, tb + ".attn1", ln(y, "norm1"), heads) + y
    # cross-attention over ONE all-zero context token: k = v = 0 (bias-free projections), softmax over one key = 1,
    # so the branch contributes exactly to_out.0(0) = its bias; computed explicitly here
    ctx = torch.zeros(bv, 1, cross_dim, dtype=y.dtype, device=y.device)
    q = F.linear(ln(y, "norm2"), sd[tb + ".attn2.to_q.weight"])
    kk = F.linear(ctx, sd[tb + ".attn2.to_k.weight"])
    vv = F.linear(ctx, sd[tb + ".attn2.to_v.weight"])
    d = c // heads
    sp = lambda t: t.reshape(bv, -1, heads, d).permute(0, 2, 1, 3)  # noqa: E731
    a2 = torch.softmax(sp(q) @ sp(kk).transpose(-1, -2) * d ** -0.5, dim=-1) @ sp(vv)
    a2 = a2.permute(0, 2, 1, 3).reshape(bv, h * w, c)
    y = F.linear(a2, sd[tb + ".attn2.to_out.0.weight"], sd[tb + ".attn2.to_out.0.bias"]) + y
    z = F.linear(ln(y, "norm3"), sd[tb + ".ff.net.0.proj.weight"], sd[tb + ".ff.net.0.proj.bias"])
    a, gate = z.chunk(2, dim=-1)
    y = F.linear(a * F.gelu(gate), sd[tb + ".ff.net.2.weight"], sd[tb + ".ff.net.2.bias"]) + y
    y = F.linear(y, sd[k + ".proj_out.weight"], sd[k + ".proj_out.bias"])
    return y.reshape(bv, h, w, c).permute(0, 3, 1, 2) + x


# --------------------------------------------------------------------------------------
# MultiViewUNet.forward (mvunet.py:90-208), Variant A (and Variant B with cfg.variant_b)
# --------------------------------------------------------------------------------------
def unet_forward(sd: Dict[str, Tensor], latents: Tensor, timestep: Tensor, cfg: OracleCfg,
                 taps: Optional[dict] = None) -> Tensor:
    """latents [B,V,C_in,h,w] fp32, timestep int64 [B] or [B,V] -> eps [B,V,C_out,h,w]."""
    B, V = latents.shape[:2]
    G = cfg.norm_groups
    nb = len(cfg.block_out_channels)
    if timestep.dim() < 2:                       # mvunet.py:102-105
        t_flat = timestep[:, None].expand(B, V).reshape(-1)
    else:
        t_flat = timestep.reshape(-1)
    emb = time_embedding(sd, t_flat, cfg)

    def tap(name, t):
        if taps is not None:
            taps[name] = t

    def mv_block(sd_, k, x_, b, v, heads, groups, taps_):       # multi_view_attention.name picks the block
        if cfg.mv_block == "standard":
            return standard_block(sd_, k, x_, b, v, heads, cfg.mv_num_layers, taps_)
        return _mv_block_3d(sd_, k, x_, b, v, heads, groups, taps_)

    x = latents.reshape(B * V, *latents.shape[2:])
    x = F.conv2d(x, sd["unet.conv_in.weight"], sd["unet.conv_in.bias"], padding=1)
    tap("conv_in", x)
    skips: List[Tensor] = [x]
    for l in range(nb):
        for i in range(cfg.layers_per_block):
            x = resnet_block(sd, f"unet.down_blocks.{l}.resnets.{i}", x, emb, G)
            if cfg.variant_b and l != nb - 1:    # CrossAttnDownBlock2D: per-view transformer after every resnet (:122-134)
                x = transformer2d(sd, f"unet.down_blocks.{l}.attentions.{i}", x, cfg.t2d_heads[l], G, cfg.cross_attention_dim)
            tap(f"down{l}.res{i}", x)
            skips.append(x)                      # recorded BEFORE the multi-view block (:135)
        if x.shape[-2] <= cfg.max_attn_res and x.shape[-1] <= cfg.max_attn_res:
            x = mv_block(sd, f"cross_attn_blocks_encoder.{l}", x, B, V, cfg.num_heads, G, taps)
            tap(f"down{l}.mv", x)
        if l != nb - 1:
            kd = f"unet.down_blocks.{l}.downsamplers.0.conv"
            x = F.conv2d(x, sd[kd + ".weight"], sd[kd + ".bias"], stride=2, padding=1)
            tap(f"down{l}.ds", x)
            skips.append(x)
    x = resnet_block(sd, "unet.mid_block.resnets.0", x, emb, G)
    if cfg.variant_b:                            # UNetMidBlock2DCrossAttn: attn, resnet (:152-159)
        x = transformer2d(sd, "unet.mid_block.attentions.0", x, cfg.t2d_heads[-1], G, cfg.cross_attention_dim)
        x = resnet_block(sd, "unet.mid_block.resnets.1", x, emb, G)
    tap("mid.res0", x)
    x = mv_block(sd, "cross_attn_blocks_mid.0", x, B, V, cfg.num_heads, G, taps)
    tap("mid.mv", x)
    for l in range(nb):
        for i in range(cfg.layers_per_block + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, f"unet.up_blocks.{l}.resnets.{i}", x, emb, G)
            tap(f"up{l}.res{i}", x)
        if x.shape[-2] <= cfg.max_attn_res and x.shape[-1] <= cfg.max_attn_res:
            x = mv_block(sd, f"cross_attn_blocks_decoder.{l}", x, B, V, cfg.num_heads, G, taps)
            tap(f"up{l}.mv", x)
        if l != nb - 1:
            ku = f"unet.up_blocks.{l}.upsamplers.0.conv"
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, sd[ku + ".weight"], sd[ku + ".bias"], padding=1)
            tap(f"up{l}.us", x)
    x = F.group_norm(x, G, sd["unet.conv_norm_out.weight"], sd["unet.conv_norm_out.bias"], eps=1e-5)
    x = F.conv2d(F.silu(x), sd["unet.conv_out.weight"], sd["unet.conv_out.bias"], padding=1)
    return x.reshape(B, V, *x.shape[1:])


# --------------------------------------------------------------------------------------
# DDIM scheduler (diffusers DDIMScheduler with config/model/scheduler/ddim.yaml:7-16)
# --------------------------------------------------------------------------------------
class DDIMOracle:
    """linear betas 1e-4..0.02, 1000 train steps, leading spacing, steps_offset 0, clip_sample False,
    set_alpha_to_one True, eta 0, prediction_type epsilon."""

    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02):
        self.num_train_timesteps = num_train_timesteps
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)
        self.num_inference_steps = None

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        self.timesteps = (torch.arange(0, n, dtype=torch.int64) * ratio).flip(0).contiguous()

    def scale_model_input(self, x, t=None):
        return x

    def coefficients(self, t: int) -> Tuple[float, float, float, float]:
        """(sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)) as python floats (from fp32 tensors)."""
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return (float(a_t ** 0.5), float((1 - a_t) ** 0.5), float(a_p ** 0.5), float((1 - a_p) ** 0.5))

    def step(self, eps: Tensor, t, x: Tensor) -> Tensor:
        t = int(t)
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        direction = (1 - a_p) ** 0.5 * eps          # eta = 0 -> variance 0
        return a_p ** 0.5 * x0 + direction

    def add_noise(self, x: Tensor, noise: Tensor, t: Tensor) -> Tensor:
        a = self.alphas_cumprod[t].to(x.dtype)
        sa = (a ** 0.5).reshape(-1, *([1] * (x.dim() - 1)))
        s1 = ((1 - a) ** 0.5).reshape(-1, *([1] * (x.dim() - 1)))
        return sa * x + s1 * noise


class DDPMOracle(DDIMOracle):
    """diffusers ``DDPMScheduler`` with config/model/scheduler/ddpm.yaml (linear betas, fixed_small variance, clip_sample true,
    leading spacing), restated from knowledge of diffusers 0.27.2 ``DDPMScheduler.step`` - parity unpinned, like every
    diffusers piece (the reference only constructs it: src/model/scheduler/__init__.py:19-22)."""

    def __init__(self, clip_sample=True, clip_sample_range=1.0, variance_type="fixed_small", **kw):
        super().__init__(**kw)
        self.clip_sample, self.clip_range, self.variance_type = clip_sample, clip_sample_range, variance_type

    def step(self, eps: Tensor, t, x: Tensor, noise: Optional[Tensor] = None) -> Tensor:
        t = int(t)
        n = len(self.timesteps)
        prev = t - 1000 // n
        a_t = self.alphas_cumprod[t].double()
        a_p = self.alphas_cumprod[prev].double() if prev >= 0 else torch.tensor(1.0, dtype=torch.float64)
        cur_alpha = a_t / a_p
        cur_beta = 1 - cur_alpha
        x0 = (x.double() - (1 - a_t).sqrt() * eps.double()) / a_t.sqrt()
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_range, self.clip_range)
        prev_sample = (a_p.sqrt() * cur_beta / (1 - a_t)) * x0 + (cur_alpha.sqrt() * (1 - a_p) / (1 - a_t)) * x.double()
        if t > 0:
            var = (1 - a_p) / (1 - a_t) * cur_beta if self.variance_type == "fixed_small" else cur_beta
            prev_sample = prev_sample + var.clamp(min=1e-20).sqrt() * noise.double()
        return prev_sample.float()


# --------------------------------------------------------------------------------------
# geometry (projection.py:74-138, camera_utils.py:7-27, diffusion_wrapper.py:169-190,301-322)
# --------------------------------------------------------------------------------------
def absolute_to_relative(extr: Tensor, index: int = 0) -> Tensor:
    ref = extr[:, index:index + 1]
    return torch.linalg.inv(ref) @ extr


def positional_encoding(x: Tensor, num_octaves: int) -> Tensor:
    """``PositionalEncoding(num_octaves)`` (src/model/encodings/positional_encoding.py:28-49): [..., d] -> [..., d*F*2],
    channel (d f p) = sin(2 pi 2^f x_d + p pi/2)"""
    freq = 2 * torch.pi * 2 ** torch.arange(num_octaves).float()                      # [F]
    phase = torch.tensor([0, 0.5 * torch.pi], dtype=torch.float32)                     # [2]
    arg = x[..., :, None, None] * freq[:, None] + phase                                # [..., d, F, 2]
    return torch.sin(arg).flatten(-3)


def srt_positional_encoding(x: Tensor, num_octaves: int) -> Tensor:
    """SRT ``PositionalEncoding(num_octaves, start_octave=0)`` (src/model/srt/layers.py:9-32): [..., d] -> [..., 2*d*F] =
    [sin(pi 2^f x_d) over (d f) | cos(pi 2^f x_d) over (d f)]"""
    mult = 2 ** torch.arange(num_octaves).float() * math.pi
    sc = x[..., :, None] * mult
    return torch.cat([torch.sin(sc).flatten(-2), torch.cos(sc).flatten(-2)], dim=-1)


def raymap(extr: Tensor, intr: Tensor, h: int, w: int, plucker: bool = False, origin_octaves: int = 0,
           direction_octaves: int = 0, srt: bool = False) -> Tensor:
    """extr [B,V,4,4] cam->world, intr [B,V,3,3] normalised -> [B,V,C,h,w]:
    channels = origin (or origin x direction if plucker) then direction; pixel centres (i+0.5)/n, x fastest.
    Octaves > 0 (use_ray_encoding: true, diffusion_wrapper.py:115-126,317-320): each triple is replaced by its encoding."""
    B, V = extr.shape[:2]
    ys = (torch.arange(h, dtype=extr.dtype) + 0.5) / h
    xs = (torch.arange(w, dtype=extr.dtype) + 0.5) / w
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    pix = torch.stack([gx, gy, torch.ones_like(gx)], dim=-1).reshape(-1, 3)          # homogeneous xy1
    kinv = torch.linalg.inv(intr)                                                      # [B,V,3,3]
    d = torch.einsum("bvij,nj->bvni", kinv, pix)
    d = d / d.norm(dim=-1, keepdim=True)
    d = torch.einsum("bvij,bvnj->bvni", extr[..., :3, :3], d)
    o = extr[..., :3, 3][:, :, None, :].expand_as(d)
    if plucker:
        o = torch.cross(o, d, dim=-1)
    if srt:                                        # RayEncoder(pos, rays) = cat(pos_enc, ray_enc)  (srt/layers.py:53-56)
        o, d = srt_positional_encoding(o, origin_octaves), srt_positional_encoding(d, direction_octaves)
    else:
        if origin_octaves > 0:
            o = positional_encoding(o, origin_octaves)
        if direction_octaves > 0:
            d = positional_encoding(d, direction_octaves)
    r = torch.cat([o, d], dim=-1)                                                      # [B,V,n,C]
    return r.reshape(B, V, h, w, r.shape[-1]).permute(0, 1, 4, 2, 3).contiguous()


# --------------------------------------------------------------------------------------
# DiffusionWrapper.step / sample (diffusion_wrapper.py:413-490), VAE excluded (latents in, latents out)
# --------------------------------------------------------------------------------------
def build_inputs(x_t, context_inputs, rays, target_mask):
    """diffusion_wrapper.py:429-432: per view [latent(4) | mask(1) | rays(6)], context views first."""
    tgt = torch.cat([x_t, target_mask], dim=2)
    inp = torch.cat([context_inputs, tgt], dim=1)
    return torch.cat([inp, rays], dim=2), tgt


def ddim_step(sd, cfg: OracleCfg, sched: DDIMOracle, x_t, ts, context_inputs, rays, target_mask,
              use_cfg: bool = False, cfg_scale: float = 3.0, forward=None):
    fwd = forward or (lambda lat, t: unet_forward(sd, lat, t, cfg))
    B, v_c = context_inputs.shape[:2]
    v_t = x_t.shape[1]
    ts = int(ts)
    t_ctx = torch.zeros(B, v_c, dtype=torch.int64)
    t_tgt = torch.full((B, v_t), ts, dtype=torch.int64)
    inputs, tgt = build_inputs(sched.scale_model_input(x_t, ts), context_inputs, rays, target_mask)
    pred_c = fwd(inputs, torch.cat([t_ctx, t_tgt], dim=1))
    if use_cfg:
        pred_u = fwd(torch.cat([tgt, rays[:, v_c:]], dim=2), t_tgt)
        pred = pred_u + cfg_scale * (pred_c[:, v_c:] - pred_u)
    else:
        pred = pred_c[:, v_c:]
    return sched.step(pred, ts, x_t), pred


def sample(sd, cfg: OracleCfg, context_latents, x_T, extr, intr, num_steps=25, use_cfg=False,
           cfg_scale=3.0, plucker=False, forward=None, record=None):
    """``DiffusionWrapper.sample`` minus the VAE: context latents and x_T are passed in
    (the reference draws them from two different RNGs, diffusion_wrapper.py:283,473)."""
    sched = DDIMOracle()
    sched.set_timesteps(num_steps)
    B, v_c, _, h, w = context_latents.shape
    v_t = x_T.shape[1]
    x_t = x_T * sched.init_noise_sigma
    target_mask = torch.ones(B, v_t, 1, h, w)
    context_inputs = torch.cat([context_latents, torch.zeros(B, v_c, 1, h, w)], dim=2)
    rays = raymap(extr, intr, h, w, plucker)
    for ts in sched.timesteps:
        x_t, pred = ddim_step(sd, cfg, sched, x_t, ts, context_inputs, rays, target_mask,
                              use_cfg, cfg_scale, forward)
        if record is not None:
            record.append((int(ts), x_t.clone(), pred.clone()))
    return x_t


# --------------------------------------------------------------------------------------
# the synthetic workload of SURVEY.md §8(d)
# --------------------------------------------------------------------------------------
def synthetic_cameras(B: int, V: int) -> Tuple[Tensor, Tensor]:
    intr = torch.tensor([[1.2, 0.0, 0.5], [0.0, 1.2, 0.5], [0.0, 0.0, 1.0]]).expand(B, V, 3, 3).contiguous()
    extr = torch.eye(4).expand(B, V, 4, 4).contiguous().clone()
    extr[:, :, 0, 3] = 0.25 * torch.arange(V, dtype=torch.float32)[None, :]
    # a mild yaw per view so rotation handling is exercised
    for v in range(V):
        a = 0.05 * v
        extr[:, v, 0, 0] = math.cos(a); extr[:, v, 0, 2] = math.sin(a)
        extr[:, v, 2, 0] = -math.sin(a); extr[:, v, 2, 2] = math.cos(a)
    return absolute_to_relative(extr, 0), intr


def synthetic_scene(B: int, v_c: int, v_t: int, h: int = 32, w: int = 32, seed: int = 1):
    g1 = torch.Generator().manual_seed(seed)
    g2 = torch.Generator().manual_seed(seed + 1)
    ctx = torch.randn(B, v_c, 4, h, w, generator=g1)
    x_T = torch.randn(B, v_t, 4, h, w, generator=g2)
    extr, intr = synthetic_cameras(B, v_c + v_t)
    return ctx, x_T, extr, intr
