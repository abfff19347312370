"""Minimal stand-in for ``diffusers==0.27.2`` (reference requirements.txt:8).  TEST INFRASTRUCTURE ONLY.

Used by ``oracle/make_golden.py`` so that the reference's own ``src/model/denoiser/mvunet.py`` imports and
runs UNCHANGED in this sandbox (diffusers itself is neither installed nor installable here).  It restates,
as ``nn.Module``s with diffusers' attribute / state-dict names, exactly the Variant-A pieces listed in
SURVEY.md Appendix A.  Nothing in the product imports this package.
"""
import math
import torch
from torch import nn
import torch.nn.functional as F


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos=True, downscale_freq_shift=0):
        super().__init__()
        self.num_channels, self.flip, self.shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - self.shift)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels=1280, groups=32, eps=1e-5):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = None
        if in_channels != out_channels:
            self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1)

    def forward(self, input_tensor, temb):
        hidden = self.conv1(self.nonlinearity(self.norm1(input_tensor)))
        temb = self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        hidden = hidden + temb
        hidden = self.conv2(self.dropout(self.nonlinearity(self.norm2(hidden))))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + hidden) / 1.0


class Downsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        dtype = x.dtype
        if dtype == torch.bfloat16:
            x = x.float()
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        if dtype == torch.bfloat16:
            x = x.to(dtype)
        return self.conv(x)


class DownBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels)
            for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None


class UNetMidBlock2D(nn.Module):
    """as built by UNet2DConditionModel for mid_block_type='UNetMidBlock2D': num_layers=0, add_attention=False"""
    def __init__(self, in_channels, temb_channels):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, in_channels, temb_channels)])
        self.attentions = nn.ModuleList([])


class UpBlock2D(nn.Module):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers, add_upsample):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            res_skip = in_channels if (i == num_layers - 1) else out_channels
            res_in = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(res_in + res_skip, out_channels, temb_channels))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None


# ---- Variant B (SD-2.1 topology) pieces: per-view Transformer2DModel with use_linear_projection=True ------------
class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = heads * dim_head
        kv = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv, inner, bias=False)
        self.to_v = nn.Linear(kv, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, encoder_hidden_states=None):
        ctx = x if encoder_hidden_states is None else encoder_hidden_states
        b, n, _ = x.shape
        sp = lambda t: t.reshape(b, t.shape[1], self.heads, -1).permute(0, 2, 1, 3)  # noqa: E731
        q, k, v = sp(self.to_q(x)), sp(self.to_k(ctx)), sp(self.to_v(ctx))
        a = torch.softmax((q.float() @ k.float().transpose(-1, -2)) * self.scale, dim=-1).to(v.dtype) @ v  # upcast_attention
        return self.to_out[1](self.to_out[0](a.permute(0, 2, 1, 3).reshape(b, n, -1)))


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states) + x
        return self.ff(self.norm3(x)) + x


class _Out:
    def __init__(self, sample):
        self.sample = sample


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, cross_attention_dim, norm_num_groups=32):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, hidden_states, encoder_hidden_states=None):
        b, c, h, w = hidden_states.shape
        res = hidden_states
        x = self.norm(hidden_states).permute(0, 2, 3, 1).reshape(b, h * w, c)
        x = self.proj_in(x)
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states)
        x = self.proj_out(x).reshape(b, h, w, c).permute(0, 3, 1, 2)
        return _Out(x + res)


class CrossAttnDownBlock2D(DownBlock2D):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers, add_downsample, heads, cross_attention_dim):
        super().__init__(in_channels, out_channels, temb_channels, num_layers, add_downsample)
        self.has_cross_attention = True
        self.attentions = nn.ModuleList([Transformer2DModel(heads, out_channels // heads, out_channels, cross_attention_dim)
                                         for _ in range(num_layers)])


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, in_channels, temb_channels, heads, cross_attention_dim):
        super().__init__()
        self.has_cross_attention = True
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, in_channels, temb_channels) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, in_channels // heads, in_channels, cross_attention_dim)])


class CrossAttnUpBlock2D(UpBlock2D):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers, add_upsample, heads,
                 cross_attention_dim):
        super().__init__(in_channels, prev_output_channel, out_channels, temb_channels, num_layers, add_upsample)
        self.has_cross_attention = True
        self.attentions = nn.ModuleList([Transformer2DModel(heads, out_channels // heads, out_channels, cross_attention_dim)
                                         for _ in range(num_layers)])


class UNet2DConditionModel(nn.Module):
    def __init__(self, in_channels=4, out_channels=4, down_block_types=(), mid_block_type="UNetMidBlock2D",
                 up_block_types=(), only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280),
                 cross_attention_dim=1280, layers_per_block=2, norm_num_groups=32, norm_eps=1e-5, attention_head_dim=None):
        super().__init__()
        if attention_head_dim is not None:
            self._build_sd21(in_channels, out_channels, list(block_out_channels), attention_head_dim, cross_attention_dim,
                             layers_per_block, norm_num_groups, norm_eps)
            return
        assert all(t == "DownBlock2D" for t in down_block_types), "shim covers Variant A only"
        assert all(t == "UpBlock2D" for t in up_block_types), "shim covers Variant A only"
        assert mid_block_type == "UNetMidBlock2D"
        boc = list(block_out_channels)
        temb = boc[0] * 4
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_proj = Timesteps(boc[0], True, 0)
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.down_blocks = nn.ModuleList()
        out_c = boc[0]
        for i, c in enumerate(boc):
            in_c, out_c = out_c, c
            self.down_blocks.append(DownBlock2D(in_c, out_c, temb, layers_per_block, i != len(boc) - 1))
        self.mid_block = UNetMidBlock2D(boc[-1], temb)
        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        out_c = rev[0]
        for i in range(len(boc)):
            prev, out_c = out_c, rev[i]
            in_c = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(UpBlock2D(in_c, prev, out_c, temb, layers_per_block + 1, i != len(boc) - 1))
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, boc[0], eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)


def _build_sd21(self, in_channels, out_channels, boc, heads, cross_dim, layers_per_block, groups, eps):
    """stabilityai/stable-diffusion-2-1 unet/config.json: (CrossAttnDownBlock2D x3, DownBlock2D), UNetMidBlock2DCrossAttn,
    (UpBlock2D, CrossAttnUpBlock2D x3), attention_head_dim [5,10,20,20] (= heads), cross_attention_dim 1024,
    use_linear_projection true.  Random init (no hub in this sandbox)."""
    temb = boc[0] * 4
    self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
    self.time_proj = Timesteps(boc[0], True, 0)
    self.time_embedding = TimestepEmbedding(boc[0], temb)
    self.down_blocks = nn.ModuleList()
    out_c = boc[0]
    for i, c in enumerate(boc):
        in_c, out_c = out_c, c
        last = i == len(boc) - 1
        self.down_blocks.append(DownBlock2D(in_c, out_c, temb, layers_per_block, False) if last else
                                CrossAttnDownBlock2D(in_c, out_c, temb, layers_per_block, True, heads[i], cross_dim))
    self.mid_block = UNetMidBlock2DCrossAttn(boc[-1], temb, heads[-1], cross_dim)
    self.up_blocks = nn.ModuleList()
    rev, rheads = boc[::-1], heads[::-1]
    out_c = rev[0]
    for i in range(len(boc)):
        prev, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, len(boc) - 1)]
        add_up = i != len(boc) - 1
        self.up_blocks.append(UpBlock2D(in_c, prev, out_c, temb, layers_per_block + 1, add_up) if i == 0 else
                              CrossAttnUpBlock2D(in_c, prev, out_c, temb, layers_per_block + 1, add_up, rheads[i], cross_dim))
    self.conv_norm_out = nn.GroupNorm(groups, boc[0], eps=eps)
    self.conv_act = nn.SiLU()
    self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)


def _from_pretrained(cls, name, subfolder=None):
    assert "stable-diffusion-2-1" in name, "shim knows the SD-2.1 UNet config only"
    return cls(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), cross_attention_dim=1024,
               attention_head_dim=[5, 10, 20, 20])


UNet2DConditionModel._build_sd21 = _build_sd21
UNet2DConditionModel.from_pretrained = classmethod(_from_pretrained)


class DDIMScheduler:  # names only: the reference imports them at module scope (scheduler/__init__.py:4)
    pass


class DDPMScheduler:
    pass


from .vae import AutoencoderKL  # noqa: E402,F401  (restated diffusers AutoencoderKL: see vae.py's header - parity unpinned)
