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


class UNet2DConditionModel(nn.Module):
    def __init__(self, in_channels=4, out_channels=4, down_block_types=(), mid_block_type="UNetMidBlock2D",
                 up_block_types=(), only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280),
                 cross_attention_dim=1280, layers_per_block=2, norm_num_groups=32, norm_eps=1e-5):
        super().__init__()
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


class DDIMScheduler:  # names only: the reference imports them at module scope (scheduler/__init__.py:4)
    pass


class DDPMScheduler:
    pass


class AutoencoderKL:
    pass
