"""``diffusers.AutoencoderKL`` (0.27.2) restated as plain ``nn.Module``s.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference holds no VAE arithmetic of its own - ``src/model/autoencoder/__init__.py:40-43`` only
constructs ``diffusers.AutoencoderKL`` and ``diffusion_wrapper.py:283,295`` calls ``encode(...).latent_dist.sample()`` /
``decode(...).sample``; diffusers is neither vendored nor installable here and the reference ships no VAE test vectors.  This file
restates the published module structure (attribute names = state-dict keys, so SD checkpoints load) from knowledge of that
release: ``Encoder`` / ``Decoder`` (``models/autoencoders/vae.py``), ``DownEncoderBlock2D`` / ``UpDecoderBlock2D`` /
``UNetMidBlock2D`` (``models/unets/unet_2d_blocks.py``), ``ResnetBlock2D(temb_channels=None)``, ``Downsample2D(padding=0)``
(asymmetric (0,1,0,1) zero pad, stride-2 conv), ``Upsample2D`` (nearest x2 + conv), ``Attention(heads=1, residual_connection,
norm_num_groups, eps=1e-6)``, ``DiagonalGaussianDistribution``.  An error here would be shared by this checker and the CUDA
path; whenever a machine with diffusers is available, compare once and freeze the result as a golden file.
"""
import torch
from torch import nn
import torch.nn.functional as F


class VaeResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))       # dropout 0.0
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h                                 # output_scale_factor 1.0


class VaeAttention(nn.Module):
    """Attention(channels, heads=1, dim_head=channels, bias=True, upcast_softmax=True, residual_connection=True)"""

    def __init__(self, channels, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.to_q, self.to_k, self.to_v = (nn.Linear(channels, channels) for _ in range(3))
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        p = torch.softmax(q @ k.transpose(1, 2) * c ** -0.5, dim=-1)
        o = self.to_out[0](p @ v)
        return o.transpose(1, 2).reshape(b, c, h, w) + x      # rescale_output_factor 1.0


class _Sampler(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = None


class VaeDownsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class VaeUpsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block(nn.Module):
    def __init__(self, resnets, sampler_name=None, sampler=None):
        super().__init__()
        self.resnets = nn.ModuleList(resnets)
        if sampler is not None:
            setattr(self, sampler_name, nn.ModuleList([sampler]))
        self._sampler = sampler_name if sampler is not None else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self._sampler:
            x = getattr(self, self._sampler)[0](x)
        return x


class _Mid(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.attentions = nn.ModuleList([VaeAttention(c, groups)])
        self.resnets = nn.ModuleList([VaeResnetBlock2D(c, c, groups), VaeResnetBlock2D(c, c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Encoder(nn.Module):
    def __init__(self, in_channels, latent_channels, boc, layers_per_block, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        blocks, c = [], boc[0]
        for l, co in enumerate(boc):
            res = [VaeResnetBlock2D(c if i == 0 else co, co, groups) for i in range(layers_per_block)]
            c = co
            blocks.append(_Block(res, "downsamplers", VaeDownsample2D(c) if l != len(boc) - 1 else None))
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = _Mid(c, groups)
        self.conv_norm_out = nn.GroupNorm(groups, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, latent_channels, out_channels, boc, layers_per_block, groups):
        super().__init__()
        rev = list(boc)[::-1]
        self.conv_in = nn.Conv2d(latent_channels, rev[0], 3, padding=1)
        self.mid_block = _Mid(rev[0], groups)
        blocks, c = [], rev[0]
        for l, co in enumerate(rev):
            res = [VaeResnetBlock2D(c if i == 0 else co, co, groups) for i in range(layers_per_block + 1)]
            c = co
            blocks.append(_Block(res, "upsamplers", VaeUpsample2D(c) if l != len(rev) - 1 else None))
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(groups, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        return self.mean + self.std * torch.randn(self.mean.shape, generator=generator, device=self.parameters.device,
                                                  dtype=self.parameters.dtype)

    def mode(self):
        return self.mean


class _Obj:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",),
                 up_block_types=("UpDecoderBlock2D",), block_out_channels=(64,), layers_per_block=1, act_fn="silu",
                 latent_channels=4, norm_num_groups=32, sample_size=32, scaling_factor=0.18215, shift_factor=None,
                 latents_mean=None, latents_std=None, force_upcast=True, use_quant_conv=True, use_post_quant_conv=True,
                 mid_block_add_attention=True):
        super().__init__()
        assert all(t == "DownEncoderBlock2D" for t in down_block_types) and all(t == "UpDecoderBlock2D" for t in up_block_types)
        assert act_fn == "silu" and use_quant_conv and use_post_quant_conv and mid_block_add_attention
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.config = _Obj(scaling_factor=scaling_factor, latent_channels=latent_channels)

    @classmethod
    def from_pretrained(cls, name, subfolder=None):
        """topology of stabilityai/stable-diffusion-2-1 vae/config.json, random init (no hub here)"""
        return cls(down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
                   block_out_channels=(128, 256, 512, 512), layers_per_block=2, sample_size=768)

    def encode(self, x):
        return _Obj(latent_dist=DiagonalGaussianDistribution(self.quant_conv(self.encoder(x))))

    def decode(self, z):
        return _Obj(sample=self.decoder(self.post_quant_conv(z)))
