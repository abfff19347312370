"""First-stage autoencoder of the MV-LDM sampling path, reference-facing side.

Mirrors what the reference builds and calls:

  * ``AutoencoderKLCfg`` / ``AutoencoderCfg`` / ``AUTOENCODERS`` / ``get_autoencoder``
    (reference src/model/autoencoder/autoencoder_kl.py:27-50, src/model/autoencoder/__init__.py:9-43)
  * ``AutoencoderKL.encode(x).latent_dist.sample()`` and ``AutoencoderKL.decode(z).sample`` as used by
    ``DiffusionWrapper.first_stage_encode`` / ``last_stage_decode`` (src/model/diffusion_wrapper.py:278-298)

The arithmetic is diffusers' ``AutoencoderKL`` (``diffusers==0.27.2``, not vendored by the reference); parameter names are
diffusers' state-dict keys, so an SD-2.1 ``vae`` checkpoint loads with ``load_state_dict`` (the pre-0.15 attention names
``query/key/value/proj_attn/norm`` are accepted too, like diffusers' own loader).  Compute runs in ``libmvldm_b200.so``
(``mvldm_vae_encode`` / ``mvldm_vae_decode``: the denoiser's conv / GroupNorm / GEMM kernels); there is no CPU fallback and
no hub access (``from_pretrained`` builds the topology only; weights arrive through ``load_state_dict``).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import asdict, dataclass
from types import SimpleNamespace
from typing import Dict, List, Literal, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib
from .denoiser import MultiViewUNet, _Handle, _Node


@dataclass
class AutoencoderKLCfg:
    """reference src/model/autoencoder/autoencoder_kl.py:27-50"""
    in_channels: int = 3
    out_channels: int = 3
    down_block_types: Tuple[str, ...] | List[str] = ("DownEncoderBlock2D",)
    up_block_types: Tuple[str, ...] | List[str] = ("UpDecoderBlock2D",)
    block_out_channels: Tuple[int, ...] | List[int] = (64,)
    layers_per_block: int = 1
    act_fn: str = "silu"
    latent_channels: int = 4
    norm_num_groups: int = 32
    sample_size: int = 32
    scaling_factor: float = 0.18215
    shift_factor: Optional[float] = None
    latents_mean: Optional[Tuple[float, ...] | List[float]] = None
    latents_std: Optional[Tuple[float, ...] | List[float]] = None
    force_upcast: float | bool = True
    use_quant_conv: bool = True
    use_post_quant_conv: bool = True
    mid_block_add_attention: bool = True


@dataclass
class AutoencoderCfg:
    """reference src/model/autoencoder/__init__.py:9-13"""
    name: Literal["kl"]
    pretrained_from: str | None
    kwargs: AutoencoderKLCfg


def sd21_vae_cfg() -> AutoencoderKLCfg:
    """stabilityai/stable-diffusion-2-1 vae/config.json (what `from_pretrained(..., subfolder="vae")` builds,
    src/model/autoencoder/__init__.py:43)"""
    return AutoencoderKLCfg(down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
                            block_out_channels=(128, 256, 512, 512), layers_per_block=2, sample_size=768)


def vae_param_shapes(cfg: AutoencoderKLCfg) -> "Dict[str, Tuple[int, ...]]":
    """diffusers AutoencoderKL state-dict keys and shapes; must equal the registry the library builds in mvldm_create."""
    P: Dict[str, Tuple[int, ...]] = {}
    boc = list(cfg.block_out_channels)
    L, lc = len(boc), cfg.latent_channels

    def conv(k, co, ci, ks):
        P[k + ".weight"] = (co, ci, ks, ks)
        P[k + ".bias"] = (co,)

    def lin(k, co, ci):
        P[k + ".weight"] = (co, ci)
        P[k + ".bias"] = (co,)

    def norm(k, c):
        P[k + ".weight"] = (c,)
        P[k + ".bias"] = (c,)

    def resnet(k, ci, co):
        norm(k + ".norm1", ci); conv(k + ".conv1", co, ci, 3); norm(k + ".norm2", co); conv(k + ".conv2", co, co, 3)
        if ci != co:
            conv(k + ".conv_shortcut", co, ci, 1)

    def attn(k, c):
        norm(k + ".group_norm", c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(f"{k}.{n}", c, c)

    conv("encoder.conv_in", boc[0], cfg.in_channels, 3)
    c = boc[0]
    for l in range(L):
        for i in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{l}.resnets.{i}", c if i == 0 else boc[l], boc[l])
        c = boc[l]
        if l != L - 1:
            conv(f"encoder.down_blocks.{l}.downsamplers.0.conv", c, c, 3)
    resnet("encoder.mid_block.resnets.0", c, c); attn("encoder.mid_block.attentions.0", c)
    resnet("encoder.mid_block.resnets.1", c, c)
    norm("encoder.conv_norm_out", c); conv("encoder.conv_out", 2 * lc, c, 3)
    conv("quant_conv", 2 * lc, 2 * lc, 1); conv("post_quant_conv", lc, lc, 1)
    conv("decoder.conv_in", boc[-1], lc, 3)
    c = boc[-1]
    resnet("decoder.mid_block.resnets.0", c, c); attn("decoder.mid_block.attentions.0", c)
    resnet("decoder.mid_block.resnets.1", c, c)
    for l in range(L):
        co = boc[L - 1 - l]
        for i in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{l}.resnets.{i}", c if i == 0 else co, co)
        c = co
        if l != L - 1:
            conv(f"decoder.up_blocks.{l}.upsamplers.0.conv", c, c, 3)
    norm("decoder.conv_norm_out", c); conv("decoder.conv_out", cfg.out_channels, c, 3)
    return P


class DiagonalGaussianDistribution:
    """diffusers.models.autoencoders.vae.DiagonalGaussianDistribution: parameters = (mean | logvar) along dim 1."""

    def __init__(self, parameters: Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self) -> Tensor:
        return self.mean


class AutoencoderKL(nn.Module):
    """Drop-in for ``diffusers.AutoencoderKL`` on the two calls the reference makes (diffusion_wrapper.py:283,295)."""

    _DEPRECATED_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0", "norm": "group_norm"}

    def __init__(self, in_channels: int = 3, out_channels: int = 3, down_block_types=("DownEncoderBlock2D",),
                 up_block_types=("UpDecoderBlock2D",), block_out_channels=(64,), layers_per_block: int = 1,
                 act_fn: str = "silu", latent_channels: int = 4, norm_num_groups: int = 32, sample_size: int = 32,
                 scaling_factor: float = 0.18215, shift_factor=None, latents_mean=None, latents_std=None,
                 force_upcast=True, use_quant_conv: bool = True, use_post_quant_conv: bool = True,
                 mid_block_add_attention: bool = True, *, use_cuda_graph: bool = True) -> None:
        super().__init__()
        cfg = AutoencoderKLCfg(in_channels, out_channels, tuple(down_block_types), tuple(up_block_types),
                               tuple(block_out_channels), layers_per_block, act_fn, latent_channels, norm_num_groups,
                               sample_size, scaling_factor, shift_factor, latents_mean, latents_std, force_upcast,
                               use_quant_conv, use_post_quant_conv, mid_block_add_attention)
        if any(t != "DownEncoderBlock2D" for t in cfg.down_block_types) or \
                any(t != "UpDecoderBlock2D" for t in cfg.up_block_types):
            raise ValueError("only DownEncoderBlock2D / UpDecoderBlock2D blocks are supported")
        if not (len(cfg.down_block_types) == len(cfg.up_block_types) == len(cfg.block_out_channels)) or \
                not (1 <= len(cfg.block_out_channels) <= _lib.MVLDM_MAX_LEVELS):
            raise ValueError("down_block_types, up_block_types and block_out_channels must have the same length (1..4)")
        if cfg.act_fn != "silu" or not (cfg.use_quant_conv and cfg.use_post_quant_conv and cfg.mid_block_add_attention):
            raise ValueError("supported: act_fn 'silu', quant / post-quant convs and the mid-block attention present")
        if any(c % 64 for c in cfg.block_out_channels) or any(c % cfg.norm_num_groups for c in cfg.block_out_channels):
            raise ValueError("block_out_channels must be multiples of 64 and of norm_num_groups")
        if cfg.shift_factor is not None or cfg.latents_mean is not None or cfg.latents_std is not None:
            raise ValueError("shift_factor / latents_mean / latents_std are not used by the reference path")
        self.cfg = cfg
        self.config = SimpleNamespace(**asdict(cfg))               # `vae.config.scaling_factor` etc., as in diffusers
        self.use_cuda_graph = use_cuda_graph
        self._shapes = vae_param_shapes(cfg)
        for key, shape in self._shapes.items():
            self._register(key, self._init_param(key, shape))
        self._h = _Handle()
        self._dirty = True
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.mark_dirty())
        self._register_load_state_dict_pre_hook(self._convert_deprecated_attention)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None, **kw) -> "AutoencoderKL":
        """Topology only (no hub access in the library): the SD-2.x VAE; load the checkpoint with `load_state_dict`."""
        if "stable-diffusion-2" not in str(pretrained_model_name_or_path):
            raise ValueError("from_pretrained: only the stabilityai/stable-diffusion-2* VAE topology is known")
        return cls(**asdict(sd21_vae_cfg()), **kw)

    # ---- parameters (same plumbing as the denoiser module) ---------------------------------------
    _register = MultiViewUNet._register
    _versions = MultiViewUNet._versions
    refresh_weights = MultiViewUNet.refresh_weights
    _launch_checked = MultiViewUNet._launch_checked

    def _init_param(self, key: str, shape) -> nn.Parameter:
        leaf = key.rsplit(".", 1)[1]
        if "norm" in key.rsplit(".", 2)[-2]:
            t = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
        else:
            wshape = self._shapes[key.rsplit(".", 1)[0] + ".weight"]
            bound = 1.0 / math.sqrt(math.prod(wshape[1:]))
            t = (torch.rand(shape) * 2 - 1) * bound
        return nn.Parameter(t)

    def mark_dirty(self) -> None:
        self._dirty = True
        self.__dict__.pop("_plist", None)

    def _apply(self, fn, *args, **kwargs):
        self._dirty = True
        self.__dict__.pop("_plist", None)
        return super()._apply(fn, *args, **kwargs)

    @classmethod
    def _convert_deprecated_attention(cls, state_dict, prefix, *args) -> None:
        """diffusers' _convert_deprecated_attention_blocks: pre-0.15 checkpoints name the mid-block attention
        query/key/value/proj_attn/norm (SD's original vae files do); Linear weights may be stored as 1x1 convs."""
        for half in ("encoder", "decoder"):
            base = f"{prefix}{half}.mid_block.attentions.0."
            for old, new in cls._DEPRECATED_ATTN.items():
                for leaf in ("weight", "bias"):
                    k = f"{base}{old}.{leaf}"
                    if k in state_dict:
                        state_dict[f"{base}{new}.{leaf}"] = state_dict.pop(k)
            for name in ("to_q", "to_k", "to_v", "to_out.0"):
                k = f"{base}{name}.weight"
                if k in state_dict and state_dict[k].dim() == 4:
                    state_dict[k] = state_dict[k][:, :, 0, 0]

    def _ensure_handle(self, device: torch.device):
        h = self._h
        if h.ptr is not None and h.device == device:
            return h
        h.close()
        lib = _lib.load()
        cfg = self.cfg
        c = _lib.Config()
        c.model = _lib.MODEL_VAE
        c.in_channels, c.out_channels, c.num_levels = cfg.in_channels, cfg.out_channels, len(cfg.block_out_channels)
        for i, v in enumerate(cfg.block_out_channels):
            c.block_out_channels[i] = v
        c.layers_per_block, c.norm_groups, c.latent_channels = cfg.layers_per_block, cfg.norm_num_groups, cfg.latent_channels
        c.num_heads, c.max_attn_res, c.impl = 1, 0, _lib.IMPL_TC
        c.use_cuda_graph = 1 if self.use_cuda_graph else 0
        ptr = ctypes.c_void_p()
        _lib.check(lib.mvldm_create(ctypes.byref(c), device.index or 0, ctypes.byref(ptr)))
        h.ptr, h.device, h.synced_versions = ptr, device, None
        names = [lib.mvldm_weight_name(ptr, i).decode() for i in range(lib.mvldm_num_weights(ptr))]
        if set(names) != set(self._shapes):
            raise RuntimeError("mvldm_b200: library/Python state-dict key mismatch (VAE)")
        return h

    # ---- the two calls of the reference --------------------------------------------------------------
    @property
    def downscale(self) -> int:
        return 2 ** (len(self.cfg.block_out_channels) - 1)

    def _check(self, x: Tensor, channels: int, what: str) -> Tensor:
        if x.dim() != 4:
            raise ValueError(f"{what} must be [batch, channel, height, width]")
        if not x.is_cuda:
            raise RuntimeError("mvldm_b200: inputs must be CUDA tensors (no CPU fallback)")
        if x.shape[1] != channels:
            raise ValueError(f"expected {channels} channels, got {x.shape[1]}")
        return x.detach().to(torch.float32).contiguous()

    @torch.no_grad()
    def encode(self, x: Tensor, return_dict: bool = True):
        """`self.autoencoder.encode(inputs).latent_dist.sample()` (diffusion_wrapper.py:283): x in [-1, 1], [N, 3, H, W]"""
        x = self._check(x, self.cfg.in_channels, "image")
        n, _, h, w = x.shape
        f = self.downscale
        if h % f or w % f:
            raise ValueError(f"image size must be divisible by {f}")

        def launch():
            moments = torch.empty((n, 2 * self.cfg.latent_channels, h // f, w // f), device=x.device, dtype=torch.float32)
            with _lib.on_device(x.device):
                _lib.check(_lib.load().mvldm_vae_encode(self._h.ptr, _lib.current_stream_ptr(x.device), x.data_ptr(), n, h, w,
                                                        moments.data_ptr()))
            return moments
        dist = DiagonalGaussianDistribution(self._launch_checked(launch))
        return SimpleNamespace(latent_dist=dist) if return_dict else (dist,)

    @torch.no_grad()
    def decode(self, z: Tensor, return_dict: bool = True, generator=None):
        """`self.autoencoder.decode(latents).sample` (diffusion_wrapper.py:295): z [N, 4, h, w] already / scaling_factor"""
        z = self._check(z, self.cfg.latent_channels, "latents")
        n, _, h, w = z.shape
        f = self.downscale

        def launch():
            img = torch.empty((n, self.cfg.out_channels, h * f, w * f), device=z.device, dtype=torch.float32)
            with _lib.on_device(z.device):
                _lib.check(_lib.load().mvldm_vae_decode(self._h.ptr, _lib.current_stream_ptr(z.device), z.data_ptr(), n, h, w,
                                                        img.data_ptr()))
            return img
        img = self._launch_checked(launch)
        return SimpleNamespace(sample=img) if return_dict else (img,)

    def forward(self, sample: Tensor, sample_posterior: bool = False, generator=None):
        post = self.encode(sample).latent_dist
        z = post.sample(generator) if sample_posterior else post.mode()
        return self.decode(z)

    def last_launch_count(self) -> int:
        return _lib.load().mvldm_last_launch_count(self._h.ptr) if self._h.ptr else 0


AUTOENCODERS = {"kl": AutoencoderKL}


def get_autoencoder(cfg: AutoencoderCfg) -> AutoencoderKL:
    """reference src/model/autoencoder/__init__.py:36-43"""
    if cfg.pretrained_from is None:
        return AUTOENCODERS[cfg.name](**asdict(cfg.kwargs))
    return AUTOENCODERS[cfg.name].from_pretrained(cfg.pretrained_from, subfolder="vae")


def first_stage_encode(autoencoder: AutoencoderKL, inputs: Tensor, generator=None) -> Tensor:
    """DiffusionWrapper.first_stage_encode (diffusion_wrapper.py:278-288): images in [0, 1] [b, v, 3, H, W] -> latents"""
    b, v = inputs.shape[:2]
    x = inputs.reshape(b * v, *inputs.shape[2:]) * 2.0 - 1.0
    lat = autoencoder.encode(x).latent_dist.sample(generator) * 0.18215
    return lat.reshape(b, v, *lat.shape[1:])


def last_stage_decode(autoencoder: AutoencoderKL, latents: Tensor) -> Tensor:
    """DiffusionWrapper.last_stage_decode (diffusion_wrapper.py:290-298): latents [b, v, 4, h, w] -> images in [0, 1]"""
    b, v = latents.shape[:2]
    z = (1 / 0.18215) * latents.reshape(b * v, *latents.shape[2:])
    img = autoencoder.decode(z).sample
    return (img.reshape(b, v, *img.shape[1:]) / 2 + 0.5).clamp(0, 1)
