"""Host-side mirror of the reference denoiser interface, backed by the sm_100a C-ABI library.

Mirrors (same names, argument meaning and error behaviour):
  * ``Denoiser``                      reference src/model/denoiser/denoiser.py:12-29
  * ``MultiViewUNetCfg`` / ``UNet2DModelCfg``  reference src/model/denoiser/mvunet.py:22-40
  * ``SpatialTransformer3DCfg``       reference src/model/denoiser/mvdream/attention.py:23-32
  * ``MultiViewUNet``                 reference src/model/denoiser/mvunet.py:43-208
  * ``DENOISER`` / ``get_denoiser``   reference src/model/denoiser/__init__.py:7-18

``MultiViewUNet`` is an ``nn.Module`` whose parameters carry exactly the reference's state-dict keys
(``unet.*``, ``cross_attn_blocks_{encoder,mid,decoder}.*``), so Lightning checkpoints load unchanged and
``.parameters()`` / ``AveragedModel`` keep working; its ``forward`` does no torch math — it hands the device
pointers to ``mvldm_forward`` on the current CUDA stream.  Inference only (no autograd through the kernels).
"""
from __future__ import annotations

import ctypes
import math
import operator
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from typing import Dict, Generic, List, Literal, Optional, Tuple, TypeVar

import torch
from torch import Tensor, nn

from . import _lib

T = TypeVar("T")


@dataclass
class UNet2DModelCfg:
    name: Literal["unet"]
    down_block_types: list | Tuple
    mid_block_type: str
    up_block_types: list | Tuple
    only_cross_attention: bool
    block_out_channels: list | Tuple


@dataclass
class SpatialTransformer3DCfg:
    name: Literal["spatial_transformer_3d"]
    num_heads: int
    num_layers: int = 1
    d_dot: int | None = None
    d_mlp: int | None = None
    d_mlp_multiplier: int | None = None
    downscale: int = 1
    pos_enc: bool = False


@dataclass
class CrossAttentionCfg:
    """reference src/model/denoiser/standard/transformer.py:34-43 (multi_view_attention/standard_attention.yaml)"""
    name: Literal["standard"]
    num_heads: int
    num_layers: int = 1
    d_dot: int | None = None
    d_mlp: int | None = None
    d_mlp_multiplier: int | None = None
    downscale: int = 1
    pos_enc: bool = False


MultiViewAttentionCfg = CrossAttentionCfg | SpatialTransformer3DCfg    # reference src/model/denoiser/attention.py:6


@dataclass
class MultiViewUNetCfg:
    name: Literal["mv_unet"]
    autoencoder: UNet2DModelCfg
    multi_view_attention: MultiViewAttentionCfg
    use_ray_encoding: bool = True
    encoder_conditioning: bool = True
    mid_conditioning: bool = True
    decoder_conditioning: bool = True
    pretrained_from: str | None = None


def default_cfg(num_heads: int = 8) -> MultiViewUNetCfg:
    """The released topology minus the hub download: config/model/denoiser/mv_unet.yaml:8-17 +
    multi_view_attention/spatial_transformer_3d.yaml."""
    return MultiViewUNetCfg(
        name="mv_unet",
        autoencoder=UNet2DModelCfg("unet", ["DownBlock2D"] * 4, "UNetMidBlock2D", ["UpBlock2D"] * 4, False,
                                   [320, 640, 1280, 1280]),
        multi_view_attention=SpatialTransformer3DCfg("spatial_transformer_3d", num_heads=num_heads),
        use_ray_encoding=False)


def standard_cfg(num_heads: int = 8, d_mlp_multiplier: int = 1, num_layers: int = 1) -> MultiViewUNetCfg:
    """mv_unet.yaml's own default `multi_view_attention: standard_attention` (config/model/denoiser/mv_unet.yaml:5,
    multi_view_attention/standard_attention.yaml): StandardTransformer blocks at the 9 multi-view positions."""
    cfg = default_cfg(num_heads)
    cfg.multi_view_attention = CrossAttentionCfg("standard", num_heads=num_heads, num_layers=num_layers,
                                                 d_mlp_multiplier=d_mlp_multiplier)
    return cfg


class Denoiser(nn.Module, ABC, Generic[T]):
    cfg: T

    def __init__(self, cfg: T) -> None:
        super().__init__()
        self.cfg = cfg

    @abstractmethod
    def forward(self, latents: Tensor, timestep: Tensor, cond_state: Optional[Tensor] = None) -> Tensor:
        pass


# stabilityai/stable-diffusion-2-1 unet/config.json, the only hub topology mvunet.py:124-128 can drive (it hard-codes a
# 1024-wide context): attention_head_dim (= heads, a diffusers quirk; every head is 64 wide) and cross_attention_dim
SD21_BLOCK_OUT_CHANNELS = (320, 640, 1280, 1280)
SD21_HEADS = (5, 10, 20, 20)
SD21_CROSS_ATTENTION_DIM = 1024


def param_shapes(block_out_channels, in_channels: int, out_channels: int, layers_per_block: int = 2,
                 variant_b: bool = False, standard: "CrossAttentionCfg | None" = None) -> "Dict[str, Tuple[int, ...]]":
    """State-dict keys and shapes of the MultiViewUNet; must equal the registry the C
    library builds in ``mvldm_create`` (checked at handle creation).  ``variant_b`` adds the SD-2.1 pieces:
    ``unet.{down_blocks.0-2,mid_block,up_blocks.1-3}.attentions.*`` and ``unet.mid_block.resnets.1``."""
    P: Dict[str, Tuple[int, ...]] = {}
    boc = list(block_out_channels)
    T_ = boc[0] * 4

    def conv(k, co, ci, ks):
        P[k + ".weight"] = (co, ci, ks, ks)
        P[k + ".bias"] = (co,)

    def lin(k, co, ci, bias=True):
        P[k + ".weight"] = (co, ci)
        if bias:
            P[k + ".bias"] = (co,)

    def norm(k, c):
        P[k + ".weight"] = (c,)
        P[k + ".bias"] = (c,)

    def resnet(k, ci, co):
        norm(k + ".norm1", ci); conv(k + ".conv1", co, ci, 3); lin(k + ".time_emb_proj", co, T_)
        norm(k + ".norm2", co); conv(k + ".conv2", co, co, 3)
        if ci != co:
            conv(k + ".conv_shortcut", co, ci, 1)

    def mv(k, c):
        if standard is not None:
            # StandardTransformer(d_in=c).transformer = Transformer(c, num_layers, heads, c // heads, d_mlp) with
            # layers.J = [PreNorm(Attention), PreNorm(FeedForward)]  (transformer/transformer.py:24-47)
            d_mlp = standard.d_mlp or c * standard.d_mlp_multiplier
            for j in range(standard.num_layers):
                lk = f"{k}.transformer.layers.{j}"
                norm(lk + ".0.norm", c); lin(lk + ".0.fn.to_qkv", 3 * c, c, bias=False); lin(lk + ".0.fn.to_out.0", c, c)
                norm(lk + ".1.norm", c); lin(lk + ".1.fn.net.0", d_mlp, c); lin(lk + ".1.fn.net.3", c, d_mlp)
            return
        norm(k + ".norm", c); conv(k + ".proj_in", c, c, 1)
        tb = k + ".transformer_blocks.0"
        for a in ("attn1", "attn2"):
            for p in ("to_q", "to_k", "to_v"):
                lin(f"{tb}.{a}.{p}", c, c, bias=False)
            lin(f"{tb}.{a}.to_out.0", c, c)
        lin(f"{tb}.ff.net.0.proj", 8 * c, c); lin(f"{tb}.ff.net.2", c, 4 * c)
        for n in ("norm1", "norm2", "norm3"):
            norm(f"{tb}.{n}", c)
        conv(k + ".proj_out", c, c, 1)

    def t2d(k, c):
        norm(k + ".norm", c); lin(k + ".proj_in", c, c)
        tb = k + ".transformer_blocks.0"
        for a, kv in (("attn1", c), ("attn2", SD21_CROSS_ATTENTION_DIM)):
            lin(f"{tb}.{a}.to_q", c, c, bias=False)
            lin(f"{tb}.{a}.to_k", c, kv, bias=False); lin(f"{tb}.{a}.to_v", c, kv, bias=False)
            lin(f"{tb}.{a}.to_out.0", c, c)
        lin(f"{tb}.ff.net.0.proj", 8 * c, c); lin(f"{tb}.ff.net.2", c, 4 * c)
        for n in ("norm1", "norm2", "norm3"):
            norm(f"{tb}.{n}", c)
        lin(k + ".proj_out", c, c)

    L = len(boc)
    conv("unet.conv_in", boc[0], in_channels, 3)
    lin("unet.time_embedding.linear_1", T_, boc[0]); lin("unet.time_embedding.linear_2", T_, T_)
    co = boc[0]
    for l in range(L):
        ci, co = co, boc[l]
        for i in range(layers_per_block):
            resnet(f"unet.down_blocks.{l}.resnets.{i}", ci if i == 0 else co, co)
            if variant_b and l != L - 1:
                t2d(f"unet.down_blocks.{l}.attentions.{i}", co)
        if l != L - 1:
            conv(f"unet.down_blocks.{l}.downsamplers.0.conv", co, co, 3)
    resnet("unet.mid_block.resnets.0", boc[-1], boc[-1])
    if variant_b:
        t2d("unet.mid_block.attentions.0", boc[-1]); resnet("unet.mid_block.resnets.1", boc[-1], boc[-1])
    rev = boc[::-1]
    oc = rev[0]
    for l in range(L):
        prev, oc = oc, rev[l]
        ic = rev[min(l + 1, L - 1)]
        for i in range(layers_per_block + 1):
            resnet(f"unet.up_blocks.{l}.resnets.{i}", (prev if i == 0 else oc) + (ic if i == layers_per_block else oc), oc)
            if variant_b and l != 0:
                t2d(f"unet.up_blocks.{l}.attentions.{i}", oc)
        if l != L - 1:
            conv(f"unet.up_blocks.{l}.upsamplers.0.conv", oc, oc, 3)
    norm("unet.conv_norm_out", boc[0]); conv("unet.conv_out", out_channels, boc[0], 3)
    for l in range(L):
        mv(f"cross_attn_blocks_encoder.{l}", boc[l])
    mv("cross_attn_blocks_mid.0", boc[-1])
    for l in range(L):
        mv(f"cross_attn_blocks_decoder.{l}", rev[l])
    return P


_VERSION_OF = operator.attrgetter("_version")


class _Node(nn.Module):
    """Parameter container node; children are registered under the reference's attribute names."""

    def enable_xformers_memory_efficient_attention(self, *a, **k):   # diffusion_wrapper.py:144-147
        return None


class _Handle:
    """Owns the C-side handle; never copied (AveragedModel deep-copies the module: the copy re-creates it)."""

    def __init__(self):
        self.ptr = None
        self.device = None
        self.synced_versions = None

    def __deepcopy__(self, memo):
        return _Handle()

    def close(self):
        if self.ptr is not None:
            try:
                _lib.load().mvldm_destroy(self.ptr)
            except Exception:
                pass
            self.ptr = None

    def __del__(self):
        self.close()


class MultiViewUNet(Denoiser[MultiViewUNetCfg]):
    """Drop-in for reference ``MultiViewUNet`` (mvunet.py:43-208).

    ``pretrained_from=None`` (Variant A) builds the topology ``cfg.autoencoder`` names (DownBlock2D / UNetMidBlock2D /
    UpBlock2D).  ``pretrained_from=<hub id>`` (Variant B) builds the SD-2.1 UNet2DConditionModel topology the reference
    downloads (mvunet.py:64-72) — WITHOUT the download: there is no hub access in the library, the parameters are
    random-initialised and the checkpoint arrives through ``load_state_dict`` like any Lightning checkpoint."""

    def __init__(self, cfg: MultiViewUNetCfg, in_channels: int, out_channels: int, *, impl: int = _lib.IMPL_TC,
                 use_cuda_graph: bool = True) -> None:
        super().__init__(cfg)
        self.variant_b = cfg.pretrained_from is not None
        mva = cfg.multi_view_attention
        if mva.name not in ("spatial_transformer_3d", "standard"):
            raise ValueError("multi_view_attention.name must be 'spatial_transformer_3d' or 'standard' "
                             "(reference src/model/denoiser/attention.py:12-27)")
        self.standard = mva.name == "standard"
        if self.standard:
            # StandardTransformer.__init__ asserts this (standard/transformer.py:63-64)
            if (mva.d_mlp is None) == (mva.d_mlp_multiplier is None):
                raise ValueError("Expected exactly one of d_mlp and d_mlp_multiplier")
            if mva.d_dot is not None or mva.downscale != 1 or mva.pos_enc or not (1 <= mva.num_layers <= 8):
                raise ValueError("standard multi_view_attention: supported are d_dot None (= d_in // num_heads), "
                                 "downscale 1, pos_enc false, 1 <= num_layers <= 8")
        elif mva.num_layers != 1 or mva.d_dot is not None:
            raise ValueError("multi_view_attention: num_layers must be 1 and d_dot None")
        if not (cfg.encoder_conditioning and cfg.mid_conditioning and cfg.decoder_conditioning):
            raise ValueError("encoder/mid/decoder_conditioning must all be true")
        ae = cfg.autoencoder
        if self.variant_b:
            # from_pretrained ignores cfg.autoencoder's block types; only block_out_channels[0] is read (conv_in/out)
            if "stable-diffusion-2" not in cfg.pretrained_from:
                raise ValueError("pretrained_from: only the stabilityai/stable-diffusion-2* UNet topology is known "
                                 "(mvunet.py:127 hard-codes its 1024-wide context)")
            if ae.block_out_channels[0] != SD21_BLOCK_OUT_CHANNELS[0]:
                raise ValueError("autoencoder.block_out_channels[0] must be 320 with an SD-2.1 UNet")
            self._boc = list(SD21_BLOCK_OUT_CHANNELS)
        else:
            if any(t != "DownBlock2D" for t in ae.down_block_types) or any(t != "UpBlock2D" for t in ae.up_block_types) \
                    or ae.mid_block_type != "UNetMidBlock2D":
                raise ValueError("pretrained_from=None: only DownBlock2D / UNetMidBlock2D / UpBlock2D topologies are "
                                 "supported")
            self._boc = list(ae.block_out_channels)
        self.use_ray_encoding = cfg.use_ray_encoding
        self.pretrained_from = cfg.pretrained_from
        self.in_channels, self.out_channels = in_channels, out_channels
        self.impl, self.use_cuda_graph = impl, use_cuda_graph
        self._shapes = param_shapes(self._boc, in_channels, out_channels, variant_b=self.variant_b,
                                    standard=mva if self.standard else None)
        for key, shape in self._shapes.items():
            self._register(key, self._init_param(key, shape))
        self._h = _Handle()
        self._dirty = True
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.mark_dirty())

    def mark_dirty(self) -> None:
        """Tell the module its parameters changed in place (the packed device copy is rebuilt on the next call).
        Done automatically after load_state_dict / .to() / .cuda(); in-place updates (optimizer steps, EMA / AveragedModel
        updates, `p.detach().copy_()`) are caught by the per-call parameter-version check.  Writes through `p.data` bypass
        autograd's version counter and cannot be seen: call this after them."""
        self._dirty = True
        self.__dict__.pop("_plist", None)

    def _apply(self, fn, *args, **kwargs):
        self._dirty = True
        self.__dict__.pop("_plist", None)
        return super()._apply(fn, *args, **kwargs)

    # ---- parameters -------------------------------------------------------------------
    def _init_param(self, key: str, shape) -> nn.Parameter:
        leaf = key.rsplit(".", 1)[1]
        is_norm = ".norm" in key or key.endswith("conv_norm_out.weight") or key.endswith("conv_norm_out.bias")
        if is_norm:
            t = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
        elif key.startswith("cross_attn_blocks_") and (key.endswith("proj_out.weight") or key.endswith("proj_out.bias")):
            t = torch.zeros(shape)                       # zero_module(): mvdream/attention.py:90-96,406-411
        else:
            wshape = self._shapes[key.rsplit(".", 1)[0] + ".weight"]
            bound = 1.0 / math.sqrt(math.prod(wshape[1:]))
            t = (torch.rand(shape) * 2 - 1) * bound
        return nn.Parameter(t)

    def _register(self, key: str, p: nn.Parameter) -> None:
        parts = key.split(".")
        node: nn.Module = self
        for name in parts[:-1]:
            if name not in node._modules:
                node.add_module(name, _Node())
            node = node._modules[name]
        node.register_parameter(parts[-1], p)

    # ---- library handle -----------------------------------------------------------------
    def _ensure_handle(self, device: torch.device):
        h = self._h
        if h.ptr is not None and h.device == device:
            return h
        h.close()
        lib = _lib.load()
        boc = self._boc
        c = _lib.Config()
        c.in_channels, c.out_channels, c.num_levels = self.in_channels, self.out_channels, len(boc)
        for i, v in enumerate(boc):
            c.block_out_channels[i] = v
        c.layers_per_block, c.norm_groups = 2, 32
        c.num_heads = self.cfg.multi_view_attention.num_heads
        c.max_attn_res = 32
        c.impl = self.impl
        c.use_cuda_graph = 1 if self.use_cuda_graph else 0
        if self.standard:
            mva = self.cfg.multi_view_attention
            c.mv_block, c.mv_num_layers = _lib.MV_STANDARD, mva.num_layers
            c.mv_d_mlp, c.mv_d_mlp_multiplier = mva.d_mlp or 0, mva.d_mlp_multiplier or 0
        if self.variant_b:
            c.variant, c.cross_attention_dim = 1, SD21_CROSS_ATTENTION_DIM
            for i, v in enumerate(SD21_HEADS):
                c.t2d_heads[i] = v
        ptr = ctypes.c_void_p()
        _lib.check(lib.mvldm_create(ctypes.byref(c), device.index or 0, ctypes.byref(ptr)))
        h.ptr, h.device, h.synced_versions = ptr, device, None
        # the C registry and the Python parameter table must describe the same state dict
        n = lib.mvldm_num_weights(ptr)
        names = [lib.mvldm_weight_name(ptr, i).decode() for i in range(n)]
        if set(names) != set(self._shapes):
            raise RuntimeError("mvldm_b200: library/Python state-dict key mismatch")
        return h

    def _versions(self):
        # (storage pointers change only through _apply / load_state_dict, which mark the module dirty and drop this cache;
        # walking the module tree on every call cost ~0.4 ms, the flat list ~40 us)
        n = self.__dict__["_vcalls"] = self.__dict__.get("_vcalls", 0) + 1
        plist = self.__dict__.get("_plist") if n % 64 else None  # (`p.data = other` swaps storage silently: re-walk now and then)
        if plist is None:
            plist = list(self.parameters())
            self.__dict__["_plist"] = plist
            self.__dict__["_pptrs"] = tuple(p.data_ptr() for p in plist)
        return self.__dict__["_pptrs"], tuple(map(_VERSION_OF, plist))     # C-level iteration: 2/3 of the genexpr's time

    def refresh_weights(self, force: bool = True) -> None:
        """(Re-)pack the module's parameters into the library's kernel layouts."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("mvldm_b200: the denoiser must live on a CUDA device (no CPU fallback)")
        h = self._ensure_handle(dev)
        # the (data_ptr, version) tuple is recomputed on every call (~700 entries, microseconds against a multi-ms
        # forward): in-place updates made in eval mode (EMA copy_to, AveragedModel.update_parameters, p.data.copy_) must
        # not leave a stale packed copy behind
        vers = self._versions()
        if not force and not self._dirty and h.synced_versions is not None and h.synced_versions == vers:
            self._dirty = False
            return
        lib = _lib.load()
        stream = _lib.current_stream_ptr(dev)
        dt = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}
        with _lib.on_device(dev):
            for key, p in self.state_dict(keep_vars=True).items():
                t = p.detach().contiguous()
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                _lib.check(lib.mvldm_set_weight(h.ptr, key.encode(), t.data_ptr(), shape, t.dim(), dt[t.dtype], stream))
            _lib.check(lib.mvldm_finalize_weights(h.ptr, stream))
        h.synced_versions = vers
        self._dirty = False

    def _launch_checked(self, launch):
        """Run ``launch()`` (enqueues the library call, returns its output) against up-to-date packed weights.  The cheap
        state (first use, load_state_dict, .to()) is checked in front of the call; the per-parameter version scan (tens of
        microseconds of Python that would otherwise sit between the caller and the GPU) runs AFTER the call is enqueued,
        under the GPU work.  If it finds an in-place update (optimizer / EMA step since the last call) the weights are
        re-packed and the call is issued again on the same stream: the result returned always reflects the current
        parameters, only the rare changed-weights call pays for a wasted forward."""
        h = self._h
        if self._dirty or h.ptr is None or h.synced_versions is None:
            self.refresh_weights(force=False)
            return launch()
        out = launch()
        if self._versions() != h.synced_versions:
            self.refresh_weights(force=True)
            out = launch()
        return out

    # ---- Denoiser.forward -----------------------------------------------------------------
    @torch.no_grad()
    def forward(self, latents: Tensor, timestep: Tensor, cond_state: Optional[Tensor] = None) -> Tensor:
        # cond_state never reaches a kernel, as in the reference: Variant A has no cross-attention block to feed
        # (mvunet.py:119), Variant B overwrites it with one zero token (mvunet.py:127-128)
        if latents.dim() != 5:
            raise ValueError("latents must be [batch, view, channel, height, width]")
        if not latents.is_cuda:
            raise RuntimeError("mvldm_b200: inputs must be CUDA tensors (no CPU fallback)")
        b, v, c, h, w = latents.shape
        if c != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {c}")
        if timestep.dtype != torch.int64:
            raise TypeError("timestep must be int64 (jaxtyping Int64 at mvunet.py:94)")
        # mvunet.py:102-105: [B] -> repeat over views, [B, V] -> flatten
        t = timestep.to(latents.device)
        t = t[:, None].expand(b, v) if t.dim() < 2 else t
        t = t.reshape(-1).contiguous()
        lat = latents.detach().to(torch.float32).contiguous()
        lib = _lib.load()

        def launch():
            out = torch.empty((b, v, self.out_channels, h, w), device=latents.device, dtype=torch.float32)
            with _lib.on_device(latents.device):
                _lib.check(lib.mvldm_forward(self._h.ptr, _lib.current_stream_ptr(latents.device), lat.data_ptr(),
                                             t.data_ptr(), b, v, h, w, out.data_ptr()))
            return out
        return self._launch_checked(launch)

    @torch.no_grad()
    def forward_scenes(self, latents: Tensor, timestep: Tensor, views_per_scene) -> Tensor:
        """Scenes with different view counts in one pass (``mvldm_forward_scenes``): ``latents`` [N, C, h, w] with the
        views of a scene contiguous, ``timestep`` int64 [N], ``sum(views_per_scene) == N``.  Returns [N, C_out, h, w].
        Used by ``DenoisingPath.step`` to run the conditional and the unconditional pass of CFG as one forward."""
        if latents.dim() != 4:
            raise ValueError("latents must be [view, channel, height, width]")
        if not latents.is_cuda:
            raise RuntimeError("mvldm_b200: inputs must be CUDA tensors (no CPU fallback)")
        n, c, h, w = latents.shape
        views = [int(v) for v in views_per_scene]
        if sum(views) != n or min(views) < 1:
            raise ValueError("views_per_scene must be positive and sum to the number of views")
        if c != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {c}")
        if timestep.dtype != torch.int64 or timestep.numel() != n:
            raise TypeError("timestep must be int64, one per view")
        t = timestep.to(latents.device).reshape(-1).contiguous()
        lat = latents.detach().to(torch.float32).contiguous()
        vp = (ctypes.c_int32 * len(views))(*views)

        def launch():
            out = torch.empty((n, self.out_channels, h, w), device=latents.device, dtype=torch.float32)
            with _lib.on_device(latents.device):
                _lib.check(_lib.load().mvldm_forward_scenes(self._h.ptr, _lib.current_stream_ptr(latents.device),
                                                            lat.data_ptr(), t.data_ptr(), len(views), vp, h, w, out.data_ptr()))
            return out
        return self._launch_checked(launch)

    @torch.no_grad()
    def forward_view_sharded(self, latents: Tensor, timestep: Tensor, v_total: int, exchange) -> Tensor:
        """View-group sharded forward: `latents` [1, V_local, C, h, w] are THIS rank's contiguous views of one scene of
        `v_total` views; `exchange` is a `ViewGroupExchange`.  Returns this rank's [1, V_local, C_out, h, w]."""
        b, v, c, h, w = latents.shape
        if b != 1:
            raise ValueError("view-group sharding handles one scene at a time (batch must be 1)")
        if not latents.is_cuda:
            raise RuntimeError("mvldm_b200: inputs must be CUDA tensors (no CPU fallback)")
        self.refresh_weights(force=False)
        t = timestep.to(latents.device)
        t = (t[:, None].expand(b, v) if t.dim() < 2 else t).reshape(-1).contiguous()
        lat = latents.detach().to(torch.float32).contiguous()
        out = torch.empty((b, v, self.out_channels, h, w), device=latents.device, dtype=torch.float32)
        lib = _lib.load()
        with _lib.on_device(latents.device):
            _lib.check(lib.mvldm_forward_sharded(
                self._h.ptr, _lib.current_stream_ptr(latents.device), lat.data_ptr(), t.data_ptr(), v, v_total,
                exchange.group_index, h, w, out.data_ptr(), exchange.send.data_ptr(), exchange.recv.data_ptr(),
                exchange.recv.numel() * exchange.recv.element_size(), exchange.callback, None))
        return out

    # ---- diagnostics ---------------------------------------------------------------------
    def last_launch_count(self) -> int:
        return _lib.load().mvldm_last_launch_count(self._h.ptr) if self._h.ptr else 0

    def set_profiling(self, on: bool = True) -> None:
        """Eager launches with a CUDA-event pair around every op (see mvldm_set_profiling)."""
        dev = next(self.parameters()).device
        _lib.check(_lib.load().mvldm_set_profiling(self._ensure_handle(dev).ptr, 1 if on else 0))

    def profile(self) -> dict:
        import json
        s = _lib.load().mvldm_profile_json(self._h.ptr)
        if s is None:
            _lib.check(1)
        return json.loads(s.decode())

    def enable_taps(self, on: bool = True) -> None:
        dev = next(self.parameters()).device
        _lib.check(_lib.load().mvldm_enable_taps(self._ensure_handle(dev).ptr, 1 if on else 0))

    def tap(self, name: str) -> Tensor:
        lib = _lib.load()
        dev = next(self.parameters()).device
        n = ctypes.c_int64()
        _lib.check(lib.mvldm_debug_tap(self._h.ptr, _lib.current_stream_ptr(dev), name.encode(), None, ctypes.byref(n)))
        out = torch.empty(n.value, device=dev, dtype=torch.float32)
        _lib.check(lib.mvldm_debug_tap(self._h.ptr, _lib.current_stream_ptr(dev), name.encode(), out.data_ptr(),
                                       ctypes.byref(n)))
        return out

    def workspace_bytes(self, b: int, v: int, h: int, w: int) -> int:
        self.refresh_weights(force=False)
        return int(_lib.load().mvldm_workspace_bytes(self._h.ptr, b, v, h, w))


DENOISER = {"mv_unet": MultiViewUNet}
DenoiserCfg = MultiViewUNetCfg


def get_denoiser(denoiser_cfg: DenoiserCfg, in_channels: int, out_channels: int) -> Denoiser:
    return DENOISER[denoiser_cfg.name](denoiser_cfg, in_channels, out_channels)
