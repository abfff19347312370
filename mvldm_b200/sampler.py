"""The denoising loop of the reference's ``DiffusionWrapper`` on the CUDA path.

Mirrors reference ``src/model/diffusion_wrapper.py``: ``ray_encode`` / ``generate_image_rays`` (:169-190,
:301-322), ``step`` (:413-453) and the loop of ``sample`` (:455-490) between the two VAE calls (latents in,
latents out).  Input concat, the denoiser forward(s), CFG and the DDIM update all run in this package's
kernels; this module only sequences them.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from .scheduler import DDIMScheduler, DDPMScheduler, fused_cfg_ddim_step, fused_cfg_ddpm_step


def ray_encode(extrinsics: Tensor, intrinsics: Tensor, h: int, w: int, use_plucker: bool = False,
               num_origin_octaves: int = 0, num_direction_octaves: int = 0, srt_ray_encoding: bool = False) -> Tensor:
    """extrinsics [B,V,4,4] (cam-to-world), intrinsics [B,V,3,3] (normalised) -> ray encodings [B,V,C,h,w]
    (diffusion_wrapper.py:301-322, srt_ray_encoding=false).  Octaves 0 / 0: ``use_ray_encoding: false`` (the released settings,
    baseline.yaml:50-51), C = 6.  Octaves > 0: ``use_ray_encoding: true`` (config/main.yaml:28-33): origins / directions are
    replaced by their ``PositionalEncoding`` (diffusion_wrapper.py:115-126), C = 6*origin_octaves + 6*direction_octaves.
    ``srt_ray_encoding``: the SRT ``RayEncoder`` instead (diffusion_wrapper.py:104-113,311-315; src/model/srt/layers.py:9-58)."""
    if not extrinsics.is_cuda:
        raise RuntimeError("mvldm_b200: ray_encode needs CUDA tensors (no CPU fallback)")
    B, V = extrinsics.shape[:2]
    e = extrinsics.detach().to(torch.float32).contiguous()
    k = intrinsics.detach().to(torch.float32).contiguous()
    C = (6 * num_origin_octaves if num_origin_octaves > 0 else 3) + (6 * num_direction_octaves if num_direction_octaves > 0 else 3)
    out = torch.empty((B, V, C, h, w), device=e.device, dtype=torch.float32)
    with _lib.on_device(e.device):
        _lib.check(_lib.load().mvldm_raymap_encoded(_lib.current_stream_ptr(e.device), e.data_ptr(), k.data_ptr(), B * V, h, w,
                                                    1 if use_plucker else 0, num_origin_octaves, num_direction_octaves,
                                                    1 if srt_ray_encoding else 0, out.data_ptr()))
    return out


def build_inputs(x_t: Tensor, context_latents: Optional[Tensor], rays: Tensor, ray_view_offset: int = 0,
                 out: Optional[Tensor] = None) -> Tensor:
    """[latent | mask | rays] per view, context views first (diffusion_wrapper.py:429-432 / :438).  ``out``: an optional
    contiguous fp32 destination of B*(v_c+v_t) views (e.g. a slice of a larger batch)."""
    B, v_t, _, h, w = x_t.shape
    v_c = 0 if context_latents is None else context_latents.shape[1]
    R = rays.shape[2]
    if out is None:
        out = torch.empty((B, v_c + v_t, 5 + R, h, w), device=x_t.device, dtype=torch.float32)
    elif out.numel() != B * (v_c + v_t) * (5 + R) * h * w or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("build_inputs: `out` must be a contiguous fp32 buffer of B*(v_c+v_t) views")
    x = x_t.detach().to(torch.float32).contiguous()
    c = context_latents.detach().to(torch.float32).contiguous() if v_c else None
    r = rays.detach().to(torch.float32).contiguous()
    with _lib.on_device(x.device):
        _lib.check(_lib.load().mvldm_build_inputs(_lib.current_stream_ptr(x.device), x.data_ptr(),
                                                  c.data_ptr() if v_c else None, r.data_ptr(), B, v_c, v_t,
                                                  rays.shape[1], ray_view_offset, R, h * w, out.data_ptr()))
    return out


class DenoisingPath:
    """``step`` / ``sample`` of the reference wrapper with the same knobs (use_cfg, cfg_scale, use_plucker)."""

    def __init__(self, denoiser, scheduler: DDIMScheduler, use_cfg: bool = False, cfg_scale: float = 3.0,
                 use_plucker: bool = False, batch_cfg: bool = True, num_origin_octaves: int = 0,
                 num_direction_octaves: int = 0, srt_ray_encoding: bool = False):
        """``batch_cfg``: run the conditional (v_c + v_t views) and unconditional (v_t views) forwards of a CFG step as
        ONE pass over 2B scenes of unequal view counts (``forward_scenes``) instead of two back-to-back forwards."""
        self.denoiser, self.scheduler = denoiser, scheduler
        self.use_cfg, self.cfg_scale, self.use_plucker = use_cfg, cfg_scale, use_plucker
        self.batch_cfg = batch_cfg
        # use_ray_encoding: true (config/main.yaml:28-33): octave counts of the positional encoding of the ray maps
        self.num_origin_octaves, self.num_direction_octaves = num_origin_octaves, num_direction_octaves
        self.srt_ray_encoding = srt_ray_encoding
        self.generator = None  # torch.Generator for the DDPM scheduler's variance noise (None: the default CUDA generator)
        self._t_cache = {}     # (B, v_c, v_t, ts, device) -> (timesteps [B, v_c+v_t], target timesteps [B, v_t])

    def set_timesteps(self, num: int) -> None:
        self.scheduler.set_timesteps(num)

    def step(self, model, x_t: Tensor, ts, context_inputs: Tensor, ray_encodings: Tensor,
             target_mask: Optional[Tensor] = None) -> Tensor:
        """diffusion_wrapper.py:413-453.  ``context_inputs`` is [B, v_c, 5, h, w] (latent + zero mask) as in the
        reference; ``target_mask`` is accepted for signature parity (it is all ones, :476)."""
        B, v_c = context_inputs.shape[:2]
        v_t = x_t.shape[1]
        ts = int(ts)
        dev = x_t.device
        x_in = self.scheduler.scale_model_input(x_t, ts)
        key = (B, v_c, v_t, ts, dev)
        if key not in self._t_cache:   # timestep 0 for context views, ts for target views (diffusion_wrapper.py:419-428)
            t_t = torch.full((B, v_t), ts, dtype=torch.long, device=dev)
            self._t_cache[key] = (torch.cat([torch.zeros((B, v_c), dtype=torch.long, device=dev), t_t], dim=1), t_t)
        t_all, t_t = self._t_cache[key]
        if self.use_cfg and self.batch_cfg and hasattr(model, "forward_scenes"):
            h, w = x_t.shape[-2:]
            V, R = v_c + v_t, ray_encodings.shape[2]
            kb = ("cfg", B, v_c, v_t, ts, dev)
            if kb not in self._t_cache:
                self._t_cache[kb] = torch.cat([t_all.reshape(-1), t_t.reshape(-1)])
            buf = torch.empty((B * (V + v_t), 5 + R, h, w), device=dev, dtype=torch.float32)
            build_inputs(x_in, context_inputs[:, :, :4], ray_encodings, out=buf[:B * V])
            build_inputs(x_in, None, ray_encodings, ray_view_offset=v_c, out=buf[B * V:])
            pred = model.forward_scenes(buf, self._t_cache[kb], [V] * B + [v_t] * B)
            pred_c = pred[:B * V].view(B, V, -1, h, w)
            pred_u = pred[B * V:].view(B, v_t, -1, h, w)
            return self._update(pred_c, pred_u, v_c, ts, x_t)
        inputs = build_inputs(x_in, context_inputs[:, :, :4], ray_encodings)
        pred_c = model.forward(inputs, t_all)
        pred_u = None
        if self.use_cfg:
            pred_u = model.forward(build_inputs(x_in, None, ray_encodings, ray_view_offset=v_c), t_t)
        return self._update(pred_c, pred_u, v_c, ts, x_t)

    def _update(self, pred_c: Tensor, pred_u: Optional[Tensor], v_c: int, ts: int, x_t: Tensor) -> Tensor:
        """CFG compose + scheduler step (diffusion_wrapper.py:444-451) as one kernel, for either scheduler of the registry"""
        if isinstance(self.scheduler, DDPMScheduler):
            return fused_cfg_ddpm_step(self.scheduler, pred_c, pred_u, self.cfg_scale, v_c, ts, x_t, self.generator)
        return fused_cfg_ddim_step(self.scheduler, pred_c, pred_u, self.cfg_scale, v_c, ts, x_t)

    @torch.no_grad()
    def sample(self, context_latents: Tensor, x_T: Tensor, extrinsics: Tensor, intrinsics: Tensor,
               record: Optional[list] = None) -> Tensor:
        """The loop of DiffusionWrapper.sample (diffusion_wrapper.py:473-488) on latents."""
        B, v_c, _, h, w = context_latents.shape
        x_t = x_T * self.scheduler.init_noise_sigma
        ctx = torch.cat([context_latents, torch.zeros_like(context_latents[:, :, :1])], dim=2)
        rays = ray_encode(extrinsics, intrinsics, h, w, self.use_plucker, self.num_origin_octaves, self.num_direction_octaves,
                          self.srt_ray_encoding)
        for ts in self.scheduler.timesteps:
            x_t = self.step(self.denoiser, x_t, ts, ctx, rays)
            if record is not None:
                record.append((int(ts), x_t.clone()))
        return x_t

    @torch.no_grad()
    def sample_images(self, autoencoder, context_images: Tensor, extrinsics: Tensor, intrinsics: Tensor,
                      num_target_views: Optional[int] = None, x_T: Optional[Tensor] = None, generator=None) -> Tensor:
        """DiffusionWrapper.sample (diffusion_wrapper.py:455-490) end to end: context images [B, v_c, 3, H, W] in [0, 1] ->
        first_stage_encode (posterior sample x 0.18215) -> DDIM / DDPM loop on the latents -> last_stage_decode -> target
        images [B, v_t, 3, H, W] in [0, 1].  ``extrinsics`` / ``intrinsics`` hold the context views first, then the targets.
        ``x_T``: the initial noise [B, v_t, 4, H/8, W/8]; when omitted it is drawn on the CPU generator and moved to the
        device, as the reference does (:473)."""
        from .autoencoder import first_stage_encode, last_stage_decode
        ctx = first_stage_encode(autoencoder, context_images, generator)
        v_t = (extrinsics.shape[1] - ctx.shape[1]) if num_target_views is None else num_target_views
        if x_T is None:
            x_T = torch.randn((ctx.shape[0], v_t, *ctx.shape[2:])).to(ctx.device)
        lat = self.sample(ctx, x_T, extrinsics, intrinsics)
        return last_stage_decode(autoencoder, lat)

