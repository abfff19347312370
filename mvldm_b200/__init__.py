"""mvldm_b200 — B200-native (sm_100a) implementation of the MV-LDM denoising hot path.

Python here is plumbing: the reference-facing interfaces (`Denoiser`, `MultiViewUNet`, `get_denoiser`,
`DDIMScheduler`, `get_scheduler`, `DenoisingPath.step/sample`) forward to the C-ABI library
`libmvldm_b200.so` (hand-written CUDA; see include/mvldm_b200.h).
"""
from . import _lib
from .denoiser import (DENOISER, CrossAttentionCfg, Denoiser, DenoiserCfg, MultiViewUNet, MultiViewUNetCfg,
                       SpatialTransformer3DCfg,
                       UNet2DModelCfg, default_cfg, get_denoiser, standard_cfg)
from .anchored import AnchoredPlan, anchored_plan, sample_anchored
from .autoencoder import (AUTOENCODERS, AutoencoderCfg, AutoencoderKL, AutoencoderKLCfg, first_stage_encode, get_autoencoder,
                          last_stage_decode, sd21_vae_cfg)
from .sampler import DenoisingPath, build_inputs, ray_encode
from .scheduler import (SCHEDULER, DDIMScheduler, DDIMSchedulerCfg, DDPMScheduler, DDPMSchedulerCfg, SchedulerCfg,
                        fused_cfg_ddim_step, fused_cfg_ddpm_step, get_scheduler)
from .sharding import ViewGroupExchange, gather_scenes, scene_slice, view_slice

__all__ = [
    "DENOISER", "CrossAttentionCfg", "Denoiser", "DenoiserCfg", "MultiViewUNet", "MultiViewUNetCfg", "SpatialTransformer3DCfg",
    "UNet2DModelCfg", "default_cfg", "standard_cfg", "get_denoiser", "DenoisingPath", "build_inputs", "ray_encode", "SCHEDULER",
    "DDIMScheduler", "DDIMSchedulerCfg", "DDPMScheduler", "DDPMSchedulerCfg", "SchedulerCfg", "fused_cfg_ddim_step",
    "fused_cfg_ddpm_step", "get_scheduler", "gather_scenes",
    "scene_slice", "view_slice", "ViewGroupExchange", "AnchoredPlan", "anchored_plan", "sample_anchored",
    "AUTOENCODERS", "AutoencoderCfg", "AutoencoderKL", "AutoencoderKLCfg", "get_autoencoder", "first_stage_encode",
    "last_stage_decode", "sd21_vae_cfg",
]
