"""DDIM (and DDPM) schedulers with the surface the reference uses from ``diffusers.DDIMScheduler`` / ``DDPMScheduler``.

Mirrors reference ``src/model/scheduler/__init__.py:12-40`` (``SchedulerCfg``, ``SCHEDULER``, ``get_scheduler``)
and ``src/model/scheduler/ddim.py:10-18`` (``DDIMSchedulerCfg``).  Call sites kept working:
``diffusion_wrapper.py:198`` (set_timesteps), ``:370`` (add_noise), ``:417`` (scale_model_input), ``:451``
(step(...).prev_sample), ``:474`` (init_noise_sigma), ``:486`` (timesteps).

The schedule is host-side scalar arithmetic (restated from diffusers 0.27.2: linear betas, leading spacing,
eta = 0, epsilon prediction, no clipping/thresholding); the tensor update runs in the fused CUDA kernel
``mvldm_ddim_step``.
"""
from __future__ import annotations

import ctypes
from dataclasses import asdict, dataclass
from types import SimpleNamespace
from typing import Literal, Optional, Union

import numpy as np
import torch
from torch import Tensor

from . import _lib


@dataclass
class DDIMSchedulerCfg:
    num_train_timesteps: int = 1000
    beta_start: float = 0.0001
    beta_end: float = 0.02
    beta_schedule: str = "linear"
    trained_betas: Optional[Union[np.ndarray, list]] = None
    clip_sample: bool = True
    set_alpha_to_one: bool = True
    steps_offset: int = 0


@dataclass
class DDPMSchedulerCfg:
    """reference src/model/scheduler/ddpm.py:9-25 (config/model/scheduler/ddpm.yaml)"""
    num_train_timesteps: int = 1000
    beta_start: float = 0.0001
    beta_end: float = 0.02
    beta_schedule: str = "linear"
    trained_betas: Optional[Union[np.ndarray, list]] = None
    variance_type: str = "fixed_small"
    clip_sample: bool = True
    prediction_type: str = "epsilon"
    thresholding: bool = False
    dynamic_thresholding_ratio: float = 0.995
    clip_sample_range: float = 1.0
    sample_max_value: float = 1.0
    timestep_spacing: str = "leading"
    steps_offset: int = 0
    rescale_betas_zero_snr: bool = False


@dataclass
class SchedulerCfg:
    name: Literal["ddim", "ddpm"]
    num_train_timesteps: int
    num_inference_steps: int
    pretrained_from: Optional[str]
    kwargs: Union[DDIMSchedulerCfg, DDPMSchedulerCfg]


@dataclass
class DDIMSchedulerOutput:
    prev_sample: Tensor
    pred_original_sample: Optional[Tensor] = None


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon"):
        if trained_betas is not None:
            betas = torch.tensor(np.asarray(trained_betas), dtype=torch.float32)
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
        if clip_sample:
            raise NotImplementedError("clip_sample=True is not supported by the fused DDIM kernel "
                                      "(the reference sets clip_sample: False, config/model/scheduler/ddim.yaml:9)")
        if prediction_type != "epsilon":
            raise NotImplementedError("only epsilon prediction is supported (reference default)")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, clip_sample=clip_sample,
                                      set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                      prediction_type=prediction_type)
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def scale_model_input(self, sample: Tensor, timestep=None) -> Tensor:
        return sample

    def coefficients(self, timestep: int):
        """(sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)) for the fused kernel."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        prev = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return float(a_t ** 0.5), float((1 - a_t) ** 0.5), float(a_p ** 0.5), float((1 - a_p) ** 0.5)

    def step(self, model_output: Tensor, timestep, sample: Tensor, eta: float = 0.0, **kwargs) -> DDIMSchedulerOutput:
        if eta != 0.0:
            raise NotImplementedError("eta != 0 is not supported (the reference never passes eta)")
        out = fused_cfg_ddim_step(self, model_output, None, 1.0, 0, int(timestep), sample)
        return DDIMSchedulerOutput(prev_sample=out)

    def add_noise(self, original_samples: Tensor, noise: Tensor, timesteps: Tensor) -> Tensor:
        a = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)[timesteps]
        sa, s1 = a ** 0.5, (1 - a) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, s1 = sa.unsqueeze(-1), s1.unsqueeze(-1)
        return sa * original_samples + s1 * noise


def fused_cfg_ddim_step(sched: DDIMScheduler, eps_c: Tensor, eps_u: Optional[Tensor], cfg_scale: float, v_c: int,
                        timestep: int, x_t: Tensor) -> Tensor:
    """CFG compose (diffusion_wrapper.py:444/447) + DDIMScheduler.step (:451) in one kernel.
    eps_c [B, v_c+v_t, C, h, w] (the first v_c views are skipped), eps_u [B, v_t, C, h, w] or None."""
    if not x_t.is_cuda:
        raise RuntimeError("mvldm_b200: DDIM step needs CUDA tensors (no CPU fallback)")
    sa, s1a, sp, s1p = sched.coefficients(timestep)
    B, v_t = x_t.shape[:2]
    chw = x_t[0, 0].numel()
    if eps_c.shape[1] != v_c + v_t:
        raise ValueError("eps_c must hold v_c + v_t views")
    x = x_t.detach().to(torch.float32).contiguous()
    ec = eps_c.detach().to(torch.float32).contiguous()
    eu = eps_u.detach().to(torch.float32).contiguous() if eps_u is not None else None
    out = torch.empty_like(x)
    lib = _lib.load()
    with _lib.on_device(x.device):
        _lib.check(lib.mvldm_ddim_step(_lib.current_stream_ptr(x.device), ec.data_ptr(),
                                       eu.data_ptr() if eu is not None else None, float(cfg_scale), B, v_c, v_t, chw,
                                       x.data_ptr(), sa, s1a, sp, s1p, out.data_ptr(), None))
    return out


class DDPMScheduler:
    """``diffusers.DDPMScheduler`` on the surface the reference uses (the "ddpm" entry of its registry,
    src/model/scheduler/__init__.py:19-22): ancestral sampling with the posterior variance, epsilon prediction, optional
    clipping of the predicted x0.  Scalars on the host (restated from diffusers 0.27.2 ``DDPMScheduler.step``), the tensor
    update in ``mvldm_ddpm_step``; the variance noise is drawn with torch on the caller's generator, like diffusers'."""
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, variance_type: str = "fixed_small",
                 clip_sample: bool = True, prediction_type: str = "epsilon", thresholding: bool = False,
                 dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0, sample_max_value: float = 1.0,
                 timestep_spacing: str = "leading", steps_offset: int = 0, rescale_betas_zero_snr: bool = False):
        if trained_betas is not None and not (isinstance(trained_betas, str) and trained_betas == "None"):
            betas = torch.tensor(np.asarray(trained_betas), dtype=torch.float32)     # (ddpm.yaml spells null as "None")
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
        if prediction_type != "epsilon" or thresholding or rescale_betas_zero_snr:
            raise NotImplementedError("supported: epsilon prediction, no dynamic thresholding, no zero-SNR rescaling")
        if variance_type not in ("fixed_small", "fixed_large") or timestep_spacing != "leading":
            raise NotImplementedError("supported: variance_type fixed_small / fixed_large, timestep_spacing 'leading'")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, variance_type=variance_type, clip_sample=clip_sample,
                                      clip_sample_range=clip_sample_range, prediction_type=prediction_type,
                                      timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    set_timesteps = DDIMScheduler.set_timesteps
    scale_model_input = DDIMScheduler.scale_model_input
    add_noise = DDIMScheduler.add_noise

    def coefficients(self, timestep: int):
        """(sqrt(a_t), sqrt(1-a_t), c_x0, c_xt, sigma) of one ancestral step"""
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        prev = timestep - self.config.num_train_timesteps // n
        a_t = float(self.alphas_cumprod[timestep])
        a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else 1.0
        cur_alpha = a_t / a_p
        cur_beta = 1.0 - cur_alpha
        c_x0 = a_p ** 0.5 * cur_beta / (1.0 - a_t)
        c_xt = cur_alpha ** 0.5 * (1.0 - a_p) / (1.0 - a_t)
        var = (1.0 - a_p) / (1.0 - a_t) * cur_beta if self.config.variance_type == "fixed_small" else cur_beta
        sigma = max(var, 1e-20) ** 0.5 if timestep > 0 else 0.0
        return a_t ** 0.5, (1.0 - a_t) ** 0.5, c_x0, c_xt, sigma

    def step(self, model_output: Tensor, timestep, sample: Tensor, generator=None, **kwargs) -> DDIMSchedulerOutput:
        out = fused_cfg_ddpm_step(self, model_output, None, 1.0, 0, int(timestep), sample, generator)
        return DDIMSchedulerOutput(prev_sample=out)


def fused_cfg_ddpm_step(sched: "DDPMScheduler", eps_c: Tensor, eps_u: Optional[Tensor], cfg_scale: float, v_c: int,
                        timestep: int, x_t: Tensor, generator=None) -> Tensor:
    """CFG compose + DDPMScheduler.step in one kernel (same argument layout as ``fused_cfg_ddim_step``)."""
    if not x_t.is_cuda:
        raise RuntimeError("mvldm_b200: DDPM step needs CUDA tensors (no CPU fallback)")
    sa, s1a, c_x0, c_xt, sigma = sched.coefficients(timestep)
    B, v_t = x_t.shape[:2]
    chw = x_t[0, 0].numel()
    if eps_c.shape[1] != v_c + v_t:
        raise ValueError("eps_c must hold v_c + v_t views")
    x = x_t.detach().to(torch.float32).contiguous()
    ec = eps_c.detach().to(torch.float32).contiguous()
    eu = eps_u.detach().to(torch.float32).contiguous() if eps_u is not None else None
    noise = torch.randn(x.shape, generator=generator, device=x.device, dtype=torch.float32) if sigma > 0.0 else None
    out = torch.empty_like(x)
    clip = float(sched.config.clip_sample_range) if sched.config.clip_sample else 0.0
    with _lib.on_device(x.device):
        _lib.check(_lib.load().mvldm_ddpm_step(_lib.current_stream_ptr(x.device), ec.data_ptr(),
                                               eu.data_ptr() if eu is not None else None, float(cfg_scale), B, v_c, v_t, chw,
                                               x.data_ptr(), noise.data_ptr() if noise is not None else None, sa, s1a, c_x0,
                                               c_xt, sigma, clip, out.data_ptr()))
    return out


SCHEDULER = {"ddim": DDIMScheduler, "ddpm": DDPMScheduler}


def get_scheduler(cfg: SchedulerCfg) -> DDIMScheduler:
    if cfg.pretrained_from is not None:
        raise ValueError("scheduler.pretrained_from needs the HF hub; construct from kwargs instead")
    kw = asdict(cfg.kwargs) if not isinstance(cfg.kwargs, dict) else dict(cfg.kwargs)
    if isinstance(kw.get("trained_betas"), list):
        kw["trained_betas"] = np.array(kw["trained_betas"])
    return SCHEDULER[cfg.name](**kw)
