"""GPU parity of the first-stage autoencoder (`first_stage_encode` / `last_stage_decode`,
reference src/model/diffusion_wrapper.py:278-298) against the restated diffusers `AutoencoderKL` of
oracle/diffusers_shim/diffusers/vae.py.

PARITY UNPINNED for this row: the reference only *calls* diffusers here and ships no VAE vectors (see the shim's header), so
the checker is a restatement of the published module, not an execution of it.  Tolerance: bf16 activations with fp32
accumulation, judged like the denoiser against the checker's own bf16-autocast drift: err <= max(2 x drift, 3e-2)
(max-abs / max-abs)."""
import os
import sys

import pytest
import torch

import mvldm_b200 as mv
from helpers import ROOT, rel_err, rms_err

sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_shim"))
from diffusers.vae import AutoencoderKL as RefVAE  # noqa: E402

pytestmark = pytest.mark.gpu
VAE_TOL = 3e-2


def _pair(kwargs, seed=0):
    torch.manual_seed(seed)
    ref = RefVAE(**kwargs).eval()
    for n, p in ref.named_parameters():            # perturb the norm affines so gamma / beta handling is exercised
        if "norm" in n:
            p.data.add_(0.1 * torch.randn_like(p))
    ours = mv.AutoencoderKL(**kwargs)
    ours.load_state_dict(ref.state_dict())         # strict
    return ref, ours.cuda().eval()


def _drift(fn, ref_out):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y = fn().float().cpu()
    return rel_err(y, ref_out)


SMALL = dict(down_block_types=("DownEncoderBlock2D",) * 2, up_block_types=("UpDecoderBlock2D",) * 2,
             block_out_channels=(64, 128), layers_per_block=1)
SD = dict(down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
          block_out_channels=(128, 256, 512, 512), layers_per_block=2)


@pytest.mark.parametrize("name,kwargs,n,size", [("small", SMALL, 3, 64), ("sd21", SD, 2, 64), ("sd21", SD, 1, 256)])
def test_vae_decode_and_encode_match_restated_diffusers(name, kwargs, n, size):
    ref, ours = _pair(kwargs)
    f = 2 ** (len(kwargs["block_out_channels"]) - 1)
    torch.manual_seed(1)
    z = torch.randn(n, 4, size // f, size // f)
    x = torch.rand(n, 3, size, size) * 2 - 1
    with torch.no_grad():
        img_ref = ref.decode(z).sample
        mom_ref = ref.encode(x).latent_dist.parameters
    img = ours.decode(z.cuda()).sample.cpu()
    mom = ours.encode(x.cuda()).latent_dist.parameters.cpu()
    assert img.shape == img_ref.shape and mom.shape == mom_ref.shape
    ref_gpu = RefVAE(**kwargs).eval()
    ref_gpu.load_state_dict(ref.state_dict())
    ref_gpu = ref_gpu.cuda()
    d_dec = _drift(lambda: ref_gpu.decode(z.cuda()).sample, img_ref)
    d_enc = _drift(lambda: ref_gpu.encode(x.cuda()).latent_dist.parameters, mom_ref)
    e_dec, e_enc = rel_err(img, img_ref), rel_err(mom, mom_ref)
    print(f"{name} {size}px: decode err {e_dec:.3e} (bf16-autocast drift {d_dec:.3e}), encode err {e_enc:.3e} (drift {d_enc:.3e}), "
          f"launches {ours.last_launch_count()}")
    assert e_dec < max(2 * d_dec, VAE_TOL) and rms_err(img, img_ref) < max(2 * d_dec, VAE_TOL)
    assert e_enc < max(2 * d_enc, VAE_TOL) and rms_err(mom, mom_ref) < max(2 * d_enc, VAE_TOL)
    # CUDA-graph replay is bit-stable
    assert torch.equal(ours.decode(z.cuda()).sample.cpu(), img)
    assert torch.equal(ours.encode(x.cuda()).latent_dist.parameters.cpu(), mom)


def test_first_and_last_stage_helpers_follow_the_wrapper():
    """DiffusionWrapper.first_stage_encode / last_stage_decode (diffusion_wrapper.py:278-298): [0,1] images -> latents x 0.18215
    (posterior sample drawn on the host generator) and latents / 0.18215 -> images clamped to [0,1]"""
    ref, ours = _pair(SMALL, seed=2)
    torch.manual_seed(3)
    imgs = torch.rand(1, 2, 3, 64, 64)
    g1 = torch.Generator(device="cuda").manual_seed(11)
    lat = mv.first_stage_encode(ours, imgs.cuda(), generator=g1)
    assert lat.shape == (1, 2, 4, 32, 32)
    with torch.no_grad():
        post = ref.encode(imgs.reshape(2, 3, 64, 64) * 2 - 1).latent_dist
    g2 = torch.Generator(device="cuda").manual_seed(11)
    noise = torch.randn(post.mean.shape, generator=g2, device="cuda").cpu()
    expect = (post.mean + post.std * noise) * 0.18215
    assert rel_err(lat.reshape(2, 4, 32, 32).cpu(), expect) < VAE_TOL
    out = mv.last_stage_decode(ours, lat)
    assert out.shape == (1, 2, 3, 64, 64) and float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    with torch.no_grad():
        want = (ref.decode(lat.reshape(2, 4, 32, 32).cpu() / 0.18215).sample / 2 + 0.5).clamp(0, 1)
    assert (out.reshape(2, 3, 64, 64).cpu() - want).abs().max() < 3e-2


def test_vae_rejects_what_it_does_not_implement():
    with pytest.raises(ValueError):
        mv.AutoencoderKL(block_out_channels=(48,))                     # channels not a multiple of 64
    with pytest.raises(ValueError):
        mv.AutoencoderKL(down_block_types=("AttnDownEncoderBlock2D",))
    ours = mv.AutoencoderKL(**SMALL).cuda()
    with pytest.raises(RuntimeError):
        ours.decode(torch.zeros(1, 4, 8, 8))                           # CPU tensor: no fallback
    with pytest.raises(ValueError):
        ours.encode(torch.zeros(1, 3, 31, 31, device="cuda"))
    with pytest.raises(RuntimeError):
        ours.decode(torch.zeros(1, 4, 4, 4, device="cuda"))            # 16 tokens in the mid attention: below the 64-token tile


def test_anchored_sampling_with_vae_hand_over(gpu_models):
    """BASELINE config 3 with the reference's image-space hand-over: anchors are decoded, clipped to [0,1] and re-encoded
    before they condition the chunk calls (diffusion_wrapper.py:722,779 + first/last_stage)."""
    ref, vae = _pair(SD, seed=5)
    m = gpu_models(0, True)
    path = mv.DenoisingPath(m, mv.DDIMScheduler(clip_sample=False), use_cfg=False)
    path.set_timesteps(3)
    torch.manual_seed(6)
    T = 10
    ctx = torch.randn(1, 1, 4, 32, 32, device="cuda")
    noise = torch.randn(1, T, 4, 32, 32, device="cuda")
    extr = torch.eye(4, device="cuda").expand(1, T + 1, 4, 4).clone()
    extr[0, :, 0, 3] = 0.1 * torch.arange(T + 1, device="cuda")
    intr = torch.tensor([[1.2, 0, 0.5], [0, 1.2, 0.5], [0, 0, 1.0]], device="cuda").expand(1, T + 1, 3, 3).clone()
    g = torch.Generator(device="cuda").manual_seed(1)
    lat_v, done_v, plan = mv.sample_anchored(path, ctx, extr[:, :1], intr[:, :1], extr[:, 1:], intr[:, 1:], noise,
                                             autoencoder=vae, generator=g)
    lat_l, done_l, _ = mv.sample_anchored(path, ctx, extr[:, :1], intr[:, :1], extr[:, 1:], intr[:, 1:], noise)
    assert torch.equal(done_v, done_l) and torch.isfinite(lat_v).all()
    A = plan.anchors
    assert torch.equal(lat_v[:, A], lat_l[:, A])                         # phase 1 is the same call
    rest = [t for _, c in plan.chunks for t in c]
    assert rest and not torch.equal(lat_v[:, rest], lat_l[:, rest])      # phase 2 saw the round-tripped anchors


def test_sample_images_is_encode_sample_decode(gpu_models):
    """DiffusionWrapper.sample end to end (diffusion_wrapper.py:455-490): images in -> images out equals the three stages
    composed by hand with the same posterior noise and the same x_T"""
    ref, vae = _pair(SD, seed=7)
    m = gpu_models(0, True)
    path = mv.DenoisingPath(m, mv.DDIMScheduler(clip_sample=False), use_cfg=False)
    path.set_timesteps(2)
    torch.manual_seed(8)
    imgs = torch.rand(1, 2, 3, 256, 256, device="cuda")
    extr = torch.eye(4, device="cuda").expand(1, 5, 4, 4).clone()
    extr[0, :, 0, 3] = 0.2 * torch.arange(5, device="cuda")
    intr = torch.tensor([[1.2, 0, 0.5], [0, 1.2, 0.5], [0, 0, 1.0]], device="cuda").expand(1, 5, 3, 3).clone()
    x_T = torch.randn(1, 3, 4, 32, 32, device="cuda")
    out = path.sample_images(vae, imgs, extr, intr, x_T=x_T, generator=torch.Generator(device="cuda").manual_seed(3))
    assert out.shape == (1, 3, 3, 256, 256) and float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    ctx = mv.first_stage_encode(vae, imgs, torch.Generator(device="cuda").manual_seed(3))
    want = mv.last_stage_decode(vae, path.sample(ctx, x_T, extr, intr))
    assert torch.equal(out, want)
