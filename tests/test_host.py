"""CPU: host-side mirror of the reference interfaces, the C-ABI surface, and scene sharding (gloo)."""
import os
import re
import socket

import numpy as np
import pytest
import torch

import mvldm_b200 as mv
from helpers import GOLD, ROOT
from mvldm_b200 import _lib
from oracle import mvldm_oracle as O


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mvldm_b200.h")).read()
    declared = set(re.findall(r"\b(mvldm_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libmvldm_b200.so does not export {name}"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.mvldm_version() >= 100


def test_struct_layouts_match_header():
    import ctypes
    assert ctypes.sizeof(_lib.Config) == 4 * (3 + 4 + 6 + 1 + 4 + 1 + 4 + 2)
    assert ctypes.sizeof(_lib.ASeg) == 8 + 6 * 4 + 9 + 9 + 2 + 9 * 4      # ptr, 6 ints, 2x9 int8 (+2 pad), 9 ints
    assert _lib.GemmDesc.seg.offset == 8 and ctypes.sizeof(_lib.GemmDesc) % 8 == 0


def test_no_gpu_is_an_error_not_a_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = mv.MultiViewUNet(mv.default_cfg(), 11, 4)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2, 11, 32, 32), torch.zeros(1, dtype=torch.int64))
    with pytest.raises(RuntimeError):
        mv.DDIMScheduler(clip_sample=False).step(torch.zeros(1, 1, 4, 8, 8), 0, torch.zeros(1, 1, 4, 8, 8))


def test_state_dict_keys_match_reference_module(oracle_cfg):
    m = mv.MultiViewUNet(mv.default_cfg(), 11, 4)
    sd = m.state_dict()
    ref = O.param_shapes(oracle_cfg)      # == the reference module's keys (make_golden loads them strict=True)
    assert set(sd) == set(ref)
    assert all(tuple(sd[k].shape) == ref[k][0] for k in ref)
    # zero-init proj_out like the reference (mvdream/attention.py:406-411)
    assert float(sd["cross_attn_blocks_mid.0.proj_out.weight"].abs().max()) == 0.0
    assert mv.get_denoiser(mv.default_cfg(), 11, 4).__class__ is mv.MultiViewUNet
    assert "mv_unet" in mv.DENOISER


def test_variant_b_state_dict_keys_match_reference_module():
    """`pretrained_from` set: the SD-2.1 topology (920 keys; make_golden loads them into the reference strict=True)."""
    cfg = mv.default_cfg()
    cfg.pretrained_from = "stabilityai/stable-diffusion-2-1"
    m = mv.MultiViewUNet(cfg, 11, 4)
    sd = m.state_dict()
    ref = O.param_shapes(O.OracleCfg(variant_b=True))
    assert set(sd) == set(ref) and len(sd) == 920
    assert all(tuple(sd[k].shape) == ref[k][0] for k in ref)
    # only the multi-view blocks' proj_out is zero-initialised, not the Transformer2DModels'
    assert float(sd["cross_attn_blocks_mid.0.proj_out.weight"].abs().max()) == 0.0
    assert float(sd["unet.mid_block.attentions.0.proj_out.weight"].abs().max()) > 0.0


def test_standard_transformer_state_dict_keys_match_reference_module():
    """multi_view_attention "standard": keys / shapes of StandardTransformer.transformer (make_golden loads the oracle's
    dict into the reference module strict=True, so the oracle's table IS the reference's)."""
    m = mv.MultiViewUNet(mv.standard_cfg(), 11, 4)
    sd = m.state_dict()
    ref = O.param_shapes(O.OracleCfg(mv_block="standard"))
    assert set(sd) == set(ref)
    assert all(tuple(sd[k].shape) == ref[k][0] for k in ref)
    assert tuple(sd["cross_attn_blocks_decoder.3.transformer.layers.0.0.fn.to_qkv.weight"].shape) == (960, 320)
    assert not any(".proj_out." in k or ".transformer_blocks." in k for k in sd if k.startswith("cross_attn_blocks_"))
    wide = mv.MultiViewUNet(mv.standard_cfg(d_mlp_multiplier=4, num_layers=2), 11, 4).state_dict()
    assert tuple(wide["cross_attn_blocks_mid.0.transformer.layers.1.1.fn.net.0.weight"].shape) == (5120, 1280)
    for bad in (dict(d_dot=32), dict(downscale=2), dict(pos_enc=True), dict(d_mlp=640), dict(num_layers=0)):
        cfg = mv.standard_cfg()
        for k, v in bad.items():
            setattr(cfg.multi_view_attention, k, v)                 # d_mlp AND d_mlp_multiplier both set: the reference asserts
        with pytest.raises(ValueError):
            mv.MultiViewUNet(cfg, 11, 4)


def test_vae_module_matches_the_diffusers_layout():
    """AutoencoderKL drop-in (reference src/model/autoencoder/__init__.py:15-43): same state-dict keys / shapes as the
    restated diffusers module (oracle/diffusers_shim/diffusers/vae.py) for the SD-2.1 VAE and for kl.yaml's default config,
    83.65 M parameters for the SD VAE, strict loading incl. the pre-0.15 attention names, registry + factory."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "diffusers_shim"))
    from diffusers.vae import AutoencoderKL as RefVAE
    ref = RefVAE.from_pretrained("stabilityai/stable-diffusion-2-1", subfolder="vae")
    ours = mv.get_autoencoder(mv.AutoencoderCfg("kl", "stabilityai/stable-diffusion-2-1", mv.AutoencoderKLCfg()))
    sd = ref.state_dict()
    assert set(sd) == set(ours.state_dict()) and len(sd) == 248
    assert all(tuple(sd[k].shape) == tuple(ours.state_dict()[k].shape) for k in sd)
    assert sum(p.numel() for p in ours.parameters()) == 83653863
    assert not ours.load_state_dict(sd, strict=True).missing_keys
    old = {}
    for k, v in sd.items():                 # a checkpoint written before diffusers 0.15: query/key/value/proj_attn/norm, 1x1 convs
        for new, dep in (("to_q", "query"), ("to_k", "key"), ("to_v", "value"), ("to_out.0", "proj_attn"), ("group_norm", "norm")):
            if f".attentions.0.{new}." in k:
                k = k.replace(f".attentions.0.{new}.", f".attentions.0.{dep}.")
                v = v[:, :, None, None] if (v.dim() == 2) else v
        old[k] = v
    assert set(old) != set(sd)
    res = ours.load_state_dict(old, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(ours.state_dict()["decoder.mid_block.attentions.0.to_q.weight"], sd["decoder.mid_block.attentions.0.to_q.weight"])
    small = mv.get_autoencoder(mv.AutoencoderCfg("kl", None, mv.AutoencoderKLCfg()))          # kl.yaml defaults: one 64-wide block
    assert set(small.state_dict()) == set(RefVAE().state_dict())
    assert small.config.scaling_factor == 0.18215 and "kl" in mv.AUTOENCODERS


def test_scheduler_registry_has_both_reference_entries():
    """reference src/model/scheduler/__init__.py:19-40: SCHEDULER = {"ddim", "ddpm"}, built from the dataclass kwargs"""
    assert set(mv.SCHEDULER) == {"ddim", "ddpm"}
    s = mv.get_scheduler(mv.SchedulerCfg("ddpm", 1000, 1000, None, mv.DDPMSchedulerCfg(trained_betas="None")))   # ddpm.yaml:11
    assert isinstance(s, mv.DDPMScheduler) and s.config.clip_sample and s.config.prediction_type == "epsilon"
    s.set_timesteps(50)
    o = O.DDPMOracle()
    o.set_timesteps(50)
    assert torch.equal(s.timesteps, o.timesteps) and s.init_noise_sigma == 1.0
    sa, s1a, c0, ct, sg = s.coefficients(0)
    assert sg == 0.0 and abs(c0 - 1.0) < 1e-12 and ct == 0.0                     # the last step returns the (clipped) x0
    x, n = torch.randn(2, 3, 4, 8, 8), torch.randn(2, 3, 4, 8, 8)
    t = torch.tensor([10, 900])
    assert torch.allclose(s.add_noise(x, n, t), o.add_noise(x, n, t))
    d = mv.get_scheduler(mv.SchedulerCfg("ddim", 1000, 25, None, mv.DDIMSchedulerCfg(clip_sample=False)))
    assert isinstance(d, mv.DDIMScheduler)


def test_unsupported_configs_raise():
    cfg = mv.default_cfg()
    cfg.pretrained_from = "runwayml/stable-diffusion-v1-5"     # 768-wide context: mvunet.py:127 cannot drive it either
    with pytest.raises(ValueError):
        mv.MultiViewUNet(cfg, 11, 4)
    cfg = mv.default_cfg()
    cfg.autoencoder.down_block_types = ["CrossAttnDownBlock2D"] * 4
    with pytest.raises(ValueError):
        mv.MultiViewUNet(cfg, 11, 4)


def test_scheduler_mirror_matches_golden():
    g = np.load(os.path.join(GOLD, "g4_ddim.npz"))
    cfg = mv.SchedulerCfg("ddim", 1000, 25, None, mv.DDIMSchedulerCfg(clip_sample=False))
    s = mv.get_scheduler(cfg)
    np.testing.assert_array_equal(s.alphas_cumprod.numpy(), g["alphas_cumprod"])
    assert s.init_noise_sigma == 1.0 and s.config.prediction_type == "epsilon"
    for n in (25, 50, 70):
        s.set_timesteps(n)
        np.testing.assert_array_equal(s.timesteps.numpy(), g[f"timesteps_{n}"])
        co = np.array([s.coefficients(int(t)) for t in s.timesteps])
        np.testing.assert_allclose(co, g[f"coef_{n}"], rtol=0, atol=0)
    x = torch.randn(3, 2, 4, 8, 8)
    n = torch.randn_like(x)
    t = torch.tensor([0, 500, 999])
    ref = O.DDIMOracle().add_noise(x, n, t)
    torch.testing.assert_close(s.add_noise(x, n, t), ref)
    assert s.scale_model_input(x, 3) is x
    with pytest.raises(NotImplementedError):
        mv.DDIMScheduler(clip_sample=True)
    with pytest.raises(ValueError):
        mv.DDIMScheduler(clip_sample=False).coefficients(10)      # set_timesteps not called


def test_scene_slices_partition():
    for n in (1, 7, 64):
        for ws in (1, 2, 4, 8):
            sl = [mv.scene_slice(n, r, ws) for r in range(ws)]
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(sl[i][1] == sl[i + 1][0] for i in range(ws - 1))
            assert max(b - a for a, b in sl) - min(b - a for a, b in sl) <= 1
    with pytest.raises(ValueError):
        mv.scene_slice(4, 2, 2)


def _gather_worker(rank, ws, port, n):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=ws)
    a, b = mv.scene_slice(n, rank, ws)
    local = torch.arange(a, b, dtype=torch.float32)[:, None].repeat(1, 3)
    full = mv.gather_scenes(local, n)
    assert full.shape == (n, 3) and torch.equal(full[:, 0], torch.arange(n, dtype=torch.float32))
    dist.destroy_process_group()


def test_gather_scenes_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_gather_worker, args=(2, port, 5), nprocs=2, join=True)


def _kv_worker(rank, ws, port):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=ws)
    ex = mv.ViewGroupExchange(v_local=1, v_total=ws, h=2, w=2, heads=8, device="cpu")
    n = 4 * 2 * 8 * 64
    assert ex.send.numel() == n and ex.recv.numel() == n * ws
    ex.send[:] = float(rank + 1)
    assert ex.group_index == rank and ex.side is None    # no CUDA device: the collective runs inline in BEGIN
    ex.begin(100)                                        # a coarser level uses a prefix of the buffers
    ex.end()
    for r in range(ws):
        assert torch.all(ex.recv[r * 100:(r + 1) * 100] == float(r + 1))     # rank order == view order
    assert ex.calls == 1 and ex.bytes_sent == 200
    assert mv.view_slice(8, rank, ws) == (rank * 4, rank * 4 + 4)
    dist.destroy_process_group()


def test_view_group_exchange_gloo_world2():
    """host side of view-group sharding: the K|V all-gather lands the ranks' slabs in view order"""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_kv_worker, args=(2, port), nprocs=2, join=True)
    with pytest.raises(ValueError):
        mv.view_slice(7, 0, 2)


def test_anchored_plan_matches_reference_index_logic():
    """BASELINE config 3 / SURVEY.md §3.1 + App. D (index logic of diffusion_wrapper.py:689-885 emulated on frame ids):
    80 frames -> 3 anchors at 20/40/60, 26 sample() calls, 75 of 77 non-anchor frames generated, 2 dropped;
    278 frames -> 4 anchors, 92 calls."""
    p = mv.anchored_plan(list(range(80)), 4)
    assert p.anchors == [20, 40, 60]
    assert p.num_sample_calls == 26 and len(p.chunks) == 25
    covered = sorted(t for _, c in p.chunks for t in c)
    assert len(covered) == 75 == len(set(covered)) and len(p.dropped) == 2
    assert not set(covered) & set(p.anchors)
    assert all(len(c) == 3 for _, c in p.chunks)
    p = mv.anchored_plan(list(range(278)), 4)
    assert p.anchors == [69, 138, 207, 276] and p.num_sample_calls == 92
    # every chunk is conditioned on an anchor of the plan; nearest-anchor assignment for an isolated example
    p = mv.anchored_plan(list(range(20)), 4)
    assert p.anchors == [5, 10, 15] and all(a in p.anchors for a, _ in p.chunks)
    with pytest.raises(ValueError):
        mv.anchored_plan([0, 1], 4)
    # more than 4 anchors = the reference's iterative anchor rounds (diffusion_wrapper.py:744-792): refused, not approximated
    with pytest.raises(ValueError):
        mv.anchored_plan(list(range(80)), 8)
    assert mv.anchored_plan(list(range(80)), 2).anchors == [40]


def test_header_is_plain_c(tmp_path):
    """include/mvldm_b200.h is the drop-in boundary: it must compile as C (no C++ or torch types in the signatures)
    and declare exactly the symbols the library exports."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "mvldm_b200.h"\n'
                   "int main(void) { mvldm_config c; mvldm_gemm_desc d; (void)c; (void)d; return (int)sizeof(mvldm_config) == 0; }\n")
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc, str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_library_sass_uses_tcgen05_tmem_and_tma():
    """Static evidence that needs no GPU: the shipped library's SASS contains the Blackwell tensor-core, tensor-memory and
    TMA instructions (B200_PROFILING.md: tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG),
    it is built for sm_100a only, and carries no legacy warp-level MMA."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mvldm_b200", "libmvldm_b200.so")
    if not os.path.exists(lib):
        pytest.skip("library not built")
    elf = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and "sm_90" not in elf and "sm_80" not in elf
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    count = lambda m: sass.count(m)  # noqa: E731
    assert count("UTCHMMA") >= 50          # tcgen05.mma in the GEMM and attention kernels
    assert count("UTMALDG") >= 20          # TMA tensor loads (3-D / 4-D / 5-D)
    assert count("LDTM") >= 10 and count("STTM") >= 10   # tcgen05.ld / tcgen05.st (TMEM <-> registers)
    assert count("UTCBAR") >= 10           # tcgen05.commit -> mbarrier
    assert count("HMMA.") == 0 or count("UTCHMMA") > 0   # no mma.sync-only fallback
    assert "HGMMA" not in sass             # wgmma is sm_90a-only


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys;
    bounded sample, no GPU needed"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0
