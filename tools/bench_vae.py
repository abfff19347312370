"""SD-2.1 VAE (83.7 M parameters, random init) on one B200: decode / encode of n views at 256x256, CUDA events around graph
replays, and the per-category profile of one decode."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mvldm_b200 as mv  # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    torch.manual_seed(0)
    vae = mv.AutoencoderKL.from_pretrained("stabilityai/stable-diffusion-2-1", subfolder="vae").cuda().eval()
    out = {}
    for n in (1, 8):
        z = torch.randn(n, 4, 32, 32, device="cuda")
        x = torch.rand(n, 3, 256, 256, device="cuda") * 2 - 1
        out[f"decode_{n}x256px_ms"] = timed(lambda: vae.decode(z))
        out[f"decode_{n}_launches"] = vae.last_launch_count()
        out[f"encode_{n}x256px_ms"] = timed(lambda: vae.encode(x))
        out[f"encode_{n}_launches"] = vae.last_launch_count()
    # algorithmic conv FLOPs of one 256-px decode / encode (2*MAC): from the library's own per-op count
    lib = mv._lib.load()
    mv._lib.check(lib.mvldm_set_profiling(vae._h.ptr, 1))
    z = torch.randn(8, 4, 32, 32, device="cuda")
    vae.decode(z)
    prof = json.loads(lib.mvldm_profile_json(vae._h.ptr).decode())["categories"]
    mv._lib.check(lib.mvldm_set_profiling(vae._h.ptr, 0))
    out["decode_8_profile"] = {k: {"launches": c["launches"], "us": round(c["us"], 1), "gflop": round(c["gflop"], 1)} for k, c in prof.items()}
    gf = sum(c["gflop"] for c in prof.values())
    out["decode_8_gflop"] = gf
    out["decode_8_tflops_graph"] = gf / out["decode_8x256px_ms"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
