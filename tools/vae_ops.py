import json, sys, os
sys.path.insert(0, "/root/repo")
import torch, mvldm_b200 as mv
vae = mv.AutoencoderKL.from_pretrained("stabilityai/stable-diffusion-2-1", subfolder="vae").cuda().eval()
lib = mv._lib.load()
z = torch.randn(8, 4, 32, 32, device="cuda")
vae.decode(z)
mv._lib.check(lib.mvldm_set_profiling(vae._h.ptr, 1))
for _ in range(2):
    vae.decode(z)
ops = json.loads(lib.mvldm_profile_json(vae._h.ptr).decode())["ops"]
for o in ops:
    if o["cat"] in ("groupnorm", "upsample"):
        print(o["cat"], o["what"], round(o["us"], 1))
