"""per-block parity of the fused-sequence-kernel forward against the oracle (debug): python tools/fused_taps_check.py [fuse]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import mvldm_b200 as mv
from oracle import mvldm_oracle as O
from helpers import GOLD, rel_err

fuse = int(sys.argv[1]) if len(sys.argv) > 1 else -1
g = np.load(os.path.join(GOLD, "g1_forward_v4.npz"))
inp, ts = torch.tensor(g["inputs"]), torch.tensor(g["timesteps"])
cfg = O.OracleCfg()
sd = O.init_weights(cfg, seed=0)
m = mv.MultiViewUNet(mv.default_cfg(), 11, 4, use_cuda_graph=False, fuse_max_tokens=fuse)
m.load_state_dict(sd)
m = m.cuda().eval()
m.enable_taps(True)
y = m(inp.cuda(), ts.cuda()).cpu()
taps = {}
with torch.no_grad():
    ref = O.unet_forward(sd, inp, ts, cfg, taps)
for k, v in taps.items():
    t = m.tap(k).cpu()
    got = t.reshape(v.shape) if v.dim() == 4 else t.reshape(v.shape[0], v.shape[2], v.shape[1]).permute(0, 2, 1)
    print(f"{k:20s} err {rel_err(got, v):.3e} nan {int(torch.isnan(got).sum())} |got| {got.abs().mean():.3e} |ref| {v.abs().mean():.3e}")
print("output err", rel_err(y, ref), "nan", int(torch.isnan(y).sum()))
