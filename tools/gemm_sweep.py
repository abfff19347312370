"""Offline check of the GEMM tile / split-K cost model: for the GEMM shapes of one 1-scene x 8-view forward, time every
(BN, splits) configuration (forced through MVLDM_GEMM_BN / MVLDM_GEMM_SPLITS) as a dependent chain inside a CUDA graph
and compare with the configuration `pick_tiles` chooses.   python tools/gemm_sweep.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import *  # noqa

N_IMG = 8
CHAIN = 10


def make(kind, side, cin, cout, mode=0, res=False, extra=0):
    """kind 'lin': 1x1 over cin; 'conv': 3x3 over cin (+ a 1x1 shortcut segment over `extra` channels)"""
    x = torch.randn(N_IMG, side, side, cin).to(torch.bfloat16).cuda()
    segs = [conv_seg(x, taps=9 if kind == "conv" else 1)]
    keep = [x]
    k = cin * (9 if kind == "conv" else 1)
    if extra:
        x2 = torch.randn(N_IMG, side, side, extra).to(torch.bfloat16).cuda()
        segs.append(conv_seg(x2, taps=1))
        keep.append(x2)
        k += extra
    w = (torch.randn(cout, k) / k ** 0.5).to(torch.bfloat16).cuda()
    b = torch.randn(cout).cuda()
    r = torch.randn(N_IMG * side * side, cout).to(torch.bfloat16).cuda() if res else None
    return dict(segs=segs, side=side, w=w, b=b, r=r, mode=mode, keep=keep, M=N_IMG * side * side, N=cout, K=k)


def time_cfg(g, bn, sp):
    for k_, v in (("MVLDM_GEMM_BN", bn), ("MVLDM_GEMM_SPLITS", sp)):
        if v is None:
            os.environ.pop(k_, None)
        else:
            os.environ[k_] = str(v)
    outs = [None, None]
    try:
        for i in range(2):
            outs[i] = run_gemm(0, g["segs"], N_IMG, g["side"], g["side"], g["w"], bias=g["b"], residual=g["r"], mode=g["mode"])
    except RuntimeError:
        return None
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for i in range(CHAIN):
                run_gemm(0, g["segs"], N_IMG, g["side"], g["side"], g["w"], bias=g["b"], residual=g["r"], mode=g["mode"],
                         out=outs[i & 1])
        gr.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / CHAIN * 1e3)
    return best


SHAPES = []
for side, c in ((32, 320), (16, 640), (8, 1280), (4, 1280)):
    d = c // 8
    hp = 8 * ((d + 63) // 64 * 64)
    SHAPES += [("lin qkv", "lin", side, c, 3 * hp, 0, False, 0, 4), ("lin attn-out", "lin", side, hp, c, 0, True, 0, 4),
               ("lin ff1 geglu", "lin", side, c, 8 * c, 1, False, 0, 2), ("lin ff2", "lin", side, 4 * c, c, 0, True, 0, 2),
               ("lin proj", "lin", side, c, c, 0, True, 0, 4)]
SHAPES += [("conv", "conv", 32, 320, 320, 0, False, 0, 9), ("conv", "conv", 32, 640, 320, 0, False, 0, 2),
           ("conv+sc", "conv", 32, 320, 320, 0, False, 640, 2), ("conv", "conv", 32, 960, 320, 0, False, 0, 1),
           ("conv", "conv", 16, 320, 640, 0, False, 0, 1), ("conv", "conv", 16, 640, 640, 0, False, 0, 8),
           ("conv", "conv", 16, 1280, 640, 0, False, 0, 2), ("conv+sc", "conv", 16, 640, 640, 0, False, 1280, 2),
           ("conv", "conv", 16, 1920, 640, 0, False, 0, 1), ("conv", "conv", 8, 640, 1280, 0, False, 0, 1),
           ("conv", "conv", 8, 1280, 1280, 0, False, 0, 6), ("conv", "conv", 8, 2560, 1280, 0, False, 0, 2),
           ("conv+sc", "conv", 8, 1280, 1280, 0, False, 2560, 2), ("conv", "conv", 4, 1280, 1280, 0, False, 0, 9),
           ("conv", "conv", 4, 2560, 1280, 0, False, 0, 3), ("conv+sc", "conv", 4, 1280, 1280, 0, False, 2560, 3)]

rows, tot_model, tot_best = [], 0.0, 0.0
for (name, kind, side, cin, cout, mode, res, extra, count) in SHAPES:
    g = make(kind, side, cin, cout, mode, res, extra)
    t_model = time_cfg(g, None, None)
    res_ = {}
    for bn in (256, 160, 128, 64, 32):
        if cout % bn:
            continue
        for sp in (1, 2, 3, 4, 6, 8, 12, 16):
            t = time_cfg(g, bn, sp)
            if t is not None:
                res_[(bn, sp)] = t
    (bbn, bsp), tb = min(res_.items(), key=lambda kv: kv[1])
    model_cfg = [k for k, v in res_.items() if abs(v - t_model) < 0.02 * t_model]
    tot_model += t_model * count
    tot_best += tb * count
    print(f"{name:14s} M{g['M']:5d} N{g['N']:5d} K{g['K']:5d} x{count}: model {t_model:6.2f} us (~{model_cfg[:2]})  best {tb:6.2f} us bn={bbn} sp={bsp}"
          f"  gain {t_model - tb:5.2f}", flush=True)
    rows.append(dict(name=name, M=g["M"], N=g["N"], K=g["K"], count=count, model_us=t_model, best_us=tb, best=[bbn, bsp],
                     all={f"{k[0]}x{k[1]}": v for k, v in res_.items()}))
print(f"weighted over the listed shapes: model {tot_model:.0f} us, best {tot_best:.0f} us")
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"))
