"""GPU parity of every kernel behind the hot path, called through the C ABI, against torch fp32 on the
same (bf16-rounded) operands.  Tolerance: outputs are stored in bf16 (half-ulp 2^-9 = 2e-3 relative), the
accumulation is fp32 -> max|err| / max|ref| <= 1e-2 per op; fp32-in/fp32-out elementwise kernels <= 1e-5."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import mvldm_b200 as mv
from helpers import (GOLD, attention_ref, conv_seg, geglu_interleave, nhwc_bf16, pack_conv_weight, pack_qkv, rel_err,
                     run_gemm, stream_ptr)
from mvldm_b200 import _lib
from oracle import mvldm_oracle as O

pytestmark = pytest.mark.gpu
BF16_TOL = 1e-2
IMPLS = [pytest.param(0, id="tcgen05"), pytest.param(1, id="simt")]


def _conv_ref(xb, wp, b, cout, cin, stride):
    w = wp.float().cpu().reshape(cout, 3, 3, cin).permute(0, 3, 1, 2)
    return F.conv2d(xb.float().permute(0, 3, 1, 2).cpu(), w, b, stride=stride, padding=1)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("n,cin,cout,hw,stride", [
    (2, 64, 64, 32, 1), (8, 320, 320, 32, 1), (4, 128, 64, 16, 1), (4, 128, 128, 8, 1), (4, 64, 128, 4, 1),
    (3, 64, 64, 8, 1),      # ragged: 192 pixels, second tile half empty
    (8, 64, 96, 4, 1),      # N = 96 -> 32-wide tiles
    (1, 64, 64, 4, 1),      # a single 16-pixel image: tile mostly out of range
    (4, 64, 64, 32, 2), (4, 128, 128, 16, 2), (8, 64, 64, 8, 2),
    (8, 1280, 256, 4, 1),   # weight-streaming shape: few tiles, K = 11520 -> split-K + fixed-order reduce
    (8, 640, 128, 8, 1),    # 4 x 1 tiles, K = 5760 -> split-K
])
def test_conv3x3(impl, n, cin, cout, hw, stride):
    torch.manual_seed(n * 1000 + cin + cout + hw + stride)
    x = torch.randn(n, cin, hw, hw)
    w = torch.randn(cout, cin, 3, 3) / (3 * cin ** 0.5)
    b = torch.randn(cout)
    xb, wp = nhwc_bf16(x).cuda(), pack_conv_weight(w).cuda()
    oh = hw // stride
    out = run_gemm(impl, [conv_seg(xb, stride)], n, oh, oh, wp, bias=b.cuda())
    got = out.float().cpu().reshape(n, oh, oh, cout).permute(0, 3, 1, 2)
    assert rel_err(got, _conv_ref(xb, wp, b, cout, cin, stride)) < BF16_TOL


@pytest.mark.parametrize("impl", IMPLS)
def test_linear_epilogues(impl):
    torch.manual_seed(1)
    M, K, N = 1024, 320, 640
    a = torch.randn(M, K).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N).cuda()
    res = torch.randn(M, N).to(torch.bfloat16).cuda()
    a4 = a.view(4, 16, 16, K)
    base = a.float() @ w.float().t() + b
    out = run_gemm(impl, [conv_seg(a4, 1, 1)], 4, 16, 16, w, bias=b, residual=res)
    assert rel_err(out.float(), base + res.float()) < BF16_TOL
    rv = torch.randn(4, 1000).cuda()                       # per-image row vector = time_emb_proj(silu(emb))
    out = run_gemm(impl, [conv_seg(a4, 1, 1)], 4, 16, 16, w, bias=b, rowvec=rv[:, 100:])
    assert rel_err(out.float(), base + rv[:, 100:100 + N].repeat_interleave(256, 0)) < BF16_TOL
    # GEGLU: x * gelu(gate), exact erf gelu (mvdream/attention.py:60-67)
    w2, b2 = torch.randn(2 * N, K) / K ** 0.5, torch.randn(2 * N)
    wi = geglu_interleave(w2).to(torch.bfloat16).cuda()
    bi = geglu_interleave(b2[:, None])[:, 0].contiguous().cuda()
    out = run_gemm(impl, [conv_seg(a4, 1, 1)], 4, 16, 16, wi, bias=bi, mode=1)
    z = a.float().cpu() @ w2.to(torch.bfloat16).float().t() + b2
    assert rel_err(out.float(), z[:, :N] * F.gelu(z[:, N:])) < BF16_TOL


@pytest.mark.parametrize("impl", IMPLS)
def test_conv_plus_shortcut_segments_and_head(impl):
    """conv2(3x3) + 1x1 shortcut over a channel concat accumulate in ONE GEMM (3 K-segments); NCHW fp32 head."""
    torch.manual_seed(2)
    n, hw, c1, c2, c3, co = 4, 8, 128, 64, 64, 128
    x1, x2, x3 = (torch.randn(n, c, hw, hw) for c in (c1, c2, c3))
    wc = torch.randn(co, c1, 3, 3) / (3 * c1 ** 0.5)
    ws = torch.randn(co, c2 + c3, 1, 1) / (c2 + c3) ** 0.5
    X1, X2, X3 = (nhwc_bf16(t).cuda() for t in (x1, x2, x3))
    wp = torch.cat([pack_conv_weight(wc), pack_conv_weight(ws)], 1).contiguous().cuda()
    out = run_gemm(impl, [conv_seg(X1), conv_seg(X2, 1, 1), conv_seg(X3, 1, 1)], n, hw, hw, wp)
    tof = lambda T: T.float().cpu().permute(0, 3, 1, 2)  # noqa: E731
    ref = F.conv2d(tof(X1), wc.to(torch.bfloat16).float(), padding=1) + \
        F.conv2d(torch.cat([tof(X2), tof(X3)], 1), ws.to(torch.bfloat16).float())
    assert rel_err(out.float().cpu().reshape(n, hw, hw, co).permute(0, 3, 1, 2), ref) < BF16_TOL
    wh = torch.zeros(32, 9 * c1)
    wh[:4] = torch.randn(4, 9 * c1) / (3 * c1 ** 0.5)
    bh = torch.zeros(32)
    bh[:4] = torch.randn(4)
    out = run_gemm(impl, [conv_seg(X1)], n, hw, hw, wh.to(torch.bfloat16).cuda(), bias=bh.cuda(), mode=2, n_valid=4)
    ref = F.conv2d(tof(X1), wh[:4].to(torch.bfloat16).float().reshape(4, 3, 3, c1).permute(0, 3, 1, 2), bh[:4], padding=1)
    assert rel_err(out, ref) < 1e-5           # fp32 out: only summation order differs


@pytest.mark.parametrize("impl", IMPLS)
def test_time_embedding_shaped_gemms(impl):
    """the time-embedding chain: M = 8 rows of 1x1 'images' (128 images per tile: bias straight from memory, no table),
    SiLU epilogue (mode 3) and fp32 row-major output (mode 4)"""
    torch.manual_seed(8)
    n, K, N = 8, 320, 1280
    a = torch.randn(n, K).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N).cuda()
    ref = a.float() @ w.float().t() + b
    out = torch.empty((n, N), device="cuda", dtype=torch.bfloat16)
    run_gemm(impl, [conv_seg(a.view(n, 1, 1, K), 1, 1)], n, 1, 1, w, bias=b, mode=3, out=out)
    assert rel_err(out.float(), F.silu(ref)) < BF16_TOL
    out32 = torch.empty((n, N), device="cuda", dtype=torch.float32)
    run_gemm(impl, [conv_seg(a.view(n, 1, 1, K), 1, 1)], n, 1, 1, w, bias=b, mode=4, out=out32)
    assert rel_err(out32, ref) < 1e-5 if impl != 1 else rel_err(out32, ref) < 1e-4


@pytest.mark.parametrize("impl", [pytest.param(0, id="tcgen05"), pytest.param(1, id="simt")])
def test_gelu_epilogue_gemm(impl):
    """mode 5: bias + exact (erf) GELU, the first Linear of the StandardTransformer FeedForward
    (reference src/model/transformer/feed_forward.py:31-33)"""
    torch.manual_seed(9)
    n, hw, K, N = 2, 16, 640, 640
    a = torch.randn(n, hw, hw, K).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16).cuda()
    b = torch.randn(N).cuda()
    ref = F.gelu(a.float().reshape(-1, K) @ w.float().t() + b)
    out = torch.empty((n * hw * hw, N), device="cuda", dtype=torch.bfloat16)
    run_gemm(impl, [conv_seg(a, 1, 1)], n, hw, hw, w, bias=b, mode=5, out=out)
    assert rel_err(out.float(), ref) < BF16_TOL


def test_splitk_epilogue_and_segments():
    """split-K path with bias + per-image row vector + residual, and with the 3-segment conv+shortcut operand"""
    torch.manual_seed(6)
    n, hw, cin, co = 8, 4, 1280, 256
    x = nhwc_bf16(torch.randn(n, cin, hw, hw)).cuda()
    sk = nhwc_bf16(torch.randn(n, 640, hw, hw)).cuda()
    wc = torch.randn(co, cin, 3, 3) / (3 * cin ** 0.5)
    ws = torch.randn(co, 640, 1, 1) / 640 ** 0.5
    b, rv = torch.randn(co).cuda(), torch.randn(n, co).cuda()
    res = torch.randn(n * hw * hw, co).to(torch.bfloat16).cuda()
    tof = lambda T: T.float().cpu().permute(0, 3, 1, 2)  # noqa: E731
    nhwc = lambda T: T.permute(0, 2, 3, 1).reshape(-1, co)  # noqa: E731
    base = F.conv2d(tof(x), wc.to(torch.bfloat16).float(), padding=1)
    for impl in (0, 1):
        out = run_gemm(impl, [conv_seg(x)], n, hw, hw, pack_conv_weight(wc).cuda(), bias=b, rowvec=rv, residual=res)
        ref = nhwc(base + b.cpu()[None, :, None, None] + rv.cpu()[:, :, None, None]) + res.float().cpu()
        assert rel_err(out.float(), ref) < BF16_TOL
        wp = torch.cat([pack_conv_weight(wc), pack_conv_weight(ws)], 1).contiguous().cuda()
        out = run_gemm(impl, [conv_seg(x), conv_seg(sk, 1, 1)], n, hw, hw, wp)
        ref = nhwc(base + F.conv2d(tof(sk), ws.to(torch.bfloat16).float()))
        assert rel_err(out.float(), ref) < BF16_TOL
    a = run_gemm(0, [conv_seg(x)], n, hw, hw, pack_conv_weight(wc).cuda())
    assert torch.equal(a, run_gemm(0, [conv_seg(x)], n, hw, hw, pack_conv_weight(wc).cuda()))   # bit-stable


def test_tcgen05_matches_simt_bitwise_close():
    """Same operands through both kernel families: differences are fp32 summation order only."""
    torch.manual_seed(3)
    x = nhwc_bf16(torch.randn(8, 320, 16, 16)).cuda()
    wp = pack_conv_weight(torch.randn(640, 320, 3, 3) / (3 * 320 ** 0.5)).cuda()
    a = run_gemm(0, [conv_seg(x)], 8, 16, 16, wp).float()
    b = run_gemm(1, [conv_seg(x)], 8, 16, 16, wp).float()
    assert rel_err(a, b) < 4e-3               # at most one bf16 ulp apart
    assert torch.equal(a, run_gemm(0, [conv_seg(x)], 8, 16, 16, wp).float())   # bit-stable rerun


@pytest.mark.parametrize("n,cin,cout,h,w,stride", [
    (2, 64, 64, 24, 24, 1),      # 24-wide rows: tile = 24 x 5 rows = 120 of the 128 MMA rows, ragged last tile (24 = 4*5 + 4)
    (3, 128, 96, 12, 12, 1),     # 12 x 10 rows
    (5, 64, 64, 6, 6, 1),        # 6 x 6 x 2 images = 72 rows, odd image count
    (7, 64, 32, 3, 3, 1),        # 9-pixel images: 3 x 1 x 8 images = 24 rows per tile
    (2, 64, 64, 5, 5, 1),        # 5 x 5
    (2, 64, 64, 48, 24, 2),      # stride 2 onto 24 x 12
    (1, 64, 64, 8, 256, 1),      # 256-wide rows split into two 128-pixel column blocks (VAE resolution)
    (1, 128, 64, 6, 192, 1),     # 192 = 2 x 96
    (2, 1280, 128, 12, 12, 1),   # K = 11520: split-K with the general geometry (fused fixed-order reduce)
    (8, 640, 256, 6, 6, 1),
])
def test_conv3x3_general_tile_geometry(n, cin, cout, h, w, stride):
    """output sizes whose width does not divide 128 (24x24 / 48x48 / 40x40 latents and their coarser levels) or exceeds it
    (256-wide feature maps): the M tile is a (columns x rows x images) TMA box covering <= 128 rows of the MMA"""
    torch.manual_seed(n + cin + cout + h + w)
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, 3, 3) / (3 * cin ** 0.5)
    b = torch.randn(cout)
    xb, wp = nhwc_bf16(x).cuda(), pack_conv_weight(wt).cuda()
    oh, ow = h // stride, w // stride
    res = torch.randn(n * oh * ow, cout).to(torch.bfloat16).cuda()
    rv = torch.randn(n, cout).cuda()
    ref = _conv_ref(xb, wp, b, cout, cin, stride)
    ref = ref + rv.cpu()[:, :, None, None] + res.float().cpu().reshape(n, oh, ow, cout).permute(0, 3, 1, 2)
    outs = []
    for impl in (0, 1):
        out = run_gemm(impl, [conv_seg(xb, stride)], n, oh, ow, wp, bias=b.cuda(), rowvec=rv, residual=res)
        got = out.float().cpu().reshape(n, oh, ow, cout).permute(0, 3, 1, 2)
        assert rel_err(got, ref) < BF16_TOL, impl
        outs.append(out)
    assert torch.equal(outs[0], run_gemm(0, [conv_seg(xb, stride)], n, oh, ow, wp, bias=b.cuda(), rowvec=rv, residual=res))
    if stride == 1 and cout == 64:     # fp32 NCHW head (mode 2) on the same geometry
        wh = torch.zeros(32, 9 * cin)
        wh[:4] = torch.randn(4, 9 * cin) / (3 * cin ** 0.5)
        o2 = run_gemm(0, [conv_seg(xb)], n, oh, ow, wh.to(torch.bfloat16).cuda(), mode=2, n_valid=4)
        r2 = F.conv2d(xb.float().permute(0, 3, 1, 2).cpu(), wh[:4].to(torch.bfloat16).float().reshape(4, 3, 3, cin).permute(0, 3, 1, 2),
                      padding=1)
        assert rel_err(o2, r2) < 1e-4


def test_gemm_rejects_unsupported_shapes():
    x = nhwc_bf16(torch.randn(1, 64, 1, 129)).cuda()          # 129-wide rows: no equal column blocks of <= 128 pixels ...
    wp = pack_conv_weight(torch.randn(64, 64, 3, 3)).cuda()
    with pytest.raises(RuntimeError):                         # ... with a row count that is a multiple of 8
        run_gemm(0, [conv_seg(x)], 1, 1, 129, wp)
    x = nhwc_bf16(torch.randn(1, 48, 8, 8)).cuda()            # channels not a multiple of 64
    wp = pack_conv_weight(torch.randn(64, 48, 3, 3)).cuda()
    with pytest.raises(RuntimeError):
        run_gemm(0, [conv_seg(x)], 1, 8, 8, wp)


@pytest.mark.parametrize("n,c0,c1,hw,silu,eps", [(4, 320, 0, 1024, 1, 1e-5), (4, 640, 320, 256, 1, 1e-5),
                                                 (8, 1280, 640, 16, 1, 1e-5), (2, 1280, 0, 64, 0, 1e-6),
                                                 (8, 1280, 0, 16, 1, 1e-5), (8, 640, 0, 256, 1, 1e-5),
                                                 (8, 1280, 1280, 16, 1, 1e-5), (3, 640, 0, 1024, 0, 1e-6),
                                                 (8, 320, 320, 1024, 1, 1e-5), (2, 1280, 0, 1024, 1, 1e-5),
                                                 (3, 1280, 1280, 64, 1, 1e-5), (5, 320, 0, 256, 1, 1e-5),
                                                 (1, 640, 0, 64, 1, 1e-5), (7, 1280, 0, 16, 0, 1e-5),
                                                 # the VAE's tensors: 4 / 8 / 16-wide groups, up to 256 x 256 pixels
                                                 (2, 128, 0, 65536, 1, 1e-6), (2, 256, 0, 16384, 1, 1e-6),
                                                 (3, 512, 0, 4096, 1, 1e-6), (2, 512, 0, 1024, 0, 1e-6), (1, 64, 0, 1024, 1, 1e-6),
                                                 # widths that do not divide 128: 24 x 24, 12 x 12, 6 x 6, 3 x 3 pixels
                                                 (3, 320, 0, 576, 1, 1e-5), (3, 640, 320, 144, 1, 1e-5), (2, 1280, 0, 36, 1, 1e-5),
                                                 (2, 1280, 1280, 9, 1, 1e-5)])
def test_groupnorm(n, c0, c1, hw, silu, eps):
    torch.manual_seed(4)
    lib = _lib.load()
    fn = lib.mvldm_op_groupnorm
    x0 = torch.randn(n, hw, c0).mul(2).add(0.5).to(torch.bfloat16).cuda()
    x1 = torch.randn(n, hw, c1).to(torch.bfloat16).cuda() if c1 else None
    C = c0 + c1
    g, b = torch.randn(C).cuda(), torch.randn(C).cuda()
    out = torch.empty(n, hw, C, dtype=torch.bfloat16, device="cuda")
    scratch = torch.empty(n * 32 * 2 * 64, device="cuda")
    _lib.check(fn(stream_ptr(), x0.data_ptr(), c0, x1.data_ptr() if c1 else None, c1, n, hw, 32,
                  eps, g.data_ptr(), b.data_ptr(), silu, out.data_ptr(), scratch.data_ptr()))
    xx = torch.cat([x0, x1], -1) if c1 else x0            # groups straddle the concat boundary for 960 / 1920
    ref = F.group_norm(xx.float().permute(0, 2, 1), 32, g, b, eps)
    ref = (F.silu(ref) if silu else ref).permute(0, 2, 1)
    assert rel_err(out.float(), ref) < BF16_TOL


def test_groupnorm_large_mean_keeps_precision():
    """|mean| >> std (real checkpoints, eps 1e-6 in the transformer norms): GroupNorm takes the variance in a centred second
    pass over the register-resident block, so it must stay accurate where E[x^2] - mean^2 in fp32 loses every digit"""
    torch.manual_seed(7)
    lib = _lib.load()
    n, c, hw = 4, 320, 256
    x = (torch.randn(n, hw, c) * 0.05 + 8.0).to(torch.bfloat16).cuda()       # mean / std of the bf16 values ~ 100
    g, b = torch.randn(c).cuda(), torch.randn(c).cuda()
    out = torch.empty(n, hw, c, dtype=torch.bfloat16, device="cuda")
    scratch = torch.empty(n * 32 * 2 * 64, device="cuda")
    _lib.check(lib.mvldm_op_groupnorm(stream_ptr(), x.data_ptr(), c, None, 0, n, hw, 32, 1e-6, g.data_ptr(),
                                      b.data_ptr(), 0, out.data_ptr(), scratch.data_ptr()))
    ref = F.group_norm(x.double().permute(0, 2, 1), 32, g.double(), b.double(), 1e-6).permute(0, 2, 1).float()
    assert rel_err(out.float(), ref) < BF16_TOL


@pytest.mark.parametrize("c", [320, 640, 1280])
def test_layernorm(c):
    torch.manual_seed(5)
    x = torch.randn(1000, c).mul(3).to(torch.bfloat16).cuda()
    g, b = torch.randn(c).cuda(), torch.randn(c).cuda()
    out = torch.empty_like(x)
    fn = _lib.load().mvldm_op_layernorm
    _lib.check(fn(stream_ptr(), x.data_ptr(), 1000, c, 1e-5, g.data_ptr(), b.data_ptr(), out.data_ptr()))
    assert rel_err(out.float(), F.layer_norm(x.float(), (c,), g, b, 1e-5)) < BF16_TOL


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("B,N,d", [(1, 2048, 40), (2, 256, 80), (3, 64, 160), (4, 16, 160), (2, 1024, 40),
                                   (1, 4096, 40), (1, 200, 40)])
def test_attention(impl, B, N, d):
    """joint (long) and per-view (short, down to 16 tokens) sequences, head dims 40/80/160, ragged N"""
    torch.manual_seed(B * 7 + N + d)
    heads, C, dpad = 8, 8 * d, (d + 63) // 64 * 64
    q, k, v = (torch.randn(B, N, C) for _ in range(3))
    qkv = pack_qkv(q, k, v, heads, dpad).cuda()
    out = torch.full((B * N, heads * dpad), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(_lib.load().mvldm_op_attention(stream_ptr(), impl, qkv.data_ptr(), out.data_ptr(), B, N, heads, d, dpad))
    r = lambda t: t.to(torch.bfloat16).float()  # noqa: E731
    ref = attention_ref(r(q), r(k), r(v), heads)
    got = out.float().cpu().view(B, N, heads, dpad)
    assert bool((got[..., d:] == 0).all()), "pad columns must be written as zeros"
    assert rel_err(got[..., :d].reshape(B, N, C), ref) < BF16_TOL


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("B,N,heads", [(4, 1024, 5), (2, 256, 10), (3, 64, 20), (1, 200, 5)])
def test_attention_unpadded_64_wide_heads(impl, B, N, heads):
    """Variant B's Transformer2DModel heads: d == dpad == 64, no pad column for the tensor-core row sum, heads != 8"""
    torch.manual_seed(B + N + heads)
    d = dpad = 64
    C = heads * d
    q, k, v = (torch.randn(B, N, C) for _ in range(3))
    qkv = pack_qkv(q, k, v, heads, dpad).cuda()
    out = torch.full((B * N, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(_lib.load().mvldm_op_attention(stream_ptr(), impl, qkv.data_ptr(), out.data_ptr(), B, N, heads, d, dpad))
    r = lambda t: t.to(torch.bfloat16).float()  # noqa: E731
    assert rel_err(out.float().cpu().view(B, N, C), attention_ref(r(q), r(k), r(v), heads)) < BF16_TOL


def test_attention_rerun_bit_stable_with_wide_score_range():
    """large score range -> the lazily moved softmax reference is rescaled often; reruns must still be bit-identical
    (this caught a barrier-phase race in the double-buffered kernel)"""
    torch.manual_seed(11)
    heads, d, dpad = 8, 40, 64
    for (B, N) in [(1, 8192), (2, 1000)]:
        q, k, v = (torch.randn(B, N, heads * d) * 3 for _ in range(3))
        qkv = pack_qkv(q, k, v, heads, dpad).cuda()
        outs = []
        for _ in range(4):
            out = torch.empty((B * N, heads * dpad), dtype=torch.bfloat16, device="cuda")
            _lib.check(_lib.load().mvldm_op_attention(stream_ptr(), 0, qkv.data_ptr(), out.data_ptr(), B, N, heads, d, dpad))
            outs.append(out)
        torch.cuda.synchronize()
        assert all(torch.equal(outs[0], o) for o in outs[1:])
        r = lambda t: t.to(torch.bfloat16).float()  # noqa: E731
        ref = attention_ref(r(q), r(k), r(v), heads)
        got = outs[0].float().cpu().view(B, N, heads, dpad)[..., :d].reshape(B, N, heads * d)
        assert rel_err(got, ref) < BF16_TOL


def test_attention_softmax_is_shift_invariant_and_uniform_for_equal_keys():
    """property checks independent of the oracle: identical keys -> output = mean of V"""
    heads, d, dpad, N = 8, 40, 64, 512
    q = torch.randn(1, N, heads * d)
    k = torch.randn(1, 1, heads * d).expand(1, N, heads * d).contiguous()
    v = torch.randn(1, N, heads * d)
    qkv = pack_qkv(q, k, v, heads, dpad).cuda()
    for impl in (0, 1):
        out = torch.empty((N, heads * dpad), dtype=torch.bfloat16, device="cuda")
        _lib.check(_lib.load().mvldm_op_attention(stream_ptr(), impl, qkv.data_ptr(), out.data_ptr(), 1, N, heads, d, dpad))
        got = out.float().cpu().view(N, heads, dpad)[..., :d].reshape(N, heads * d)
        ref = v.to(torch.bfloat16).float().mean(1).expand(N, -1)
        assert (got - ref).abs().max() < 2e-2


def test_ddim_step_and_cfg():
    g = np.load(os.path.join(GOLD, "g4_ddim.npz"))
    s = mv.DDIMScheduler(clip_sample=False)
    s.set_timesteps(25)
    x, e = torch.tensor(g["step_x"]).cuda(), torch.tensor(g["step_eps"]).cuda()
    for t in (960, 0):
        assert rel_err(s.step(e, t, x).prev_sample, torch.tensor(g[f"step_out_{t}"])) < 1e-5
    # CFG compose fused with the update, context views skipped
    B, v_c, v_t = 2, 2, 3
    ec, eu, xt = torch.randn(B, v_c + v_t, 4, 32, 32), torch.randn(B, v_t, 4, 32, 32), torch.randn(B, v_t, 4, 32, 32)
    o = O.DDIMOracle()
    o.set_timesteps(25)
    ref = o.step(eu + 3.0 * (ec[:, v_c:] - eu), 480, xt)
    got = mv.fused_cfg_ddim_step(s, ec.cuda(), eu.cuda(), 3.0, v_c, 480, xt.cuda())
    assert rel_err(got, ref) < 1e-5
    # linearity of the update in (x, eps)
    got2 = mv.fused_cfg_ddim_step(s, 2 * ec.cuda(), 2 * eu.cuda(), 3.0, v_c, 480, 2 * xt.cuda())
    assert rel_err(got2, 2 * ref) < 1e-5


def test_ddpm_step_and_cfg():
    """the "ddpm" entry of the reference's scheduler registry (src/model/scheduler/__init__.py:19-22, ddpm.yaml: clip_sample
    true, fixed_small variance): fused CFG compose + ancestral step against the restated diffusers DDPMScheduler.step, with the
    variance noise drawn from a seeded CUDA generator on both sides"""
    torch.manual_seed(12)
    B, v_c, v_t = 2, 2, 3
    ec, eu = torch.randn(B, v_c + v_t, 4, 32, 32), torch.randn(B, v_t, 4, 32, 32)
    xt = torch.randn(B, v_t, 4, 32, 32) * 2.0                     # wide enough that the x0 clamp bites
    for n_steps, t, vt in ((50, 480, "fixed_small"), (1000, 999, "fixed_small"), (50, 0, "fixed_small"), (25, 440, "fixed_large")):
        s = mv.DDPMScheduler(variance_type=vt)
        s.set_timesteps(n_steps)
        o = O.DDPMOracle(variance_type=vt)
        o.set_timesteps(n_steps)
        assert torch.equal(s.timesteps, o.timesteps)
        g1 = torch.Generator(device="cuda").manual_seed(5)
        got = mv.fused_cfg_ddpm_step(s, ec.cuda(), eu.cuda(), 3.0, v_c, t, xt.cuda(), g1)
        g2 = torch.Generator(device="cuda").manual_seed(5)
        z = torch.randn(xt.shape, generator=g2, device="cuda").cpu() if t > 0 else None
        ref = o.step(eu + 3.0 * (ec[:, v_c:] - eu), t, xt, z)
        assert rel_err(got, ref) < 1e-5, (n_steps, t, vt)
        # scheduler.step surface (no CFG): `.prev_sample`
        g3 = torch.Generator(device="cuda").manual_seed(5)
        got1 = s.step(eu.cuda(), t, xt.cuda(), generator=g3).prev_sample
        assert rel_err(got1, o.step(eu, t, xt, z)) < 1e-5
    # no clipping: linear in (x, eps, z)
    s = mv.DDPMScheduler(clip_sample=False)
    s.set_timesteps(50)
    o = O.DDPMOracle(clip_sample=False)
    o.set_timesteps(50)
    g1 = torch.Generator(device="cuda").manual_seed(6)
    got = s.step(eu.cuda(), 480, xt.cuda(), generator=g1).prev_sample
    g2 = torch.Generator(device="cuda").manual_seed(6)
    z = torch.randn(xt.shape, generator=g2, device="cuda").cpu()
    assert rel_err(got, o.step(eu, 480, xt, z)) < 1e-5
    with pytest.raises(NotImplementedError):
        mv.DDPMScheduler(thresholding=True)
    with pytest.raises(RuntimeError):
        s.step(eu, 480, xt)                                         # CPU tensors: no fallback


def test_raymap_and_build_inputs():
    g = np.load(os.path.join(GOLD, "g5_rays.npz"))
    extr, intr = torch.tensor(g["extr"]).cuda(), torch.tensor(g["intr"]).cuda()
    for pl in (0, 1):
        r = mv.ray_encode(extr, intr, 32, 32, bool(pl))
        assert (r.cpu() - torch.tensor(g[f"rays_plucker{pl}"])).abs().max() < 1e-5
    ctx, xT, extr, intr = O.synthetic_scene(2, 2, 3)
    rays = O.raymap(extr, intr, 32, 32)
    ref, tgt = O.build_inputs(xT, torch.cat([ctx, torch.zeros(2, 2, 1, 32, 32)], 2), rays, torch.ones(2, 3, 1, 32, 32))
    assert torch.equal(mv.build_inputs(xT.cuda(), ctx.cuda(), rays.cuda()).cpu(), ref)          # byte-exact concat
    assert torch.equal(mv.build_inputs(xT.cuda(), None, rays.cuda(), 2).cpu(), torch.cat([tgt, rays[:, 2:]], 2))


def test_positionally_encoded_ray_maps_match_reference():
    """use_ray_encoding: true (config/main.yaml:28-33): golden g8 comes from the reference's own projection.py and
    PositionalEncoding; sin() of arguments up to 2 pi 2^9 in fp32 - the tolerance is fp32 argument rounding, not bf16"""
    g = np.load(os.path.join(GOLD, "g8_ray_encoding.npz"))
    extr, intr = torch.tensor(g["extr"]).cuda(), torch.tensor(g["intr"]).cuda()
    for fo, fd in ((10, 8), (4, 0)):
        ref = torch.tensor(g[f"rays_{fo}_{fd}"])
        got = mv.ray_encode(extr, intr, ref.shape[-2], ref.shape[-1], False, fo, fd).cpu()
        assert got.shape == ref.shape
        assert (got - ref).abs().max() < 2e-3, (fo, fd, (got - ref).abs().max())
        assert (got[:, :, :6 * min(fo, 4)] - ref[:, :, :6 * min(fo, 4)]).abs().max() < 2e-5   # low octaves: exact to fp32 noise
    ref = torch.tensor(g["rays_srt_6_5"])                           # srt_ray_encoding: true, the reference's own RayEncoder
    got = mv.ray_encode(extr, intr, ref.shape[-2], ref.shape[-1], False, 6, 5, srt_ray_encoding=True).cpu()
    assert got.shape == ref.shape and (got - ref).abs().max() < 2e-4
    # a denoiser with the matching in_channels (4 latent + 1 mask + 108 ray channels) takes these inputs
    rays = mv.ray_encode(extr, intr, 16, 24, False, 10, 8)
    m = mv.MultiViewUNet(mv.default_cfg(), 4 + 1 + 108, 4).cuda().eval()
    x = mv.build_inputs(torch.randn(1, 2, 4, 16, 24, device="cuda"), torch.randn(1, 2, 4, 16, 24, device="cuda"), rays)
    assert x.shape == (1, 4, 113, 16, 24)
    y = m(x, torch.tensor([[0, 0, 500, 500]], device="cuda"))
    assert y.shape == (1, 4, 4, 16, 24) and torch.isfinite(y).all()
