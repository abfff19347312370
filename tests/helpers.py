"""Shared helpers for the parity tests: build C-ABI descriptors from torch tensors, torch-fp32 references
for single ops, error metrics."""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mvldm_b200 import _lib  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|"""
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rms_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """rms(a-b) / rms(b)"""
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30)).item()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def nhwc_bf16(x_nchw: torch.Tensor) -> torch.Tensor:
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def conv_seg(t_nhwc: torch.Tensor, stride: int = 1, taps: int = 9, c: int | None = None) -> _lib.ASeg:
    n, h, w, ctot = t_nhwc.shape
    s = _lib.ASeg()
    s.ptr = t_nhwc.data_ptr()
    s.c = ctot if c is None else c
    s.ctot = ctot
    s.sh, s.sw, s.stride, s.ntaps = h, w, stride, taps
    for t in range(taps):
        s.dh[t] = (t // 3 - 1) if taps == 9 else 0
        s.dw[t] = (t % 3 - 1) if taps == 9 else 0
        s.coff[t] = 0
    return s


def pack_conv_weight(w: torch.Tensor) -> torch.Tensor:
    """[cout, cin, kh, kw] fp32 -> bf16 [cout, taps*cin] tap-major (the library's K order)"""
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous().to(torch.bfloat16)


def run_gemm(impl, segs, n_img, oh, ow, w_packed, bias=None, rowvec=None, residual=None, mode=0, n_valid=None,
             out=None):
    d = _lib.GemmDesc()
    d.nseg = len(segs)
    for i, s in enumerate(segs):
        d.seg[i] = s
    d.n_img, d.oh, d.ow = n_img, oh, ow
    d.w = w_packed.data_ptr()
    d.n, d.k = w_packed.shape
    d.bias = bias.data_ptr() if bias is not None else None
    if rowvec is not None:
        d.rowvec, d.rowvec_ld = rowvec.data_ptr(), rowvec.stride(0)
    if residual is not None:
        d.residual, d.res_ld = residual.data_ptr(), residual.shape[-1]
    d.mode = mode
    M = n_img * oh * ow
    if out is None:
        if mode == 0:
            out = torch.empty((M, d.n), device="cuda", dtype=torch.bfloat16)
        elif mode == 1:
            out = torch.empty((M, d.n // 2), device="cuda", dtype=torch.bfloat16)
        else:
            out = torch.empty((n_img, n_valid, oh, ow), device="cuda", dtype=torch.float32)
    d.out = out.data_ptr()
    d.ldo = out.shape[-1] if mode != 2 else 0
    d.n_valid = n_valid if n_valid is not None else d.n
    _lib.check(_lib.load().mvldm_op_gemm(stream_ptr(), impl, ctypes.byref(d)))
    return out


def geglu_interleave(w: torch.Tensor) -> torch.Tensor:
    """rows [x(0..4C) | gate(0..4C)] -> 8-row interleave used by the GEGLU epilogue"""
    c4 = w.shape[0] // 2
    idx = torch.empty(2 * c4, dtype=torch.long)
    ch = torch.arange(c4)
    idx[(ch // 8) * 16 + ch % 8] = ch
    idx[(ch // 8) * 16 + 8 + ch % 8] = c4 + ch
    return w[idx]


def pack_qkv(q, k, v, heads, dpad):
    """q,k,v fp32 [B, N, C] -> packed head-padded bf16 [B*N, 3*heads*dpad]"""
    B, N, C = q.shape
    d = C // heads
    out = torch.zeros(B * N, 3 * heads * dpad)
    for t, x in enumerate((q, k, v)):
        xs = x.reshape(B * N, heads, d)
        out.view(B * N, 3, heads, dpad)[:, t, :, :d] = xs
    if d < dpad:
        out.view(B * N, 3, heads, dpad)[:, 2, :, d] = 1.0     # V's ones column: row sums come out of P.V (include/mvldm_b200.h)
    return out.to(torch.bfloat16)


def attention_ref(q, k, v, heads):
    """fp32 reference on (bf16-rounded) q,k,v [B,N,C]"""
    B, N, C = q.shape
    d = C // heads
    sp = lambda t: t.reshape(B, N, heads, d).permute(0, 2, 1, 3)  # noqa: E731
    o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v))
    return o.permute(0, 2, 1, 3).reshape(B, N, C)
