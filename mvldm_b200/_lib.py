"""ctypes binding of ``libmvldm_b200.so`` (the C ABI declared in ``include/mvldm_b200.h``).

The product path has no CPU / PyTorch fallback: if the CUDA extension is missing or fails to load this
module raises, it never substitutes another implementation.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int8, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmvldm_b200.so")

MVLDM_MAX_LEVELS = 4
MVLDM_MAX_SEGS = 3
F32, BF16, F16 = 0, 1, 2
IMPL_TC, IMPL_SIMT, IMPL_TC_GEMM_SIMT_ATTN = 0, 1, 2
MV_SPATIAL_TRANSFORMER_3D, MV_STANDARD = 0, 1
MODEL_DENOISER, MODEL_VAE = 0, 1


class Config(Structure):
    _fields_ = [
        ("in_channels", c_int32), ("out_channels", c_int32), ("num_levels", c_int32),
        ("block_out_channels", c_int32 * MVLDM_MAX_LEVELS), ("layers_per_block", c_int32),
        ("norm_groups", c_int32), ("num_heads", c_int32), ("max_attn_res", c_int32), ("impl", c_int32),
        ("use_cuda_graph", c_int32),
        ("variant", c_int32), ("t2d_heads", c_int32 * MVLDM_MAX_LEVELS), ("cross_attention_dim", c_int32),
        ("mv_block", c_int32), ("mv_num_layers", c_int32), ("mv_d_mlp", c_int32), ("mv_d_mlp_multiplier", c_int32),
        ("model", c_int32), ("latent_channels", c_int32),
    ]


class ASeg(Structure):
    _fields_ = [
        ("ptr", c_void_p), ("c", c_int32), ("ctot", c_int32), ("sh", c_int32), ("sw", c_int32),
        ("stride", c_int32), ("ntaps", c_int32), ("dh", c_int8 * 9), ("dw", c_int8 * 9), ("coff", c_int32 * 9),
    ]


class GemmDesc(Structure):
    _fields_ = [
        ("nseg", c_int32), ("seg", ASeg * MVLDM_MAX_SEGS), ("n_img", c_int32), ("oh", c_int32), ("ow", c_int32),
        ("w", c_void_p), ("n", c_int32), ("k", c_int32), ("bias", c_void_p), ("rowvec", c_void_p),
        ("rowvec_ld", c_int32), ("residual", c_void_p), ("res_ld", c_int32), ("mode", c_int32), ("out", c_void_p),
        ("ldo", c_int32), ("n_valid", c_int32),
    ]


# every symbol include/mvldm_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mvldm_last_error": (c_char_p, []),
    "mvldm_version": (c_int, []),
    "mvldm_create": (c_int, [POINTER(Config), c_int, POINTER(c_void_p)]),
    "mvldm_destroy": (c_int, [c_void_p]),
    "mvldm_num_weights": (c_int, [c_void_p]),
    "mvldm_weight_name": (c_char_p, [c_void_p, c_int]),
    "mvldm_weight_shape": (c_int, [c_void_p, c_int, POINTER(c_int64), POINTER(c_int)]),
    "mvldm_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int, c_int, c_void_p]),
    "mvldm_finalize_weights": (c_int, [c_void_p, c_void_p]),
    "mvldm_workspace_bytes": (c_int64, [c_void_p, c_int, c_int, c_int, c_int]),
    "mvldm_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mvldm_last_launch_count": (c_int, [c_void_p]),
    "mvldm_set_profiling": (c_int, [c_void_p, c_int]),
    "mvldm_profile_json": (c_char_p, [c_void_p]),
    "mvldm_debug_tap": (c_int, [c_void_p, c_void_p, c_char_p, c_void_p, POINTER(c_int64)]),
    "mvldm_enable_taps": (c_int, [c_void_p, c_int]),
    "mvldm_build_inputs": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_void_p]),
    "mvldm_ddim_step": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_void_p, c_float,
                                c_float, c_float, c_float, c_void_p, c_void_p]),
    "mvldm_raymap": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mvldm_op_gemm": (c_int, [c_void_p, c_int, POINTER(GemmDesc)]),
    "mvldm_op_attention": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "mvldm_debug_attn_trace": (c_int, [POINTER(c_int64), c_int]),
    "mvldm_op_groupnorm": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                   c_void_p, c_int, c_void_p, c_void_p]),
    "mvldm_op_layernorm": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
}

EXCHANGE_BEGIN, EXCHANGE_END, EXCHANGE_DONE = 0, 1, 2
KV_EXCHANGE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int)
SYMBOLS["mvldm_forward_sharded"] = (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                            c_void_p, c_void_p, c_void_p, c_int64, KV_EXCHANGE_FN, c_void_p])
SYMBOLS["mvldm_forward_scenes"] = (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, POINTER(c_int32), c_int, c_int,
                                           c_void_p])
SYMBOLS["mvldm_op_attention_kv"] = (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                            c_int, c_int, c_int, c_int, c_int, c_void_p])
SYMBOLS["mvldm_raymap_encoded"] = (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p])
SYMBOLS["mvldm_ddpm_step"] = (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                      c_float, c_float, c_float, c_float, c_float, c_float, c_void_p])
SYMBOLS["mvldm_vae_decode"] = (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p])
SYMBOLS["mvldm_vae_encode"] = (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p])
SYMBOLS["mvldm_op_attention_merge"] = (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), c_int64, c_int, c_int,
                                               c_void_p])

_lib = None


def load() -> ctypes.CDLL:
    """Load the extension (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"mvldm_b200: CUDA extension not built ({LIB_PATH} missing). Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C mvldm_b200/csrc`. There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("mvldm_b200: " + load().mvldm_last_error().decode("utf-8", "replace"))


def current_stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def on_device(device):
    """`with on_device(dev):` == `with torch.cuda.device(dev):` when dev is not the current device, and free otherwise (the
    context manager's cudaGetDevice / cudaSetDevice pair is ~5 us per library call on the host's critical path)."""
    import torch
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return _NO_GUARD if idx == torch.cuda.current_device() else torch.cuda.device(idx)
