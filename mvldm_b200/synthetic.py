"""Seeded synthetic inputs for benchmarks and smoke runs (no dataset, no checkpoint, no oracle).

The recipe is SURVEY.md §8d's: random-init weights with the zero-initialised multi-view ``proj_out`` layers
re-randomised (otherwise the nine multi-view blocks contribute nothing to the output), unit-scale latents, an arc of
pinhole cameras expressed relative to the first view.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch
from torch import Tensor, nn


def randomise_weights(module: nn.Module, seed: int = 0) -> nn.Module:
    """In-place seeded init of a ``MultiViewUNet``: default fan-in uniform everywhere, norm layers (1, 0), and
    non-zero ``proj_out`` in the multi-view blocks (mvdream/attention.py:406-411 zero-initialises them)."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    with torch.no_grad():
        for key, p in sd.items():
            leaf = key.rsplit(".", 1)[1]
            if ".norm" in key or "conv_norm_out" in key:
                p.fill_(1.0 if leaf == "weight" else 0.0)
                continue
            w = sd[key.rsplit(".", 1)[0] + ".weight"]
            bound = 1.0 / math.sqrt(w[0].numel())
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
    if hasattr(module, "mark_dirty"):
        module.mark_dirty()
    return module


def cameras(scenes: int, views: int) -> Tuple[Tensor, Tensor]:
    """cam-to-world extrinsics [S, V, 4, 4] (view 0 = identity; 0.25 baseline and a 0.05 rad yaw per view) and
    normalised intrinsics [S, V, 3, 3] (focal 1.2, principal point at the centre)."""
    intr = torch.tensor([[1.2, 0.0, 0.5], [0.0, 1.2, 0.5], [0.0, 0.0, 1.0]]).expand(scenes, views, 3, 3).contiguous()
    extr = torch.eye(4).repeat(scenes, views, 1, 1)
    for v in range(views):
        a = 0.05 * v
        extr[:, v, 0, 0] = extr[:, v, 2, 2] = math.cos(a)
        extr[:, v, 0, 2] = math.sin(a)
        extr[:, v, 2, 0] = -math.sin(a)
        extr[:, v, 0, 3] = 0.25 * v
    return extr, intr


def scene(scenes: int, v_c: int, v_t: int, h: int = 32, w: int = 32, seed: int = 1):
    """(context latents [S, v_c, 4, h, w], x_T [S, v_t, 4, h, w], extrinsics, intrinsics), all fp32 on the host."""
    g = torch.Generator().manual_seed(seed)
    ctx = torch.randn(scenes, v_c, 4, h, w, generator=g)
    x_T = torch.randn(scenes, v_t, 4, h, w, generator=g)
    extr, intr = cameras(scenes, v_c + v_t)
    return ctx, x_T, extr, intr
