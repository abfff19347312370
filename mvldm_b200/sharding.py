"""Multi-GPU partitioning of the denoising path, one process per GPU.

* Scene-batch sharding: scenes never interact, so each rank takes a contiguous slice of the scene batch and runs
  the unmodified single-GPU path; no data-path collective, only the final latents are gathered.
* View-group sharding: one scene whose views are split in contiguous groups over the ranks.  Everything is per
  view except the joint multi-view attention, where every rank needs all views' K and V: `ViewGroupExchange` is the
  host side of `mvldm_forward_sharded` - an NCCL all-gather (ring over NVLink/NVSwitch) of the packed K|V slab at
  each of the 9 multi-view blocks, in view order so the softmax sums in the same order on every rank."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def scene_slice(num_scenes: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of the scenes owned by `rank`; earlier ranks take the remainder."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(num_scenes, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_scenes(local: torch.Tensor, num_scenes: int, group=None) -> torch.Tensor:
    """all-gather the per-rank results (dim 0 = scenes) back into scene order on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [scene_slice(num_scenes, r, ws) for r in range(ws)]
    max_n = max(b - a for a, b in sizes)
    pad = torch.zeros((max_n, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: b - a] for r, (a, b) in enumerate(sizes)], dim=0)


def view_slice(num_views: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of the views owned by `rank`; view groups must be equal (the K/V all-gather is unpadded)."""
    if num_views % world_size != 0:
        raise ValueError("num_views must be a multiple of world_size for view-group sharding")
    per = num_views // world_size
    return rank * per, (rank + 1) * per


class ViewGroupExchange:
    """Owns the K|V send / receive buffers and the callback the library invokes around every joint attention.

    ``overlap=False`` (default): the all-gather runs on the compute stream inside BEGIN and the callback answers
    MVLDM_EXCHANGE_DONE; the library makes one pass over all slabs in view order.
    ``overlap=True``: two-phase protocol (include/mvldm_b200.h, MVLDM_EXCHANGE_BEGIN / _END): BEGIN starts the all-gather on a
    side stream (ordered after the K|V pack on the compute stream) and returns; the library runs the attention over this
    rank's own keys meanwhile; END makes the compute stream wait for the gathered slabs; the partial softmaxes (own / before /
    after slabs) are merged in a fixed order.  Measured on 2 / 4 / 8 B200 over NVLink 5 the exchange is a few percent of the
    forward and the split costs about what it hides (bench.py's ``view_sharded`` object reports both), hence the default.
    ``overlap="split"`` keeps the three-way split but runs the collective on the compute stream (measurement only)."""

    def __init__(self, v_local: int, v_total: int, h: int, w: int, heads: int, device, group=None, overlap=False):
        from . import _lib
        self.group = group
        self.world = v_total // v_local
        self.group_index = dist.get_rank(group) if (dist.is_initialized() and self.world > 1) else 0
        per_rank = v_local * h * w * 2 * heads * 64          # bf16 elements at the finest level (head_dim_pad 64)
        self.send = torch.empty(per_rank, dtype=torch.bfloat16, device=device)
        self.recv = torch.empty(per_rank * self.world, dtype=torch.bfloat16, device=device)
        self.overlap = overlap
        self.side = torch.cuda.Stream(device=device) if (overlap is True and torch.device(device).type == "cuda") else None
        self.done = torch.cuda.Event() if self.side is not None else None
        self.calls = 0
        self.bytes_sent = 0

        def _cb(user, send_ptr, recv_ptr, nbytes, stream, phase):
            try:
                if phase == _lib.EXCHANGE_BEGIN:
                    self.begin(nbytes // 2)
                    return _lib.EXCHANGE_DONE if (self.overlap is False and self.world > 1) else 0
                self.end()
                return 0
            except Exception:                                  # never unwind through the C frame
                import traceback
                traceback.print_exc()
                return 1
        self.callback = _lib.KV_EXCHANGE_FN(_cb)               # keep a reference: ctypes callbacks are not owned by C

    def all_gather(self, n_elems: int) -> None:
        """recv[r * n : (r + 1) * n] = rank r's send[:n]  (rank order == view order)"""
        if self.world == 1 or not dist.is_initialized():
            self.recv[:n_elems].copy_(self.send[:n_elems])
            return
        dist.all_gather_into_tensor(self.recv[: n_elems * self.world], self.send[:n_elems], group=self.group)

    def begin(self, n_elems: int) -> None:
        self.calls += 1
        self.bytes_sent += n_elems * 2
        if self.world == 1:
            return                                             # the library reads the local slab straight from `send`
        if self.side is None:
            self.all_gather(n_elems)
            return
        cur = torch.cuda.current_stream(self.send.device)
        self.side.wait_stream(cur)                             # after the K|V pack (and the previous block's readers)
        with torch.cuda.stream(self.side):
            self.all_gather(n_elems)
            self.done.record(self.side)

    def end(self) -> None:
        if self.world > 1 and self.side is not None:
            torch.cuda.current_stream(self.send.device).wait_event(self.done)
