"""Multi-GPU plumbing for the bootstrapping path: one process per GPU over torch.distributed.

Gates are independent (`bootstrap` is a pure function of the key, two LWEs and the draws), so the batch is
sharded with no data-path collective.  The only exchange is the one-off NCCL broadcast of
the pre-transformed key.  The reference is single-process; nothing here has an upstream counterpart.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


def shard_bounds(total: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced [start, stop) per rank; the first `total % world` ranks get one extra gate."""
    if world < 1 or total < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total, world)
    out, start = [], 0
    for r in range(world):
        stop = start + base + (1 if r < extra else 0)
        out.append((start, stop))
        start = stop
    return out


def bootstrap_sharded(compute, lwes1: np.ndarray, lwes2: np.ndarray, dist=None, gather: bool = True):
    """Run `compute(lwe1_shard, lwe2_shard) -> (and, or, xor)` on this rank's shard of the batch.

    `compute` is normally `lambda a, b: sg.bootstrap_batch(bkey, None, a, b)`.  With `gather` every rank
    returns the full [batch, n+1] outputs (one all_gather of small integer arrays); without it each rank
    returns its shard only (what a layered circuit with rank-local wiring needs)."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    lo, hi = shard_bounds(lwes1.shape[0], world)[rank]
    outs = compute(np.ascontiguousarray(lwes1[lo:hi]), np.ascontiguousarray(lwes2[lo:hi]))
    if not gather or world == 1:
        return outs
    full = []
    for o in outs:
        parts = [None] * world
        dist.all_gather_object(parts, np.ascontiguousarray(o))
        full.append(np.concatenate(parts, axis=0))
    return tuple(full)


def exchange_layer(t_and, t_xor, dist, shift: int = 1):
    """Layered-circuit wiring across ranks: all-gather the (AND, XOR) outputs of one layer -- torch tensors
    [W, n+1] on every rank, CUDA under NCCL or CPU under gloo -- and return the pair of rank (rank + shift) % world as this
    rank's next inputs.  One collective per layer; the two tensors travel as one buffer."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.stack([t_and, t_xor]).contiguous()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    src = parts[(rank + shift) % world]
    return src[0].contiguous(), src[1].contiguous()


def broadcast_key(params, rows: int, dist, src: int = 0):
    """NCCL-broadcast the pre-transformed key from `src` into every rank's library-owned device buffer.
    Rank `src` must have uploaded its key (BootstrapKey.upload) before the call.  The broadcast gives every context a
    new key token: afterwards use `BootstrapKey.resident(params)` (or `bkey.rebind()` on the root's own key object)."""
    import torch
    L = _lib.lib()
    dptr, nbytes = C.c_void_p(), C.c_uint64()
    check(L.sgfhe_bkey_device_buffer(params.ctx, rows, C.byref(dptr), C.byref(nbytes)))

    class _Alias:
        __cuda_array_interface__ = {"shape": (nbytes.value,), "typestr": "|u1", "data": (dptr.value, False), "version": 3}

    t = torch.as_tensor(_Alias(), device="cuda")
    dist.broadcast(t, src=src)
    torch.cuda.synchronize()
    check(L.sgfhe_bkey_adopt(params.ctx, rows))
    return nbytes.value
