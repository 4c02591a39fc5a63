"""Chained bootstraps with the ciphertexts resident on the device (examples/depth.jl:63-78).

Output LWEs of `bootstrap` are over Z_r like its inputs (src/fhe.jl:616-618), so layers chain with no conversion
and no host round trip: layer l+1 bootstraps (AND_l, XOR_l) exactly as the reference example feeds
`enc_y1 = enc_and; enc_y2 = enc_xor` back in.  torch is used for device memory and the stream only.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check


def bootstrap_chain(bkey, lwes1: np.ndarray, lwes2: np.ndarray, layers: int, keep_layers: bool = False, dist=None,
                    wiring: str = "local"):
    """Run `layers` sequential gate layers over a batch of independent bit pairs.

    lwes1, lwes2: uint64[batch, n+1].  Returns the last layer's (and, or, xor) as numpy arrays, or with
    `keep_layers` a list of such triples, one per layer (for per-layer decrypt checks as in depth.jl:65-69).

    Multi-GPU (one process per GPU, `dist` = torch.distributed): with wiring="local" every rank chains its own gates and
    nothing is exchanged; with wiring="allgather" the (AND, XOR) outputs of a layer are all-gathered (one NCCL collective
    per layer, W x 2 x (n+1) small integers per rank) and rank r takes the outputs of rank r+1 as its next inputs
    (parallel.exchange_layer), i.e. the wires of the layered circuit cross GPUs."""
    import torch
    P = bkey.params
    if layers < 1:
        raise _lib.SgfheError("layers must be >= 1")
    bkey.upload()
    L = _lib.lib()
    dev = torch.device("cuda", P.device)
    with torch.cuda.device(dev):
        a = torch.from_numpy(np.ascontiguousarray(lwes1, np.uint64).view(np.int64)).to(dev)
        b = torch.from_numpy(np.ascontiguousarray(lwes2, np.uint64).view(np.int64)).to(dev)
        batch = a.shape[0]
        bufs = [[torch.empty_like(a) for _ in range(3)] for _ in range(2)]
        stream = torch.cuda.current_stream()
        kept = []
        for layer in range(layers):
            o = bufs[layer & 1]
            check(L.sgfhe_bootstrap_batch_device(P.ctx, batch, a.data_ptr(), b.data_ptr(), None,
                                                 o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), stream.cuda_stream))
            if keep_layers:
                stream.synchronize()
                kept.append(tuple(t.cpu().numpy().view(np.uint64).copy() for t in o))
            a, b = o[0], o[2]                                   # enc_y1 = enc_and; enc_y2 = enc_xor  (depth.jl:71-72)
            if wiring == "allgather" and dist is not None and layer + 1 < layers:
                from .parallel import exchange_layer
                a, b = exchange_layer(a, b, dist)
        stream.synchronize()
        last = tuple(t.cpu().numpy().view(np.uint64).copy() for t in bufs[(layers - 1) & 1])
    return kept if keep_layers else last
