"""world_size-2 CPU tests (gloo) of the N>1 host path: shard bounds, the sharded driver with gather, and
rank-local shards.  The compute engine injected here is the CPU oracle (test infrastructure), because there
is no GPU on the CPU test box; on a GPU box the same driver is handed sg.bootstrap_batch."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition(sg):
    for total in (0, 1, 7, 148, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            b = sg.shard_bounds(total, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sg.shard_bounds(4, 0)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import sgfhe_jl_b200 as sg
    import sgfhe_oracle as so
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        OP = so.Params(64)
        sk = so.make_secret(OP, 0)
        steps = 3                                        # truncated gates keep the CPU test fast
        key = so.make_bkey(OP, sk, 0, rows=steps)
        _, lwes = so.make_lwes(OP, sk, 0)
        l1, l2 = lwes[:5], lwes[5:10]                    # 5 gates over 2 ranks: shards of 3 and 2

        def compute(a, b):
            return so.bootstrap_batch(OP, key, a, b, n_steps=steps, literal=False, threads=1)

        full = sg.bootstrap_sharded(compute, l1, l2, dist=dist, gather=True)
        local = sg.bootstrap_sharded(compute, l1, l2, dist=dist, gather=False)
        ref = compute(l1, l2)
        lo, hi = sg.shard_bounds(5, world)[rank]
        ok = all(np.array_equal(f, r) for f, r in zip(full, ref)) and \
            all(np.array_equal(x, r[lo:hi]) for x, r in zip(local, ref))
        # layered-circuit wiring across ranks: after the exchange rank r holds rank r+1's (AND, XOR)
        import torch
        mine = compute(l1[lo:hi][:2], l2[lo:hi][:2])     # two gates per rank so both ranks send the same shape
        t_and, t_xor = torch.from_numpy(mine[0].view(np.int64).copy()), torch.from_numpy(mine[2].view(np.int64).copy())
        n_and, n_xor = sg.exchange_layer(t_and, t_xor, dist)
        olo, ohi = sg.shard_bounds(5, world)[(rank + 1) % world]
        other = compute(l1[olo:ohi][:2], l2[olo:ohi][:2])
        ok = ok and np.array_equal(n_and.numpy().view(np.uint64), other[0]) and np.array_equal(n_xor.numpy().view(np.uint64), other[2])
        q.put((rank, bool(ok), (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_sharded_bootstrap_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, (0, 3)), (1, True, (3, 5))]
