#!/usr/bin/env python3
"""bench_kernels.py -- kernel-level sweeps for BASELINE.json configs[2]: standalone batched NTT polymuls
(DarkIntegers `Polynomial *`, src/fhe.jl:527-528) over the ring degrees / moduli of Params(64..1024), and the
two-limb RNS modmul of Scheme 2 (src/rns.jl:51-52) over the moduli of Scheme2.Params(1..5).

    python bench_kernels.py [--batch 2048] [--count 16777216] [--reps 5]      # prints one JSON line per case

Inputs are resident in HBM; timing with CUDA events on the launching stream after warm-up.  Operands of the
polymul sweep are larger than L2 at the larger sizes; the RNS sweep streams 6 x 128 MiB per launch.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--count", type=int, default=1 << 24)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import sgfhe_jl_b200 as sg
    from sgfhe_jl_b200 import _lib
    L = _lib.lib()
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6650.0
    try:
        imad = json.load(open(os.path.join(ROOT, "profiles", "int_peaks_r01.json")))["imad_lo"]["thread_ops_per_s"]
    except Exception:
        imad = 148 * 64 * 1.965e9
    stream = torch.cuda.current_stream()

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.reps

    for n in (64, 128, 256, 512, 1024, 2048):
        P = sg.Params(n)
        m, qbits = P.m, P.Q.bit_length()
        rng = np.random.default_rng(n)
        batch = max(64, (args.batch * 8192) // m)            # the same operand bytes (268 MB per operand array) at every ring degree
        shape = (batch, m, 2)
        a = rng.integers(0, 1 << 62, size=shape, dtype=np.uint64)
        a[..., 1] = 0 if qbits <= 64 else a[..., 1] & np.uint64((1 << (qbits - 65)) - 1)     # canonical: below Q
        if qbits <= 64:
            a[..., 0] %= np.uint64(P.Q)
        da = torch.from_numpy(a.view(np.int64)).cuda()
        db = da.flip(0).contiguous()
        do = torch.empty_like(da)
        ms = timed(lambda: _lib.check(L.sgfhe_polymul_device(P.ctx, batch, da.data_ptr(), db.data_ptr(), do.data_ptr(), stream.cuda_stream)))
        w = (qbits + 31) // 32
        modmuls = 3 * (m // 2) * (m.bit_length() - 1) + m                     # standalone product
        rate = batch / (ms / 1e3)
        print(json.dumps({"kernel": "polymul_kernel_v4" if 4096 <= m <= 8192 else "polymul_kernel", "workload": f"Params({n}): m={m}, {qbits}-bit Q, batch {batch}, both operands full size",
                          "polymuls_per_s": rate, "ms": ms,
                          "roofline": {"bound": "int32-pipe", "achieved": rate * modmuls * (2 * w * w + w) / 1e9, "peak": imad / 1e9, "unit": "GIMAD/s",
                                       "frac": rate * modmuls * (2 * w * w + w) / imad},
                          "hbm_gbs": rate * 3 * m * 16 / 1e9}))
        P.close()
    for k in (1, 2, 3, 4, 5):
        S2 = sg.Scheme2Params(k)
        rng = np.random.default_rng(k)
        t = [torch.from_numpy(rng.integers(0, M, size=args.count, dtype=np.uint64).view(np.int64)).cuda() for M in (S2.B, S2.Bp, S2.B, S2.Bp)]
        o = [torch.empty_like(t[0]) for _ in range(2)]
        ms = timed(lambda: _lib.check(L.sgfhe_rns2_op_device(0, 0, args.count, *[x.data_ptr() for x in t], S2.B, S2.Bp, o[0].data_ptr(), o[1].data_ptr(), stream.cuda_stream)))
        gbs = args.count * 6 * 8 / (ms / 1e3) / 1e9
        print(json.dumps({"kernel": "rns2_kernel", "workload": f"Scheme2.Params({k}): B={S2.B}, Bp={S2.Bp}, {args.count} operand pairs (RNS2Number *)",
                          "modmuls_per_s": 2 * args.count / (ms / 1e3), "ms": ms,
                          "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm}}))

    # Scheme 2 (src/fhe2.jl; no reference bootstrap exists): two-limb negacyclic products, transforms and the transform-domain
    # MAC of an external product at the ring degree of Scheme2.Params(k).  Algorithmic work of a product: 3 transforms of
    # (m/2) log2 m modmuls + m pointwise, per limb; a 64-bit word modmul counted as a 2-limb Montgomery product (10 IMAD).
    for k in (1, 2, 3, 4, 5):
        ctx = sg.Scheme2Context(k)
        S2, m = ctx.params, ctx.params.m
        batch = max(64, (args.batch * 8192) // (4 * m))
        rng = np.random.default_rng(10 + k)
        mk = lambda M, shape: torch.from_numpy(rng.integers(0, M, size=shape, dtype=np.uint64).view(np.int64)).cuda()
        a = [mk(S2.B, (batch, m)), mk(S2.Bp, (batch, m))]
        b = [mk(S2.B, (batch, m)), mk(S2.Bp, (batch, m))]
        wa, wb = [x.clone() for x in a], [x.clone() for x in b]
        o = [torch.empty_like(a[0]) for _ in range(2)]

        def product():
            for w_, src in zip(wa + wb, a + b):
                w_.copy_(src)                                            # the product overwrites its operands with their transforms
            _lib.check(L.sgfhe_s2_polymul_device(ctx._h, batch, wa[0].data_ptr(), wa[1].data_ptr(), wb[0].data_ptr(), wb[1].data_ptr(), 0,
                                                 o[0].data_ptr(), o[1].data_ptr(), stream.cuda_stream))
        ms_copy = timed(lambda: [w_.copy_(src) for w_, src in zip(wa + wb, a + b)])
        ms = timed(product) - ms_copy
        modmuls = 2 * (3 * (m // 2) * (m.bit_length() - 1) + m)
        rate = batch / (ms / 1e3)
        print(json.dumps({"kernel": "s2_fwd_top/_local + s2_mulinv_local + s2_inv_top", "workload": f"Scheme2.Params({k}): m={m}, B={S2.B}, Bp={S2.Bp}, batch {batch} products of Polynomial{{RNS2Number}} (no reference bootstrap for this scheme)",
                          "polymuls_per_s": rate, "ms": ms,
                          "roofline": {"bound": "int32-pipe", "achieved": rate * modmuls * 10 / 1e9, "peak": imad / 1e9, "unit": "GIMAD/s", "frac": rate * modmuls * 10 / imad},
                          "hbm_gbs": rate * 2 * 3 * m * 8 / 1e9}))
        d = [mk(S2.B, (batch, 4, m)), mk(S2.Bp, (batch, 4, m))]
        K = [mk(S2.B, (batch, 4, 2, m)), mk(S2.Bp, (batch, 4, 2, m))]
        om = [torch.empty((batch, 2, m), dtype=torch.int64, device="cuda") for _ in range(2)]
        ms = timed(lambda: _lib.check(L.sgfhe_s2_mac8_device(ctx._h, batch, d[0].data_ptr(), d[1].data_ptr(), K[0].data_ptr(), K[1].data_ptr(),
                                                            om[0].data_ptr(), om[1].data_ptr(), stream.cuda_stream)))
        gbs = batch * 2 * (4 + 8 + 2) * m * 8 / (ms / 1e3) / 1e9
        print(json.dumps({"kernel": "s2_mac8", "workload": f"Scheme2.Params({k}): transform-domain MAC of {batch} external products (4 digit x 4x2 key polynomials, two limbs)",
                          "external_macs_per_s": batch / (ms / 1e3), "ms": ms,
                          "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm}}))
        ctx.close()


if __name__ == "__main__":
    main()
