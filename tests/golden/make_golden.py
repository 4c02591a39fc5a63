#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the CPU oracle (oracle/sgfhe_oracle.c), cross-checked against the
independent big-integer model (oracle/model.py) before anything is written.

The reference holds no golden vectors and cannot run here (Julia + DarkIntegers absent), so these vectors
pin THIS repository's oracle against regressions and pin the GPU path at sizes the CPU tests do not rerun;
parity with the reference itself rests on exact-ring uniqueness (oracle/sgfhe_oracle.h).

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import model as md            # noqa: E402
import sgfhe_oracle as so     # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gate_case(n, seed, pairs, steps, with_draws, check_model_steps):
    OP = so.Params(n)
    sk = so.make_secret(OP, seed)
    key = so.make_bkey(OP, sk, seed, rows=steps)
    bits, lwes = so.make_lwes(OP, sk, seed)
    rng = np.random.default_rng([seed, 9])
    xmax = OP.B // 2 * 3
    out = {"n": n, "seed": seed, "steps": steps, "Q": so.pack([OP.Q])[0], "B": so.pack([OP.B])[0],
           "key_sha256": sha(key), "pairs": np.array(pairs), "sk": sk, "bits": bits}
    for pi, (i, j) in enumerate(pairs):
        for mode in (["det", "rnd"] if with_draws else ["det"]):
            draws = rng.integers(-xmax, xmax + 1, size=(steps, 2, OP.m, 2), dtype=np.int64) if mode == "rnd" else None
            a, o, x, tr = so.bootstrap_internal(OP, key, lwes[i], lwes[j], draws=draws, n_steps=steps, trace=True)
            fa, fo, fx, ftr = so.bootstrap_internal(OP, key, lwes[i], lwes[j], draws=draws, n_steps=steps, trace=True, fast=True)
            assert np.array_equal(tr, ftr) and np.array_equal(a, fa) and np.array_equal(o, fo) and np.array_equal(x, fx)
            if check_model_steps and pi == 0:
                M = md.params(n)
                mk = so.unpack(key[:check_model_steps])
                mtr = []
                md.bootstrap_internal(M, mk, lwes[i].tolist(), lwes[j].tolist(),
                                      None if draws is None else draws[:check_model_steps].tolist(),
                                      n_steps=check_model_steps, trace=mtr)
                for k in range(check_model_steps):
                    assert so.unpack(tr[k, 0]) == mtr[k][0] and so.unpack(tr[k, 1]) == mtr[k][1], "oracle != model"
            tag = f"p{pi}_{mode}"
            out[tag + "_lwe1"], out[tag + "_lwe2"] = lwes[i], lwes[j]
            if draws is not None:
                out[tag + "_draws_sha256"] = sha(draws)
                out[tag + "_draws_seed"] = np.array([seed, 9])
            out[tag + "_and_Q"], out[tag + "_or_Q"], out[tag + "_xor_Q"] = a, o, x
            out[tag + "_trace_sha256"] = np.array([sha(tr[k]) for k in range(steps)])
            out[tag + "_trace_last_head"] = tr[steps - 1][:, :8]
            if steps == n:
                r = so.bootstrap(OP, key, lwes[i], lwes[j], draws)
                out[tag + "_and_r"], out[tag + "_or_r"], out[tag + "_xor_r"] = r
                y1, y2 = int(bits[i]), int(bits[j])
                assert tuple(so.decrypt_lwe(OP, sk, v) for v in r) == (y1 & y2, y1 | y2, y1 ^ y2)
    return out


CASES = {
    "golden_p64.npz": lambda: gate_case(64, 0, [(10, 20), (0, 63), (5, 6)], 64, True, 2),
    "golden_p1024_trunc.npz": lambda: gate_case(1024, 1, [(3, 700)], 2, True, 1),
    "golden_p512_trunc.npz": lambda: gate_case(512, 2, [(1, 2)], 2, False, 0),
    "golden_p2048_trunc.npz": lambda: gate_case(2048, 3, [(5, 1500)], 2, True, 1),      # m = 16384, 93-bit Q, six primes on the GPU
}

if __name__ == "__main__":
    # existing files are left alone (a zip archive is not byte-reproducible) unless --all is given
    for name, make in CASES.items():
        path = os.path.join(HERE, name)
        if "--all" in sys.argv or not os.path.exists(path):
            np.savez_compressed(path, **make())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
