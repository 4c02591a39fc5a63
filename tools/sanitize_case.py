#!/usr/bin/env python3
"""Small cases for compute-sanitizer: one truncated Params(512) trace (fused v4 kernel, m = 4096), one truncated
Params(1024) trace, one Params(64) gate (v3 kernel), one polymul; each checked against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import sgfhe_jl_b200 as sg
import sgfhe_oracle as so

for n, steps in ((512, 2), (1024, 1), (64, 2)):
    P, OP = sg.Params(n), so.Params(n)
    sk = so.make_secret(OP, 1); key = so.make_bkey(OP, sk, 1, rows=steps); _, lwes = so.make_lwes(OP, sk, 1)
    bkey = sg.BootstrapKey(params=P, key=key)
    ga, go, gx, gtr = sg.bootstrap_trace(bkey, None, lwes[1], lwes[2], n_steps=steps)
    ra, ro, rx, rtr = so.bootstrap_internal(OP, key, lwes[1], lwes[2], n_steps=steps, trace=True, fast=True)
    assert np.array_equal(gtr, rtr) and np.array_equal(ga, ra), n
    a = so.rand_below(np.random.default_rng(n), OP.Q, (1, OP.m))
    assert np.array_equal(sg.polymul(P, a, a)[0], so.polymul(a[0], a[0], OP.Q))
    P.close()
    print("ok", n)
