"""Determinism / correctness probe for large batches: tools/dbg_verify.py <batch>

Runs Params(1024) bootstrap_batch twice on the same inputs, lists gates whose outputs differ between the runs and gates
that decrypt to the wrong AND / OR / XOR.  (Found the barrier-less twiddle publication that only showed up when the
persistent CTAs had drifted apart, i.e. at batches of several waves.)"""
import os, sys, numpy as np, ctypes as C
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import importlib
sg = importlib.import_module("sgfhe_jl_b200")
import torch
n = 1024; batch = int(sys.argv[1])
P = sg.Params(n, device=0)
sk = sg.PrivateKey(P, np.random.default_rng([0, 1]))
bkey = sg.BootstrapKey(np.random.default_rng([0, 2]), sk); bkey.upload()
rng = np.random.default_rng([0, 3, 0])
blocks = (2 * batch + n - 1) // n
bits, lw = [], []
for _ in range(blocks):
    msg = rng.integers(0, 2, size=n, dtype=np.uint8)
    ct = sg.encrypt(sk, rng, msg)
    lw.append(np.stack([e.lwe.flat() for e in sg.split_ciphertext(ct)])); bits.append(msg)
bits = np.concatenate(bits)[: 2 * batch]; lw = np.concatenate(lw)[: 2 * batch]
outs = sg.bootstrap_batch(bkey, None, lw[:batch], lw[batch:])
outs2 = sg.bootstrap_batch(bkey, None, lw[:batch], lw[batch:])
skb = sk.key.astype(bool)
y1, y2 = bits[:batch].astype(np.int64), bits[batch:].astype(np.int64)
for name, arr, arr2, want in zip("and or xor".split(), outs, outs2, (y1 & y2, y1 | y2, y1 ^ y2)):
    b1 = (arr[:, n].astype(np.int64) - arr[:, :n][:, skb].astype(np.int64).sum(axis=1)) % P.r
    got = ((b1 + P.Dr // 2) % P.r) // P.Dr
    bad = np.nonzero(got != want)[0]
    nd = np.nonzero((arr != arr2).any(axis=1))[0]
    print(name, "bad gates:", bad[:40], "count", len(bad), "| nondeterministic:", nd[:20], len(nd))
