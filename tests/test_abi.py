"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/sgfhe_cuda.h
declares, derives Params exactly like the oracle, mirrors the reference's error behaviour, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(sg):
    from sgfhe_jl_b200 import _lib
    sg.build()
    header = open(os.path.join(ROOT, "include", "sgfhe_cuda.h")).read()
    declared = set(re.findall(r"\b(sgfhe_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS)
    L = C.CDLL(_lib.SO_PATH)
    for name in sorted(declared):
        assert getattr(L, name) is not None


def test_library_is_sm100a_only(sg):
    import subprocess
    from sgfhe_jl_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048])
def test_params_derive_matches_oracle(sg, so, n):
    P, OP = sg.Params(n), so.Params(n)
    assert (P.n, P.t, P.m, P.r, P.q, P.Dr, P.Dq, P.Q, P.B, P.DQ_tilde) == \
           (OP.n, OP.t, OP.m, OP.r, OP.q, OP.Dr, OP.Dq, OP.Q, OP.B, OP.DQ)


def test_params_errors_mirror_reference(sg):
    for n in (0, 32, 63, 100):
        with pytest.raises(sg.SgfheError, match="power of two"):     # @assert at src/fhe.jl:45-46
            sg.Params(n)
    with pytest.raises(sg.SgfheError):
        sg.Params(1 << 20)                                            # error("n=... is too large") src/fhe.jl:77


def test_no_cpu_fallback(sg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    P = sg.Params(64)
    with pytest.raises(sg.SgfheError, match="no CUDA device"):
        P.ctx
    rng = np.random.default_rng(0)
    sk = sg.PrivateKey(P, rng)
    with pytest.raises(sg.SgfheError):
        sg.BootstrapKey(rng, sk)                                      # products run on the GPU only


def test_product_path_does_not_touch_the_oracle():
    """the oracle is test infrastructure: nothing under sgfhe.jl_b200/ or include/ may reference it"""
    for base in ("sgfhe.jl_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "sgfhe_oracle" not in txt and "libsgfhe_oracle" not in txt and "import model" not in txt, f


def test_host_encrypt_split_decrypt_roundtrip(sg):
    """host-side mirror (no GPU needed): test/api.test.jl:8-18 and :33-42 at Params(512)"""
    rng = np.random.default_rng(12)
    P = sg.Params(512)
    sk = sg.PrivateKey(P, rng)
    msg = rng.integers(0, 2, size=P.n, dtype=np.uint8)
    ct = sg.encrypt(sk, rng, msg)
    assert np.array_equal(sg.decrypt(sk, ct), msg.astype(bool))
    for e, b in zip(sg.split_ciphertext(ct), msg):
        assert sg.decrypt(sk, e) == bool(b)
    with pytest.raises(sg.SgfheError):
        sg.encrypt(sk, rng, msg[:-1])                                 # @assert length(message) == n, src/fhe.jl:313


def test_host_split_matches_oracle(sg, so):
    rng = np.random.default_rng(13)
    P, OP = sg.Params(64), so.Params(64)
    a = rng.integers(0, P.r, size=64, dtype=np.uint64)
    b = rng.integers(0, P.r, size=64, dtype=np.uint64)
    got = np.stack([e.lwe.flat() for e in sg.split_ciphertext(sg.PackedCiphertext(P, a, b))])
    assert np.array_equal(got, so.split_ciphertext(OP, a, b))


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_scheme2_params_match_oracle(sg, so, k):
    """Scheme2.Params(k), src/fhe2.jl:36-70"""
    P, OP = sg.Scheme2Params(k), so.scheme2_params(k)
    for f in ("n", "k", "t", "r", "m", "q", "tau", "B", "Bp", "Dr", "Dq"):
        assert getattr(P, f) == getattr(OP, f), f
    assert (P.B - 1) % P.r == 0 and (P.Bp - 1) % P.r == 0 and P.Bp < P.B
    with pytest.raises(sg.SgfheError):
        sg.Scheme2Params(6)                                           # @assert 1 <= k <= 5, src/fhe2.jl:39
