"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle on identical inputs.

Bar: bit-exact (all arithmetic on the path is integer).  Mirrors test/internals.test.jl and
test/api.test.jl of the reference, plus per-step accumulator equality the reference cannot check.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env64(so, sg):
    P = sg.Params(64)
    OP = so.Params(64)
    sk = so.make_secret(OP, 0)
    key = so.make_bkey(OP, sk, 0)
    bits, lwes = so.make_lwes(OP, sk, 0)
    bkey = sg.BootstrapKey(params=P, key=key)
    return P, OP, sk, key, bits, lwes, bkey


def _edge_polys(m, Q, so):
    z = np.zeros((m, 2), np.uint64)
    one = z.copy(); one[0, 0] = 1
    top = z.copy(); top[m - 1, 0] = 1                       # x^(m-1)
    full = np.broadcast_to(so.pack([Q - 1]), (m, 2)).copy()  # all coefficients Q-1
    return [z, one, top, full]


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024])
def test_polymul_matches_oracle(so, sg, n):
    """DarkIntegers `Polynomial *` seam (called at src/fhe.jl:527-528): random and edge operands."""
    P, OP = sg.Params(n), so.Params(n)
    rng = np.random.default_rng(100 + n)
    a = so.rand_below(rng, OP.Q, (6, OP.m))
    b = so.rand_below(rng, OP.Q, (6, OP.m))
    edges = _edge_polys(OP.m, OP.Q, so)
    a = np.concatenate([a, np.stack(edges), np.stack(edges[::-1])])
    b = np.concatenate([b, np.stack(edges), np.stack(edges)])
    got = sg.polymul(P, a, b)
    for i in range(a.shape[0]):
        assert np.array_equal(got[i], so.polymul(a[i], b[i], OP.Q)), f"product {i}"
    P.close()


@pytest.mark.parametrize("n", [64, 512, 1024])
@pytest.mark.parametrize("use_rng", [False, True])
def test_flatten_poly_matches_oracle(so, sg, n, use_rng):
    """flatten_poly (src/utils.jl:253-264), port of test/internals.test.jl:115-141 plus exact equality."""
    P, OP = sg.Params(n), so.Params(n)
    rng = np.random.default_rng(200 + n)
    a = so.rand_below(rng, OP.Q, (OP.m,))
    a[0] = 0; a[1] = so.pack([OP.Q - 1])[0]; a[2] = so.pack([OP.B])[0]; a[3] = so.pack([OP.B - 1])[0]
    xmax = OP.B // 2 * 3
    draws = rng.integers(-xmax, xmax + 1, size=(OP.m, 2), dtype=np.int64) if use_rng else None
    if use_rng:
        draws[0] = (-xmax, xmax); draws[1] = (xmax, -xmax)
    got = sg.flatten_poly(P, draws, a)
    ref = so.flatten_poly(a, OP.B, 2, OP.Q, draws)
    assert np.array_equal(got, ref)
    # recomposition (test/internals.test.jl:138-140) and limits (:50-66)
    gi = so.unpack(got)
    for j in range(0, OP.m, 97):
        assert (gi[0][j] + gi[1][j] * OP.B) % OP.Q == so.unpack(a[j:j + 1])[0]
    P.close()


@pytest.mark.parametrize("use_rng", [False, True])
def test_external_product_with_gadget_is_identity(so, sg, use_rng):
    """port of test/internals.test.jl:144-166: (a,b) (.) G == (a,b)"""
    P, OP = sg.Params(64), so.Params(64)
    rng = np.random.default_rng(7)
    a, b = so.rand_below(rng, OP.Q, (OP.m,)), so.rand_below(rng, OP.Q, (OP.m,))
    G = np.zeros((4, 2, OP.m, 2), np.uint64)
    G[0, 0, 0, 0] = 1; G[1, 0, 0, 0] = OP.B; G[2, 1, 0, 0] = 1; G[3, 1, 0, 0] = OP.B
    xmax = OP.B // 2 * 3
    draws = rng.integers(-xmax, xmax + 1, size=(2, OP.m, 2), dtype=np.int64) if use_rng else None
    oa, ob = sg.external_product(P, draws, a, b, G)
    assert np.array_equal(oa, a) and np.array_equal(ob, b)
    P.close()


@pytest.mark.parametrize("n", [64, 1024])
@pytest.mark.parametrize("use_rng", [False, True])
def test_external_product_matches_oracle(so, sg, n, use_rng):
    """external_product (src/fhe.jl:519-530) with a random full-size A"""
    P, OP = sg.Params(n), so.Params(n)
    rng = np.random.default_rng(300 + n)
    a, b = so.rand_below(rng, OP.Q, (OP.m,)), so.rand_below(rng, OP.Q, (OP.m,))
    A = so.rand_below(rng, OP.Q, (4, 2, OP.m))
    xmax = OP.B // 2 * 3
    draws = rng.integers(-xmax, xmax + 1, size=(2, OP.m, 2), dtype=np.int64) if use_rng else None
    oa, ob = sg.external_product(P, draws, a, b, A)
    ra, rb = so.external_product(a, b, A, OP.B, OP.Q, draws)
    assert np.array_equal(oa, ra) and np.array_equal(ob, rb)
    P.close()


@pytest.mark.parametrize("use_rng", [False, True])
def test_bootstrap_trace_p64_every_step(env64, so, sg, use_rng):
    """_bootstrap_internal (src/fhe.jl:559-595): accumulator after every one of the n steps and the three
    output LWEs over Z_Q equal the oracle's LITERAL formulation."""
    P, OP, sk, key, bits, lwes, bkey = env64
    rng = np.random.default_rng(11)
    xmax = OP.B // 2 * 3
    draws = rng.integers(-xmax, xmax + 1, size=(OP.n, 2, OP.m, 2), dtype=np.int64) if use_rng else None
    for (i, j) in [(10, 20), (0, 63)]:
        ga, go, gx, gtr = sg.bootstrap_trace(bkey, draws, lwes[i], lwes[j])
        ra, ro, rx, rtr = so.bootstrap_internal(OP, key, lwes[i], lwes[j], draws=draws, trace=True)
        for k in range(OP.n):
            assert np.array_equal(gtr[k], rtr[k]), f"accumulator differs after step {k}"
        assert np.array_equal(ga, ra) and np.array_equal(go, ro) and np.array_equal(gx, rx)


@pytest.mark.parametrize("use_rng", [False, True])
def test_bootstrap_gates_p64(env64, so, sg, use_rng):
    """port of test/api.test.jl:45-83: 32 disjoint bit pairs, rng and nothing; decrypt(AND/OR/XOR) is the
    plaintext gate, and every output LWE equals the oracle's bootstrap() bit for bit."""
    P, OP, sk, key, bits, lwes, bkey = env64
    l1, l2 = lwes[:32], lwes[32:64]
    if use_rng:
        rng = np.random.default_rng(21)
        xmax = OP.B // 2 * 3
        draws = rng.integers(-xmax, xmax + 1, size=(32, OP.n, 2, OP.m, 2), dtype=np.int64)
        import ctypes as C
        outs = [np.zeros_like(l1) for _ in range(3)]
        from sgfhe_jl_b200 import _lib
        bkey.upload()
        _lib.check(_lib.lib().sgfhe_bootstrap_batch(P.ctx, 32, l1.ctypes.data_as(C.c_void_p), l2.ctypes.data_as(C.c_void_p),
                                                    draws.ctypes.data_as(C.c_void_p), *[o.ctypes.data_as(C.c_void_p) for o in outs]))
    else:
        draws = None
        outs = sg.bootstrap_batch(bkey, None, l1, l2)
    for g in range(32):
        y1, y2 = int(bits[g]), int(bits[32 + g])
        assert so.decrypt_lwe(OP, sk, outs[0][g]) == (y1 & y2)
        assert so.decrypt_lwe(OP, sk, outs[1][g]) == (y1 | y2)
        assert so.decrypt_lwe(OP, sk, outs[2][g]) == (y1 ^ y2)
    for g in range(0, 32, 5):
        ref = so.bootstrap(OP, key, l1[g], l2[g], None if draws is None else draws[g])
        for o, r in zip(outs, ref):
            assert np.array_equal(o[g], r)


def test_more_gates_than_resident_ctas_p64(env64, so, sg):
    """a batch larger than the number of resident CTAs (dynamic gate distribution through the work counter, small-m
    kernel): same rows as one-gate launches, every gate decrypts to the plaintext gate, run to run identical"""
    P, OP, sk, key, bits, lwes, bkey = env64
    W = 3000
    idx1 = np.arange(W) % 32
    idx2 = 32 + (np.arange(W) * 7) % 32
    l1, l2 = lwes[idx1], lwes[idx2]
    o1 = sg.bootstrap_batch(bkey, None, l1, l2)
    o2 = sg.bootstrap_batch(bkey, None, l1, l2)
    small = sg.bootstrap_batch(bkey, None, l1[:32], l2[:32])
    skb = np.asarray(sk, dtype=bool)
    y1, y2 = bits[idx1].astype(np.int64), bits[idx2].astype(np.int64)
    for a, b, s_, want in zip(o1, o2, small, (y1 & y2, y1 | y2, y1 ^ y2)):
        assert np.array_equal(a, b)
        assert np.array_equal(a[:32], s_)
        assert np.array_equal(a[32:64], a[2432:2464])          # gates 32 + k and 2432 + k have the same inputs (2400 = 75 * 32, 7 * 2400 = 525 * 32)
        b1 = (a[:, OP.n].astype(np.int64) - a[:, :OP.n][:, skb].astype(np.int64).sum(axis=1)) % OP.r
        assert np.array_equal(((b1 + OP.Dr // 2) % OP.r) // OP.Dr, want)


def test_bootstrap_is_deterministic_without_rng(env64, sg):
    """docs/src/manual.md:155-169"""
    P, OP, sk, key, bits, lwes, bkey = env64
    o1 = sg.bootstrap_batch(bkey, None, lwes[:4], lwes[4:8])
    o2 = sg.bootstrap_batch(bkey, None, lwes[:4], lwes[4:8])
    assert all(np.array_equal(a, b) for a, b in zip(o1, o2))


def test_public_api_roundtrip_p64(sg):
    """docs/src/index.md:14-39 through the mirrored API only (no oracle): keygen on the GPU, encrypt, split,
    bootstrap bits 10 and 20, decrypt."""
    rng = np.random.default_rng(5)
    P = sg.Params(64)
    sk = sg.PrivateKey(P, rng)
    bkey = sg.BootstrapKey(rng, sk)
    msg = rng.integers(0, 2, size=64, dtype=np.uint8)
    ct = sg.encrypt(sk, rng, msg)
    assert np.array_equal(sg.decrypt(sk, ct), msg.astype(bool))
    ebits = sg.split_ciphertext(ct)
    assert all(sg.decrypt(sk, e) == bool(b) for e, b in zip(ebits, msg))     # test/api.test.jl:33-42
    for r in (None, rng):
        a, o, x = sg.bootstrap(bkey, r, ebits[10], ebits[20])
        y1, y2 = bool(msg[10]), bool(msg[20])
        assert (sg.decrypt(sk, a), sg.decrypt(sk, o), sg.decrypt(sk, x)) == (y1 and y2, y1 or y2, y1 != y2)
    P.close()


@pytest.mark.parametrize("n", [128, 256, 512, 1024])
def test_bootstrap_trace_truncated_large(so, sg, n):
    """every supported transform shape (m = 1024 ... 8192; 86-bit Q at n = 1024): first steps of the loop against the
    oracle, both flatten modes"""
    P, OP = sg.Params(n), so.Params(n)
    steps = 3
    sk = so.make_secret(OP, 1)
    key = so.make_bkey(OP, sk, 1, rows=steps)
    bits, lwes = so.make_lwes(OP, sk, 1)
    bkey = sg.BootstrapKey(params=P, key=key)
    rng = np.random.default_rng(31)
    xmax = OP.B // 2 * 3
    for draws in (None, rng.integers(-xmax, xmax + 1, size=(steps, 2, OP.m, 2), dtype=np.int64)):
        ga, go, gx, gtr = sg.bootstrap_trace(bkey, draws, lwes[3], lwes[700 % n], n_steps=steps)
        ra, ro, rx, rtr = so.bootstrap_internal(OP, key, lwes[3], lwes[700 % n], draws=draws, n_steps=steps, trace=True)
        for k in range(steps):
            assert np.array_equal(gtr[k], rtr[k]), f"accumulator differs after step {k}"
        assert np.array_equal(ga, ra) and np.array_equal(go, ro) and np.array_equal(gx, rx)
    P.close()


@pytest.mark.parametrize("use_rng", [False, True])
def test_pack_encrypted_bits_p64(env64, so, sg, use_rng):
    """port of test/api.test.jl:86-108 (packing): pack -> split -> decrypt and pack -> decrypt both give the message;
    and the packed RLWE equals the oracle's pack_encrypted_bits (src/fhe.jl:660-696) bit for bit."""
    P, OP, sk, key, bits, lwes, bkey = env64

    class _Rng:                                       # hands the library pre-drawn values in the reference's order
        def __init__(self, arrays): self.arrays = list(arrays)
        def integers(self, lo, hi, size, dtype): return self.arrays.pop(0)

    xmax = OP.B // 2 * 3
    gen = np.random.default_rng(41)
    db = gen.integers(-xmax, xmax + 1, size=(OP.n, OP.n, 2, OP.m, 2), dtype=np.int64) if use_rng else None
    ds = gen.integers(-xmax, xmax + 1, size=(OP.n, OP.m, 2), dtype=np.int64) if use_rng else None
    ebits = [sg.EncryptedBit(sg.LWE(l[:-1], l[-1])) for l in lwes]
    ct = sg.pack_encrypted_bits(bkey, _Rng([db, ds]) if use_rng else None, ebits)
    rw, rv = so.pack_encrypted_bits(OP, key, lwes, db, ds)
    assert np.array_equal(ct.a, rw) and np.array_equal(ct.b, rv)
    key_obj = type("K", (), {"params": P, "key": sk})()
    assert np.array_equal(sg.decrypt(key_obj, ct), bits.astype(bool))
    assert [sg.decrypt(key_obj, e) for e in sg.split_ciphertext(ct)] == [bool(b) for b in bits]


def test_shortened_products_match_oracle_p1024(so, sg):
    """shortened_external_product at paper size (src/fhe.jl:632-641) against the oracle, both flatten modes"""
    import ctypes as C
    from sgfhe_jl_b200 import _lib
    P, OP = sg.Params(1024), so.Params(1024)
    rows = 2
    sk = so.make_secret(OP, 1)
    key = so.make_bkey(OP, sk, 1, rows=rows)
    bkey = sg.BootstrapKey(params=P, key=key)
    bkey.upload()
    rng = np.random.default_rng(51)
    polys = so.rand_below(rng, OP.Q, (rows, OP.m))
    xmax = OP.B // 2 * 3
    for draws in (None, rng.integers(-xmax, xmax + 1, size=(rows, OP.m, 2), dtype=np.int64)):
        out = np.zeros((rows, 2, OP.m, 2), np.uint64)
        _lib.check(_lib.lib().sgfhe_shortened_products(P.ctx, rows, polys.ctypes.data_as(C.c_void_p),
                                                       None if draws is None else draws.ctypes.data_as(C.c_void_p),
                                                       out.ctypes.data_as(C.c_void_p)))
        for i in range(rows):
            w, v = so.shortened_external_product(polys[i], key[i], OP.B, OP.Q, None if draws is None else draws[i])
            assert np.array_equal(out[i, 0], w) and np.array_equal(out[i, 1], v)
    P.close()


def test_chained_layers_match_oracle_p64(env64, so, sg):
    """examples/depth.jl:63-78 pattern: 4 layers of 6 gates with (AND, XOR) fed back in, ciphertexts resident on the
    device; every layer equals the oracle's bootstrap() and decrypts to the plaintext circuit."""
    P, OP, sk, key, bits, lwes, bkey = env64
    W, layers = 6, 4
    l1, l2 = lwes[:W].copy(), lwes[W:2 * W].copy()
    y1, y2 = bits[:W].astype(int), bits[W:2 * W].astype(int)
    got = sg.bootstrap_chain(bkey, l1, l2, layers, keep_layers=True)
    for layer in range(layers):
        ref = so.bootstrap_batch(OP, key, l1, l2, literal=False, threads=4)
        for g, r in zip(got[layer], ref):
            assert np.array_equal(g, r), f"layer {layer}"
        want = (y1 & y2, y1 | y2, y1 ^ y2)
        for arr, w in zip(ref, want):
            assert [so.decrypt_lwe(OP, sk, arr[i]) for i in range(W)] == w.tolist()
        l1, l2, y1, y2 = ref[0], ref[2], want[0], want[2]
    last = sg.bootstrap_chain(bkey, lwes[:W], lwes[W:2 * W], layers)
    assert all(np.array_equal(a, b) for a, b in zip(last, got[-1]))


@pytest.mark.parametrize("k", [1, 3, 5])
def test_rns2_arithmetic_matches_oracle(so, sg, k):
    """RNS2Number * + - (src/rns.jl:51-60) with the moduli of Scheme2.Params(k): random and edge operands"""
    S2 = sg.Scheme2Params(k)
    rng = np.random.default_rng(60 + k)
    N = 1 << 16
    a1 = rng.integers(0, S2.B, size=N, dtype=np.uint64); a2 = rng.integers(0, S2.Bp, size=N, dtype=np.uint64)
    b1 = rng.integers(0, S2.B, size=N, dtype=np.uint64); b2 = rng.integers(0, S2.Bp, size=N, dtype=np.uint64)
    for arr, M in ((a1, S2.B), (a2, S2.Bp), (b1, S2.B), (b2, S2.Bp)):
        arr[:4] = (0, 1, M - 1, M - 2)
    b1[:2] = (S2.B - 1, S2.B - 1); b2[:2] = (S2.Bp - 1, S2.Bp - 1)
    for code, op in enumerate("*+-"):
        g1, g2 = sg.rns2_op(op, (a1, a2), (b1, b2), S2.B, S2.Bp)
        r1, r2 = so.rns2_op(code, a1, a2, b1, b2, S2.B, S2.Bp)
        assert np.array_equal(g1, r1) and np.array_equal(g2, r2), op


def test_error_behaviour_and_empty_batch(sg, so):
    """boundary errors: no key, truncated key, bad shapes, empty batch"""
    P, OP = sg.Params(64), so.Params(64)
    sk = so.make_secret(OP, 0)
    _, lwes = so.make_lwes(OP, sk, 0)
    part = sg.BootstrapKey(params=P, key=so.make_bkey(OP, sk, 0, rows=3))
    with pytest.raises(sg.SgfheError, match="no complete bootstrap key"):
        sg.bootstrap_batch(part, None, lwes[:2], lwes[2:4])            # only 3 of n rows uploaded
    with pytest.raises(sg.SgfheError):
        sg.bootstrap_trace(part, None, lwes[0], lwes[1], n_steps=4)    # more steps than rows
    with pytest.raises(sg.SgfheError, match=r"\[batch, n\+1\]"):
        sg.bootstrap_batch(part, None, lwes[:2, :-1], lwes[2:4, :-1])
    full = sg.BootstrapKey(params=P, key=so.make_bkey(OP, sk, 0))
    outs = sg.bootstrap_batch(full, None, lwes[:0], lwes[:0])
    assert all(o.shape == (0, OP.n + 1) for o in outs)
    P.close()


def test_full_gates_paper_size(so, sg):
    """Params(1024), full n = 1024 steps with a real key: 2 gates equal the oracle bit for bit, 12 more decrypt to
    the plaintext gates (the 4096-gate batch of BASELINE.json is checked the same way inside bench.py)."""
    P, OP = sg.Params(1024), so.Params(1024)
    so.set_setup_threads(16)
    sk = so.make_secret(OP, 1)
    key = so.make_bkey(OP, sk, 1)
    bits, lwes = so.make_lwes(OP, sk, 1)
    bkey = sg.BootstrapKey(params=P, key=key)
    G = 14
    l1, l2 = lwes[:G], lwes[G:2 * G]
    outs = sg.bootstrap_batch(bkey, None, l1, l2)
    ref = so.bootstrap_batch(OP, key, l1[:2], l2[:2], literal=False, threads=2)
    for o, r in zip(outs, ref):
        assert np.array_equal(o[:2], r)
    for g in range(G):
        y1, y2 = int(bits[g]), int(bits[G + g])
        assert tuple(so.decrypt_lwe(OP, sk, o[g]) for o in outs) == (y1 & y2, y1 | y2, y1 ^ y2)
    # stress: eight waves of gates per SM, twice.  The persistent CTAs drift apart and every warp runs at its own pace;
    # a missing barrier shows up as run-to-run differences (this caught a top-stage twiddle slot published without one).
    W = 1184
    reps = (2 * W + len(lwes) - 1) // len(lwes)
    big, bb = np.concatenate([lwes] * reps)[: 2 * W], np.concatenate([bits] * reps)[: 2 * W]
    o1 = sg.bootstrap_batch(bkey, None, big[:W], big[W:])
    o2 = sg.bootstrap_batch(bkey, None, big[:W], big[W:])
    skb = np.asarray(sk, dtype=bool)
    y1, y2 = bb[:W].astype(np.int64), bb[W:].astype(np.int64)
    for a, b, want in zip(o1, o2, (y1 & y2, y1 | y2, y1 ^ y2)):
        assert np.array_equal(a, b)
        b1 = (a[:, OP.n].astype(np.int64) - a[:, :OP.n][:, skb].astype(np.int64).sum(axis=1)) % OP.r
        assert np.array_equal(((b1 + OP.Dr // 2) % OP.r) // OP.Dr, want)
    P.close()


def test_split_and_decrypt_on_device(env64, sg):
    """split_ciphertext / extract (src/fhe.jl:237-244, 287-290) and decrypt(::EncryptedBit)
    (src/fhe.jl:504-507) on the GPU equal the host mirror, for packed ciphertexts (N = n) and for the length-m ciphertext
    of pack_encrypted_bits (N = m, wrap-around taken from the tail of the long polynomial)."""
    P, OP, sk, key, bits, lwes, bkey = env64
    rng = np.random.default_rng(11)
    skey = sg.PrivateKey(P, rng)
    msgs = [rng.integers(0, 2, size=P.n, dtype=np.uint8) for _ in range(3)]
    cts = [sg.encrypt(skey, rng, m_) for m_ in msgs]
    got = sg.split_ciphertexts(cts)
    want = np.stack([e.lwe.flat() for c in cts for e in sg.split_ciphertext(c)])
    assert np.array_equal(got, want)
    dec = sg.decrypt_bits(skey, got)
    assert np.array_equal(dec, np.concatenate(msgs).astype(bool))
    assert [bool(x) for x in dec[:8]] == [sg.decrypt(skey, sg.EncryptedBit(sg.LWE(r_[:-1], r_[-1]))) for r_ in got[:8]]
    long_ct = sg.Ciphertext(P, rng.integers(0, P.r, size=P.m, dtype=np.uint64), rng.integers(0, P.r, size=P.m, dtype=np.uint64))
    got2 = sg.split_ciphertexts([long_ct])
    want2 = np.stack([e.lwe.flat() for e in sg.split_ciphertext(long_ct)])
    assert np.array_equal(got2, want2)
    with pytest.raises(sg.SgfheError):
        sg.decrypt_bits(skey, got[:, :-1])


def test_transformed_key_roundtrip(env64, sg, tmp_path):
    """serialise the pre-transformed key, load it into a fresh context, same ciphertexts out; wrong parameters rejected"""
    P, OP, sk, key, bits, lwes, bkey = env64
    ref = sg.bootstrap_batch(bkey, None, lwes[:3], lwes[3:6])
    path = str(tmp_path / "key.sgk")
    bkey.save_transformed(path)
    P2 = sg.Params(64)
    bk2 = sg.BootstrapKey.load_transformed(P2, path)
    got = sg.bootstrap_batch(bk2, None, lwes[:3], lwes[3:6])
    assert all(np.array_equal(a, b) for a, b in zip(ref, got))
    P3 = sg.Params(128)
    with pytest.raises(sg.SgfheError, match="other parameters"):
        sg.BootstrapKey.load_transformed(P3, path)
    P2.close(); P3.close()
