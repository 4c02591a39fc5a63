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


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048])
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


@pytest.mark.parametrize("n", [64, 512, 1024, 2048])
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


@pytest.mark.parametrize("n", [64, 1024, 2048])
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


@pytest.mark.parametrize("n", [128, 256, 512, 1024, 2048])
def test_bootstrap_trace_truncated_large(so, sg, n):
    """every supported transform shape (m = 1024 ... 16384; 87-bit Q at n = 1024, 93-bit Q and six primes at n = 2048, the
    largest n whose Q fits 96 bits): first steps of the loop against the oracle, both flatten modes"""
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
    # a blob of an older format (before the CRT pre-scaling moved into the key words: version 2) must not load
    blob = bytearray(open(path, "rb").read())
    assert int.from_bytes(blob[4:8], "little") == 3
    blob[4:8] = (2).to_bytes(4, "little")
    old = str(tmp_path / "old.sgk")
    open(old, "wb").write(bytes(blob))
    with pytest.raises(sg.SgfheError, match="not a serialised sgfhe key"):
        sg.BootstrapKey.load_transformed(P2, old)
    P2.close(); P3.close()


# ---- round 2: holes named by the round-1 review ----------------------------------------------------------------------
def _key_randomness(so, OP, seed, rows):
    """the draws of BootstrapKey (src/fhe.jl:193-194) exactly as so.make_bkey makes them"""
    a_rand = np.zeros((rows, 4, OP.m, 2), np.uint64)
    e_rand = np.zeros((rows, 4, OP.m), np.int64)
    for i in range(rows):
        rng = np.random.default_rng([seed, 2, i])
        a_rand[i] = so.rand_below(rng, OP.Q, (4, OP.m))
        e_rand[i] = rng.integers(-OP.n, OP.n + 1, size=(4, OP.m), dtype=np.int64)
    return a_rand, e_rand


def _generate_on_device(sg, P, sk, a_rand, e_rand, chunks=1):
    import ctypes as C
    from sgfhe_jl_b200 import _lib
    rows = a_rand.shape[0]
    out = np.zeros((rows, 4, 2, P.m, 2), np.uint64)
    skb = np.ascontiguousarray(sk, np.uint8)
    step = -(-rows // chunks)
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        a, e, o = np.ascontiguousarray(a_rand[r0:r1]), np.ascontiguousarray(e_rand[r0:r1]), out[r0:r1]
        _lib.check(_lib.lib().sgfhe_bkey_generate(P.ctx, skb.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                                                  e.ctypes.data_as(C.c_void_p), r0, r1 - r0, o.ctypes.data_as(C.c_void_p)))
    return out


@pytest.mark.parametrize("n,rows,chunks", [(64, 64, 3), (128, 5, 1), (512, 3, 1), (1024, 2, 2)])
def test_bkey_generate_matches_oracle(so, sg, n, rows, chunks):
    """BootstrapKey(rng, sk) (src/fhe.jl:181-201) on the device against sgo_bkey_generate on the SAME pre-drawn a_j, e_j:
    the coefficient form is bit-equal, and the transform-domain key the device kept drives the loop to the oracle's
    accumulators (so the key never has to visit the host)."""
    P, OP = sg.Params(n), so.Params(n)
    sk = so.make_secret(OP, 7)
    sk[:4] = (1, 0, 1, 1)                                             # both branches of + s_i G in the first rows
    a_rand, e_rand = _key_randomness(so, OP, 7, rows)
    e_rand[0, 0, :3] = (-n, n, 0); a_rand[0, 1, 0] = so.pack([OP.Q - 1])[0]      # edges: extreme errors, wrap at + B
    ref = so.bkey_generate(OP, sk, np.concatenate([a_rand, np.zeros((n - rows,) + a_rand.shape[1:], np.uint64)]),
                           np.concatenate([e_rand, np.zeros((n - rows,) + e_rand.shape[1:], np.int64)]), 0, rows)
    got = _generate_on_device(sg, P, sk, a_rand, e_rand, chunks)
    assert np.array_equal(got, ref)
    bits, lwes = so.make_lwes(OP, sk, 7)
    bk = sg.BootstrapKey.resident(P)                                  # the device copy sgfhe_bkey_generate left behind
    assert bk.rows == rows
    steps = min(rows, 3)
    ga, go, gx, gtr = sg.bootstrap_trace(bk, None, lwes[1], lwes[2], n_steps=steps)
    ra, ro, rx, rtr = so.bootstrap_internal(OP, ref, lwes[1], lwes[2], n_steps=steps, trace=True, fast=True)
    assert np.array_equal(gtr, rtr) and np.array_equal(ga, ra) and np.array_equal(go, ro) and np.array_equal(gx, rx)
    P.close()


def test_bkey_generate_then_gates_p64(so, sg):
    """the mirrored constructor BootstrapKey(rng, sk) end to end: gates on the device-generated key decrypt correctly and
    equal the oracle run on the coefficient form the device reports (keep_coefficients=True)"""
    P, OP = sg.Params(64), so.Params(64)
    rng = np.random.default_rng(77)
    sk = sg.PrivateKey(P, rng)
    bkey = sg.BootstrapKey(rng, sk, keep_coefficients=True)
    bits, lwes = so.make_lwes(OP, sk.key, 3)
    outs = sg.bootstrap_batch(bkey, None, lwes[:8], lwes[8:16])
    ref = so.bootstrap_batch(OP, bkey.key, lwes[:8], lwes[8:16], literal=False, threads=4)
    for o, r in zip(outs, ref):
        assert np.array_equal(o, r)
    for g in range(8):
        y1, y2 = int(bits[g]), int(bits[8 + g])
        assert tuple(so.decrypt_lwe(OP, sk.key, o[g]) for o in outs) == (y1 & y2, y1 | y2, y1 ^ y2)
    P.close()


def test_two_keys_on_one_params(so, sg):
    """a context holds one key: using a key object whose device copy was replaced re-uploads it (coefficient form kept) or
    raises (none kept) -- never runs silently with the other key (round-1 advisor finding)"""
    P, OP = sg.Params(64), so.Params(64)
    sk = so.make_secret(OP, 0)
    _, lwes = so.make_lwes(OP, sk, 0)
    k1, k2 = so.make_bkey(OP, sk, 0), so.make_bkey(OP, sk, 5)
    b1, b2 = sg.BootstrapKey(params=P, key=k1), sg.BootstrapKey(params=P, key=k2)
    o1 = sg.bootstrap_batch(b1, None, lwes[:2], lwes[2:4])
    o2 = sg.bootstrap_batch(b2, None, lwes[:2], lwes[2:4])
    o1b = sg.bootstrap_batch(b1, None, lwes[:2], lwes[2:4])          # b2 replaced b1 on the device in between
    assert all(np.array_equal(a, b) for a, b in zip(o1, o1b))
    assert not all(np.array_equal(a, b) for a, b in zip(o1, o2))
    for o, k in ((o1, k1), (o2, k2)):
        ref = so.bootstrap(OP, k, lwes[0], lwes[2])
        assert all(np.array_equal(x[0], r) for x, r in zip(o, ref))
    rng = np.random.default_rng(3)
    skg = sg.PrivateKey(P, rng)
    g = sg.BootstrapKey(rng, skg)                                    # device-generated, no coefficient form kept
    sg.bootstrap_batch(b1, None, lwes[:1], lwes[1:2])
    with pytest.raises(sg.SgfheError, match="replaced by another key"):
        sg.bootstrap_batch(g, None, lwes[:1], lwes[1:2])
    P.close()


def test_split_n_equals_m_and_decrypt_match_oracle(env64, so, sg):
    """split_ciphertext of a length-m Ciphertext (src/fhe.jl:287-290 with extract's wrap-around branch, :237-244) against
    sgo_split_rlwe, and decrypt(::EncryptedBit) (:504-507) against sgo_decrypt_lwe, row by row"""
    P, OP, sk, key, bits, lwes, bkey = env64
    rng = np.random.default_rng(12)
    for N in (OP.n, OP.m):
        a = rng.integers(0, OP.r, size=N, dtype=np.uint64)
        b = rng.integers(0, OP.r, size=N, dtype=np.uint64)
        a[:3] = (0, OP.r - 1, 1); a[-2:] = (0, OP.r - 1)             # zero stays zero under the negated wrap-around
        ct = sg.Ciphertext(P, a, b) if N == OP.m else sg.PackedCiphertext(P, a, b)
        got = sg.split_ciphertexts([ct])
        assert np.array_equal(got, so.split_rlwe(OP, a, b))
    key_obj = type("K", (), {"params": P, "key": sk})()
    dec = sg.decrypt_bits(key_obj, lwes)
    assert dec.tolist() == [bool(so.decrypt_lwe(OP, sk, l)) for l in lwes]
    assert dec.tolist() == [bool(b_) for b_ in bits]


@pytest.fixture(scope="module")
def env1024(so, sg):
    """one full paper-size key (1 GiB) shared by the tests below"""
    OP = so.Params(1024)
    so.set_setup_threads(16)
    sk = so.make_secret(OP, 1)
    key = so.make_bkey(OP, sk, 1)
    bits, lwes = so.make_lwes(OP, sk, 1)
    return OP, sk, key, bits, lwes


def test_full_randomised_gate_p1024(env1024, so, sg):
    """one FULL gate at Params(1024) with flatten(rng, ...) (src/utils.jl:198-241): 1024 x 2 x 8192 x 2 host-supplied draws
    (268 MB, including the extreme draws +-xmax), outputs over Z_Q and over Z_r equal the oracle's"""
    import ctypes as C
    from sgfhe_jl_b200 import _lib
    OP, sk, key, bits, lwes = env1024
    P = sg.Params(1024)
    bkey = sg.BootstrapKey(params=P, key=key)
    bkey.upload()
    rng = np.random.default_rng(91)
    xmax = OP.B // 2 * 3
    draws = rng.integers(-xmax, xmax + 1, size=(1, OP.n, 2, OP.m, 2), dtype=np.int64)
    draws[0, 0, 0, :4] = ((xmax, xmax), (-xmax, -xmax), (xmax, -xmax), (0, 0))
    draws[0, 500, 1, :] = xmax; draws[0, 501, 0, :] = -xmax           # whole steps at the extremes
    l1, l2 = lwes[5:6], lwes[900:901]
    raw = [np.zeros((1, OP.n + 1, 2), np.uint64) for _ in range(3)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(_lib.lib().sgfhe_bootstrap_internal_batch(P.ctx, 1, p(l1), p(l2), p(draws), *[p(o) for o in raw]))
    ref = so.bootstrap_internal(OP, key, l1[0], l2[0], draws=draws[0], fast=True)
    for g, r in zip(raw, ref):
        assert np.array_equal(g[0], r)
    outs = [np.zeros((1, OP.n + 1), np.uint64) for _ in range(3)]
    _lib.check(_lib.lib().sgfhe_bootstrap_batch(P.ctx, 1, p(l1), p(l2), p(draws), *[p(o) for o in outs]))
    y1, y2 = int(bits[5]), int(bits[900])
    for o, r, want in zip(outs, ref, (y1 & y2, y1 | y2, y1 ^ y2)):
        assert o[0].tolist() == [so.rescale(OP.r, v, OP.Q, True) for v in so.unpack(r)]      # reduce_modulus, src/fhe.jl:616-618
        assert so.decrypt_lwe(OP, sk, o[0]) == want
    P.close()


@pytest.mark.parametrize("n", [128, 256, 512])
def test_full_deterministic_gate_mid_sizes(so, sg, n):
    """one full gate (all n steps) at every transform shape between the test size and the paper size vs so.bootstrap"""
    P, OP = sg.Params(n), so.Params(n)
    so.set_setup_threads(16)
    sk = so.make_secret(OP, 2)
    key = so.make_bkey(OP, sk, 2)
    bits, lwes = so.make_lwes(OP, sk, 2)
    bkey = sg.BootstrapKey(params=P, key=key)
    outs = sg.bootstrap_batch(bkey, None, lwes[:3], lwes[3:6])
    ref = so.bootstrap_batch(OP, key, lwes[:3], lwes[3:6], literal=False, threads=3)
    for o, r in zip(outs, ref):
        assert np.array_equal(o, r)
    for g in range(3):
        y1, y2 = int(bits[g]), int(bits[3 + g])
        assert tuple(so.decrypt_lwe(OP, sk, o[g]) for o in outs) == (y1 & y2, y1 | y2, y1 ^ y2)
    if n == 512:                                                      # examples/depth.jl runs at Params(512): chained layers
        W, layers = 3, 3
        got = sg.bootstrap_chain(bkey, lwes[:W], lwes[W:2 * W], layers, keep_layers=True)
        l1, l2 = lwes[:W].copy(), lwes[W:2 * W].copy()
        for layer in range(layers):
            r = so.bootstrap_batch(OP, key, l1, l2, literal=False, threads=3)
            for g_, r_ in zip(got[layer], r):
                assert np.array_equal(g_, r_), f"layer {layer}"
            l1, l2 = r[0], r[2]
    P.close()


@pytest.mark.parametrize("n", [64, 1024, 2048])
def test_worst_case_magnitudes(so, sg, n):
    """the approximate CRT (v from the top bits of each residue) and the FP64-assisted Barrett step at the edge of their
    bounds: every draw +-xmax, every key coefficient floor(Q/2) or floor(Q/2)+1 (centred +-Q/2), accumulator all Q-1 or
    chosen so that every deterministic digit is at its extreme -- |sum digit * key| reaches 4m * 2B * Q/2."""
    P, OP = sg.Params(n), so.Params(n)
    Q, B, m = OP.Q, OP.B, OP.m
    xmax = B // 2 * 3
    half = so.pack([Q // 2])[0]; half1 = so.pack([Q // 2 + 1])[0]
    full = np.broadcast_to(so.pack([Q - 1]), (m, 2)).copy()
    s = B // 2 - 1
    # a with both deterministic digits at +B/2 (the largest digit): a + s (1 + B) = (B - 1) + (B - 1) B
    top = np.broadcast_to(so.pack([((B - 1) + (B - 1) * B - s * (1 + B)) % Q]), (m, 2)).copy()
    cases = []
    for kv in (half, half1):
        A = np.broadcast_to(kv, (4, 2, m, 2)).copy()
        for sign in (1, -1):
            cases.append((full, full, A, np.full((2, m, 2), sign * xmax, np.int64)))
        cases.append((top, top, A, None))
        cases.append((full, top, A, None))
    alt = np.broadcast_to(half, (4, 2, m, 2)).copy(); alt[:, :, 1::2] = half1        # alternating signs: cancellation pattern
    d_alt = np.full((2, m, 2), xmax, np.int64); d_alt[:, 1::2] = -xmax
    cases.append((full, full, alt, d_alt))
    for a, b, A, draws in cases:
        oa, ob = sg.external_product(P, draws, a, b, A)
        ra, rb = so.external_product(a, b, A, B, Q, draws)
        assert np.array_equal(oa, ra) and np.array_equal(ob, rb)
    P.close()


@pytest.mark.parametrize("name", ["golden_p64.npz", "golden_p1024_trunc.npz", "golden_p512_trunc.npz", "golden_p2048_trunc.npz"])
def test_gpu_reproduces_golden(so, sg, name):
    """the committed fixtures (tests/golden/, generator make_golden.py): the GPU path reproduces every stored output and
    the hash of every accumulator state"""
    import hashlib, os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    n, seed, steps = int(g["n"]), int(g["seed"]), int(g["steps"])
    P, OP = sg.Params(n), so.Params(n)
    key = so.make_bkey(OP, g["sk"], seed, rows=steps)
    assert sha(key) == str(g["key_sha256"])
    bkey = sg.BootstrapKey(params=P, key=key)
    xmax = OP.B // 2 * 3
    rng = np.random.default_rng([seed, 9])
    for pi in range(len(g["pairs"])):
        for mode in ("det", "rnd"):
            tag = f"p{pi}_{mode}"
            if tag + "_lwe1" not in g:
                continue
            draws = rng.integers(-xmax, xmax + 1, size=(steps, 2, OP.m, 2), dtype=np.int64) if mode == "rnd" else None
            a, o, x, tr = sg.bootstrap_trace(bkey, draws, g[tag + "_lwe1"], g[tag + "_lwe2"], n_steps=steps)
            assert np.array_equal(a, g[tag + "_and_Q"]) and np.array_equal(o, g[tag + "_or_Q"]) and np.array_equal(x, g[tag + "_xor_Q"])
            assert [sha(tr[k]) for k in range(steps)] == [str(s_) for s_ in g[tag + "_trace_sha256"]]
            if tag + "_and_r" in g and mode == "det":
                outs = sg.bootstrap_batch(bkey, None, g[tag + "_lwe1"][None], g[tag + "_lwe2"][None])
                for got, want in zip(outs, (g[tag + "_and_r"], g[tag + "_or_r"], g[tag + "_xor_r"])):
                    assert np.array_equal(got[0], want)
    P.close()


def test_gpu_matches_bigint_model_p64(so, sg):
    """second, independent pin: the first two accumulation steps on the GPU against oracle/model.py (plain Python big
    integers, Kronecker-substitution products -- shares no code with the C oracle), both flatten modes"""
    import model as md
    P, OP, M = sg.Params(64), so.Params(64), md.params(64)
    sk = so.make_secret(OP, 4)
    steps = 2
    key = so.make_bkey(OP, sk, 4, rows=steps)
    _, lwes = so.make_lwes(OP, sk, 4)
    bkey = sg.BootstrapKey(params=P, key=key)
    rng = np.random.default_rng(5)
    xmax = OP.B // 2 * 3
    for draws in (None, rng.integers(-xmax, xmax + 1, size=(steps, 2, OP.m, 2), dtype=np.int64)):
        ga, go, gx, gtr = sg.bootstrap_trace(bkey, draws, lwes[7], lwes[9], n_steps=steps)
        mtr = []
        ma, mo, mx = md.bootstrap_internal(M, so.unpack(key), lwes[7].tolist(), lwes[9].tolist(),
                                           None if draws is None else draws.tolist(), n_steps=steps, trace=mtr)
        for k in range(steps):
            assert so.unpack(gtr[k, 0]) == mtr[k][0] and so.unpack(gtr[k, 1]) == mtr[k][1]
        assert so.unpack(ga) == ma and so.unpack(go) == mo and so.unpack(gx) == mx
    P.close()


def test_pack_tail_p1024(env1024, so, sg):
    """pack_encrypted_bits' stages after the n bootstraps at paper size, on the device (sgfhe_pack_from_lwes): the
    transposition, the n shortened products, the two sums, negate / subtract and ModRed of src/fhe.jl:675-693 from given
    pre-ModRed LWEs equal the oracle's pack_from_lwes, both flatten modes (the bootstraps in front of it are covered at
    Params(64) and by the gate tests)"""
    import ctypes as C
    from sgfhe_jl_b200 import _lib
    OP, sk, key, bits, lwes = env1024
    P = sg.Params(1024)
    bkey = sg.BootstrapKey(params=P, key=key)
    bkey.upload()
    rng = np.random.default_rng(17)
    new_lwes = so.rand_below(rng, OP.Q, (OP.n, OP.n + 1))
    xmax = OP.B // 2 * 3
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    for ds in (None, rng.integers(-xmax, xmax + 1, size=(OP.n, OP.m, 2), dtype=np.int64)):
        w, v = np.zeros(OP.m, np.uint64), np.zeros(OP.m, np.uint64)
        _lib.check(_lib.lib().sgfhe_pack_from_lwes(P.ctx, p(new_lwes), p(ds), p(w), p(v)))
        rw, rv = so.pack_from_lwes(OP, key, new_lwes, ds)
        assert np.array_equal(w, rw) and np.array_equal(v, rv)
    P.close()


# ---- Scheme 2 (src/fhe2.jl, src/rns.jl): what exists upstream -- there is no reference bootstrap for it -----------------
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_scheme2_polymul_matches_oracle(so, sg, k):
    """Polynomial{RNS2Number} `*` at the ring degree of Scheme2.Params(k) (m = 2048 .. 32768; one- and two-pass
    transforms): random and edge operands against the oracle's limb-wise product"""
    ctx = sg.Scheme2Context(k)
    S = ctx.params
    rng = np.random.default_rng(70 + k)
    nb = 3
    a1 = rng.integers(0, S.B, size=(nb, S.m), dtype=np.uint64); a2 = rng.integers(0, S.Bp, size=(nb, S.m), dtype=np.uint64)
    b1 = rng.integers(0, S.B, size=(nb, S.m), dtype=np.uint64); b2 = rng.integers(0, S.Bp, size=(nb, S.m), dtype=np.uint64)
    a1[1] = S.B - 1; a2[1] = S.Bp - 1; b1[1] = S.B - 1; b2[1] = S.Bp - 1        # all coefficients at the maximum
    a1[2] = 0; a2[2] = 0; a1[2, S.m - 1] = 1; a2[2, S.m - 1] = 1                # x^(m-1): a rotation with wrap-around sign
    o1, o2 = ctx.polymul((a1, a2), (b1, b2))
    for i in range(nb):
        r1, r2 = so.rns2_polymul(a1[i], a2[i], b1[i], b2[i], S.B, S.Bp)
        assert np.array_equal(o1[i], r1) and np.array_equal(o2[i], r2), f"product {i}"
    ctx.close()


@pytest.mark.parametrize("k", [1, 4])
def test_scheme2_transform_and_mac(so, sg, k):
    """forward + inverse transform is the identity; the 8 multiply-accumulates of an external product done in the transform
    domain and transformed back equal the sum of the oracle's products (src/fhe.jl:527-528 shape over RNS2Number)"""
    import ctypes as C
    import torch
    from sgfhe_jl_b200 import _lib
    ctx = sg.Scheme2Context(k)
    S, L = ctx.params, _lib.lib()
    rng = np.random.default_rng(80 + k)
    mods = (S.B, S.Bp)
    d = [rng.integers(0, M, size=(1, 4, S.m), dtype=np.uint64) for M in mods]           # digit polynomials
    K = [rng.integers(0, M, size=(1, 4, 2, S.m), dtype=np.uint64) for M in mods]        # key tile
    dev = lambda a: torch.from_numpy(a.view(np.int64).copy()).cuda()
    td, tK = [dev(x) for x in d], [dev(x) for x in K]
    t0 = [x.clone() for x in td]
    _lib.check(L.sgfhe_s2_ntt_device(ctx._h, 0, 4, td[0].data_ptr(), td[1].data_ptr(), None))
    back = [x.clone() for x in td]
    _lib.check(L.sgfhe_s2_ntt_device(ctx._h, 1, 4, back[0].data_ptr(), back[1].data_ptr(), None))
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(back, t0))
    _lib.check(L.sgfhe_s2_ntt_device(ctx._h, 0, 8, tK[0].data_ptr(), tK[1].data_ptr(), None))
    out = [torch.zeros((1, 2, S.m), dtype=torch.int64, device="cuda") for _ in mods]
    _lib.check(L.sgfhe_s2_mac8_device(ctx._h, 1, td[0].data_ptr(), td[1].data_ptr(), tK[0].data_ptr(), tK[1].data_ptr(),
                                      out[0].data_ptr(), out[1].data_ptr(), None))
    _lib.check(L.sgfhe_s2_ntt_device(ctx._h, 1, 2, out[0].data_ptr(), out[1].data_ptr(), None))
    torch.cuda.synchronize()
    got = [o.cpu().numpy().view(np.uint64) for o in out]
    for c in range(2):
        acc = [np.zeros(S.m, object), np.zeros(S.m, object)]
        for j in range(4):
            r = so.rns2_polymul(d[0][0, j], d[1][0, j], K[0][0, j, c], K[1][0, j, c], S.B, S.Bp)
            for l in range(2):
                acc[l] = (acc[l] + r[l].astype(object)) % mods[l]
        for l in range(2):
            assert got[l][0, c].tolist() == acc[l].tolist()
    ctx.close()


@pytest.mark.parametrize("k,rows", [(1, 3), (3, 2)])
def test_scheme2_bootstrap_key_matches_oracle(so, sg, k, rows):
    """Scheme2.BootstrapKey (src/fhe2.jl:104-131) generated on the device from pre-drawn a_j, e_j against the oracle"""
    ctx = sg.Scheme2Context(k)
    S = ctx.params
    rng = np.random.default_rng(90 + k)
    sk = rng.integers(0, 2, size=S.n, dtype=np.uint8)
    sk[:3] = (1, 0, 1)
    a = so.rand_below(rng, S.B * S.Bp, (rows, 4, S.m))
    e = rng.integers(-S.tau, S.tau + 1, size=(rows, 4, S.m), dtype=np.int64)
    e[0, 0, :2] = (-S.tau, S.tau)
    got = ctx.bootstrap_key(sk, a, e)
    assert np.array_equal(got, so.scheme2_bkey_generate(k, sk, a, e, rows))
    ctx.close()


# ---- randomised flatten with draws made on the device (round-1 advisor finding: host draws only fit toy sizes) ---------
def _philox4x32_10(c, k):
    """numpy model of the device generator (Salmon et al., SC'11): c uint32[4, N] counters, k uint32[2] key"""
    c = [x.astype(np.uint64) for x in c]
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & mask, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & mask]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & mask, (k1 + np.uint64(0xBB67AE85)) & mask
    return c


def _device_draws(sg, P, seed, gate, step0, steps):
    import ctypes as C
    from sgfhe_jl_b200 import _lib
    out = np.zeros((steps, 2, P.m, 2), np.int64)
    _lib.check(_lib.lib().sgfhe_device_draws(P.ctx, seed, gate, step0, steps, out.ctypes.data_as(C.c_void_p)))
    return out


def test_device_rng_stream_is_philox(sg):
    """known-answer check of the counter-based generator and of the counter layout (coefficient, 2 step + polynomial, gate),
    the multiply-shift map onto [-xmax, xmax], and the range / spread of the draws"""
    kat = _philox4x32_10([np.zeros(1, np.uint32)] * 4, np.zeros(2, np.uint32))
    assert [int(x[0]) for x in kat] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]       # Random123 kat_vectors, philox4x32-10
    P = sg.Params(64)
    seed, gate, step0, steps = 0x1234567890ABCDEF, (5 << 32) + 77, 3, 2
    got = _device_draws(sg, P, seed, gate, step0, steps)
    xmax = P.B // 2 * 3
    j = np.tile(np.arange(P.m, dtype=np.uint32), steps * 2)
    sc = np.repeat(np.arange(steps * 2, dtype=np.uint32) + 2 * step0, P.m)                   # 2 step + polynomial
    r = _philox4x32_10([j, sc, np.full_like(j, gate & 0xFFFFFFFF), np.full_like(j, 0x53474648 ^ (gate >> 32))],
                       np.array([seed & 0xFFFFFFFF, seed >> 32], np.uint32))
    for d, (lo, hi) in enumerate(((r[0], r[1]), (r[2], r[3]))):
        v = [((int(a) | (int(b) << 32)) * (2 * xmax + 1) >> 64) - xmax for a, b in zip(lo, hi)]
        assert got.reshape(-1, 2)[:, d].tolist() == v
    assert np.abs(got).max() <= xmax and np.abs(got).max() > 0.99 * xmax and abs(float(got.mean())) < 0.05 * xmax
    P.close()


def test_device_rng_gates_match_oracle_p64(env64, so, sg):
    """bootstrap with a DeviceRng: every gate equals the oracle's bootstrap run on the draws the device generator makes for
    that gate (fetched through the seam), successive batches continue the gate counter, and every output decrypts"""
    P, OP, sk, key, bits, lwes, bkey = env64
    rng = sg.DeviceRng(20261018)
    l1, l2 = lwes[:5], lwes[32:37]
    first = sg.bootstrap_batch(bkey, rng, l1[:2], l2[:2])
    rest = sg.bootstrap_batch(bkey, rng, l1[2:], l2[2:])             # gates 2..4 of the stream
    outs = [np.concatenate([a, b]) for a, b in zip(first, rest)]
    det = sg.bootstrap_batch(bkey, None, l1, l2)
    assert not np.array_equal(outs[0], det[0])
    for g in range(5):
        draws = _device_draws(sg, P, rng.seed, g, 0, OP.n)
        ref = so.bootstrap(OP, key, l1[g], l2[g], draws)
        y1, y2 = int(bits[g]), int(bits[32 + g])
        for o, r_, want in zip(outs, ref, (y1 & y2, y1 | y2, y1 ^ y2)):
            assert np.array_equal(o[g], r_)
            assert so.decrypt_lwe(OP, sk, o[g]) == want
    again = sg.bootstrap_batch(bkey, sg.DeviceRng(20261018), l1, l2)
    assert all(np.array_equal(a, b) for a, b in zip(outs, again))


def test_device_rng_paper_size(env1024, so, sg):
    """the randomised mode at Params(1024) without host draws: a batch wider than one wave decrypts to the plaintext gates,
    is reproducible from the seed, and one gate equals the oracle on the device's own draws"""
    OP, sk, key, bits, lwes = env1024
    P = sg.Params(1024)
    bkey = sg.BootstrapKey(params=P, key=key)
    W = 160
    l1, l2 = lwes[:W], lwes[W:2 * W]
    o1 = sg.bootstrap_batch(bkey, sg.DeviceRng(99), l1, l2)
    o2 = sg.bootstrap_batch(bkey, sg.DeviceRng(99), l1, l2)
    assert all(np.array_equal(a, b) for a, b in zip(o1, o2))
    skb = np.asarray(sk, dtype=bool)
    y1, y2 = bits[:W].astype(np.int64), bits[W:2 * W].astype(np.int64)
    for a, want in zip(o1, (y1 & y2, y1 | y2, y1 ^ y2)):
        b1 = (a[:, OP.n].astype(np.int64) - a[:, :OP.n][:, skb].astype(np.int64).sum(axis=1)) % OP.r
        assert np.array_equal(((b1 + OP.Dr // 2) % OP.r) // OP.Dr, want)
    g = 151                                                           # a gate of the second wave
    ref = so.bootstrap_internal(OP, key, l1[g], l2[g], draws=_device_draws(sg, P, 99, g, 0, OP.n), fast=True)
    for o, r_ in zip(o1, ref):
        assert o[g].tolist() == [so.rescale(OP.r, v, OP.Q, True) for v in so.unpack(r_)]
    P.close()


def test_two_gates_per_sm_variant_p1024(env1024, so, sg, monkeypatch):
    """bootstrap_kernel_v5 (experimental, SGFHE_V5=1: 256-thread CTAs, two gates per SM, half the slices at a time) gives
    the default kernel's ciphertexts bit for bit, over more gates than resident CTAs, and the oracle's accumulators"""
    OP, sk, key, bits, lwes = env1024
    P4 = sg.Params(1024)
    k4 = sg.BootstrapKey(params=P4, key=key)
    W = 310
    l1, l2 = lwes[:W], lwes[W:2 * W]
    ref = sg.bootstrap_batch(k4, None, l1, l2)
    P4.close()
    monkeypatch.setenv("SGFHE_V5", "1")
    P5 = sg.Params(1024)
    k5 = sg.BootstrapKey(params=P5, key=key)
    got = sg.bootstrap_batch(k5, None, l1, l2)
    assert all(np.array_equal(a, b) for a, b in zip(ref, got))
    rng = np.random.default_rng(8)
    xmax = OP.B // 2 * 3
    for draws in (None, rng.integers(-xmax, xmax + 1, size=(3, 2, OP.m, 2), dtype=np.int64)):
        ga, go, gx, gtr = sg.bootstrap_trace(k5, draws, lwes[3], lwes[700], n_steps=3)
        ra, ro, rx, rtr = so.bootstrap_internal(OP, key[:3], lwes[3], lwes[700], draws=draws, n_steps=3, trace=True, fast=True)
        assert np.array_equal(gtr, rtr) and np.array_equal(ga, ra) and np.array_equal(go, ro) and np.array_equal(gx, rx)
    P5.close()


@pytest.mark.parametrize("n", [512, 1024])
@pytest.mark.parametrize("switch", ["SGFHE_HEAD_INT", "SGFHE_FORCE_V3"])
def test_selectable_kernel_variants_match_oracle(so, sg, monkeypatch, switch, n):
    """the A/B variants that stay selectable in the library -- the all-integer head of the v4 step (SGFHE_HEAD_INT) and the
    generic one-polynomial-per-thread-group step at m >= 4096 (SGFHE_FORCE_V3) -- reach the oracle's accumulators too"""
    monkeypatch.setenv(switch, "1")
    P, OP = sg.Params(n), so.Params(n)
    steps = 3
    sk = so.make_secret(OP, 5)
    key = so.make_bkey(OP, sk, 5, rows=steps)
    _, lwes = so.make_lwes(OP, sk, 5)
    bkey = sg.BootstrapKey(params=P, key=key)
    rng = np.random.default_rng(32)
    xmax = OP.B // 2 * 3
    for draws in (None, rng.integers(-xmax, xmax + 1, size=(steps, 2, OP.m, 2), dtype=np.int64)):
        ga, go, gx, gtr = sg.bootstrap_trace(bkey, draws, lwes[7], lwes[9], n_steps=steps)
        ra, ro, rx, rtr = so.bootstrap_internal(OP, key, lwes[7], lwes[9], draws=draws, n_steps=steps, trace=True, fast=True)
        assert np.array_equal(gtr, rtr) and np.array_equal(ga, ra) and np.array_equal(go, ro) and np.array_equal(gx, rx)
    P.close()
