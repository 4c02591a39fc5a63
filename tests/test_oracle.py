"""CPU tests of the oracle (test infrastructure): the reference's own test properties ported one for one
(test/internals.test.jl, test/api.test.jl), the C oracle against the independent big-integer model, the
literal against the rewritten loop, and the committed golden vectors."""
import hashlib
import os

import numpy as np
import pytest

import model as md

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_params_match_survey_table(so):
    """SURVEY.md 8(d) table, derived by hand from src/fhe.jl:43-97"""
    P = so.Params(64)
    assert (P.r, P.m, P.q, P.Q, P.B) == (1024, 512, 65537, 5494391545392009217, 2348810240)
    P = so.Params(512)
    assert (P.r, P.m, P.q, P.Q, P.B) == (8192, 4096, 4205569, 1440321777275241790332929, 1202590842880)
    P = so.Params(1024)
    assert (P.r, P.m, P.q, P.Q, P.B) == (16384, 8192, 16801793, 92180593745615474572738561, 9620726743040)
    assert P.Q == 0x4c40000000000000154001 and P.Dr == 4096 and P.DQ == P.Q // 8


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048])
def test_params_oracle_equals_model(so, n):
    P, M = so.Params(n), md.params(n)
    assert (P.n, P.r, P.q, P.Q, P.t, P.m, P.B, P.Dr, P.Dq, P.DQ) == (M.n, M.r, M.q, M.Q, M.t, M.m, M.B, M.Dr, M.Dq, M.DQ)
    assert (P.Q - 1) % (2 * P.m) == 0 and (P.q - 1) % (2 * P.n) == 0


def test_params_rejects_bad_n(so):
    for n in (0, 32, 100, 96):
        with pytest.raises(ValueError):
            so.Params(n)                                   # src/fhe.jl:45-46


@pytest.mark.parametrize("odd_new_max", [False, True])
@pytest.mark.parametrize("round_result", [False, True])
def test_rescale(so, odd_new_max, round_result):
    """port of test/internals.test.jl:26-47 (exhaustive) against the BigInt model at :6-20"""
    old_max = 2 ** 12 + 1
    new_max = 2 ** 4 + 1 if odd_new_max else 2 ** 4
    for i in range(old_max):
        res = so.rescale(new_max, i, old_max, round_result)
        assert res == md.rescale_ref(new_max, i, old_max, round_result) == md.rescale(new_max, i, old_max, round_result)


def test_rescale_wide(so):
    P = so.Params(1024)
    rng = np.random.default_rng(3)
    for x in [0, 1, P.Q - 1, P.Q // 2, P.Q // 2 + 1] + [int(v) for v in so.unpack(so.rand_below(rng, P.Q, (200,)))]:
        assert so.rescale(P.r, x, P.Q, True) == md.rescale(P.r, x, P.Q, True)
        assert so.rescale(P.r, x, P.Q, False) == md.rescale(P.r, x, P.Q, False)


def decomposition_limits(B, q, use_rng):
    """test/internals.test.jl:50-66"""
    if use_rng:
        s = 2 * B
        return q - s, s
    s = (B - 1) // 2 if B & 1 else B // 2 - 1
    return q - s, B - s - 1


@pytest.mark.parametrize("use_rng", [False, True])
@pytest.mark.parametrize("l", [3, 4])
@pytest.mark.parametrize("B", [4, 5])
@pytest.mark.parametrize("odd_q", [False, True])
def test_flatten_exhaustive_small(so, use_rng, l, B, odd_q):
    """port of test/internals.test.jl:69-112 (ModUInt: q = B^l - 1; MgModUInt: forced odd)"""
    q = B ** l - 1
    if odd_q and q % 2 == 0:
        q -= 1
    lim_lo, lim_hi = decomposition_limits(B, q, use_rng)
    rng = np.random.default_rng(l * 10 + B)
    xmax = (B - 1) // 2 * 3 if B & 1 else B // 2 * 3
    for a in range(q):
        draws = rng.integers(-xmax, xmax + 1, size=l, dtype=np.int64) if use_rng else None
        d = so.flatten(a, B, l, q, draws)
        assert sum(x * B ** i for i, x in enumerate(d)) % q == a
        assert all(x <= lim_hi or x >= lim_lo for x in d)
        assert d == md.flatten(a, B, l, q, None if draws is None else draws.tolist())


@pytest.mark.parametrize("use_rng", [False, True])
def test_flatten_poly(so, use_rng):
    """port of test/internals.test.jl:115-141: N=64, B=2^30, l=2, q=B^2-1"""
    B, l = 2 ** 30, 2
    q = B ** l - 1
    rng = np.random.default_rng(4)
    a = so.rand_below(rng, q, (64,))
    xmax = B // 2 * 3
    draws = rng.integers(-xmax, xmax + 1, size=(64, l), dtype=np.int64) if use_rng else None
    u = so.unpack(so.flatten_poly(a, B, l, q, draws))
    lim_lo, lim_hi = decomposition_limits(B, q, use_rng)
    ai = so.unpack(a)
    for j in range(64):
        assert u[0][j] <= lim_hi or u[0][j] >= lim_lo
        assert u[1][j] <= lim_hi or u[1][j] >= lim_lo
        assert (u[0][j] + u[1][j] * B) % q == ai[j]
    assert u == md.flatten_poly(ai, B, l, q, None if draws is None else draws.tolist())


@pytest.mark.parametrize("use_rng", [False, True])
def test_external_product_with_gadget_is_identity(so, use_rng):
    """port of test/internals.test.jl:144-166 at the scheme's own modulus (an NTT exists, as in bootstrap)"""
    P = so.Params(64)
    rng = np.random.default_rng(8)
    a, b = so.rand_below(rng, P.Q, (P.m,)), so.rand_below(rng, P.Q, (P.m,))
    G = np.zeros((4, 2, P.m, 2), np.uint64)
    G[0, 0, 0, 0] = 1; G[1, 0, 0, 0] = P.B; G[2, 1, 0, 0] = 1; G[3, 1, 0, 0] = P.B
    xmax = P.B // 2 * 3
    draws = rng.integers(-xmax, xmax + 1, size=(2, P.m, 2), dtype=np.int64) if use_rng else None
    oa, ob = so.external_product(a, b, G, P.B, P.Q, draws)
    assert np.array_equal(oa, a) and np.array_equal(ob, b)


@pytest.mark.parametrize("N,Q", [(64, 5494391545392009217), (256, 92180593745615474572738561)])
def test_polymul_three_ways(so, N, Q):
    """NTT product == schoolbook product == Kronecker product (the ring fixes the result)"""
    assert (Q - 1) % (2 * N) == 0
    rng = np.random.default_rng(N)
    a, b = so.rand_below(rng, Q, (N,)), so.rand_below(rng, Q, (N,))
    r1, r2 = so.polymul(a, b, Q), so.polymul(a, b, Q, schoolbook=True)
    assert np.array_equal(r1, r2)
    assert so.unpack(r1) == md.polymul(so.unpack(a), so.unpack(b), Q)


def test_monomial_initial_extract(so):
    P, M = so.Params(64), md.params(64)
    rng = np.random.default_rng(1)
    p = so.rand_below(rng, P.Q, (P.m,))
    pi = so.unpack(p)
    for shift in (0, 1, -1, 511, 512, 513, 1023, -700, 5000):
        assert so.unpack(so.mul_by_monomial(p, shift, P.Q)) == md.mul_by_monomial(pi, shift, P.Q)
    t = so.unpack(so.initial_poly(P))
    assert t == md.initial_poly(M)
    assert t[:P.m // 2] == [1] * (P.m // 2) and t[P.m // 2] == 0 and t[P.m // 2 + 1:] == [P.Q - 1] * (P.m // 2 - 1)
    for i in (1, 5, 63, 64, 65, 385, 512):
        assert so.unpack(so.extract(p, i, 64, P.Q)) == md.extract(pi, i, 64, P.Q)


def test_encrypt_split_decrypt(so):
    """port of test/api.test.jl:33-42 (split ciphertext) at Params(512)"""
    P, M = so.Params(512), md.params(512)
    sk = so.make_secret(P, 3)
    bits, lwes = so.make_lwes(P, sk, 3)
    for i in range(P.n):
        assert so.decrypt_lwe(P, sk, lwes[i]) == bits[i]
    for i in (0, 17, 511):
        assert md.decrypt_lwe(M, sk.tolist(), lwes[i].tolist()) == bits[i]


@pytest.fixture(scope="module")
def gate64(so):
    P = so.Params(64)
    sk = so.make_secret(P, 0)
    key = so.make_bkey(P, sk, 0)
    bits, lwes = so.make_lwes(P, sk, 0)
    return P, sk, key, bits, lwes


@pytest.mark.parametrize("use_rng", [False, True])
def test_bootstrap_gates(so, gate64, use_rng):
    """port of test/api.test.jl:45-83: Params(64), 32 disjoint pairs; literal formulation (sgo_bootstrap)"""
    P, sk, key, bits, lwes = gate64
    rng = np.random.default_rng(2)
    xmax = P.B // 2 * 3
    for g in range(0, 32, 1 if not use_rng else 4):
        draws = rng.integers(-xmax, xmax + 1, size=(P.n, 2, P.m, 2), dtype=np.int64) if use_rng else None
        r = so.bootstrap(P, key, lwes[g], lwes[32 + g], draws)
        y1, y2 = int(bits[g]), int(bits[32 + g])
        assert tuple(so.decrypt_lwe(P, sk, v) for v in r) == (y1 & y2, y1 | y2, y1 ^ y2)


def test_bootstrap_deterministic(so, gate64):
    """docs/src/manual.md:155-169"""
    P, sk, key, bits, lwes = gate64
    r1, r2 = so.bootstrap(P, key, lwes[1], lwes[2]), so.bootstrap(P, key, lwes[1], lwes[2])
    assert all(np.array_equal(a, b) for a, b in zip(r1, r2))


@pytest.mark.parametrize("use_rng", [False, True])
def test_literal_equals_rewritten_every_step(so, gate64, use_rng):
    """(a,b) (.) (G + (x^u-1) C) == (a,b) + (x^u-1) ([flatten a; flatten b] . C), every step"""
    P, sk, key, bits, lwes = gate64
    xmax = P.B // 2 * 3
    draws = np.random.default_rng(6).integers(-xmax, xmax + 1, size=(P.n, 2, P.m, 2), dtype=np.int64) if use_rng else None
    a = so.bootstrap_internal(P, key, lwes[7], lwes[9], draws=draws, trace=True)
    b = so.bootstrap_internal(P, key, lwes[7], lwes[9], draws=draws, trace=True, fast=True)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_oracle_equals_model_two_steps(so, gate64):
    P, sk, key, bits, lwes = gate64
    M = md.params(64)
    xmax = P.B // 2 * 3
    draws = np.random.default_rng(5).integers(-xmax, xmax + 1, size=(2, 2, P.m, 2), dtype=np.int64)
    mk = so.unpack(key[:2])
    for d in (None, draws):
        oa, oo, ox, tr = so.bootstrap_internal(P, key, lwes[10], lwes[20], draws=d, n_steps=2, trace=True)
        mtr = []
        ma, mo, mx = md.bootstrap_internal(M, mk, lwes[10].tolist(), lwes[20].tolist(), None if d is None else d.tolist(), 2, mtr)
        assert (so.unpack(oa), so.unpack(oo), so.unpack(ox)) == (ma, mo, mx)
        assert so.unpack(tr[1, 0]) == mtr[1][0] and so.unpack(tr[1, 1]) == mtr[1][1]


def test_batch_threads_equal_single(so, gate64):
    P, sk, key, bits, lwes = gate64
    o1 = so.bootstrap_batch(P, key, lwes[:6], lwes[6:12], literal=True, threads=3)
    o2 = so.bootstrap_batch(P, key, lwes[:6], lwes[6:12], literal=False, threads=1)
    assert all(np.array_equal(a, b) for a, b in zip(o1, o2))
    r = so.bootstrap(P, key, lwes[4], lwes[10])
    assert all(np.array_equal(o[4], x) for o, x in zip(o1, r))


@pytest.mark.parametrize("name", ["golden_p64.npz", "golden_p1024_trunc.npz", "golden_p512_trunc.npz", "golden_p2048_trunc.npz"])
def test_oracle_reproduces_golden(so, name):
    """committed fixtures (tests/golden/make_golden.py): inputs regenerate from the seed, outputs match"""
    g = np.load(os.path.join(GOLD, name))
    n, seed, steps = int(g["n"]), int(g["seed"]), int(g["steps"])
    P = so.Params(n)
    sk = so.make_secret(P, seed)
    assert np.array_equal(sk, g["sk"])
    key = so.make_bkey(P, sk, seed, rows=steps)
    assert sha(key) == str(g["key_sha256"])
    xmax = P.B // 2 * 3
    rng = np.random.default_rng([seed, 9])
    for pi in range(len(g["pairs"])):
        for mode in ("det", "rnd"):
            tag = f"p{pi}_{mode}"
            if tag + "_lwe1" not in g:
                continue
            draws = rng.integers(-xmax, xmax + 1, size=(steps, 2, P.m, 2), dtype=np.int64) if mode == "rnd" else None
            if draws is not None:
                assert sha(draws) == str(g[tag + "_draws_sha256"])
            a, o, x, tr = so.bootstrap_internal(P, key, g[tag + "_lwe1"], g[tag + "_lwe2"], draws=draws, n_steps=steps,
                                                trace=True, fast=(n > 64))
            assert np.array_equal(a, g[tag + "_and_Q"]) and np.array_equal(o, g[tag + "_or_Q"]) and np.array_equal(x, g[tag + "_xor_Q"])
            assert [sha(tr[k]) for k in range(steps)] == [str(s) for s in g[tag + "_trace_sha256"]]


@pytest.mark.parametrize("use_rng", [False, True])
def test_pack_encrypted_bits(so, gate64, use_rng):
    """port of test/api.test.jl:86-108: pack -> split -> decrypt and pack -> decrypt give the message"""
    P, sk, key, bits, lwes = gate64
    xmax = P.B // 2 * 3
    gen = np.random.default_rng(17)
    db = gen.integers(-xmax, xmax + 1, size=(P.n, P.n, 2, P.m, 2), dtype=np.int64) if use_rng else None
    ds = gen.integers(-xmax, xmax + 1, size=(P.n, P.m, 2), dtype=np.int64) if use_rng else None
    w, v = so.pack_encrypted_bits(P, key, lwes, db, ds)
    assert np.array_equal(so.decrypt_ciphertext(P, sk, w, v), bits)
    l2 = so.split_rlwe(P, w, v)
    assert [so.decrypt_lwe(P, sk, l2[i]) for i in range(P.n)] == bits.tolist()


def test_pack_oracle_equals_model(so, gate64):
    """shortened_external_product + the assembly of src/fhe.jl:675-693 against the big-integer model"""
    P, sk, key, bits, lwes = gate64
    M = md.params(64)
    triv = np.zeros(P.n + 1, np.uint64); triv[P.n] = P.Dr
    nl = np.stack([so.bootstrap_internal(P, key, triv, lwes[j], fast=True)[0] for j in range(P.n)])
    w, v = so.pack_from_lwes(P, key, nl)
    mw, mv = md.pack_from_lwes(M, so.unpack(key), so.unpack(nl))
    assert mw == w.tolist() and mv == v.tolist()
    rng = np.random.default_rng(9)
    a = so.rand_below(rng, P.Q, (P.m,))
    xmax = P.B // 2 * 3
    d = rng.integers(-xmax, xmax + 1, size=(P.m, 2), dtype=np.int64)
    for dr in (None, d):
        oa, ob = so.shortened_external_product(a, key[5], P.B, P.Q, dr)
        ma, mb = md.shortened_external_product(so.unpack(a), so.unpack(key[5]), P.B, P.Q, None if dr is None else dr.tolist())
        assert so.unpack(oa) == ma and so.unpack(ob) == mb


def test_rns2_polymul_oracle_vs_model(so):
    """the oracle's two-limb negacyclic product (Scheme 2, src/fhe2.jl:124 / src/rns.jl:51-52) against the big-integer
    model's Kronecker product, with the moduli of Scheme2.Params(1) at a short length the model handles quickly"""
    S = so.scheme2_params(1)
    N = 256                                           # B - 1 and B' - 1 are divisible by r = 4096, so by 2 N
    rng = np.random.default_rng(3)
    a1 = rng.integers(0, S.B, size=N, dtype=np.uint64); a2 = rng.integers(0, S.Bp, size=N, dtype=np.uint64)
    b1 = rng.integers(0, S.B, size=N, dtype=np.uint64); b2 = rng.integers(0, S.Bp, size=N, dtype=np.uint64)
    a1[:2] = (S.B - 1, 0); b1[:2] = (S.B - 1, S.B - 1)
    o1, o2 = so.rns2_polymul(a1, a2, b1, b2, S.B, S.Bp)
    assert o1.tolist() == md.polymul(a1.tolist(), b1.tolist(), S.B)
    assert o2.tolist() == md.polymul(a2.tolist(), b2.tolist(), S.Bp)
