"""ctypes front end of oracle/libsgfhe_oracle.so (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  Wide values are numpy uint64 arrays with a trailing axis of 2 (lo, hi).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class U128(C.Structure):
    _fields_ = [("lo", C.c_uint64), ("hi", C.c_uint64)]


class ParamsC(C.Structure):
    _fields_ = [("n", C.c_int32), ("t", C.c_int32), ("m", C.c_int32), ("large", C.c_int32),
                ("r", C.c_uint64), ("q", C.c_uint64), ("Dr", C.c_uint64), ("Dq", C.c_uint64),
                ("Q", U128), ("B", U128), ("DQ", U128)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libsgfhe_oracle.so")
    src = os.path.join(_HERE, "sgfhe_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libsgfhe_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.sgo_rescale.restype = U128
        _LIB.sgo_rescale.argtypes = [U128, U128, U128, C.c_int]
        _LIB.sgo_find_modulus.argtypes = [U128, U128, U128, C.POINTER(U128)]
        _LIB.sgo_is_prime.argtypes = [U128]
        _LIB.sgo_flatten.argtypes = [U128, U128, C.c_int, U128, C.c_void_p, C.c_void_p]
        _LIB.sgo_flatten_poly.argtypes = [C.c_void_p, C.c_int, U128, C.c_int, U128, C.c_void_p, C.c_void_p]
        _LIB.sgo_polymul_schoolbook.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, U128]
        _LIB.sgo_polymul_ntt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, U128]
        _LIB.sgo_mul_by_monomial.argtypes = [C.c_void_p, C.c_int, C.c_int64, U128, C.c_void_p]
        _LIB.sgo_extract.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, U128, C.c_void_p]
        _LIB.sgo_external_product.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, U128, U128,
                                              C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.sgo_decrypt_lwe.restype = C.c_uint64
        _LIB.sgo_shortened_external_product.argtypes = [C.c_void_p, C.c_void_p, C.c_int, U128, U128, C.c_void_p, C.c_void_p, C.c_void_p]
    return _LIB


def u(x: int) -> U128:
    return U128(x & 0xFFFFFFFFFFFFFFFF, x >> 64)


def to_int(x: U128) -> int:
    return x.lo | (x.hi << 64)


def pack(vals) -> np.ndarray:
    """ints (any nesting, via numpy object array) -> uint64[..., 2]"""
    a = np.asarray(vals, dtype=object)
    out = np.empty(a.shape + (2,), dtype=np.uint64)
    flat = a.reshape(-1)
    o = out.reshape(-1, 2)
    for i, v in enumerate(flat):
        v = int(v)
        o[i, 0] = v & 0xFFFFFFFFFFFFFFFF
        o[i, 1] = v >> 64
    return out


def unpack(arr: np.ndarray):
    """uint64[..., 2] -> nested lists of Python ints"""
    a = np.asarray(arr, dtype=np.uint64)
    obj = a[..., 0].astype(object) + (a[..., 1].astype(object) << 64)
    return obj.tolist()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Params:
    """fhe.jl:27-99 via sgo_params_init"""

    def __init__(self, n: int):
        self.c = ParamsC()
        rc = lib().sgo_params_init(n, C.byref(self.c))
        if rc:
            raise ValueError(f"sgo_params_init({n}) -> {rc}")
        c = self.c
        self.n, self.t, self.m, self.large = c.n, c.t, c.m, c.large
        self.r, self.q, self.Dr, self.Dq = c.r, c.q, c.Dr, c.Dq
        self.Q, self.B, self.DQ = to_int(c.Q), to_int(c.B), to_int(c.DQ)


def find_modulus(n, qmin, qmax=None):
    out = U128()
    rc = lib().sgo_find_modulus(u(n), u(qmin), u(qmax or 0), C.byref(out))
    if rc:
        raise ValueError("no modulus")
    return to_int(out)


def rescale(new_max, x, old_max, round_result):
    return to_int(lib().sgo_rescale(u(new_max), u(x), u(old_max), int(round_result)))


def flatten(a, B, l, q, draws=None):
    out = np.zeros((l, 2), np.uint64)
    d = None if draws is None else np.ascontiguousarray(draws, np.int64)
    lib().sgo_flatten(u(a), u(B), l, u(q), _p(d), _p(out))
    return unpack(out)


def flatten_poly(a, B, l, q, draws=None):
    a = np.ascontiguousarray(a, np.uint64)
    N = a.shape[0]
    out = np.zeros((l, N, 2), np.uint64)
    d = None if draws is None else np.ascontiguousarray(draws, np.int64)
    lib().sgo_flatten_poly(_p(a), N, u(B), l, u(q), _p(d), _p(out))
    return out


def polymul(a, b, Q, schoolbook=False):
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    out = np.zeros_like(a)
    if schoolbook:
        lib().sgo_polymul_schoolbook(_p(a), _p(b), _p(out), a.shape[0], u(Q))
    else:
        rc = lib().sgo_polymul_ntt(_p(a), _p(b), _p(out), a.shape[0], u(Q))
        if rc:
            raise ValueError("no NTT for this (N, Q)")
    return out


def mul_by_monomial(p, shift, Q):
    p = np.ascontiguousarray(p, np.uint64)
    out = np.zeros_like(p)
    lib().sgo_mul_by_monomial(_p(p), p.shape[0], int(shift), u(Q), _p(out))
    return out


def initial_poly(P: Params):
    out = np.zeros((P.m, 2), np.uint64)
    lib().sgo_initial_poly(C.byref(P.c), _p(out))
    return out


def extract(a, i, n, modulus):
    a = np.ascontiguousarray(a, np.uint64)
    out = np.zeros((n, 2), np.uint64)
    lib().sgo_extract(_p(a), a.shape[0], i, n, u(modulus), _p(out))
    return out


def bkey_generate(P: Params, sk, a_rand, e_rand, row0=0, row1=None):
    """sk uint8[n]; a_rand uint64[n,4,m,2]; e_rand int64[n,4,m] -> key uint64[rows,4,2,m,2]"""
    row1 = P.n if row1 is None else row1
    sk = np.ascontiguousarray(sk, np.uint8)
    a_rand = np.ascontiguousarray(a_rand, np.uint64)
    e_rand = np.ascontiguousarray(e_rand, np.int64)
    key = np.zeros((row1 - row0, 4, 2, P.m, 2), np.uint64)
    rc = lib().sgo_bkey_generate(C.byref(P.c), _p(sk), _p(a_rand), _p(e_rand), row0, row1, _p(key))
    if rc:
        raise RuntimeError(f"sgo_bkey_generate -> {rc}")
    return key


def external_product(a, b, A, B, Q, draws=None):
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    A = np.ascontiguousarray(A, np.uint64)
    d = None if draws is None else np.ascontiguousarray(draws, np.int64)
    oa, ob = np.zeros_like(a), np.zeros_like(b)
    rc = lib().sgo_external_product(_p(a), _p(b), _p(A), a.shape[0], u(B), u(Q), _p(d), _p(oa), _p(ob))
    if rc:
        raise ValueError("no NTT for this (N, Q)")
    return oa, ob


def bootstrap_internal(P: Params, key, lwe1, lwe2, draws=None, n_steps=None, trace=False, fast=False):
    n_steps = P.n if n_steps is None else n_steps
    key = np.ascontiguousarray(key, np.uint64)
    lwe1 = np.ascontiguousarray(lwe1, np.uint64)
    lwe2 = np.ascontiguousarray(lwe2, np.uint64)
    d = None if draws is None else np.ascontiguousarray(draws, np.int64)
    tr = np.zeros((n_steps, 2, P.m, 2), np.uint64) if trace else None
    outs = [np.zeros((P.n + 1, 2), np.uint64) for _ in range(3)]
    fn = lib().sgo_bootstrap_internal_fast if fast else lib().sgo_bootstrap_internal
    rc = fn(C.byref(P.c), _p(key), _p(lwe1), _p(lwe2), _p(d), n_steps, _p(tr), *[_p(o) for o in outs])
    if rc:
        raise RuntimeError(f"bootstrap_internal -> {rc}")
    return (outs[0], outs[1], outs[2], tr) if trace else tuple(outs)


def bootstrap(P: Params, key, lwe1, lwe2, draws=None):
    key = np.ascontiguousarray(key, np.uint64)
    lwe1 = np.ascontiguousarray(lwe1, np.uint64)
    lwe2 = np.ascontiguousarray(lwe2, np.uint64)
    d = None if draws is None else np.ascontiguousarray(draws, np.int64)
    outs = [np.zeros(P.n + 1, np.uint64) for _ in range(3)]
    rc = lib().sgo_bootstrap(C.byref(P.c), _p(key), _p(lwe1), _p(lwe2), _p(d), *[_p(o) for o in outs])
    if rc:
        raise RuntimeError(f"sgo_bootstrap -> {rc}")
    return tuple(outs)


def bootstrap_batch(P: Params, key, lwe1, lwe2, n_steps=None, literal=True, threads=1):
    n_steps = P.n if n_steps is None else n_steps
    key = np.ascontiguousarray(key, np.uint64)
    lwe1 = np.ascontiguousarray(lwe1, np.uint64)
    lwe2 = np.ascontiguousarray(lwe2, np.uint64)
    batch = lwe1.shape[0]
    outs = [np.zeros((batch, P.n + 1), np.uint64) for _ in range(3)]
    rc = lib().sgo_bootstrap_batch(C.byref(P.c), _p(key), batch, _p(lwe1), _p(lwe2), n_steps,
                                   int(literal), threads, *[_p(o) for o in outs])
    if rc:
        raise RuntimeError(f"sgo_bootstrap_batch -> {rc}")
    return tuple(outs)


def encrypt_private(P: Params, sk, a, w, message):
    sk = np.ascontiguousarray(sk, np.uint8)
    a = np.ascontiguousarray(a, np.uint64)
    w = np.ascontiguousarray(w, np.int64)
    message = np.ascontiguousarray(message, np.uint8)
    b = np.zeros(P.n, np.uint64)
    lib().sgo_encrypt_private(C.byref(P.c), _p(sk), _p(a), _p(w), _p(message), _p(b))
    return b


def split_ciphertext(P: Params, a, b):
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    out = np.zeros((P.n, P.n + 1), np.uint64)
    lib().sgo_split_ciphertext(C.byref(P.c), _p(a), _p(b), _p(out))
    return out


def decrypt_lwe(P: Params, sk, lwe) -> int:
    sk = np.ascontiguousarray(sk, np.uint8)
    lwe = np.ascontiguousarray(lwe, np.uint64)
    return int(lib().sgo_decrypt_lwe(C.byref(P.c), _p(sk), _p(lwe)))


def shortened_external_product(a, A, B, Q, draws=None):
    a = np.ascontiguousarray(a, np.uint64)
    A = np.ascontiguousarray(A, np.uint64)
    d = None if draws is None else np.ascontiguousarray(draws, np.int64)
    oa, ob = np.zeros_like(a), np.zeros_like(a)
    rc = lib().sgo_shortened_external_product(_p(a), _p(A), a.shape[0], u(B), u(Q), _p(d), _p(oa), _p(ob))
    if rc:
        raise ValueError("no NTT for this (N, Q)")
    return oa, ob


def pack_from_lwes(P: Params, key, new_lwes, draws_short=None):
    key = np.ascontiguousarray(key, np.uint64)
    new_lwes = np.ascontiguousarray(new_lwes, np.uint64)
    d = None if draws_short is None else np.ascontiguousarray(draws_short, np.int64)
    w, v = np.zeros(P.m, np.uint64), np.zeros(P.m, np.uint64)
    rc = lib().sgo_pack_from_lwes(C.byref(P.c), _p(key), _p(new_lwes), _p(d), _p(w), _p(v))
    if rc:
        raise RuntimeError(f"sgo_pack_from_lwes -> {rc}")
    return w, v


def pack_encrypted_bits(P: Params, key, enc_bits, draws_boot=None, draws_short=None):
    key = np.ascontiguousarray(key, np.uint64)
    enc_bits = np.ascontiguousarray(enc_bits, np.uint64)
    db = None if draws_boot is None else np.ascontiguousarray(draws_boot, np.int64)
    ds = None if draws_short is None else np.ascontiguousarray(draws_short, np.int64)
    w, v = np.zeros(P.m, np.uint64), np.zeros(P.m, np.uint64)
    rc = lib().sgo_pack_encrypted_bits(C.byref(P.c), _p(key), _p(enc_bits), _p(db), _p(ds), _p(w), _p(v))
    if rc:
        raise RuntimeError(f"sgo_pack_encrypted_bits -> {rc}")
    return w, v


def split_rlwe(P: Params, a, b):
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    out = np.zeros((P.n, P.n + 1), np.uint64)
    lib().sgo_split_rlwe(C.byref(P.c), a.shape[0], _p(a), _p(b), _p(out))
    return out


def decrypt_ciphertext(P: Params, sk, a, b):
    sk = np.ascontiguousarray(sk, np.uint8)
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    out = np.zeros(P.n, np.uint64)
    lib().sgo_decrypt_ciphertext(C.byref(P.c), _p(sk), _p(a), _p(b), _p(out))
    return out


class Scheme2ParamsC(C.Structure):
    _fields_ = [("n", C.c_int32), ("k", C.c_int32), ("t", C.c_int32), ("pad", C.c_int32)] + \
               [(f, C.c_uint64) for f in ("r", "m", "q", "tau", "B", "Bp", "Dr", "Dq")]


def scheme2_params(k: int) -> Scheme2ParamsC:
    out = Scheme2ParamsC()
    rc = lib().sgo_scheme2_params(k, C.byref(out))
    if rc:
        raise ValueError(f"sgo_scheme2_params({k}) -> {rc}")
    return out


def rns2_op(op: int, a1, a2, b1, b2, M1: int, M2: int):
    arrs = [np.ascontiguousarray(x, np.uint64) for x in (a1, a2, b1, b2)]
    o1, o2 = np.zeros_like(arrs[0]), np.zeros_like(arrs[0])
    lib().sgo_rns2_op.argtypes = [C.c_int, C.c_size_t] + [C.c_void_p] * 4 + [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    lib().sgo_rns2_op(op, arrs[0].size, *[_p(x) for x in arrs], M1, M2, _p(o1), _p(o2))
    return o1, o2


def rns2_polymul(a1, a2, b1, b2, M1: int, M2: int):
    """Polynomial{RNS2Number} `*` (negacyclic; src/fhe2.jl:124 with the limb-wise product of src/rns.jl:51-52): each limb
    is a product in Z_Mi[x]/(x^N+1), computed by the Z_Q oracle product with Q = Mi (Mi is an NTT prime, src/fhe2.jl:57-60).
    a*, b*: uint64[N] -> (o1, o2) uint64[N]"""
    outs = []
    for a, b, M in ((a1, b1, M1), (a2, b2, M2)):
        wa = np.stack([np.asarray(a, np.uint64), np.zeros(len(a), np.uint64)], axis=-1)
        wb = np.stack([np.asarray(b, np.uint64), np.zeros(len(b), np.uint64)], axis=-1)
        outs.append(np.ascontiguousarray(polymul(wa, wb, M)[:, 0]))
    return outs[0], outs[1]


def scheme2_bkey_generate(k: int, sk, a_rand, e_rand, rows: int):
    """Scheme2.BootstrapKey (src/fhe2.jl:104-131) for key rows 0..rows-1.  sk uint8[n]; a_rand uint64[rows,4,m,2] wide
    integers below Q = B B'; e_rand int64[rows,4,m] -> uint64[rows,4,2,m,2] (v1, v2) pairs of key[i][j,c].coeffs."""
    S = scheme2_params(k)
    n, m, mods = S.n, S.m, (S.B, S.Bp)
    ext = np.zeros(m, np.uint64); ext[:n] = np.asarray(sk, np.uint64)                       # fhe2.jl:113
    key = np.zeros((rows, 4, 2, m, 2), np.uint64)
    aw = np.asarray(a_rand, np.uint64)
    a_int = aw[..., 0].astype(object) + (aw[..., 1].astype(object) << 64)
    for i in range(rows):
        for j in range(4):
            for l, M in enumerate(mods):
                aj = np.array([int(v) % M for v in a_int[i, j]], np.uint64)                 # polynomial_Q, fhe2.jl:81-82
                wa = np.stack([aj, np.zeros(m, np.uint64)], axis=-1)
                we = np.stack([ext, np.zeros(m, np.uint64)], axis=-1)
                pr = polymul(wa, we, M)[:, 0].astype(object)
                bj = np.array([(int(v) + int(e)) % M for v, e in zip(pr, e_rand[i, j])], np.uint64)   # fhe2.jl:124
                if sk[i]:                                                                   # + s_i G, fhe2.jl:125, 98-101
                    g = (S.B % M) if (j & 1) else 1
                    if j < 2:
                        aj[0] = (int(aj[0]) + g) % M
                    else:
                        bj[0] = (int(bj[0]) + g) % M
                key[i, j, 0, :, l] = aj
                key[i, j, 1, :, l] = bj
    return key


def set_setup_threads(t: int):
    lib().sgo_set_setup_threads(int(t))


# ---- seeded synthetic inputs in the reference's formats (numpy PCG64; the reference's MersenneTwister
# ---- streams are not reproducible without Julia, and parity is defined on identical inputs) ----------
def rand_below(rng: np.random.Generator, bound: int, shape) -> np.ndarray:
    """uniform integers in [0, bound) as uint64[..., 2]; bound < 2^128 (rejection sampling)"""
    bits = bound.bit_length()
    n = int(np.prod(shape))
    out = np.zeros((n, 2), np.uint64)
    todo = np.arange(n)
    hi_mask = np.uint64((1 << max(bits - 64, 0)) - 1) if bits > 64 else np.uint64(0)
    lo_mask = np.uint64((1 << min(bits, 64)) - 1)
    bh, bl = bound >> 64, bound & 0xFFFFFFFFFFFFFFFF
    while todo.size:
        lo = rng.integers(0, 1 << 64, size=todo.size, dtype=np.uint64) & lo_mask
        hi = (rng.integers(0, 1 << 64, size=todo.size, dtype=np.uint64) & hi_mask) if bits > 64 else np.zeros(todo.size, np.uint64)
        ok = (hi < np.uint64(bh)) | ((hi == np.uint64(bh)) & (lo < np.uint64(bl)))
        out[todo[ok], 0] = lo[ok]
        out[todo[ok], 1] = hi[ok]
        todo = todo[~ok]
    return out.reshape(tuple(shape) + (2,))


def make_secret(P: Params, seed: int) -> np.ndarray:
    """PrivateKey, fhe.jl:134-137: n random bits"""
    return np.random.default_rng([seed, 1]).integers(0, 2, size=P.n, dtype=np.uint8)


def make_bkey(P: Params, sk, seed: int, rows: int | None = None) -> np.ndarray:
    """BootstrapKey, fhe.jl:181-201, draw order a_1..a_4 then e_1..e_4 per row (fhe.jl:193-194)"""
    rows = P.n if rows is None else rows
    a_rand = np.zeros((P.n, 4, P.m, 2), np.uint64)
    e_rand = np.zeros((P.n, 4, P.m), np.int64)
    for i in range(rows):
        rng = np.random.default_rng([seed, 2, i])
        a_rand[i] = rand_below(rng, P.Q, (4, P.m))
        e_rand[i] = rng.integers(-P.n, P.n + 1, size=(4, P.m), dtype=np.int64)
    return bkey_generate(P, sk, a_rand, e_rand, 0, rows)


def make_lwes(P: Params, sk, seed: int, blocks: int = 1):
    """encrypt (fhe.jl:369, 310-328) + split_ciphertext (fhe.jl:287-290): blocks*n (bit, LWE) pairs"""
    bits, lwes = [], []
    for blk in range(blocks):
        rng = np.random.default_rng([seed, 3, blk])
        msg = rng.integers(0, 2, size=P.n, dtype=np.uint8)
        a = rng.integers(0, P.r, size=P.n, dtype=np.uint64)     # stands in for deterministic_expand, utils.jl:63-68
        wr = P.Dr // 8
        w = rng.integers(-wr, wr + 1, size=P.n, dtype=np.int64)  # fhe.jl:318-319
        b = encrypt_private(P, sk, a, w, msg)
        lwes.append(split_ciphertext(P, a, b))
        bits.append(msg)
    return np.concatenate(bits), np.concatenate(lwes)
