"""Host-side mirror of SGFHE.jl's public API for the bootstrapping path, over the C ABI.

Same names, argument meaning and error behaviour as the reference (exports: src/SGFHE.jl:10-20):
Params, PrivateKey, BootstrapKey, encrypt, decrypt, split_ciphertext, bootstrap.  The reference's
`rng::AbstractRNG` is a `numpy.random.Generator` here and `nothing` is `None`.  Everything that is
arithmetic on the hot path (polynomial products, flatten, the accumulation loop, extract, ModRed)
runs in libsgfhe_cuda.so; this module only prepares inputs in the reference's formats.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import SgfheError, check


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _wide_to_int(w) -> int:
    return int(w[0]) | (int(w[1]) << 64)


class Params:
    """Params(n) -- src/fhe.jl:27-99.  Raises SgfheError where the reference asserts/errors."""

    def __init__(self, n: int, device: int = 0):
        pc = _lib.ParamsC()
        check(_lib.lib().sgfhe_params_derive(int(n), C.byref(pc)))
        self.n, self.t, self.m = pc.n, pc.t, pc.m
        self.r, self.q, self.Dr, self.Dq = pc.r, pc.q, pc.Dr, pc.Dq
        self.Q, self.B, self.DQ_tilde = _wide_to_int(pc.Q), _wide_to_int(pc.B), _wide_to_int(pc.DQ_tilde)
        self.device = device
        self._ctx = None

    @property
    def ctx(self):
        """Device context (created on first use; fails loudly without a GPU)."""
        if self._ctx is None:
            h = C.c_void_p()
            check(_lib.lib().sgfhe_ctx_create(self.n, self.device, C.byref(h)))
            self._ctx = h
        return self._ctx

    def close(self):
        if self._ctx is not None:
            _lib.lib().sgfhe_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PrivateKey:
    """PrivateKey(params, rng) -- src/fhe.jl:130-138: n random bits."""

    def __init__(self, params: Params, rng: np.random.Generator):
        self.params = params
        self.key = rng.integers(0, 2, size=params.n, dtype=np.uint8)


class LWE:
    """src/fhe.jl:206-223.  a: uint64[n], b: int, both over Z_r."""

    def __init__(self, a, b):
        self.a = np.asarray(a, dtype=np.uint64)
        self.b = int(b)

    def flat(self) -> np.ndarray:
        return np.concatenate([self.a, np.array([self.b], np.uint64)])

    def __eq__(self, other):
        return np.array_equal(self.a, other.a) and self.b == other.b


class EncryptedBit:
    """src/fhe.jl:272-278"""

    def __init__(self, lwe: LWE):
        self.lwe = lwe

    def __eq__(self, other):
        return self.lwe == other.lwe


class PackedCiphertext:
    """src/fhe.jl:251-254: RLWE (a, b) over (x^n+1, r)."""

    def __init__(self, params: Params, a, b):
        self.params = params
        self.a = np.asarray(a, np.uint64)
        self.b = np.asarray(b, np.uint64)


class Ciphertext:
    """src/fhe.jl:261-264: RLWE (a, b) over (x^m+1, r), produced by pack_encrypted_bits."""

    def __init__(self, params: Params, a, b):
        self.params = params
        self.a = np.asarray(a, np.uint64)
        self.b = np.asarray(b, np.uint64)


def _negacyclic_small(a: np.ndarray, s: np.ndarray, modulus: int) -> np.ndarray:
    """a * s in Z_modulus[x]/(x^n+1) for a binary s (host; length-n ciphertext work, not the hot path)."""
    n = a.shape[0]
    full = np.convolve(a.astype(np.int64), s.astype(np.int64))
    res = full[:n].copy()
    res[: n - 1] -= full[n:]
    return np.mod(res, modulus).astype(np.uint64)


def encrypt(key: PrivateKey, rng: np.random.Generator, message) -> PackedCiphertext:
    """encrypt(key::PrivateKey, rng, message) -- src/fhe.jl:369-372 -> _encrypt_private src/fhe.jl:310-328.

    `a` is drawn from rng directly: the reference expands a random seed with a MersenneTwister
    (deterministic_expand, src/utils.jl:63-68, marked TODO upstream); any uniform a over Z_r is a valid ciphertext."""
    P = key.params
    message = np.asarray(message, dtype=np.uint8)
    if message.shape != (P.n,):
        raise SgfheError("message must have length n (src/fhe.jl:313)")
    a = rng.integers(0, P.r, size=P.n, dtype=np.uint64)
    wr = P.Dr // 8
    w = rng.integers(-wr, wr + 1, size=P.n, dtype=np.int64)                       # src/fhe.jl:318-319
    b = (_negacyclic_small(a, key.key, P.r).astype(np.int64) + w + message.astype(np.int64) * P.Dr) % P.r   # :322
    step = 1 << (P.t - 4)
    b = (b // step) * step                                                         # src/fhe.jl:325
    return PackedCiphertext(P, a, b.astype(np.uint64))


def _extract_r(a: np.ndarray, i: int, n: int, r: int) -> np.ndarray:
    """extract(a, i, n) over Z_r -- src/fhe.jl:237-244 (i is 1-based)."""
    N = a.shape[0]
    if i < n:
        tail = a[N - (n - i):][::-1]
        return np.concatenate([a[:i][::-1], np.where(tail == 0, 0, r - tail).astype(np.uint64)])
    return a[i - n:i][::-1].copy()


def split_ciphertext(ct) -> list[EncryptedBit]:
    """split_ciphertext(ct::Union{Ciphertext, PackedCiphertext}) -- src/fhe.jl:287-290"""
    P = ct.params
    return [EncryptedBit(LWE(_extract_r(ct.a, i, P.n, P.r), ct.b[i - 1])) for i in range(1, P.n + 1)]


def split_ciphertexts(cts) -> np.ndarray:
    """split_ciphertext (src/fhe.jl:287-290) for a list of ciphertexts on the GPU: uint64[len(cts) * n, n + 1], LWE rows in
    the reference's order, ready for `bootstrap_batch`."""
    P = cts[0].params
    a = np.ascontiguousarray(np.stack([c.a for c in cts]), np.uint64)
    b = np.ascontiguousarray(np.stack([c.b for c in cts]), np.uint64)
    out = np.zeros((len(cts) * P.n, P.n + 1), np.uint64)
    check(_lib.lib().sgfhe_split_ciphertext(P.ctx, len(cts), a.shape[1], _ptr(a), _ptr(b), _ptr(out)))
    return out


def decrypt_bits(key: PrivateKey, lwes: np.ndarray) -> np.ndarray:
    """decrypt(key, ::EncryptedBit) (src/fhe.jl:504-507) for an array of LWEs uint64[count, n + 1] on the GPU -> bool[count]."""
    P = key.params
    lwes = np.ascontiguousarray(lwes, np.uint64)
    if lwes.ndim != 2 or lwes.shape[1] != P.n + 1:
        raise SgfheError("LWE array must be [count, n+1]")
    sk = np.ascontiguousarray(key.key, np.uint8)
    out = np.zeros(lwes.shape[0], np.uint8)
    check(_lib.lib().sgfhe_decrypt_bits(P.ctx, lwes.shape[0], _ptr(lwes), _ptr(sk), _ptr(out)))
    if (out > 1).any():
        raise SgfheError("InexactError: decrypted value is not a Bool (src/fhe.jl:506)")
    return out.astype(bool)


def decrypt(key: PrivateKey, ct):
    """decrypt(key, ::EncryptedBit) src/fhe.jl:504-507; decrypt(key, ::PackedCiphertext) src/fhe.jl:471-494."""
    P = key.params
    if isinstance(ct, EncryptedBit):
        b1 = (ct.lwe.b - int(ct.lwe.a[key.key.astype(bool)].sum())) % P.r
        v = ((b1 + P.Dr // 2) % P.r) // P.Dr
        if v > 1:
            raise SgfheError("InexactError: decrypted value is not a Bool (src/fhe.jl:506)")
        return bool(v)
    if isinstance(ct, Ciphertext):                     # key resized to m, first n coefficients kept (src/fhe.jl:474-485)
        ext = np.zeros(P.m, np.uint8)
        ext[: P.n] = key.key
        b1 = ((ct.b.astype(np.int64) - _negacyclic_small(ct.a, ext, P.r).astype(np.int64)) % P.r)[: P.n]
    else:
        b1 = (ct.b.astype(np.int64) - _negacyclic_small(ct.a, key.key, P.r).astype(np.int64)) % P.r
    v = ((b1 + P.Dr // 2) % P.r) // P.Dr
    if (v > 1).any():
        raise SgfheError("InexactError: decrypted value is not a Bool (src/fhe.jl:493)")
    return v.astype(bool)


def _rand_below(rng: np.random.Generator, bound: int, shape) -> np.ndarray:
    """uniform on [0, bound) as wide uint64[..., 2]: rejection sampling of bit_length(bound)-bit candidates, drawn in bulk
    (the accepted candidates are used in the order they were drawn)"""
    bits = bound.bit_length()
    n = int(np.prod(shape))
    out = np.empty((n, 2), np.uint64)
    lo_mask = np.uint64((1 << min(bits, 64)) - 1)
    bh, bl = np.uint64(bound >> 64), np.uint64(bound & 0xFFFFFFFFFFFFFFFF)
    accept = bound / float(1 << bits)
    filled = 0
    while filled < n:
        k = int((n - filled) / accept * 1.05) + 64
        lo = rng.integers(0, 1 << 64, size=k, dtype=np.uint64) & lo_mask
        hi = rng.integers(0, 1 << (bits - 64), size=k, dtype=np.uint64) if bits > 64 else np.zeros(k, np.uint64)
        idx = np.flatnonzero((hi < bh) | ((hi == bh) & (lo < bl)))[: n - filled]
        out[filled:filled + idx.size, 0] = lo[idx]
        out[filled:filled + idx.size, 1] = hi[idx]
        filled += idx.size
    return out.reshape(tuple(shape) + (2,))


class BootstrapKey:
    """BootstrapKey(rng, sk) -- src/fhe.jl:176-203, generated on the device.

    Per row i the draws are a_1..a_4 (uniform on [0,Q), m each) then e_1..e_4 (uniform on [-n,n]) as at
    src/fhe.jl:193-194, made by the caller's rng on the host; the 4n products a_j * ext_key, + e_j (src/fhe.jl:195),
    + s_i G (src/fhe.jl:196) and the pre-transform run in libsgfhe_cuda.so (sgfhe_bkey_generate), which leaves the key
    in the context in transform-domain form.  `key` (uint64[n,4,2,m,2], canonical wide residues, the reference's
    bkey.key[i][j,c].coeffs[k]) is kept on the host only with keep_coefficients=True or when an existing array is adopted
    with BootstrapKey(params=P, key=array).

    A Params context holds one key at a time; every key object remembers the token of its device copy and re-uploads
    (or, with no coefficient form to upload from, raises) when another key has replaced it."""

    def __init__(self, rng: np.random.Generator | None = None, sk: PrivateKey | None = None, *,
                 params: Params | None = None, key: np.ndarray | None = None, rows: int | None = None,
                 keep_coefficients: bool = False):
        self._token = 0
        if key is not None:                       # adopt an existing key[i][j,c] array
            self.params = params
            self.key = np.ascontiguousarray(key, np.uint64)
            self.rows = self.key.shape[0]
            return
        if sk is None:                            # device-resident key placed by import / broadcast: see resident()
            self.params, self.key, self.rows = params, None, 0
            return
        P = sk.params
        self.params = P
        nrows = P.n if rows is None else rows
        self.rows = nrows
        self.key = np.zeros((nrows, 4, 2, P.m, 2), np.uint64) if keep_coefficients else None
        skb = np.ascontiguousarray(sk.key, np.uint8)
        chunk = max(1, min(nrows, (1 << 22) // (4 * P.m)))
        L = _lib.lib()
        for i0 in range(0, nrows, chunk):
            i1 = min(nrows, i0 + chunk)
            aj = np.zeros((i1 - i0, 4, P.m, 2), np.uint64)
            ej = np.zeros((i1 - i0, 4, P.m), np.int64)
            for i in range(i0, i1):
                aj[i - i0] = _rand_below(rng, P.Q, (4, P.m))                     # src/fhe.jl:193
                ej[i - i0] = rng.integers(-P.n, P.n + 1, size=(4, P.m), dtype=np.int64)   # src/fhe.jl:194
            out = self.key[i0:i1] if keep_coefficients else None
            check(L.sgfhe_bkey_generate(P.ctx, _ptr(skb), _ptr(aj), _ptr(ej), i0, i1 - i0, _ptr(out)))
        self._token = _ctx_key_state(P)[0]

    @classmethod
    def resident(cls, params: "Params") -> "BootstrapKey":
        """The key currently in the context (after BootstrapKey.load_transformed or parallel.broadcast_key)."""
        bk = cls(params=params)
        bk._token, bk.rows = _ctx_key_state(params)
        if not bk._token:
            raise SgfheError("no bootstrap key in this context")
        return bk

    def rebind(self):
        """Adopt the context's current key as this object's device copy (the root rank after broadcast_key)."""
        self._token = _ctx_key_state(self.params)[0]

    def save_transformed(self, path: str):
        """Write the pre-transformed (device) key to `path`; reload with BootstrapKey.load_transformed."""
        self.upload()
        L = _lib.lib()
        nbytes = C.c_uint64()
        check(L.sgfhe_bkey_export_size(self.params.ctx, self.rows, C.byref(nbytes)))
        buf = np.empty(nbytes.value, np.uint8)
        check(L.sgfhe_bkey_export(self.params.ctx, self.rows, _ptr(buf), nbytes.value))
        buf.tofile(path)

    @classmethod
    def load_transformed(cls, params: "Params", path: str) -> "BootstrapKey":
        """A key usable for bootstrap straight from a file written by save_transformed (no coefficient form kept)."""
        buf = np.fromfile(path, np.uint8)
        check(_lib.lib().sgfhe_bkey_import(params.ctx, _ptr(buf), buf.size))
        return cls.resident(params)

    def upload(self):
        """Make sure the context holds THIS key (src/fhe.jl:608: bootstrap is a pure function of bkey)."""
        tok, _ = _ctx_key_state(self.params)
        if self._token and tok == self._token:
            return
        if self.key is None:
            raise SgfheError("this key's device copy was replaced by another key on the same Params and no coefficient "
                             "form was kept (keep_coefficients=True) to upload it again")
        check(_lib.lib().sgfhe_bkey_upload(self.params.ctx, _ptr(self.key), self.key.shape[0]))
        self._token = _ctx_key_state(self.params)[0]


def _ctx_key_state(P: Params) -> tuple[int, int]:
    tok, rows = C.c_uint64(), C.c_int32()
    check(_lib.lib().sgfhe_bkey_token(P.ctx, C.byref(tok), C.byref(rows)))
    return tok.value, rows.value


def _draws(P: Params, rng: np.random.Generator, shape) -> np.ndarray:
    """rand(rng, -xmax:xmax) with xmax = 3 (B / 2) -- src/utils.jl:210-216, 229"""
    xmax = (P.B // 2) * 3
    return rng.integers(-xmax, xmax + 1, size=shape, dtype=np.int64)


class DeviceRng:
    """An `rng` argument that makes the flatten draws on the GPU (Philox4x32-10 keyed by `seed`, one stream per gate):
    the randomised mode at sizes where host-drawn values do not fit (268 MB per gate at Params(1024)).  Not the stream of
    any host RNG: use a numpy Generator where the caller's exact draws matter."""

    def __init__(self, seed: int):
        if not 0 < seed < (1 << 64):
            raise SgfheError("seed must be in [1, 2^64)")
        self.seed, self.gates_drawn = int(seed), 0


def bootstrap_batch(bkey: BootstrapKey, rng, lwes1: np.ndarray, lwes2: np.ndarray):
    """Batched form of `bootstrap`: lwes1, lwes2 uint64[batch, n+1] -> (and, or, xor) uint64[batch, n+1].
    rng: None (deterministic flatten), a numpy Generator (draws made here in the reference's order) or a DeviceRng."""
    P = bkey.params
    lwes1 = np.ascontiguousarray(lwes1, np.uint64)
    lwes2 = np.ascontiguousarray(lwes2, np.uint64)
    if lwes1.shape != lwes2.shape or lwes1.ndim != 2 or lwes1.shape[1] != P.n + 1:
        raise SgfheError("LWE arrays must be [batch, n+1]")
    bkey.upload()
    batch = lwes1.shape[0]
    outs = [np.zeros_like(lwes1) for _ in range(3)]
    if isinstance(rng, DeviceRng):
        check(_lib.lib().sgfhe_bootstrap_batch_rng(P.ctx, batch, _ptr(lwes1), _ptr(lwes2), rng.seed, rng.gates_drawn,
                                                   *[_ptr(o) for o in outs]))
        rng.gates_drawn += batch
        return tuple(outs)
    draws = None if rng is None else _draws(P, rng, (batch, P.n, 2, P.m, 2))   # order: src/fhe.jl:524-525, utils.jl:257-258
    check(_lib.lib().sgfhe_bootstrap_batch(P.ctx, batch, _ptr(lwes1), _ptr(lwes2), _ptr(draws), *[_ptr(o) for o in outs]))
    return tuple(outs)


def bootstrap(bkey: BootstrapKey, rng, enc_bit1: EncryptedBit, enc_bit2: EncryptedBit):
    """bootstrap(bkey, rng|nothing, enc_bit1, enc_bit2) -> (AND, OR, XOR) -- src/fhe.jl:608-621."""
    outs = bootstrap_batch(bkey, rng, enc_bit1.lwe.flat()[None, :], enc_bit2.lwe.flat()[None, :])
    return tuple(EncryptedBit(LWE(o[0, :-1], o[0, -1])) for o in outs)


def pack_encrypted_bits(bkey: BootstrapKey, rng, enc_bits) -> Ciphertext:
    """pack_encrypted_bits(bkey, rng|nothing, enc_bits) -- src/fhe.jl:660-696, every stage on the device
    (sgfhe_pack_encrypted_bits): the n internal bootstraps (src/fhe.jl:673), the transposition (:675-678), the n
    shortened external products (:683-684), the two sums, negate / subtract (:686-690) and the final ModRed (:692-693).
    With an rng the draws are made here in the reference's order: the n bootstraps first, then the n products."""
    P = bkey.params
    if len(enc_bits) != P.n:
        raise SgfheError("expected n encrypted bits (src/fhe.jl:667)")
    bkey.upload()
    n, m = P.n, P.m
    bits = np.ascontiguousarray(np.stack([e.lwe.flat() for e in enc_bits]), np.uint64)
    draws = None if rng is None else _draws(P, rng, (n, n, 2, m, 2))
    ds = None if rng is None else _draws(P, rng, (n, m, 2))
    w, v = np.zeros(m, np.uint64), np.zeros(m, np.uint64)
    check(_lib.lib().sgfhe_pack_encrypted_bits(P.ctx, _ptr(bits), _ptr(draws), _ptr(ds), _ptr(w), _ptr(v)))
    return Ciphertext(P, w, v)


# ---- inner seams (test/internals.test.jl level) ---------------------------------------------------------
def polymul(P: Params, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """DarkIntegers `Polynomial *` in Z_Q[x]/(x^m+1), batched: wide uint64[batch, m, 2]."""
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    single = a.ndim == 2
    if single:
        a, b = a[None], b[None]
    if a.shape != b.shape or a.shape[1:] != (P.m, 2):
        raise SgfheError("operands must be [batch, m, 2]")
    out = np.zeros_like(a)
    check(_lib.lib().sgfhe_polymul(P.ctx, a.shape[0], _ptr(a), _ptr(b), _ptr(out)))
    return out[0] if single else out


def flatten_poly(P: Params, rng_draws, a: np.ndarray) -> np.ndarray:
    """flatten_poly(rng|nothing, a, Val(B), Val(2)) -- src/utils.jl:253-264.  rng_draws: None or int64[m,2]."""
    a = np.ascontiguousarray(a, np.uint64)
    d = None if rng_draws is None else np.ascontiguousarray(rng_draws, np.int64)
    out = np.zeros((2, P.m, 2), np.uint64)
    check(_lib.lib().sgfhe_flatten_poly(P.ctx, _ptr(a), _ptr(d), _ptr(out)))
    return out


def external_product(P: Params, rng_draws, a, b, A):
    """external_product(rng|nothing, a, b, A, Val(B), Val(2)) -- src/fhe.jl:519-530.  A: wide [4,2,m,2]."""
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    A = np.ascontiguousarray(A, np.uint64)
    d = None if rng_draws is None else np.ascontiguousarray(rng_draws, np.int64)
    oa, ob = np.zeros_like(a), np.zeros_like(b)
    check(_lib.lib().sgfhe_external_product(P.ctx, _ptr(a), _ptr(b), _ptr(A), _ptr(d), _ptr(oa), _ptr(ob)))
    return oa, ob


def bootstrap_trace(bkey: BootstrapKey, draws, lwe1, lwe2, n_steps: int | None = None, trace: bool = True):
    """_bootstrap_internal (src/fhe.jl:559-595) for one gate: outputs over Z_Q plus the accumulator after each step."""
    P = bkey.params
    bkey.upload()
    n_steps = bkey.rows if n_steps is None else n_steps
    lwe1 = np.ascontiguousarray(lwe1, np.uint64)
    lwe2 = np.ascontiguousarray(lwe2, np.uint64)
    d = None if draws is None else np.ascontiguousarray(draws, np.int64)
    tr = np.zeros((n_steps, 2, P.m, 2), np.uint64) if trace else None
    outs = [np.zeros((P.n + 1, 2), np.uint64) for _ in range(3)]
    check(_lib.lib().sgfhe_bootstrap_trace(P.ctx, _ptr(lwe1), _ptr(lwe2), _ptr(d), n_steps, _ptr(tr), *[_ptr(o) for o in outs]))
    return outs[0], outs[1], outs[2], tr


def launch_count() -> int:
    return int(_lib.lib().sgfhe_launch_count())


# ---- Scheme 2 (src/fhe2.jl, src/rns.jl): parameters and element arithmetic only (upstream has no bootstrap) ----------
class Scheme2Params:
    """Scheme2.Params(k) -- src/fhe2.jl:17-70"""

    def __init__(self, k: int):
        pc = _lib.Scheme2ParamsC()
        check(_lib.lib().sgfhe_scheme2_params_derive(int(k), C.byref(pc)))
        for f in ("n", "k", "t", "r", "m", "q", "tau", "B", "Bp", "Dr", "Dq"):
            setattr(self, f, getattr(pc, f))
        self.Q = self.B * self.Bp


def rns2_op(op: str, a, b, M1: int, M2: int, device: int = 0):
    """RNS2Number{UInt64, M1, M2} arithmetic, batched (src/rns.jl:51-60).  a, b: pairs (v1, v2) of uint64 arrays."""
    code = {"*": 0, "+": 1, "-": 2}[op]
    arrs = [np.ascontiguousarray(x, np.uint64) for x in (a[0], a[1], b[0], b[1])]
    o1, o2 = np.zeros_like(arrs[0]), np.zeros_like(arrs[0])
    check(_lib.lib().sgfhe_rns2_op(device, code, arrs[0].size, *[_ptr(x) for x in arrs], M1, M2, _ptr(o1), _ptr(o2)))
    return o1, o2


class Scheme2Context:
    """Device context for Scheme2.Params(k) (src/fhe2.jl:36-70): polynomial arithmetic over RNS2Number and key generation.
    Upstream defines no bootstrap for this scheme (src/fhe2.jl:1-7), so neither does this."""

    def __init__(self, k: int, device: int = 0):
        self.params = Scheme2Params(k)
        self.device = device
        h = C.c_void_p()
        check(_lib.lib().sgfhe_s2_ctx_create(int(k), device, C.byref(h)))
        self._h = h

    def close(self):
        if self._h is not None:
            _lib.lib().sgfhe_s2_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def polymul(self, a, b):
        """a * b in (Z_B x Z_B')[x]/(x^m+1), batched: a, b are pairs (v1, v2) of uint64[batch, m]."""
        arrs = [np.ascontiguousarray(x, np.uint64) for x in (a[0], a[1], b[0], b[1])]
        if any(x.shape != arrs[0].shape for x in arrs) or arrs[0].ndim != 2 or arrs[0].shape[1] != self.params.m:
            raise SgfheError("operands must be pairs of [batch, m]")
        o1, o2 = np.zeros_like(arrs[0]), np.zeros_like(arrs[0])
        check(_lib.lib().sgfhe_s2_polymul(self._h, arrs[0].shape[0], *[_ptr(x) for x in arrs], _ptr(o1), _ptr(o2)))
        return o1, o2

    def bootstrap_key(self, sk: np.ndarray, a_rand: np.ndarray, e_rand: np.ndarray) -> np.ndarray:
        """Scheme2.BootstrapKey from pre-drawn randomness (src/fhe2.jl:119-126): uint64[rows, 4, 2, m, 2]."""
        a_rand = np.ascontiguousarray(a_rand, np.uint64)
        e_rand = np.ascontiguousarray(e_rand, np.int64)
        rows = a_rand.shape[0]
        out = np.zeros((rows, 4, 2, self.params.m, 2), np.uint64)
        skb = np.ascontiguousarray(sk, np.uint8)
        check(_lib.lib().sgfhe_s2_bkey_generate(self._h, _ptr(skb), _ptr(a_rand), _ptr(e_rand), 0, rows, _ptr(out)))
        return out


def scheme2_bootstrap_key(ctx: Scheme2Context, rng: np.random.Generator, sk: np.ndarray, rows: int | None = None) -> np.ndarray:
    """Scheme2.BootstrapKey(rng, sk) -- src/fhe2.jl:104-131: per row a_1..a_4 uniform below Q = B B', then e_1..e_4 uniform on
    [-tau, tau], drawn here in the reference's order; the arithmetic runs on the device."""
    S = ctx.params
    rows = S.n if rows is None else rows
    a = np.zeros((rows, 4, S.m, 2), np.uint64)
    e = np.zeros((rows, 4, S.m), np.int64)
    for i in range(rows):
        a[i] = _rand_below(rng, S.Q, (4, S.m))                                     # src/fhe2.jl:122
        e[i] = rng.integers(-S.tau, S.tau + 1, size=(4, S.m), dtype=np.int64)      # src/fhe2.jl:123
    return ctx.bootstrap_key(sk, a, e)
