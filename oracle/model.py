"""Independent big-integer model of the SGFHE.jl bootstrapping path (TEST INFRASTRUCTURE ONLY).

A second, deliberately different restatement of the reference's algorithm, used to cross-check
oracle/sgfhe_oracle.c: Python integers, Kronecker-substitution polynomial products (no NTT, no
Montgomery), sympy's primality test.  Parity status: unpinned by upstream vectors (none exist);
see oracle/sgfhe_oracle.h.  Only tests/ may import this module.

Citations are reference file:line (paths under /root/reference).
"""
from __future__ import annotations

from dataclasses import dataclass

try:
    from sympy import isprime as _isprime
except Exception:  # pragma: no cover - sympy is in the image; keep a fallback anyway
    def _isprime(x: int) -> bool:
        if x < 2:
            return False
        for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            if x % p == 0:
                return x == p
        d, s = x - 1, 0
        while d % 2 == 0:
            d //= 2
            s += 1
        for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71):
            y = pow(a, d, x)
            if y in (1, x - 1):
                continue
            for _ in range(s - 1):
                y = y * y % x
                if y == x - 1:
                    break
            else:
                return False
        return True


def find_modulus(n: int, qmin: int, qmax: int | None = None) -> int:
    """utils.jl:7-28"""
    j = -((-(qmin - 1)) // n)  # cld
    while True:
        q = j * n + 1
        if qmax is not None and q > qmax:
            break
        if _isprime(q):
            return q
        j += 1
    raise ValueError("Could not find a modulus")


@dataclass(frozen=True)
class Params:
    """fhe.jl:27-99"""
    n: int
    r: int
    q: int
    Q: int
    t: int
    m: int
    B: int
    Dr: int
    Dq: int
    DQ: int


def params(n: int) -> Params:
    assert n >= 64 and n & (n - 1) == 0            # fhe.jl:45-46
    r = 16 * n                                     # fhe.jl:53
    q = find_modulus(2 * n, r * n)                 # fhe.jl:57
    t = r.bit_length() - 2                         # fhe.jl:61
    m = r // 2                                     # fhe.jl:62
    Q = find_modulus(2 * m, r ** 4 * n ** 2 * 1220, r ** 4 * n ** 2 * 1225)   # fhe.jl:64-69
    return Params(n, r, q, Q, t, m, r * r * n * 35, r // 4, q // 4, Q // 8)   # fhe.jl:87-90


def rescale(new_max: int, x: int, old_max: int, round_result: bool) -> int:
    """utils.jl:78-92"""
    q, rem = divmod(x * new_max, old_max)
    if round_result and rem >= old_max // 2 + (old_max & 1):
        q += 1
        if q == new_max:
            q = 0
    return q


def rescale_ref(new_max: int, x: int, old_max: int, round_result: bool) -> int:
    """test/internals.test.jl:6-20 (exact rational instead of BigFloat)"""
    num = x * new_max
    if round_result:
        res = (2 * num + old_max) // (2 * old_max)   # round half up; ties: see test (odd old_max has none)
        if res == new_max:
            res = 0
        return res
    return num // old_max


def flatten(a: int, B: int, l: int, Q: int, draws=None) -> list[int]:
    """utils.jl:155-189 (draws None) / utils.jl:198-241"""
    x = [0] * l
    if draws is not None:
        x = [int(d) % Q for d in draws]            # utils.jl:229
        for i in range(l):
            a = (a - x[i] * B ** i) % Q            # utils.jl:222,232-233
    s = (B - 1) // 2 if B & 1 else B // 2 - 1      # utils.jl:162-166
    a = (a + s * sum(B ** i for i in range(l))) % Q   # utils.jl:169,179
    out = [0] * l
    for i in range(l - 1, 0, -1):                  # utils.jl:170-176
        out[i], a = divmod(a, B ** i)
    out[0] = a
    return [(o - s + xi) % Q for o, xi in zip(out, x)]   # utils.jl:183-185, 236-238


def flatten_poly(a: list[int], B: int, l: int, Q: int, draws=None) -> list[list[int]]:
    """utils.jl:253-264; draws[j][i]"""
    res = [[0] * len(a) for _ in range(l)]
    for j, c in enumerate(a):
        d = flatten(c, B, l, Q, None if draws is None else draws[j])
        for i in range(l):
            res[i][j] = d[i]
    return res


def polymul(a: list[int], b: list[int], Q: int) -> list[int]:
    """DarkIntegers `Polynomial *` (not in tree), negacyclic; Kronecker substitution."""
    N = len(a)
    slot = (2 * Q.bit_length() + N.bit_length() + 8) // 8 + 1
    pa = int.from_bytes(b"".join(int(c).to_bytes(slot, "little") for c in a), "little")
    pb = int.from_bytes(b"".join(int(c).to_bytes(slot, "little") for c in b), "little")
    raw = (pa * pb).to_bytes(2 * N * slot, "little")
    c = [int.from_bytes(raw[i * slot:(i + 1) * slot], "little") for i in range(2 * N)]
    return [(c[i] - c[i + N]) % Q for i in range(N)]


def mul_by_monomial(p: list[int], shift: int, Q: int) -> list[int]:
    """DarkIntegers mul_by_monomial (used fhe.jl:555,573)"""
    N = len(p)
    s = shift % (2 * N)
    out = [0] * N
    for i, c in enumerate(p):
        k = i + s
        sign = 1
        while k >= N:
            k -= N
            sign = -sign
        out[k] = (sign * c) % Q
    return out


def initial_poly(P: Params) -> list[int]:
    """fhe.jl:535-548"""
    coeffs = [0] * P.m
    for i in range(-(P.Dr - 1), P.Dr):
        coeffs[i % P.m] += 1 if (i // P.m) % 2 == 0 else -1
    return [c % P.Q for c in coeffs]


def extract(a: list[int], i: int, n: int, Q: int) -> list[int]:
    """fhe.jl:237-244, i 1-based"""
    N = len(a)
    if i < n:
        return [a[j - 1] for j in range(i, 0, -1)] + [(-a[j - 1]) % Q for j in range(N, N - (n - i - 1) - 1, -1)]
    return [a[j - 1] for j in range(i, i - n, -1)]


def bkey_generate(P: Params, sk, a_rand, e_rand, rows=None):
    """fhe.jl:181-201.  a_rand[i][j], e_rand[i][j]: lists of m ints.  key[i][j][c] = poly."""
    ext = [int(b) for b in sk] + [0] * (P.m - P.n)  # fhe.jl:185
    key = []
    for i in (range(P.n) if rows is None else rows):
        C = []
        for j in range(4):
            aj = [int(v) for v in a_rand[i][j]]
            bj = [(v + int(e)) % P.Q for v, e in zip(polymul(aj, ext, P.Q), e_rand[i][j])]   # fhe.jl:195
            if sk[i]:                              # fhe.jl:196, G = [1 0; B 0; 0 1; 0 B] fhe.jl:119-122
                g = P.B if j & 1 else 1
                if j < 2:
                    aj[0] = (aj[0] + g) % P.Q
                else:
                    bj[0] = (bj[0] + g) % P.Q
            C.append([aj, bj])
        key.append(C)
    return key


def external_product(a, b, A, B, Q, draws=None):
    """fhe.jl:519-530.  A[j][c]; draws[2][N][2] or None."""
    u = flatten_poly(a, B, 2, Q, None if draws is None else draws[0]) + \
        flatten_poly(b, B, 2, Q, None if draws is None else draws[1])
    N = len(a)
    res = []
    for c in range(2):
        acc = [0] * N
        for j in range(4):
            pr = polymul(u[j], A[j][c], Q)
            acc = [(x + y) % Q for x, y in zip(acc, pr)]
        res.append(acc)
    return res[0], res[1]


def bootstrap_internal(P: Params, key, lwe1, lwe2, draws=None, n_steps=None, trace=None):
    """fhe.jl:559-595, literal.  lwe = n+1 ints (a..., b).  Returns three LWEs over Z_Q."""
    n, m, Q = P.n, P.m, P.Q
    u = [(int(x) + int(y)) % P.r for x, y in zip(lwe1, lwe2)]   # fhe.jl:566
    a = [0] * m                                     # fhe.jl:570
    b = [c * P.DQ % Q for c in mul_by_monomial(initial_poly(P), -u[n], Q)]   # fhe.jl:572-573
    steps = n if n_steps is None else n_steps
    for k in range(steps):                          # fhe.jl:579
        A = []
        for j in range(4):
            row = []
            for c in range(2):
                p = key[k][j][c]
                x = [(r - s) % Q for r, s in zip(mul_by_monomial(p, u[k], Q), p)]   # fhe.jl:554-556
                if c == j // 2:
                    x[0] = (x[0] + (P.B if j & 1 else 1)) % Q   # .+ G  fhe.jl:580
                row.append(x)
            A.append(row)
        a, b = external_product(a, b, A, P.B, Q, None if draws is None else draws[k])   # fhe.jl:581
        if trace is not None:
            trace.append((list(a), list(b)))
    i_and, i_or = 3 * m // 4 + 1, m // 4 + 1
    l_and = extract(a, i_and, n, Q) + [(P.DQ + b[i_and - 1]) % Q]            # fhe.jl:585-587
    l_or = [(-v) % Q for v in extract(a, i_or, n, Q)] + [(P.DQ - b[i_or - 1]) % Q]   # fhe.jl:588-590
    l_xor = [(x - y) % Q for x, y in zip(l_or, l_and)]                       # fhe.jl:592
    return l_and, l_or, l_xor


def bootstrap(P: Params, key, lwe1, lwe2, draws=None):
    """fhe.jl:608-621"""
    return tuple([rescale(P.r, v, P.Q, True) for v in l]          # utils.jl:114-116
                 for l in bootstrap_internal(P, key, lwe1, lwe2, draws))


def encrypt_private(P: Params, sk, a, w, message):
    """fhe.jl:310-328 with expanded `a` and noise `w` supplied; everything over Z_r."""
    s = [int(x) for x in sk]
    prod = polymul([int(x) for x in a], s, 1 << 80)   # exact negacyclic integers first ...
    step = 1 << (P.t - 4)
    out = []
    for k in range(P.n):
        v = prod[k]
        if v >= 1 << 79:
            v -= 1 << 80
        v = (v + int(w[k]) + (P.Dr if message[k] else 0)) % P.r   # fhe.jl:322
        out.append(v // step * step)                              # fhe.jl:325
    return out


def split_ciphertext(P: Params, a, b):
    """fhe.jl:287-290"""
    return [extract([int(x) for x in a], i, P.n, P.r) + [int(b[i - 1])] for i in range(1, P.n + 1)]


def decrypt_lwe(P: Params, sk, lwe) -> int:
    """fhe.jl:504-507"""
    b1 = (int(lwe[P.n]) - sum(int(x) for x, s in zip(lwe, sk) if s)) % P.r
    return ((b1 + P.Dr // 2) % P.r) // P.Dr


def shortened_external_product(a, A, B, Q, draws=None):
    """fhe.jl:632-641.  A[j][c]; uses rows l+1..2l (0-based 2, 3); draws [N][2] or None."""
    u = flatten_poly(a, B, 2, Q, draws)
    N = len(a)
    res = []
    for c in range(2):
        acc = [0] * N
        for j in range(2):
            acc = [(x + y) % Q for x, y in zip(acc, polymul(u[j], A[2 + j][c], Q))]
        res.append(acc)
    return res[0], res[1]


def pack_from_lwes(P: Params, key, new_lwes, draws_short=None):
    """fhe.jl:675-693 given new_lwes[j] = (a..., b) over Z_Q; key[i][j][c]; returns (w, v) over Z_r, length m."""
    n, m, Q = P.n, P.m, P.Q
    w_t, v_t = [0] * m, [0] * m
    for i in range(n):
        as_i = [int(new_lwes[j][i]) for j in range(n)] + [0] * (m - n)          # fhe.jl:675-677
        w, v = shortened_external_product(as_i, key[i], P.B, Q, None if draws_short is None else draws_short[i])
        w_t = [(x + y) % Q for x, y in zip(w_t, w)]
        v_t = [(x + y) % Q for x, y in zip(v_t, v)]
    b = [int(new_lwes[j][n]) for j in range(n)] + [0] * (m - n)                 # fhe.jl:678
    w1 = [(-x) % Q for x in w_t]                                                # fhe.jl:689
    v1 = [(x - y) % Q for x, y in zip(b, v_t)]                                  # fhe.jl:690
    return [rescale(P.r, x, Q, True) for x in w1], [rescale(P.r, x, Q, True) for x in v1]
