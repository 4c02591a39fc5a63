/*
 * sgfhe_oracle.c -- CPU oracle for the bootstrapping hot path of nucypher/SGFHE.jl.
 *
 * TEST INFRASTRUCTURE ONLY (see sgfhe_oracle.h).  PARITY STATUS: "parity unpinned" by the
 * reference's own tests (no golden vectors upstream, Julia + DarkIntegers absent here);
 * pinned instead by exact-ring uniqueness, the ported reference properties in tests/, and the
 * independent big-integer model oracle/model.py.
 *
 * Every function cites the reference file:line it restates (paths under /root/reference).
 * Nothing here is copied: the reference is Julia on top of DarkIntegers; this is plain C on
 * unsigned __int128 with a 128-bit Montgomery core.
 */
#include "sgfhe_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef unsigned __int128 u128;
typedef __int128 i128;

static inline u128 U(sgo_u128 x) { return ((u128)x.hi << 64) | x.lo; }
static inline sgo_u128 S(u128 x) { sgo_u128 r; r.lo = (uint64_t)x; r.hi = (uint64_t)(x >> 64); return r; }

/* ------------------------------------------------------------------------------------------
 * Arithmetic in Z_Q, Q < 2^127 odd: stands in for DarkIntegers' MgModUInt/ModUInt (not in tree;
 * selected at fhe.jl:83-85,104).  Only canonical values cross function boundaries.
 * ------------------------------------------------------------------------------------------ */
static inline void mul_wide(u128 a, u128 b, u128* hi, u128* lo) {
  uint64_t a0 = (uint64_t)a, a1 = (uint64_t)(a >> 64), b0 = (uint64_t)b, b1 = (uint64_t)(b >> 64);
  u128 p00 = (u128)a0 * b0, p01 = (u128)a0 * b1, p10 = (u128)a1 * b0, p11 = (u128)a1 * b1;
  u128 mid = (p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;
  *lo = (u128)(uint64_t)p00 | (mid << 64);
  *hi = p11 + (p01 >> 64) + (p10 >> 64) + (mid >> 64);
}

/* 256-by-128 division, bit serial; only used off the hot loop (rescale, setup). hi < d required. */
static void divrem_wide(u128 hi, u128 lo, u128 d, u128* q, u128* r) {
  u128 rem = hi, quo = 0;
  for (int i = 127; i >= 0; --i) {
    int top = (int)(rem >> 127);
    rem = (rem << 1) | ((lo >> i) & 1);
    if (top || rem >= d) { rem -= d; quo |= (u128)1 << i; }
  }
  *q = quo; *r = rem;
}

static inline u128 addmod(u128 a, u128 b, u128 Q) { u128 s = a + b; return s >= Q ? s - Q : s; }
static inline u128 submod(u128 a, u128 b, u128 Q) { return a >= b ? a - b : a + Q - b; }
static inline u128 negmod(u128 a, u128 Q) { return a ? Q - a : 0; }

static u128 mulmod_slow(u128 a, u128 b, u128 Q) {
  u128 hi, lo, q, r; mul_wide(a, b, &hi, &lo);
  if (hi == 0) return lo % Q;
  divrem_wide(hi % Q, lo, Q, &q, &r); return r;
}
static u128 powmod_slow(u128 a, u128 e, u128 Q) {
  u128 r = 1 % Q; a %= Q;
  while (e) { if (e & 1) r = mulmod_slow(r, a, Q); a = mulmod_slow(a, a, Q); e >>= 1; }
  return r;
}

typedef struct { u128 Q, ninv, r2, one; } mctx;   /* Montgomery context, R = 2^128 */

static void mctx_init(mctx* c, u128 Q) {
  c->Q = Q;
  u128 inv = Q;                                  /* Newton: inv = Q^-1 mod 2^128 */
  for (int i = 0; i < 7; ++i) inv *= 2 - Q * inv;
  c->ninv = (u128)0 - inv;
  u128 q, r; divrem_wide(1, 0, Q, &q, &r);       /* 2^128 mod Q */
  c->one = r;
  c->r2 = mulmod_slow(r, r, Q);
}
static inline u128 mmul(const mctx* c, u128 a, u128 b) {
  u128 hi, lo, mh, ml; mul_wide(a, b, &hi, &lo);
  u128 m = lo * c->ninv; mul_wide(m, c->Q, &mh, &ml);
  u128 t = hi + mh + (lo != 0);
  return t >= c->Q ? t - c->Q : t;
}
static inline u128 to_m(const mctx* c, u128 a) { return mmul(c, a, c->r2); }
static inline u128 from_m(const mctx* c, u128 a) { return mmul(c, a, 1); }

/* ------------------------------------------------------------------------------------------
 * Primes.isprime stand-in (used at utils.jl:19): Miller-Rabin with the first 40 primes as bases
 * (deterministic for every modulus this scheme can produce; cross-checked by oracle/model.py).
 * ------------------------------------------------------------------------------------------ */
int sgo_is_prime(sgo_u128 xs) {
  static const int small[] = {2,3,5,7,11,13,17,19,23,29,31,37,41,43,47,53,59,61,67,71,73,79,83,89,97,
                              101,103,107,109,113,127,131,137,139,149,151,157,163,167,173};
  u128 x = U(xs);
  if (x < 2) return 0;
  for (int i = 0; i < 40; ++i) { if (x == (u128)small[i]) return 1; if (x % small[i] == 0) return 0; }
  u128 d = x - 1; int s = 0;
  while (!(d & 1)) { d >>= 1; ++s; }
  for (int i = 0; i < 40; ++i) {
    u128 y = powmod_slow(small[i], d, x);
    if (y == 1 || y == x - 1) continue;
    int comp = 1;
    for (int k = 1; k < s; ++k) { y = mulmod_slow(y, y, x); if (y == x - 1) { comp = 0; break; } }
    if (comp) return 0;
  }
  return 1;
}

/* utils.jl:7-28 */
int sgo_find_modulus(sgo_u128 ns, sgo_u128 qmins, sgo_u128 qmaxs, sgo_u128* out) {
  u128 n = U(ns), qmin = U(qmins), qmax = U(qmaxs);
  u128 j = (qmin - 1 + n - 1) / n;               /* cld(qmin - 1, n)            utils.jl:10 */
  for (;;) {
    u128 q = j * n + 1;                          /* utils.jl:13 */
    if (qmax != 0 && q > qmax) break;            /* utils.jl:15-17 */
    if (sgo_is_prime(S(q))) { *out = S(q); return 0; }
    ++j;
  }
  return -1;                                     /* utils.jl:26 error(...) */
}

/* fhe.jl:43-97 */
int sgo_params_init(int n, sgo_params* P) {
  if (n < 64 || (n & (n - 1))) return -1;        /* fhe.jl:45-46 */
  if (n > 2048) return -2;                       /* Q must stay below 2^127 here (fhe.jl:71-78) */
  memset(P, 0, sizeof *P);
  u128 bn = (u128)n, r = bn * 16;                /* fhe.jl:53 */
  sgo_u128 q;
  if (sgo_find_modulus(S(2 * bn), S(r * bn), S(0), &q)) return -3;   /* fhe.jl:57 */
  int lg = 0; while (((u128)1 << lg) < r) ++lg;
  u128 m = r / 2;                                /* fhe.jl:62 */
  u128 r4n2 = r * r * r * r * bn * bn;
  sgo_u128 Q;
  if (sgo_find_modulus(S(2 * m), S(r4n2 * 1220), S(r4n2 * 1225), &Q)) return -3;   /* fhe.jl:64-69 */
  P->n = n; P->t = lg - 1; P->m = (int32_t)m;    /* fhe.jl:61 */
  P->large = (U(Q) >> 64) != 0;                  /* fhe.jl:71-78 */
  P->r = (uint64_t)r; P->q = q.lo;
  P->Dr = (uint64_t)(r / 4); P->Dq = q.lo / 4;   /* fhe.jl:88-89 */
  P->Q = Q; P->B = S(r * r * bn * 35);           /* fhe.jl:87 */
  P->DQ = S(U(Q) / 8);                           /* fhe.jl:90 */
  return 0;
}

/* utils.jl:78-92 */
sgo_u128 sgo_rescale(sgo_u128 new_max, sgo_u128 xs, sgo_u128 old_maxs, int round_result) {
  u128 nm = U(new_max), x = U(xs), om = U(old_maxs), hi, lo, q, r;
  mul_wide(x, nm, &hi, &lo);                     /* utils.jl:81 mulhilo */
  if (hi == 0) { q = lo / om; r = lo % om; }     /* utils.jl:82 divremhilo */
  else divrem_wide(hi, lo, om, &q, &r);
  if (round_result) {                            /* utils.jl:83-90 */
    if (r >= om / 2 + (om & 1)) { q += 1; if (q == nm) q = 0; }
  }
  return S(q);
}

/* utils.jl:155-189 / 198-241 */
static void flatten_u(u128 a, u128 B, int l, u128 Q, const int64_t* draws, u128* out) {
  u128 pw[8]; pw[0] = 1; for (int i = 1; i < l; ++i) pw[i] = pw[i - 1] * B;
  u128 x[8];
  if (draws) {                                   /* utils.jl:227-233 */
    for (int i = 0; i < l; ++i) {
      x[i] = draws[i] >= 0 ? (u128)draws[i] % Q : negmod((u128)(-(i128)draws[i]) % Q, Q);
      a = submod(a, mulmod_slow(x[i], pw[i] % Q, Q), Q);
    }
  }
  u128 s = (B & 1) ? (B - 1) / 2 : B / 2 - 1;    /* utils.jl:162-166 */
  u128 sumpw = 0; for (int i = 0; i < l; ++i) sumpw += pw[i];
  u128 offset = mulmod_slow(sumpw % Q, s % Q, Q);
  a = addmod(a, offset, Q);                      /* utils.jl:179  a += offset (in Z_Q) */
  for (int i = l - 1; i >= 1; --i) {             /* utils.jl:170-176  r, a = divrem(a, B^(i-1)) */
    out[i] = a / pw[i]; a = a % pw[i];
  }
  out[0] = a;                                    /* utils.jl:181 */
  for (int i = 0; i < l; ++i) out[i] = submod(out[i] % Q, s % Q, Q);   /* utils.jl:183-185 */
  if (draws) for (int i = 0; i < l; ++i) out[i] = addmod(out[i], x[i], Q);   /* utils.jl:236-238 */
}

void sgo_flatten(sgo_u128 a, sgo_u128 B, int l, sgo_u128 q, const int64_t* draws, sgo_u128* out) {
  u128 o[8]; flatten_u(U(a), U(B), l, U(q), draws, o);
  for (int i = 0; i < l; ++i) out[i] = S(o[i]);
}

/* utils.jl:253-264 */
void sgo_flatten_poly(const sgo_u128* a, int N, sgo_u128 B, int l, sgo_u128 q,
                      const int64_t* draws, sgo_u128* out) {
  u128 o[8];
  for (int j = 0; j < N; ++j) {                  /* utils.jl:257 coefficient order */
    flatten_u(U(a[j]), U(B), l, U(q), draws ? draws + (size_t)j * l : NULL, o);
    for (int i = 0; i < l; ++i) out[(size_t)i * N + j] = S(o[i]);     /* utils.jl:259-261 */
  }
}

/* ------------------------------------------------------------------------------------------
 * Negacyclic product in Z_Q[x]/(x^N+1): DarkIntegers `Polynomial *` (not in tree; called at
 * fhe.jl:195,527-528,638-639).  Schoolbook and NTT forms; the ring fixes the result.
 * ------------------------------------------------------------------------------------------ */
void sgo_polymul_schoolbook(const sgo_u128* a, const sgo_u128* b, sgo_u128* out, int N, sgo_u128 Qs) {
  u128 Q = U(Qs); mctx c; mctx_init(&c, Q);
  u128* acc = (u128*)calloc(N, sizeof(u128));
  u128* bm = (u128*)malloc(N * sizeof(u128));
  for (int j = 0; j < N; ++j) bm[j] = to_m(&c, U(b[j]));
  for (int i = 0; i < N; ++i) {
    u128 ai = U(a[i]); if (!ai) continue;
    for (int j = 0; j < N; ++j) {
      u128 p = mmul(&c, ai, bm[j]);              /* canonical a * mont b -> canonical */
      int k = i + j;
      if (k < N) acc[k] = addmod(acc[k], p, Q); else acc[k - N] = submod(acc[k - N], p, Q);
    }
  }
  for (int i = 0; i < N; ++i) out[i] = S(acc[i]);
  free(acc); free(bm);
}

typedef struct {
  int N; u128 Q; mctx c;
  u128* fw;    /* psi^bitrev(i) in Montgomery form, i in [0,N)   (forward CT, merged twist) */
  u128* iw;    /* psi^-bitrev(i)                                  (inverse GS) */
  u128 ninv;   /* N^-1, Montgomery form */
} ntt_plan;

#define MAX_PLANS 16
static ntt_plan g_plans[MAX_PLANS];
static int g_nplans = 0;
static pthread_mutex_t g_plan_lock = PTHREAD_MUTEX_INITIALIZER;

/* minimal parallel-for over [0,count) with dynamic scheduling (pthreads; no OpenMP dependency) */
typedef struct { void (*fn)(long, void*); void* arg; long count; long next; pthread_mutex_t lock; } pfor_t;
static void* pfor_worker(void* vp) {
  pfor_t* pf = (pfor_t*)vp;
  for (;;) {
    pthread_mutex_lock(&pf->lock); long i = pf->next++; pthread_mutex_unlock(&pf->lock);
    if (i >= pf->count) break;
    pf->fn(i, pf->arg);
  }
  return NULL;
}
static void parallel_for(long count, int threads, void (*fn)(long, void*), void* arg) {
  if (threads < 1) threads = 1;
  if (threads > count) threads = (int)count;
  if (threads <= 1) { for (long i = 0; i < count; ++i) fn(i, arg); return; }
  pfor_t pf; pf.fn = fn; pf.arg = arg; pf.count = count; pf.next = 0; pthread_mutex_init(&pf.lock, NULL);
  pthread_t* th = (pthread_t*)malloc(threads * sizeof(pthread_t));
  for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, pfor_worker, &pf);
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  free(th); pthread_mutex_destroy(&pf.lock);
}
static int g_setup_threads = 8;   /* threads for one-off setup work (key generation / key transform) */
void sgo_set_setup_threads(int t) { g_setup_threads = t < 1 ? 1 : t; }

static int bitrev(int x, int bits) { int r = 0; for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; } return r; }

static const ntt_plan* get_plan(int N, u128 Q) {
  const ntt_plan* found = NULL;
  pthread_mutex_lock(&g_plan_lock);
  {
    for (int i = 0; i < g_nplans; ++i) if (g_plans[i].N == N && g_plans[i].Q == Q) found = &g_plans[i];
    if (!found && g_nplans < MAX_PLANS && N >= 2 && !(N & (N - 1)) && (Q - 1) % (2 * (u128)N) == 0) {
      ntt_plan* p = &g_plans[g_nplans];
      p->N = N; p->Q = Q; mctx_init(&p->c, Q);
      u128 psi = 0;
      for (u128 g = 2; g < 1000; ++g) {          /* psi^N = -1  <=> psi has order exactly 2N */
        u128 w = powmod_slow(g, (Q - 1) / (2 * (u128)N), Q);
        if (powmod_slow(w, N, Q) == Q - 1) { psi = w; break; }
      }
      if (psi) {
        int bits = 0; while ((1 << bits) < N) ++bits;
        u128 psi_inv = powmod_slow(psi, Q - 2, Q);
        p->fw = (u128*)malloc(N * sizeof(u128)); p->iw = (u128*)malloc(N * sizeof(u128));
        u128 f = 1, g2 = 1;
        u128* pf = (u128*)malloc(N * sizeof(u128)); u128* pi = (u128*)malloc(N * sizeof(u128));
        for (int i = 0; i < N; ++i) { pf[i] = f; pi[i] = g2; f = mulmod_slow(f, psi, Q); g2 = mulmod_slow(g2, psi_inv, Q); }
        for (int i = 0; i < N; ++i) { int r = bitrev(i, bits); p->fw[i] = to_m(&p->c, pf[r]); p->iw[i] = to_m(&p->c, pi[r]); }
        free(pf); free(pi);
        p->ninv = to_m(&p->c, powmod_slow((u128)N, Q - 2, Q));
        ++g_nplans; found = p;
      }
    }
  }
  pthread_mutex_unlock(&g_plan_lock);
  return found;
}

/* forward: natural order in, bit-reversed out; values may be canonical or Montgomery (linear) */
static void ntt_fwd(const ntt_plan* p, u128* x) {
  int N = p->N; u128 Q = p->Q; const mctx* c = &p->c;
  for (int len = N >> 1, mth = 1; len >= 1; len >>= 1, mth <<= 1)
    for (int i = 0; i < mth; ++i) {
      u128 w = p->fw[mth + i]; u128* a = x + 2 * i * len; u128* b = a + len;
      for (int j = 0; j < len; ++j) { u128 t = mmul(c, b[j], w); b[j] = submod(a[j], t, Q); a[j] = addmod(a[j], t, Q); }
    }
}
/* inverse: bit-reversed in, natural out, includes N^-1 */
static void ntt_inv(const ntt_plan* p, u128* x) {
  int N = p->N; u128 Q = p->Q; const mctx* c = &p->c;
  for (int len = 1, mth = N >> 1; len < N; len <<= 1, mth >>= 1)
    for (int i = 0; i < mth; ++i) {
      u128 w = p->iw[mth + i]; u128* a = x + 2 * i * len; u128* b = a + len;
      for (int j = 0; j < len; ++j) { u128 u = a[j], v = b[j]; a[j] = addmod(u, v, Q); b[j] = mmul(c, submod(u, v, Q), w); }
    }
  for (int j = 0; j < N; ++j) x[j] = mmul(c, x[j], p->ninv);
}

int sgo_polymul_ntt(const sgo_u128* a, const sgo_u128* b, sgo_u128* out, int N, sgo_u128 Qs) {
  u128 Q = U(Qs); const ntt_plan* p = get_plan(N, Q); if (!p) return -1;
  u128* fa = (u128*)malloc(N * sizeof(u128)); u128* fb = (u128*)malloc(N * sizeof(u128));
  for (int i = 0; i < N; ++i) { fa[i] = U(a[i]); fb[i] = to_m(&p->c, U(b[i])); }
  ntt_fwd(p, fa); ntt_fwd(p, fb);
  for (int i = 0; i < N; ++i) fa[i] = mmul(&p->c, fa[i], fb[i]);
  ntt_inv(p, fa);
  for (int i = 0; i < N; ++i) out[i] = S(fa[i]);
  free(fa); free(fb); return 0;
}

/* DarkIntegers mul_by_monomial (used fhe.jl:555,573): p * x^shift in Z_Q[x]/(x^N+1) */
static void monomial_u(const u128* p, int N, int64_t shift, u128 Q, u128* out) {
  int64_t s = shift % (2 * (int64_t)N); if (s < 0) s += 2 * N;
  for (int i = 0; i < N; ++i) {
    int64_t k = i + s; int neg = 0;
    while (k >= N) { k -= N; neg ^= 1; }
    out[k] = neg ? negmod(p[i], Q) : p[i];
  }
}
void sgo_mul_by_monomial(const sgo_u128* p, int N, int64_t shift, sgo_u128 Q, sgo_u128* out) {
  u128* a = (u128*)malloc(N * sizeof(u128)); u128* b = (u128*)malloc(N * sizeof(u128));
  for (int i = 0; i < N; ++i) a[i] = U(p[i]);
  monomial_u(a, N, shift, U(Q), b);
  for (int i = 0; i < N; ++i) out[i] = S(b[i]);
  free(a); free(b);
}

/* fhe.jl:535-548 */
static void initial_poly_u(const sgo_params* P, u128* out) {
  int m = P->m; u128 Q = U(P->Q);
  for (int i = 0; i < m; ++i) out[i] = 0;
  int64_t D = (int64_t)P->Dr;
  for (int64_t i = -(D - 1); i <= D - 1; ++i) {  /* fhe.jl:547 powers */
    int64_t md = ((i % m) + m) % m;              /* fhe.jl:540 mod(i, len) */
    int64_t fl = (i - md) / m;                   /* fld(i, len) */
    if (((fl % 2) + 2) % 2 == 0) out[md] = addmod(out[md], 1, Q); else out[md] = submod(out[md], 1, Q);
  }
}
void sgo_initial_poly(const sgo_params* P, sgo_u128* out) {
  u128* t = (u128*)malloc(P->m * sizeof(u128)); initial_poly_u(P, t);
  for (int i = 0; i < P->m; ++i) out[i] = S(t[i]);
  free(t);
}

/* fhe.jl:237-244 (i is 1-based) */
static void extract_u(const u128* a, int N, int i, int n, u128 Q, u128* out) {
  if (i < n) {                                   /* fhe.jl:239-240 */
    int k = 0;
    for (int j = i; j >= 1; --j) out[k++] = a[j - 1];
    for (int j = N; j >= N - (n - i - 1); --j) out[k++] = negmod(a[j - 1], Q);
  } else {                                       /* fhe.jl:242 */
    for (int k = 0; k < n; ++k) out[k] = a[i - 1 - k];
  }
}
void sgo_extract(const sgo_u128* a, int N, int i, int n, sgo_u128 modulus, sgo_u128* out) {
  u128* t = (u128*)malloc(N * sizeof(u128)); u128* o = (u128*)malloc(n * sizeof(u128));
  for (int k = 0; k < N; ++k) t[k] = U(a[k]);
  extract_u(t, N, i, n, U(modulus), o);
  for (int k = 0; k < n; ++k) out[k] = S(o[k]);
  free(t); free(o);
}

/* fhe.jl:181-201 */
typedef struct { const sgo_params* P; const ntt_plan* p; const uint8_t* sk; const sgo_u128* a_rand;
                 const int64_t* e_rand; int row0; sgo_u128* key_out; const u128* ek; } bkey_job;
static void bkey_row(long idx, void* vp) {
  bkey_job* J = (bkey_job*)vp; const ntt_plan* p = J->p;
  int m = J->P->m; u128 Q = U(J->P->Q), B = U(J->P->B);
  int i = J->row0 + (int)idx;
  u128* t = (u128*)malloc(m * sizeof(u128));
  for (int j = 0; j < 4; ++j) {
    const sgo_u128* aj = J->a_rand + ((size_t)i * 4 + j) * m;   /* fhe.jl:193 */
    const int64_t* ej = J->e_rand + ((size_t)i * 4 + j) * m;    /* fhe.jl:194 */
    sgo_u128* oa = J->key_out + (((size_t)idx * 4 + j) * 2 + 0) * m;
    sgo_u128* ob = J->key_out + (((size_t)idx * 4 + j) * 2 + 1) * m;
    for (int k = 0; k < m; ++k) t[k] = U(aj[k]);
    ntt_fwd(p, t);
    for (int k = 0; k < m; ++k) t[k] = mmul(&p->c, t[k], J->ek[k]);
    ntt_inv(p, t);                               /* a_j * ext_key                 fhe.jl:195 */
    for (int k = 0; k < m; ++k) {
      int64_t e = ej[k];
      u128 ev = e >= 0 ? (u128)e : Q - (u128)(-e);
      oa[k] = aj[k];
      ob[k] = S(addmod(t[k], ev, Q));            /* + e_j                         fhe.jl:195 */
    }
    if (J->sk[i]) {                              /* + s_i * G, G = [1 0; B 0; 0 1; 0 B]   fhe.jl:119-122,196 */
      u128 g = (j & 1) ? B % Q : 1;
      if (j < 2) oa[0] = S(addmod(U(oa[0]), g, Q)); else ob[0] = S(addmod(U(ob[0]), g, Q));
    }
  }
  free(t);
}
int sgo_bkey_generate(const sgo_params* P, const uint8_t* sk, const sgo_u128* a_rand,
                      const int64_t* e_rand, int row0, int row1, sgo_u128* key_out) {
  int m = P->m, n = P->n; u128 Q = U(P->Q);
  const ntt_plan* p = get_plan(m, Q); if (!p) return -1;
  u128* ek = (u128*)calloc(m, sizeof(u128));     /* fhe.jl:185 resize(sk, m), Montgomery+NTT */
  for (int i = 0; i < n; ++i) ek[i] = to_m(&p->c, sk[i] ? 1 : 0);
  ntt_fwd(p, ek);
  bkey_job J = {P, p, sk, a_rand, e_rand, row0, key_out, ek};
  parallel_for(row1 - row0, g_setup_threads, bkey_row, &J);
  free(ek);
  return 0;
}

/* flatten_poly on canonical u128 arrays: out[l][N] as signed-free residues */
static void flatten_poly_u(const u128* a, int N, u128 B, u128 Q, const int64_t* draws, u128* d0, u128* d1) {
  u128 o[2];
  for (int j = 0; j < N; ++j) {
    flatten_u(a[j], B, 2, Q, draws ? draws + (size_t)j * 2 : NULL, o);
    d0[j] = o[0]; d1[j] = o[1];
  }
}

/* one product via the plan, canonical in/out, fresh transforms of both operands (what every
 * `Polynomial * Polynomial` at fhe.jl:527-528 costs the reference) */
static void polymul_u(const ntt_plan* p, const u128* a, const u128* b, u128* out, u128* tmp) {
  int N = p->N;
  for (int i = 0; i < N; ++i) { out[i] = a[i]; tmp[i] = to_m(&p->c, b[i]); }
  ntt_fwd(p, out); ntt_fwd(p, tmp);
  for (int i = 0; i < N; ++i) out[i] = mmul(&p->c, out[i], tmp[i]);
  ntt_inv(p, out);
}

/* fhe.jl:519-530 on u128 arrays; A is [4][2][N] canonical */
static void external_product_u(const ntt_plan* p, const u128* a, const u128* b, const u128* A, u128 B,
                               const int64_t* draws, u128* a_out, u128* b_out, u128* work /* 6N */) {
  int N = p->N; u128 Q = p->Q;
  u128* u = work;                 /* [4][N]: a_decomp then b_decomp        fhe.jl:524-526 */
  u128* prod = work + 4 * (size_t)N; u128* tmp = work + 5 * (size_t)N;
  flatten_poly_u(a, N, B, Q, draws, u, u + N);
  flatten_poly_u(b, N, B, Q, draws ? draws + 2 * (size_t)N : NULL, u + 2 * (size_t)N, u + 3 * (size_t)N);
  for (int c = 0; c < 2; ++c) {
    u128* res = c ? b_out : a_out;
    for (int j = 0; j < 4; ++j) {                /* sum(u .* A[:,c])              fhe.jl:527-528 */
      polymul_u(p, u + (size_t)j * N, A + ((size_t)j * 2 + c) * N, prod, tmp);
      if (j == 0) memcpy(res, prod, N * sizeof(u128));
      else for (int k = 0; k < N; ++k) res[k] = addmod(res[k], prod[k], Q);
    }
  }
}

int sgo_external_product(const sgo_u128* a, const sgo_u128* b, const sgo_u128* A, int N,
                         sgo_u128 Bs, sgo_u128 Qs, const int64_t* draws,
                         sgo_u128* a_out, sgo_u128* b_out) {
  u128 Q = U(Qs); const ntt_plan* p = get_plan(N, Q); if (!p) return -1;
  size_t n = N;
  u128* buf = (u128*)malloc((2 + 8 + 2 + 6) * n * sizeof(u128));
  u128 *ua = buf, *ub = buf + n, *uA = buf + 2 * n, *oa = buf + 10 * n, *ob = buf + 11 * n, *work = buf + 12 * n;
  for (size_t i = 0; i < n; ++i) { ua[i] = U(a[i]); ub[i] = U(b[i]); }
  for (size_t i = 0; i < 8 * n; ++i) uA[i] = U(A[i]);
  external_product_u(p, ua, ub, uA, U(Bs), draws, oa, ob, work);
  for (size_t i = 0; i < n; ++i) { a_out[i] = S(oa[i]); b_out[i] = S(ob[i]); }
  free(buf); return 0;
}

/* fhe.jl:566-573: accumulator init */
static void acc_init_u(const sgo_params* P, const uint64_t* lwe1, const uint64_t* lwe2, uint64_t* ua,
                       u128* a, u128* b) {
  int n = P->n, m = P->m; u128 Q = U(P->Q), DQ = U(P->DQ);
  for (int i = 0; i <= n; ++i) ua[i] = (lwe1[i] + lwe2[i]) % P->r;   /* fhe.jl:566, LWE + fhe.jl:216 */
  u128* t = (u128*)malloc(m * sizeof(u128));
  initial_poly_u(P, t);                          /* fhe.jl:568 */
  for (int i = 0; i < m; ++i) a[i] = 0;          /* fhe.jl:570 */
  monomial_u(t, m, -(int64_t)ua[n], Q, b);       /* fhe.jl:572-573 */
  mctx c; mctx_init(&c, Q); u128 dq = to_m(&c, DQ);
  for (int i = 0; i < m; ++i) b[i] = mmul(&c, b[i], dq);
  free(t);
}

/* fhe.jl:585-592 */
static void assemble_u(const sgo_params* P, const u128* a, const u128* b,
                       sgo_u128* out_and, sgo_u128* out_or, sgo_u128* out_xor) {
  int n = P->n, m = P->m; u128 Q = U(P->Q), DQ = U(P->DQ);
  u128* e = (u128*)malloc((size_t)n * sizeof(u128));
  extract_u(a, m, 3 * m / 4 + 1, n, Q, e);       /* fhe.jl:586 */
  for (int k = 0; k < n; ++k) out_and[k] = S(e[k]);
  out_and[n] = S(addmod(DQ, b[3 * m / 4], Q));   /* fhe.jl:587 (1-based 3m/4+1) */
  extract_u(a, m, m / 4 + 1, n, Q, e);           /* fhe.jl:589 */
  for (int k = 0; k < n; ++k) out_or[k] = S(negmod(e[k], Q));
  out_or[n] = S(submod(DQ, b[m / 4], Q));        /* fhe.jl:590 */
  for (int k = 0; k <= n; ++k) out_xor[k] = S(submod(U(out_or[k]), U(out_and[k]), Q));   /* fhe.jl:592, LWE - fhe.jl:221 */
  free(e);
}

/* fhe.jl:559-595, literal */
int sgo_bootstrap_internal(const sgo_params* P, const sgo_u128* key, const uint64_t* lwe1,
                           const uint64_t* lwe2, const int64_t* draws, int n_steps,
                           sgo_u128* trace, sgo_u128* out_and, sgo_u128* out_or, sgo_u128* out_xor) {
  int n = P->n, m = P->m; u128 Q = U(P->Q), B = U(P->B);
  if (n_steps < 0 || n_steps > n) return -2;
  const ntt_plan* p = get_plan(m, Q); if (!p) return -1;
  size_t M = m;
  uint64_t* ua = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
  u128* buf = (u128*)malloc((2 + 2 + 8 + 2 + 6) * M * sizeof(u128));
  u128 *a = buf, *b = buf + M, *a2 = buf + 2 * M, *b2 = buf + 3 * M, *A = buf + 4 * M, *rot = buf + 12 * M,
       *work = buf + 14 * M;
  acc_init_u(P, lwe1, lwe2, ua, a, b);
  for (int k = 0; k < n_steps; ++k) {            /* fhe.jl:579 */
    const sgo_u128* C = key + (size_t)k * 8 * M;
    for (int j = 0; j < 4; ++j)
      for (int c = 0; c < 2; ++c) {              /* A = (x^u - 1) C + G           fhe.jl:580, 554-556 */
        u128* Ajc = A + ((size_t)j * 2 + c) * M; const sgo_u128* Cjc = C + ((size_t)j * 2 + c) * M;
        for (size_t i = 0; i < M; ++i) rot[M + i] = U(Cjc[i]);
        monomial_u(rot + M, m, (int64_t)ua[k], Q, rot);
        for (size_t i = 0; i < M; ++i) Ajc[i] = submod(rot[i], rot[M + i], Q);
        if (c == (j >> 1)) Ajc[0] = addmod(Ajc[0], (j & 1) ? B % Q : 1, Q);
      }
    external_product_u(p, a, b, A, B, draws ? draws + (size_t)k * 4 * M : NULL, a2, b2, work);   /* fhe.jl:581 */
    memcpy(a, a2, M * sizeof(u128)); memcpy(b, b2, M * sizeof(u128));
    if (trace) for (size_t i = 0; i < M; ++i) { trace[((size_t)k * 2) * M + i] = S(a[i]); trace[((size_t)k * 2 + 1) * M + i] = S(b[i]); }
  }
  assemble_u(P, a, b, out_and, out_or, out_xor);
  free(buf); free(ua);
  return 0;
}

/* Rewritten form: (a,b) += (x^u - 1) * ([flatten(a); flatten(b)] . C^(k)),
 * using sum_i u_i G_i = (a,b) (test/internals.test.jl:144-166).  Bit-identical to the above. */
typedef struct { const sgo_params* P; const ntt_plan* p; u128* keyhat; int rows; const sgo_u128* src; } fast_key;
static void fast_key_poly(long poly, void* vp) {
  fast_key* fk = (fast_key*)vp; int m = fk->P->m;
  u128* t = fk->keyhat + (size_t)poly * m;
  for (int i = 0; i < m; ++i) t[i] = to_m(&fk->p->c, U(fk->src[(size_t)poly * m + i]));
  ntt_fwd(fk->p, t);
}

static int fast_key_build(fast_key* fk, const sgo_params* P, const sgo_u128* key, int rows) {
  int m = P->m; u128 Q = U(P->Q);
  fk->P = P; fk->p = get_plan(m, Q); fk->rows = rows; if (!fk->p) return -1;
  size_t total = (size_t)rows * 8 * m;
  fk->keyhat = (u128*)malloc(total * sizeof(u128));
  if (!fk->keyhat) return -4;
  fk->src = key;
  parallel_for((long)rows * 8, g_setup_threads, fast_key_poly, fk);
  return 0;
}

static void fast_gate(const fast_key* fk, const uint64_t* lwe1, const uint64_t* lwe2, const int64_t* draws,
                      int n_steps, sgo_u128* trace, sgo_u128* out_and, sgo_u128* out_or, sgo_u128* out_xor) {
  const sgo_params* P = fk->P; const ntt_plan* p = fk->p;
  int n = P->n, m = P->m; u128 Q = U(P->Q), B = U(P->B); size_t M = m;
  uint64_t* ua = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
  u128* buf = (u128*)malloc(10 * M * sizeof(u128));
  u128 *a = buf, *b = buf + M, *d = buf + 2 * M, *pa = buf + 6 * M, *pb = buf + 7 * M, *ra = buf + 8 * M, *rb = buf + 9 * M;
  acc_init_u(P, lwe1, lwe2, ua, a, b);
  for (int k = 0; k < n_steps; ++k) {
    const u128* C = fk->keyhat + (size_t)k * 8 * M;
    const int64_t* dr = draws ? draws + (size_t)k * 4 * M : NULL;
    flatten_poly_u(a, m, B, Q, dr, d, d + M);
    flatten_poly_u(b, m, B, Q, dr ? dr + 2 * M : NULL, d + 2 * M, d + 3 * M);
    for (int j = 0; j < 4; ++j) ntt_fwd(p, d + (size_t)j * M);
    for (size_t i = 0; i < M; ++i) {
      u128 sa = 0, sb = 0;
      for (int j = 0; j < 4; ++j) {
        sa = addmod(sa, mmul(&p->c, d[(size_t)j * M + i], C[((size_t)j * 2) * M + i]), Q);
        sb = addmod(sb, mmul(&p->c, d[(size_t)j * M + i], C[((size_t)j * 2 + 1) * M + i]), Q);
      }
      pa[i] = sa; pb[i] = sb;
    }
    ntt_inv(p, pa); ntt_inv(p, pb);
    monomial_u(pa, m, (int64_t)ua[k], Q, ra); monomial_u(pb, m, (int64_t)ua[k], Q, rb);
    for (size_t i = 0; i < M; ++i) {
      a[i] = addmod(a[i], submod(ra[i], pa[i], Q), Q);
      b[i] = addmod(b[i], submod(rb[i], pb[i], Q), Q);
    }
    if (trace) for (size_t i = 0; i < M; ++i) { trace[((size_t)k * 2) * M + i] = S(a[i]); trace[((size_t)k * 2 + 1) * M + i] = S(b[i]); }
  }
  assemble_u(P, a, b, out_and, out_or, out_xor);
  free(buf); free(ua);
}

int sgo_bootstrap_internal_fast(const sgo_params* P, const sgo_u128* key, const uint64_t* lwe1,
                                const uint64_t* lwe2, const int64_t* draws, int n_steps,
                                sgo_u128* trace, sgo_u128* out_and, sgo_u128* out_or, sgo_u128* out_xor) {
  if (n_steps < 0 || n_steps > P->n) return -2;
  fast_key fk; int rc = fast_key_build(&fk, P, key, n_steps); if (rc) return rc;
  fast_gate(&fk, lwe1, lwe2, draws, n_steps, trace, out_and, out_or, out_xor);
  free(fk.keyhat); return 0;
}

/* fhe.jl:644-648 + utils.jl:107-117 on one LWE of n+1 elements */
static void modred_lwe(const sgo_params* P, const sgo_u128* in, uint64_t* out) {
  sgo_u128 r = {P->r, 0};
  for (int k = 0; k <= P->n; ++k) out[k] = sgo_rescale(r, in[k], P->Q, 1).lo;    /* utils.jl:115 round */
}

/* fhe.jl:608-621 */
int sgo_bootstrap(const sgo_params* P, const sgo_u128* key, const uint64_t* lwe1,
                  const uint64_t* lwe2, const int64_t* draws,
                  uint64_t* out_and, uint64_t* out_or, uint64_t* out_xor) {
  int n = P->n;
  sgo_u128* t = (sgo_u128*)malloc(3 * (size_t)(n + 1) * sizeof(sgo_u128));
  int rc = sgo_bootstrap_internal(P, key, lwe1, lwe2, draws, n, NULL, t, t + (n + 1), t + 2 * (n + 1));   /* fhe.jl:614 */
  if (!rc) { modred_lwe(P, t, out_and); modred_lwe(P, t + (n + 1), out_or); modred_lwe(P, t + 2 * (n + 1), out_xor); }   /* fhe.jl:616-618 */
  free(t); return rc;
}

typedef struct { const sgo_params* P; const sgo_u128* key; const fast_key* fk; const uint64_t *lwe1, *lwe2;
                 int n_steps, literal; uint64_t *oand, *oor, *oxor; int err; } batch_job;
static void batch_gate(long g, void* vp);

int sgo_bootstrap_batch(const sgo_params* P, const sgo_u128* key, int batch, const uint64_t* lwe1,
                        const uint64_t* lwe2, int n_steps, int literal, int threads,
                        uint64_t* out_and, uint64_t* out_or, uint64_t* out_xor) {
  int n = P->n;
  if (n_steps < 0 || n_steps > n) return -2;
  if (!get_plan(P->m, U(P->Q))) return -1;
  fast_key fk; fk.keyhat = NULL;
  if (!literal) { int rc = fast_key_build(&fk, P, key, n_steps); if (rc) return rc; }
  batch_job J = {P, key, &fk, lwe1, lwe2, n_steps, literal, out_and, out_or, out_xor, 0};
  parallel_for(batch, threads, batch_gate, &J);
  if (!literal) free(fk.keyhat);
  return J.err;
}
static void batch_gate(long g, void* vp) {
  batch_job* J = (batch_job*)vp; const sgo_params* P = J->P; size_t L = P->n + 1;
  sgo_u128* t = (sgo_u128*)malloc(3 * L * sizeof(sgo_u128));
  if (J->literal) {
    int rc = sgo_bootstrap_internal(P, J->key, J->lwe1 + g * L, J->lwe2 + g * L, NULL, J->n_steps, NULL, t, t + L, t + 2 * L);
    if (rc) J->err = rc;
  } else {
    fast_gate(J->fk, J->lwe1 + g * L, J->lwe2 + g * L, NULL, J->n_steps, NULL, t, t + L, t + 2 * L);
  }
  modred_lwe(P, t, J->oand + g * L); modred_lwe(P, t + L, J->oor + g * L); modred_lwe(P, t + 2 * L, J->oxor + g * L);
  free(t);
}

/* fhe.jl:310-328, with `a` (deterministic_expand, utils.jl:63-68) and noise `w` supplied.  All over
 * Z_r (power of two), negacyclic length n. */
void sgo_encrypt_private(const sgo_params* P, const uint8_t* sk, const uint64_t* a,
                         const int64_t* w, const uint8_t* message, uint64_t* b_out) {
  int n = P->n; uint64_t r = P->r, mask = r - 1;
  for (int k = 0; k < n; ++k) b_out[k] = 0;
  for (int i = 0; i < n; ++i) {                  /* a * key                        fhe.jl:322 */
    if (!sk[i]) continue;
    for (int j = 0; j < n; ++j) {
      int k = i + j;
      if (k < n) b_out[k] = (b_out[k] + a[j]) & mask; else b_out[k - n] = (b_out[k - n] - a[j]) & mask;
    }
  }
  uint64_t step = (uint64_t)1 << (P->t - 4);
  for (int k = 0; k < n; ++k) {
    uint64_t v = (b_out[k] + (uint64_t)w[k] + (message[k] ? P->Dr : 0)) & mask;   /* fhe.jl:322 */
    b_out[k] = (v / step) * step;                /* keep top 5 bits                fhe.jl:325 */
  }
}

/* fhe.jl:287-290 */
void sgo_split_ciphertext(const sgo_params* P, const uint64_t* a, const uint64_t* b, uint64_t* lwes) {
  int n = P->n; uint64_t r = P->r;
  for (int i = 1; i <= n; ++i) {
    uint64_t* o = lwes + (size_t)(i - 1) * (n + 1);
    int k = 0;                                   /* extract(a, i, n), fhe.jl:237-244 over Z_r */
    if (i < n) {
      for (int j = i; j >= 1; --j) o[k++] = a[j - 1];
      for (int j = n; j >= n - (n - i - 1); --j) o[k++] = a[j - 1] ? r - a[j - 1] : 0;
    } else {
      for (k = 0; k < n; ++k) o[k] = a[i - 1 - k];
    }
    o[n] = b[i - 1];
  }
}

/* fhe.jl:504-507 */
uint64_t sgo_decrypt_lwe(const sgo_params* P, const uint8_t* sk, const uint64_t* lwe) {
  uint64_t mask = P->r - 1, s = 0;
  for (int i = 0; i < P->n; ++i) if (sk[i]) s += lwe[i];
  uint64_t b1 = (lwe[P->n] - s) & mask;
  return ((b1 + P->Dr / 2) & mask) / P->Dr;
}

/* fhe.jl:471-494, PackedCiphertext branch */
void sgo_decrypt_packed(const sgo_params* P, const uint8_t* sk, const uint64_t* a, const uint64_t* b,
                        uint64_t* bits_out) {
  int n = P->n; uint64_t mask = P->r - 1;
  uint64_t* ks = (uint64_t*)calloc(n, sizeof(uint64_t));
  for (int i = 0; i < n; ++i) {
    if (!sk[i]) continue;
    for (int j = 0; j < n; ++j) {
      int k = i + j;
      if (k < n) ks[k] = (ks[k] + a[j]) & mask; else ks[k - n] = (ks[k - n] - a[j]) & mask;
    }
  }
  for (int k = 0; k < n; ++k) bits_out[k] = ((((b[k] - ks[k]) & mask) + P->Dr / 2) & mask) / P->Dr;
  free(ks);
}

/* fhe.jl:632-641: flatten(a) * A[l+1:2l, :] -- rows 3,4 (1-based) of the 4x2 matrix.  draws [N][2] or NULL. */
static void shortened_product_u(const ntt_plan* p, const u128* a, const u128* A, u128 B, const int64_t* draws,
                                u128* a_out, u128* b_out, u128* work /* 4N */) {
  int N = p->N; u128 Q = p->Q;
  u128* u = work; u128* prod = work + 2 * (size_t)N; u128* tmp = work + 3 * (size_t)N;
  flatten_poly_u(a, N, B, Q, draws, u, u + N);                 /* fhe.jl:637 */
  for (int c = 0; c < 2; ++c) {
    u128* res = c ? b_out : a_out;
    for (int j = 0; j < 2; ++j) {                              /* fhe.jl:638-639 */
      polymul_u(p, u + (size_t)j * N, A + ((size_t)(2 + j) * 2 + c) * N, prod, tmp);
      if (j == 0) memcpy(res, prod, N * sizeof(u128));
      else for (int k = 0; k < N; ++k) res[k] = addmod(res[k], prod[k], Q);
    }
  }
}

int sgo_shortened_external_product(const sgo_u128* a, const sgo_u128* A, int N, sgo_u128 Bs, sgo_u128 Qs,
                                   const int64_t* draws, sgo_u128* a_out, sgo_u128* b_out) {
  u128 Q = U(Qs); const ntt_plan* p = get_plan(N, Q); if (!p) return -1;
  size_t n = N;
  u128* buf = (u128*)malloc((1 + 8 + 2 + 4) * n * sizeof(u128));
  u128 *ua = buf, *uA = buf + n, *oa = buf + 9 * n, *ob = buf + 10 * n, *work = buf + 11 * n;
  for (size_t i = 0; i < n; ++i) ua[i] = U(a[i]);
  for (size_t i = 0; i < 8 * n; ++i) uA[i] = U(A[i]);
  shortened_product_u(p, ua, uA, U(Bs), draws, oa, ob, work);
  for (size_t i = 0; i < n; ++i) { a_out[i] = S(oa[i]); b_out[i] = S(ob[i]); }
  free(buf); return 0;
}

/* fhe.jl:675-693 given the n AND-outputs of _bootstrap_internal(bkey, rng, trivial(1), bit_j) over Z_Q
 * (new_lwes [n][n+1] wide): transpose, n shortened products against bkey.key[i], sum, negate/subtract, ModRed.
 * draws_short: [n][m][2] or NULL.  w_out, v_out: [m] over Z_r. */
typedef struct { const sgo_params* P; const ntt_plan* p; const sgo_u128* key; const sgo_u128* lwes;
                 const int64_t* draws; u128* w_acc; u128* v_acc; pthread_mutex_t lock; } pack_job;
static void pack_row(long i, void* vp) {
  pack_job* J = (pack_job*)vp; const sgo_params* P = J->P; int n = P->n, m = P->m; size_t M = m; u128 Q = U(P->Q);
  u128* buf = (u128*)malloc((1 + 8 + 2 + 4) * M * sizeof(u128));
  u128 *as = buf, *A = buf + M, *w = buf + 9 * M, *v = buf + 10 * M, *work = buf + 11 * M;
  for (size_t k = 0; k < M; ++k) as[k] = 0;                                     /* resize(..., m)  fhe.jl:676 */
  for (int j = 0; j < n; ++j) as[j] = U(J->lwes[(size_t)j * (n + 1) + i]);      /* new_lwes[j].a[i] */
  for (size_t k = 0; k < 8 * M; ++k) A[k] = U(J->key[(size_t)i * 8 * M + k]);
  shortened_product_u(J->p, as, A, U(P->B), J->draws ? J->draws + (size_t)i * 2 * M : NULL, w, v, work);   /* fhe.jl:683-684 */
  pthread_mutex_lock(&J->lock);
  for (size_t k = 0; k < M; ++k) { J->w_acc[k] = addmod(J->w_acc[k], w[k], Q); J->v_acc[k] = addmod(J->v_acc[k], v[k], Q); }   /* fhe.jl:686-687 */
  pthread_mutex_unlock(&J->lock);
  free(buf);
}
int sgo_pack_from_lwes(const sgo_params* P, const sgo_u128* key, const sgo_u128* new_lwes, const int64_t* draws_short,
                       uint64_t* w_out, uint64_t* v_out) {
  int n = P->n, m = P->m; u128 Q = U(P->Q);
  const ntt_plan* p = get_plan(m, Q); if (!p) return -1;
  pack_job J; J.P = P; J.p = p; J.key = key; J.lwes = new_lwes; J.draws = draws_short;
  J.w_acc = (u128*)calloc(m, sizeof(u128)); J.v_acc = (u128*)calloc(m, sizeof(u128));
  pthread_mutex_init(&J.lock, NULL);
  parallel_for(n, g_setup_threads, pack_row, &J);
  sgo_u128 r = {P->r, 0};
  for (int k = 0; k < m; ++k) {
    u128 bk = k < n ? U(new_lwes[(size_t)k * (n + 1) + n]) : 0;                 /* b = resize([new_lwe.b ...], m)  fhe.jl:678 */
    u128 w1 = negmod(J.w_acc[k], Q);                                            /* fhe.jl:689 */
    u128 v1 = submod(bk, J.v_acc[k], Q);                                        /* fhe.jl:690 */
    w_out[k] = sgo_rescale(r, S(w1), P->Q, 1).lo;                               /* fhe.jl:692-693 */
    v_out[k] = sgo_rescale(r, S(v1), P->Q, 1).lo;
  }
  free(J.w_acc); free(J.v_acc); pthread_mutex_destroy(&J.lock);
  return 0;
}

/* fhe.jl:660-696 in full.  enc_bits [n][n+1] over Z_r; draws_boot [n][n][2][m][2] or NULL; draws_short [n][m][2] or NULL. */
int sgo_pack_encrypted_bits(const sgo_params* P, const sgo_u128* key, const uint64_t* enc_bits, const int64_t* draws_boot,
                            const int64_t* draws_short, uint64_t* w_out, uint64_t* v_out) {
  int n = P->n, m = P->m; size_t L = n + 1;
  uint64_t* triv = (uint64_t*)calloc(L, sizeof(uint64_t)); triv[n] = P->Dr;       /* fhe.jl:670-671 */
  sgo_u128* lw = (sgo_u128*)malloc((size_t)n * L * sizeof(sgo_u128));
  sgo_u128* scratch = (sgo_u128*)malloc(2 * L * sizeof(sgo_u128));
  int rc = 0;
  for (int j = 0; j < n && !rc; ++j)                                              /* fhe.jl:673, keeps [1] = AND before ModRed */
    rc = sgo_bootstrap_internal_fast(P, key, triv, enc_bits + (size_t)j * L,
                                     draws_boot ? draws_boot + (size_t)j * n * 4 * m : NULL, n, NULL,
                                     lw + (size_t)j * L, scratch, scratch + L);
  if (!rc) rc = sgo_pack_from_lwes(P, key, lw, draws_short, w_out, v_out);
  free(triv); free(lw); free(scratch);
  return rc;
}

/* split_ciphertext on a length-N RLWE over Z_r (Ciphertext: N = m; PackedCiphertext: N = n)  fhe.jl:287-290 */
void sgo_split_rlwe(const sgo_params* P, int N, const uint64_t* a, const uint64_t* b, uint64_t* lwes) {
  int n = P->n; uint64_t r = P->r;
  for (int i = 1; i <= n; ++i) {
    uint64_t* o = lwes + (size_t)(i - 1) * (n + 1);
    int k = 0;
    if (i < n) {
      for (int j = i; j >= 1; --j) o[k++] = a[j - 1];
      for (int j = N; j >= N - (n - i - 1); --j) o[k++] = a[j - 1] ? r - a[j - 1] : 0;
    } else {
      for (k = 0; k < n; ++k) o[k] = a[i - 1 - k];
    }
    o[n] = b[i - 1];
  }
}

/* decrypt(key, ct::Ciphertext)  fhe.jl:471-494 (Ciphertext branch: key resized to m, first n coefficients kept) */
void sgo_decrypt_ciphertext(const sgo_params* P, const uint8_t* sk, const uint64_t* a, const uint64_t* b, uint64_t* bits_out) {
  int n = P->n, m = P->m; uint64_t mask = P->r - 1;
  uint64_t* ks = (uint64_t*)calloc(m, sizeof(uint64_t));
  for (int i = 0; i < n; ++i) {
    if (!sk[i]) continue;
    for (int j = 0; j < m; ++j) {
      int k = i + j;
      if (k < m) ks[k] = (ks[k] + a[j]) & mask; else ks[k - m] = (ks[k - m] - a[j]) & mask;
    }
  }
  for (int k = 0; k < n; ++k) bits_out[k] = ((((b[k] - ks[k]) & mask) + P->Dr / 2) & mask) / P->Dr;
  free(ks);
}

/* ---- Scheme 2 (src/fhe2.jl, src/rns.jl): parameters and the two-residue element type only; upstream has no
 * flatten / external product / bootstrap for this type (src/fhe2.jl:1-7). ------------------------------------ */
static uint64_t isqrt_u64(uint64_t x) { uint64_t r = 0; while ((r + 1) * (r + 1) <= x) ++r; return r; }

/* src/fhe2.jl:36-70.  returns 0 / -1 (k out of 1..5, src/fhe2.jl:39) */
int sgo_scheme2_params(int k, sgo_scheme2_params_t* out) {
  if (k < 1 || k > 5) return -1;
  uint64_t n = 1024;                                               /* fhe2.jl:41 */
  uint64_t r = ((uint64_t)1 << (k + 6)) * isqrt_u64(n);            /* fhe2.jl:43 */
  uint64_t m = r / 2, l = 2;                                       /* fhe2.jl:44-45 */
  int t = 0; while (((uint64_t)1 << t) < r) ++t; t -= 1;           /* fhe2.jl:46 */
  sgo_u128 q, Bp, B;
  if (sgo_find_modulus(S(2 * n), S((u128)128 * r * n), S(0), &q)) return -2;          /* fhe2.jl:48 */
  uint64_t tau = 2 * isqrt_u64(n);                                 /* fhe2.jl:51 */
  u128 bmin = (u128)15 * ((u128)1 << (2 * k + 2)) * r * tau * isqrt_u64(2 * l * m);   /* fhe2.jl:57 */
  if (sgo_find_modulus(S(r), S(bmin), S(0), &Bp)) return -2;
  if (sgo_find_modulus(S(r), S(U(Bp) + 1), S(0), &B)) return -2;  /* fhe2.jl:58 */
  out->n = (int32_t)n; out->k = k; out->r = r; out->m = m; out->t = t; out->q = q.lo; out->tau = tau;
  out->B = B.lo; out->Bp = Bp.lo;
  out->Dr = r >> (k + 2); out->Dq = q.lo >> (k + 2);               /* fhe2.jl:62-63 */
  return 0;
}

/* src/rns.jl:51-60: op 0 = *, 1 = +, 2 = -, limb-wise on (v1 mod M1, v2 mod M2) */
void sgo_rns2_op(int op, size_t count, const uint64_t* a1, const uint64_t* a2, const uint64_t* b1, const uint64_t* b2,
                 uint64_t M1, uint64_t M2, uint64_t* o1, uint64_t* o2) {
  for (size_t i = 0; i < count; ++i) {
    if (op == 0) { o1[i] = (uint64_t)((u128)a1[i] * b1[i] % M1); o2[i] = (uint64_t)((u128)a2[i] * b2[i] % M2); }          /* rns.jl:51-52 */
    else if (op == 1) { o1[i] = (uint64_t)(((u128)a1[i] + b1[i]) % M1); o2[i] = (uint64_t)(((u128)a2[i] + b2[i]) % M2); } /* rns.jl:55-56 */
    else { o1[i] = (uint64_t)(((u128)a1[i] + M1 - b1[i]) % M1); o2[i] = (uint64_t)(((u128)a2[i] + M2 - b2[i]) % M2); }   /* rns.jl:59-60 */
  }
}
