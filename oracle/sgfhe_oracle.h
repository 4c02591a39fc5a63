/*
 * sgfhe_oracle.h -- CPU oracle for the bootstrapping hot path of nucypher/SGFHE.jl.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product path (sgfhe.jl_b200/, libsgfhe_cuda.so).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may use it, and there only as the checker / CPU baseline.
 *
 * This is a plain-C restatement (unsigned __int128) of the reference's
 * algorithm, function by function, each citing the reference file:line it
 * follows (paths relative to the reference tree, e.g. src/fhe.jl:559).
 *
 * PARITY STATUS: "parity unpinned" by the reference's own tests -- the
 * reference holds no golden vectors (every test seeds MersenneTwister from
 * entropy and checks a property), Julia is not installed here, and the ring
 * arithmetic lives in the un-vendored dependency DarkIntegers ~0.1.0
 * (Project.toml:7,20).  What pins this oracle instead: (1) every operation on
 * the path is exact arithmetic in Z_Q or Z_Q[x]/(x^m+1), whose result is
 * unique whatever algorithm computes it; (2) all bit-affecting conventions
 * (digit ranges, rounding, extract indices/signs, draw order) are in-tree and
 * restated here literally; (3) an independent Python big-integer model
 * (oracle/model.py, Kronecker-substitution products) agrees with this file on
 * every function; (4) every property the reference's tests check
 * (test/internals.test.jl, test/api.test.jl) is ported in tests/.
 *
 * ABI: wide values cross as {lo,hi} pairs of uint64 (little-endian 128-bit).
 */
#ifndef SGFHE_ORACLE_H
#define SGFHE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t lo, hi; } sgo_u128;

/* src/fhe.jl:27-99 */
typedef struct {
  int32_t n;        /* polynomial length                     fhe.jl:30 */
  int32_t t;        /* log2(r) - 1                           fhe.jl:61 */
  int32_t m;        /* r / 2                                 fhe.jl:62 */
  int32_t large;    /* 0: Q fits UInt64, 1: UInt128          fhe.jl:71-78 */
  uint64_t r;       /* 16 n                                  fhe.jl:53 */
  uint64_t q;       /* least prime >= r n, = 1 mod 2n        fhe.jl:57 */
  uint64_t Dr;      /* r / 4                                 fhe.jl:88 */
  uint64_t Dq;      /* q / 4                                 fhe.jl:89 */
  sgo_u128 Q;       /* least prime in [1220,1225] r^4 n^2, = 1 mod 2m   fhe.jl:64-69 */
  sgo_u128 B;       /* 35 r^2 n                              fhe.jl:87 */
  sgo_u128 DQ;      /* Q / 8  (DQ_tilde)                     fhe.jl:90 */
} sgo_params;

/* returns 0 on success; <0: n invalid (fhe.jl:45-46) or no modulus found (utils.jl:26) */
int sgo_params_init(int n, sgo_params* out);

/* utils.jl:7-28.  qmax may be {0,0} for "nothing".  returns 0 / -1 (not found). */
int sgo_find_modulus(sgo_u128 n, sgo_u128 qmin, sgo_u128 qmax, sgo_u128* out);
int sgo_is_prime(sgo_u128 x);

/* utils.jl:78-92 */
sgo_u128 sgo_rescale(sgo_u128 new_max, sgo_u128 x, sgo_u128 old_max, int round_result);

/* utils.jl:155-189 (draws == NULL) and utils.jl:198-241 (draws: l signed values, as int64).
 * a, B are residues mod q; out[l]. */
void sgo_flatten(sgo_u128 a, sgo_u128 B, int l, sgo_u128 q, const int64_t* draws, sgo_u128* out);

/* utils.jl:253-264: out is [l][N]; draws is [N][l] or NULL */
void sgo_flatten_poly(const sgo_u128* a, int N, sgo_u128 B, int l, sgo_u128 q,
                      const int64_t* draws, sgo_u128* out);

/* Negacyclic product in Z_Q[x]/(x^N+1) (DarkIntegers Polynomial *, not in tree; called at
 * fhe.jl:527-528).  Two independent algorithms. */
void sgo_polymul_schoolbook(const sgo_u128* a, const sgo_u128* b, sgo_u128* out, int N, sgo_u128 Q);
int  sgo_polymul_ntt(const sgo_u128* a, const sgo_u128* b, sgo_u128* out, int N, sgo_u128 Q);

/* mul_by_monomial (DarkIntegers; used fhe.jl:555,573): out = p * x^shift mod (x^N+1), shift any sign */
void sgo_mul_by_monomial(const sgo_u128* p, int N, int64_t shift, sgo_u128 Q, sgo_u128* out);

/* fhe.jl:535-548 */
void sgo_initial_poly(const sgo_params* P, sgo_u128* out /* [m] */);

/* fhe.jl:237-244, 1-based i as in the reference.  modulus for negation. */
void sgo_extract(const sgo_u128* a, int N, int i, int n, sgo_u128 modulus, sgo_u128* out);

/* fhe.jl:181-201.  sk[n] bits; a_rand [n][4][m] uniform in [0,Q); e_rand [n][4][m] in [-n,n];
 * key_out [n][4][2][m] = key[i][j,c].coeffs[k].  Only rows [row0,row1) are produced. */
int sgo_bkey_generate(const sgo_params* P, const uint8_t* sk, const sgo_u128* a_rand,
                      const int64_t* e_rand, int row0, int row1, sgo_u128* key_out);

/* fhe.jl:519-530.  A is [4][2][N]; draws [2][N][2] (a first, then b) or NULL. */
int sgo_external_product(const sgo_u128* a, const sgo_u128* b, const sgo_u128* A, int N,
                         sgo_u128 B, sgo_u128 Q, const int64_t* draws,
                         sgo_u128* a_out, sgo_u128* b_out);

/* fhe.jl:559-595, literal formulation (A rebuilt every step, 8 independent products/step).
 * key: rows [0,n_steps) of [n][4][2][m].  lwe1/lwe2: [n+1] (a then b) over Z_r.
 * draws: [n_steps][2][m][2] or NULL.  n_steps <= n (n_steps < n gives a truncated run for
 * traces).  trace: [n_steps][2][m] accumulator after every step, or NULL.
 * out_and/or/xor: [n+1] over Z_Q (a then b), before ModRed. */
int sgo_bootstrap_internal(const sgo_params* P, const sgo_u128* key, const uint64_t* lwe1,
                           const uint64_t* lwe2, const int64_t* draws, int n_steps,
                           sgo_u128* trace, sgo_u128* out_and, sgo_u128* out_or, sgo_u128* out_xor);

/* Same function in the rewritten form of SURVEY.md 3.1: acc += (x^u - 1) * (flatten(acc) . C^(k)),
 * key pre-transformed once.  Must equal sgo_bootstrap_internal bit for bit. */
int sgo_bootstrap_internal_fast(const sgo_params* P, const sgo_u128* key, const uint64_t* lwe1,
                                const uint64_t* lwe2, const int64_t* draws, int n_steps,
                                sgo_u128* trace, sgo_u128* out_and, sgo_u128* out_or, sgo_u128* out_xor);

/* fhe.jl:608-621 incl. reduce_modulus (fhe.jl:644-648, utils.jl:107-117).  outputs [n+1] over Z_r. */
int sgo_bootstrap(const sgo_params* P, const sgo_u128* key, const uint64_t* lwe1,
                  const uint64_t* lwe2, const int64_t* draws,
                  uint64_t* out_and, uint64_t* out_or, uint64_t* out_xor);

/* threads used by one-off setup work (sgo_bkey_generate, key transform); default 8 */
void sgo_set_setup_threads(int t);

/* Batch of independent gates over `threads` pthreads (bench cpu_baseline / reference arm).
 * lwe1/lwe2: [batch][n+1]; outs: [batch][n+1].  literal != 0 selects the reference formulation. */
int sgo_bootstrap_batch(const sgo_params* P, const sgo_u128* key, int batch, const uint64_t* lwe1,
                        const uint64_t* lwe2, int n_steps, int literal, int threads,
                        uint64_t* out_and, uint64_t* out_or, uint64_t* out_xor);

/* fhe.jl:310-328 with the expanded `a` and the noise `w` supplied by the caller
 * (deterministic_expand uses Julia's MersenneTwister; utils.jl:63-68).  All over Z_r, length n. */
void sgo_encrypt_private(const sgo_params* P, const uint8_t* sk, const uint64_t* a,
                         const int64_t* w, const uint8_t* message, uint64_t* b_out);
/* fhe.jl:287-290: lwes [n][n+1] */
void sgo_split_ciphertext(const sgo_params* P, const uint64_t* a, const uint64_t* b, uint64_t* lwes);
/* fhe.jl:504-507: returns the snapped quotient (0/1 for a valid bit; Julia would throw otherwise) */
uint64_t sgo_decrypt_lwe(const sgo_params* P, const uint8_t* sk, const uint64_t* lwe);
/* fhe.jl:471-494 for PackedCiphertext (length n over Z_r) */
void sgo_decrypt_packed(const sgo_params* P, const uint8_t* sk, const uint64_t* a, const uint64_t* b,
                        uint64_t* bits_out);

/* fhe.jl:632-641: A [4][2][N] coefficient form (rows 3,4 used); draws [N][2] or NULL */
int sgo_shortened_external_product(const sgo_u128* a, const sgo_u128* A, int N, sgo_u128 B, sgo_u128 Q,
                                   const int64_t* draws, sgo_u128* a_out, sgo_u128* b_out);
/* fhe.jl:675-693 given new_lwes [n][n+1] wide over Z_Q; draws_short [n][m][2] or NULL; w_out, v_out [m] over Z_r */
int sgo_pack_from_lwes(const sgo_params* P, const sgo_u128* key, const sgo_u128* new_lwes, const int64_t* draws_short,
                       uint64_t* w_out, uint64_t* v_out);
/* fhe.jl:660-696: enc_bits [n][n+1]; draws_boot [n][n][2][m][2] or NULL; draws_short [n][m][2] or NULL */
int sgo_pack_encrypted_bits(const sgo_params* P, const sgo_u128* key, const uint64_t* enc_bits, const int64_t* draws_boot,
                            const int64_t* draws_short, uint64_t* w_out, uint64_t* v_out);
/* fhe.jl:287-290 on a length-N RLWE */
void sgo_split_rlwe(const sgo_params* P, int N, const uint64_t* a, const uint64_t* b, uint64_t* lwes);
/* fhe.jl:471-494, Ciphertext branch (length m) */
void sgo_decrypt_ciphertext(const sgo_params* P, const uint8_t* sk, const uint64_t* a, const uint64_t* b, uint64_t* bits_out);

/* Scheme 2 (src/fhe2.jl:17-70) */
typedef struct { int32_t n, k, t, pad; uint64_t r, m, q, tau, B, Bp, Dr, Dq; } sgo_scheme2_params_t;
int sgo_scheme2_params(int k, sgo_scheme2_params_t* out);
/* RNS2Number arithmetic (src/rns.jl:51-60): op 0 = *, 1 = +, 2 = - */
void sgo_rns2_op(int op, size_t count, const uint64_t* a1, const uint64_t* a2, const uint64_t* b1, const uint64_t* b2,
                 uint64_t M1, uint64_t M2, uint64_t* o1, uint64_t* o2);

#ifdef __cplusplus
}
#endif
#endif
