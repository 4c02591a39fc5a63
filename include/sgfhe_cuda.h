/*
 * sgfhe_cuda.h -- C ABI of libsgfhe_cuda.so, the B200 (sm_100a) backend for the bootstrapping hot
 * path of nucypher/SGFHE.jl (Gao's scheme, src/fhe.jl).
 *
 * The reference has no FFI; the seam is the call
 *     bootstrap(bkey::BootstrapKey, rng, enc_bit1::EncryptedBit, enc_bit2::EncryptedBit)
 * (reference src/fhe.jl:608-621), with external_product (src/fhe.jl:519-530), flatten_poly
 * (src/utils.jl:253-264) and DarkIntegers' `Polynomial *` (called at src/fhe.jl:527-528) as inner
 * test seams.  A Julia maintainer binds these with `ccall((:sgfhe_..., "libsgfhe_cuda"), ...)`;
 * see INTEGRATION.md for the shim.
 *
 * Conventions:
 *  - every value crossing the boundary is a canonical residue (`value(x)` in the reference, never
 *    the raw Montgomery word), little-endian unsigned;
 *  - elements of Z_Q are two uint64 (lo, hi) regardless of n ("wide");
 *  - elements of Z_r (LWE ciphertexts, the type of EncryptedBit.lwe, src/fhe.jl:206-209,272-274)
 *    are one uint64; an LWE is n+1 values: a[0..n-1] then b;
 *  - the caller owns every host buffer; the library owns device memory behind the context;
 *  - every entry point returns 0 on success, a negative sgfhe_status otherwise, never aborts;
 *    sgfhe_last_error() returns the message of the calling thread's last failure;
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SGFHE_CUDA_H
#define SGFHE_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgfhe_ctx sgfhe_ctx;

enum sgfhe_status {
  SGFHE_OK = 0,
  SGFHE_ERR_ARG = -1,      /* mirrors the reference's @assert failures (src/fhe.jl:45-47,238,313) */
  SGFHE_ERR_MODULUS = -2,  /* mirrors error("Cound not find a modulus") (src/utils.jl:26) / "n is too large" (src/fhe.jl:77) */
  SGFHE_ERR_CUDA = -3,     /* CUDA runtime failure or no device */
  SGFHE_ERR_STATE = -4,    /* e.g. bootstrap before a key upload */
  SGFHE_ERR_NOMEM = -5
};

/* Scheme parameters, derived exactly as Params(n) does (src/fhe.jl:43-97, src/utils.jl:7-28). */
typedef struct {
  int32_t n;            /* polynomial length                         fhe.jl:30    */
  int32_t t;            /* log2(r) - 1                               fhe.jl:61    */
  int32_t m;            /* r / 2                                     fhe.jl:62    */
  int32_t rns_primes;   /* backend detail: 30-bit NTT primes used for the exact convolution */
  uint64_t r;           /* 16 n                                      fhe.jl:53    */
  uint64_t q;           /* fhe.jl:57 */
  uint64_t Dr, Dq;      /* fhe.jl:88-89 */
  uint64_t Q[2];        /* fhe.jl:64-69 */
  uint64_t B[2];        /* fhe.jl:87 */
  uint64_t DQ_tilde[2]; /* fhe.jl:90 */
} sgfhe_params;

/* Params(n) + device context.  n: power of two, 64 <= n <= 2048 (the reference accepts every n whose Q fits
 * 128 bits, src/fhe.jl:71-78; this backend carries Z_Q in three 32-bit limbs, Q < 2^96, which n = 2048 with its
 * 93-bit Q is the last to satisfy).  device: CUDA ordinal.
 * Replaces: Params(n; ...) src/fhe.jl:43. */
int sgfhe_ctx_create(int32_t n, int32_t device, sgfhe_ctx** out);
int sgfhe_ctx_destroy(sgfhe_ctx* ctx);
int sgfhe_params_get(const sgfhe_ctx* ctx, sgfhe_params* out);

/* Params(n) without a device (host arithmetic only; usable on a CPU-only box). */
int sgfhe_params_derive(int32_t n, sgfhe_params* out);

const char* sgfhe_last_error(void);

/* Upload BootstrapKey.key (src/fhe.jl:176-201) and pre-transform it once into the NTT domain.
 * key: host, C order [rows][4][2][m][2] uint64 = value(bkey.key[i][j,c].coeffs[k]) (lo,hi),
 * i = step, j = gadget row (a-digit0, a-digit1, b-digit0, b-digit1), c = column.
 * rows == n for a usable key; rows < n is accepted for truncated traces only. */
int sgfhe_bkey_upload(sgfhe_ctx* ctx, const uint64_t* key, int32_t rows);

/* BootstrapKey(rng, sk) (src/fhe.jl:181-201) on the device, leaving the key in the context already pre-transformed.
 * sk: n bytes (0/1).  a_rand: host [rows][4][m][2] wide, the uniform polynomials a_1..a_4 of key rows row0..row0+rows-1
 * (src/fhe.jl:193); e_rand: host int64 [rows][4][m], the errors in [-n, n] (src/fhe.jl:194) -- both drawn by the caller's
 * RNG in the reference's order (per row: a_1..a_4, then e_1..e_4).  The products a_j * ext_key, + e_j (src/fhe.jl:195),
 * + s_i G (src/fhe.jl:196) and the pre-transform run on the device.  Rows must be generated in order (row0 <= rows
 * present); key_out: NULL, or host [rows][4][2][m][2] wide receiving the coefficient form (= bkey.key) on request. */
int sgfhe_bkey_generate(sgfhe_ctx* ctx, const uint8_t* sk, const uint64_t* a_rand, const int64_t* e_rand,
                        int32_t row0, int32_t rows, uint64_t* key_out);

/* A context holds ONE key.  Every call that changes it (upload, generate, import, adopt) gives it a new non-zero token;
 * a host-side key object remembers the token it obtained and compares before use, so that a second key uploaded to the
 * same context is noticed instead of silently used (the reference's bootstrap is a pure function of bkey).
 * token = 0: no key.  rows (may be NULL): key rows present. */
int sgfhe_bkey_token(const sgfhe_ctx* ctx, uint64_t* token, int32_t* rows);

/* bootstrap(bkey, rng|nothing, enc_bit1, enc_bit2) for a batch of independent gates
 * (src/fhe.jl:608-621 -> _bootstrap_internal src/fhe.jl:559-595 -> reduce_modulus src/fhe.jl:644-648).
 * lwe1, lwe2: host [batch][n+1] over Z_r.  out_*: host [batch][n+1] over Z_r.
 * draws: NULL selects flatten(rng::Nothing, ...) (src/utils.jl:155-189).  Otherwise host int64
 * [batch][n][2][m][2]: the values rand(rng, -xmax:xmax) the caller's RNG yields in the reference's
 * order -- step k, polynomial a then b, coefficient, digit (src/fhe.jl:524-525, src/utils.jl:228-230,
 * 257-258). */
int sgfhe_bootstrap_batch(sgfhe_ctx* ctx, int32_t batch, const uint64_t* lwe1, const uint64_t* lwe2,
                          const int64_t* draws, uint64_t* out_and, uint64_t* out_or, uint64_t* out_xor);

/* Same, all six buffers already in device memory of ctx's device (draws may be NULL).
 * Asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream).  Launches on one context must not
 * overlap in time (the per-gate scratch and the work counter belong to the context): use one stream per context, or one
 * context per stream. */
int sgfhe_bootstrap_batch_device(sgfhe_ctx* ctx, int32_t batch, const uint64_t* d_lwe1,
                                 const uint64_t* d_lwe2, const int64_t* d_draws, uint64_t* d_out_and,
                                 uint64_t* d_out_or, uint64_t* d_out_xor, void* stream);

/* bootstrap(bkey, rng, ...) with the flatten draws made ON THE DEVICE by a counter-based generator (Philox4x32-10 keyed by
 * `seed`, counter = (coefficient, 2 step + polynomial, 64-bit gate0 + gate index)): the randomised mode without 268 MB of host
 * draws per gate at Params(1024).  Each draw is uniform on [-xmax, xmax] as at src/utils.jl:210-216, 229, but the stream is
 * not that of any Julia RNG -- use sgfhe_bootstrap_batch with host draws where the reference's exact stream matters.
 * seed != 0; gate0 lets sharded or successive batches use disjoint streams. */
int sgfhe_bootstrap_batch_rng(sgfhe_ctx* ctx, int32_t batch, const uint64_t* lwe1, const uint64_t* lwe2, uint64_t seed,
                              uint64_t gate0, uint64_t* out_and, uint64_t* out_or, uint64_t* out_xor);
int sgfhe_bootstrap_batch_rng_device(sgfhe_ctx* ctx, int32_t batch, const uint64_t* d_lwe1, const uint64_t* d_lwe2,
                                     uint64_t seed, uint64_t gate0, uint64_t* d_out_and, uint64_t* d_out_or,
                                     uint64_t* d_out_xor, void* stream);
/* Test seam: the draws that generator makes for steps step0..step0+steps-1 of gate `gate`: host int64 [steps][2][m][2],
 * the layout sgfhe_bootstrap_batch takes for one gate. */
int sgfhe_device_draws(sgfhe_ctx* ctx, uint64_t seed, uint64_t gate, int32_t step0, int32_t steps, int64_t* out);

/* _bootstrap_internal for one gate with the accumulator after every step
 * (src/fhe.jl:559-595; loop body src/fhe.jl:579-582).  n_steps <= uploaded rows.
 * draws: NULL or [n_steps][2][m][2].  trace: NULL or [n_steps][2][m][2] uint64 (a then b, wide).
 * out_*: [n+1][2] wide, over Z_Q, i.e. before reduce_modulus. */
int sgfhe_bootstrap_trace(sgfhe_ctx* ctx, const uint64_t* lwe1, const uint64_t* lwe2, const int64_t* draws,
                          int32_t n_steps, uint64_t* trace, uint64_t* out_and, uint64_t* out_or,
                          uint64_t* out_xor);

/* _bootstrap_internal for a batch (src/fhe.jl:559-595): the three LWEs over Z_Q, i.e. BEFORE reduce_modulus,
 * as pack_encrypted_bits needs them (src/fhe.jl:673).  out_*: host [batch][n+1][2] wide. */
int sgfhe_bootstrap_internal_batch(sgfhe_ctx* ctx, int32_t batch, const uint64_t* lwe1, const uint64_t* lwe2,
                                   const int64_t* draws, uint64_t* out_and, uint64_t* out_or, uint64_t* out_xor);

/* shortened_external_product(rng|nothing, polys[i], bkey.key[i], Val(B), Val(2)) for i = 0..count-1
 * (src/fhe.jl:632-641 as called at src/fhe.jl:683-684): flatten(polys[i]) times rows 3,4 of the uploaded key row i.
 * polys: host [count][m][2] wide; draws: NULL or [count][m][2]; out: [count][2][m][2] wide (w_i then v_i). */
int sgfhe_shortened_products(sgfhe_ctx* ctx, int32_t count, const uint64_t* polys, const int64_t* draws,
                             uint64_t* out);

/* pack_encrypted_bits(bkey, rng|nothing, enc_bits) (src/fhe.jl:660-696), every stage on the device: n internal
 * bootstraps of (trivial(1), enc_bits[i]) keeping the AND output over Z_Q (:670-673), the transposition (:675-678), n
 * shortened external products (:683-684), the sums over i, negate / subtract (:686-690) and ModRed Q -> r (:692-693).
 * enc_bits: host [n][n+1] over Z_r.  draws_boot: NULL or int64 [n][n][2][m][2] (the n bootstraps, in call order);
 * draws_short: NULL or int64 [n][m][2] (the n shortened products) -- both NULL or both given, as one rng serves both.
 * out_w, out_v: host [m] over Z_r = the RLWE ciphertext Ciphertext(params, w, v) (src/fhe.jl:695). */
int sgfhe_pack_encrypted_bits(sgfhe_ctx* ctx, const uint64_t* enc_bits, const int64_t* draws_boot,
                              const int64_t* draws_short, uint64_t* out_w, uint64_t* out_v);

/* Test seam: the stages of pack_encrypted_bits AFTER the n bootstraps (src/fhe.jl:675-693) from given LWEs over Z_Q.
 * new_lwes: host [n][n+1][2] wide; draws_short: NULL or int64 [n][m][2]; out_w, out_v: host [m] over Z_r. */
int sgfhe_pack_from_lwes(sgfhe_ctx* ctx, const uint64_t* new_lwes, const int64_t* draws_short, uint64_t* out_w,
                         uint64_t* out_v);

/* Negacyclic products in Z_Q[x]/(x^m+1): out[i] = a[i] * b[i], DarkIntegers `Polynomial *` as called at
 * src/fhe.jl:195,527-528.  a, b, out: host [batch][m][2] wide canonical. */
int sgfhe_polymul(sgfhe_ctx* ctx, int32_t batch, const uint64_t* a, const uint64_t* b, uint64_t* out);
/* Same with device buffers, asynchronous on `stream`. */
int sgfhe_polymul_device(sgfhe_ctx* ctx, int32_t batch, const uint64_t* d_a, const uint64_t* d_b,
                         uint64_t* d_out, void* stream);

/* flatten_poly(rng|nothing, a, Val(B), Val(2)) (src/utils.jl:253-264): a host [m][2] wide;
 * draws NULL or [m][2]; out [2][m][2] wide (digit polynomials as residues mod Q). */
int sgfhe_flatten_poly(sgfhe_ctx* ctx, const uint64_t* a, const int64_t* draws, uint64_t* out);

/* external_product(rng|nothing, a, b, A, Val(B), Val(2)) (src/fhe.jl:519-530): a, b [m][2] wide,
 * A [4][2][m][2] wide coefficient form, draws NULL or [2][m][2]; a_out, b_out [m][2] wide. */
int sgfhe_external_product(sgfhe_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* A,
                           const int64_t* draws, uint64_t* a_out, uint64_t* b_out);

/* Serialisation of the pre-transformed key (the reference has none: `Serialization` is imported and unused at
 * examples/test_scheme2.jl:3).  The blob carries n, m, Q and the RNS basis and is rejected by a context of other
 * parameters.  export_size -> export into a caller buffer; import replaces sgfhe_bkey_upload. */
int sgfhe_bkey_export_size(sgfhe_ctx* ctx, int32_t rows, uint64_t* bytes);
int sgfhe_bkey_export(sgfhe_ctx* ctx, int32_t rows, void* blob, uint64_t bytes);
int sgfhe_bkey_import(sgfhe_ctx* ctx, const void* blob, uint64_t bytes);

/* Scheme 2 (src/fhe2.jl, "experimental, not finished" upstream: it defines Params, keys, encrypt/decrypt and NO
 * bootstrap).  What exists upstream and is mirrored here: Params(k) (src/fhe2.jl:36-70) and the arithmetic of
 * its ring element type RNS2Number{UInt64, B, Bp} (src/rns.jl:51-60), batched. */
typedef struct { int32_t n, k, t, pad; uint64_t r, m, q, tau, B, Bp, Dr, Dq; } sgfhe_scheme2_params;
int sgfhe_scheme2_params_derive(int32_t k, sgfhe_scheme2_params* out);
/* out = a op b limb-wise, op 0 = * (rns.jl:51-52), 1 = + (rns.jl:55-56), 2 = - (rns.jl:59-60); host buffers of `count` */
int sgfhe_rns2_op(int32_t device, int32_t op, uint64_t count, const uint64_t* a1, const uint64_t* a2,
                  const uint64_t* b1, const uint64_t* b2, uint64_t M1, uint64_t M2, uint64_t* o1, uint64_t* o2);
/* same with device buffers, asynchronous on `stream` */
int sgfhe_rns2_op_device(int32_t device, int32_t op, uint64_t count, const uint64_t* d_a1, const uint64_t* d_a2,
                         const uint64_t* d_b1, const uint64_t* d_b2, uint64_t M1, uint64_t M2, uint64_t* d_o1,
                         uint64_t* d_o2, void* stream);

/* Scheme 2 polynomial arithmetic and key generation (K11).  Ring elements are RNS2Number{UInt64, B, B'} = (v mod B,
 * v mod B') (src/rns.jl:8-24); polynomials live in (Z_B x Z_B')[x]/(x^m+1) with m = Scheme2.Params(k).m (2048..32768).
 * Planar layout: limb 1 and limb 2 of a batch of polynomials are separate arrays [count][m] of canonical residues.
 * NO bootstrap exists upstream for this scheme (src/fhe2.jl:1-7); these entry points cover what does exist. */
typedef struct sgfhe_s2_ctx sgfhe_s2_ctx;
int sgfhe_s2_ctx_create(int32_t k, int32_t device, sgfhe_s2_ctx** out);      /* Scheme2.Params(k) + twiddles, src/fhe2.jl:36-70 */
int sgfhe_s2_ctx_destroy(sgfhe_s2_ctx* ctx);
int sgfhe_s2_params_get(const sgfhe_s2_ctx* ctx, sgfhe_scheme2_params* out);
/* out = a * b, negacyclic, limb-wise (Polynomial{RNS2Number} `*` as called at src/fhe2.jl:124 with src/rns.jl:51-52): host */
int sgfhe_s2_polymul(sgfhe_s2_ctx* ctx, int32_t batch, const uint64_t* a1, const uint64_t* a2, const uint64_t* b1,
                     const uint64_t* b2, uint64_t* o1, uint64_t* o2);
/* same with device buffers, asynchronous on `stream`; a and b are OVERWRITTEN by their transforms; b_broadcast != 0: one
 * polynomial b multiplies every a[i] */
int sgfhe_s2_polymul_device(sgfhe_s2_ctx* ctx, int32_t batch, uint64_t* d_a1, uint64_t* d_a2, uint64_t* d_b1, uint64_t* d_b2,
                            int32_t b_broadcast, uint64_t* d_o1, uint64_t* d_o2, void* stream);
/* forward (inverse = 0) or inverse negacyclic transform of `count` polynomials per limb, in place, device buffers; the
 * transform domain is in bit-reversed order, residues canonical */
int sgfhe_s2_ntt_device(sgfhe_s2_ctx* ctx, int32_t inverse, int32_t count, uint64_t* d_x1, uint64_t* d_x2, void* stream);
/* the 8 multiply-accumulates of an external product in the transform domain (the shape of src/fhe.jl:527-528 over
 * RNS2Number): out[g][c] = sum_j d[g][j] . K[g][j][c]; d [count][4][m], K [count][4][2][m], out [count][2][m] per limb */
int sgfhe_s2_mac8_device(sgfhe_s2_ctx* ctx, int32_t count, const uint64_t* d_d1, const uint64_t* d_d2, const uint64_t* d_k1,
                         const uint64_t* d_k2, uint64_t* d_o1, uint64_t* d_o2, void* stream);
/* Scheme2.BootstrapKey(rng, sk) (src/fhe2.jl:104-131) on the device.  sk: n = 1024 bytes (0/1); a_rand: host
 * [rows][4][m][2] wide integers below Q = B B' (rand(rng, range_Q, m), src/fhe2.jl:122); e_rand: host int64 [rows][4][m]
 * in [-tau, tau] (src/fhe2.jl:123), both in the reference's draw order.  key_out: host [rows][4][2][m][2] uint64, the
 * (v1, v2) pairs of key[i][j,c].coeffs[k] -- the memory layout of an Array of RNS2Number. */
int sgfhe_s2_bkey_generate(sgfhe_s2_ctx* ctx, const uint8_t* sk, const uint64_t* a_rand, const int64_t* e_rand,
                           int32_t row0, int32_t rows, uint64_t* key_out);

/* split_ciphertext (src/fhe.jl:287-290) on the device: `count` RLWE ciphertexts over Z_r, polynomials of length N
 * (N = n for a PackedCiphertext, N = m for a Ciphertext), a and b as [count][N] uint64 -> count * n LWEs [count*n][n+1]:
 * LWE i = (extract(a, i, n), b[i]) with extract as in src/fhe.jl:237-244 (reversed window, negated wrap-around).
 * Device buffers, asynchronous on `stream`; outputs feed sgfhe_bootstrap_batch_device directly. */
int sgfhe_split_ciphertext_device(sgfhe_ctx* ctx, int32_t count, int32_t N, const uint64_t* d_a, const uint64_t* d_b,
                                  uint64_t* d_lwes, void* stream);
/* same with host buffers */
int sgfhe_split_ciphertext(sgfhe_ctx* ctx, int32_t count, int32_t N, const uint64_t* a, const uint64_t* b, uint64_t* lwes);

/* decrypt(key, ::EncryptedBit) (src/fhe.jl:504-507) for a batch: lwes [count][n+1] over Z_r, sk n bytes (0/1);
 * out[i] = ((b - <a, s>) + Dr/2 mod r) / Dr as one byte (0 or 1; larger values are the reference's InexactError and
 * are returned as they are for the caller to reject).  Device buffers, asynchronous on `stream`. */
int sgfhe_decrypt_bits_device(sgfhe_ctx* ctx, int32_t count, const uint64_t* d_lwes, const uint8_t* d_sk, uint8_t* d_out,
                              void* stream);
/* same with host buffers */
int sgfhe_decrypt_bits(sgfhe_ctx* ctx, int32_t count, const uint64_t* lwes, const uint8_t* sk, uint8_t* out);

/* Kernel launches issued by this library in the calling process since load (for bench accounting). */
uint64_t sgfhe_launch_count(void);

/* Device pointer / size of the pre-transformed key (NCCL broadcast between ranks, src: none --
 * the reference is single-process).  After a broadcast into this buffer call sgfhe_bkey_adopt. */
int sgfhe_bkey_device_buffer(sgfhe_ctx* ctx, int32_t rows, void** d_ptr, uint64_t* bytes);
int sgfhe_bkey_adopt(sgfhe_ctx* ctx, int32_t rows);

#ifdef __cplusplus
}
#endif
#endif
