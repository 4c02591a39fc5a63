// device_math.cuh -- device arithmetic for the bootstrapping hot path (sm_100a).
//
// Design: the reference multiplies polynomials in Z_Q[x]/(x^m+1) with a single wide prime
// Q (62..93 bits; src/fhe.jl:64-69).  B200's integer pipe is 32 bits wide (IMAD.lo 64 lanes/clk/SM,
// IMAD.WIDE/HI ~2.5x slower; tools/microbench), so every product here is computed as an EXACT integer
// negacyclic convolution of small signed digits with centred key coefficients, carried in an RNS basis
// of 30-bit NTT primes (Shoup/Harvey lazy butterflies: 1 IMAD.HI + 2 IMAD.lo per modmul), then lifted
// back to Z_Q by a fixed-point CRT.  All values are exact, so results equal the reference bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sgfhe {

typedef unsigned __int128 u128;
typedef __int128 i128;

constexpr int MAXP = 8;

struct DevConst {
  int n, m, logm, logr, kB;     // r = 2^logr, B = 35 << kB (src/fhe.jl:87)
  int L;                        // primes used by the bootstrap external product
  int LM;                       // primes used by a general product of two full-size operands
  int sbits;                    // bits(Q) - 1
  int pm_uncentred;             // the LM-prime basis holds m Q^2 with the CRT margin: standalone products skip the centring
  uint32_t zero;                // always 0; a constant-bank operand ptxas cannot fold (forces 3-input IADD3 on the ALU pipe)
  u128 Q, DQ, offs, B;          // offs = s * (1 + B) mod Q (src/utils.jl:169)
  uint64_t s;                   // digit offset (src/utils.jl:162-166)
  uint64_t s46;                 // dig_bias - s: bias of the stored digit words minus the digit offset
  uint64_t dig_bias;            // 2^46 (n <= 1024: |digit| <= 2B < 2^45) or 2^48 (n = 2048: 2B < 2^47.2)
  uint64_t xmax;                // 3 (B / 2): draws of the randomised flatten are uniform on [-xmax, xmax] (src/utils.jl:210-216)
  double barrett_inv;           // 2^(sbits-16) / Q, scaled by (1 - 2^-40): never above the true value
  double inv35;                 // 1/35 rounded up
  uint32_t Ql[3], offl[3];      // Q and offs as 32-bit limbs
  uint32_t KQ[4];               // K Q with K = L 2^30 + L + 1: an upper bound of every unreduced CRT sum of the bootstrap basis
  uint32_t p[MAXP], pinv_neg[MAXP], dig_mu[MAXP], vinv[MAXP], vk[MAXP];
  uint32_t r32[MAXP], r64[MAXP];            // 2^32 mod p, 2^64 mod p
  uint32_t qmodp[MAXP];                     // Q mod p
  uint32_t r32_sh[MAXP], r64_sh[MAXP];      // their Shoup companions
  uint64_t Qhalf[2];                        // floor(Q / 2) as (lo, hi)
  uint32_t mont[MAXP], mont_sh[MAXP];       // 2^32 mod p as a Shoup constant
  // Pre-transformed key words carry 2^32 (Montgomery form) AND scale[0] = -m^-1 (P_L/p)^-1: the external product is linear in
  // the key, so the CRT pre-scaling costs nothing per step (it used to be one extra multiplication per coefficient pair in
  // the last inverse stage).  lastw = psi^(-m/2), the bare twiddle of that stage.
  uint32_t keymul[MAXP], keymul_sh[MAXP];
  uint32_t lastw[MAXP], lastw_sh[MAXP];
  uint32_t dig_negc[MAXP];                  // p - (2^46 mod p) - 4p (mod 2^32), see digit_mod
  uint32_t scale[2][MAXP], scale_sh[2][MAXP];   // [0]: -m^-1 (P_L/p)^-1 (digits are transformed negated) ; [1]: 2^32 m^-1 (P_LM/p)^-1  (mod p)
  uint32_t scale_w1[MAXP], scale_w1_sh[MAXP];   // the same for scale[1] (standalone products)
  uint2 topf[MAXP][15], topi[MAXP][15];     // v4 kernels: twiddles of the top stages held in registers (forward, inverse), index k - 1 for tw[k]
  uint2 topf_h[MAXP][2][7], topi_h[MAXP][2][7];   // v5: per half h of a radix-16 group, the radix-8 after / before its first stage (direct / inverse table entries)
  uint32_t crt_c[2][MAXP][3];               // [1]: (P/p_i) mod Q; [0]: -(P/p_i) mod Q (the bootstrap sums represent -z), 32-bit limbs
  uint32_t negP[2][3];                      // [1]: (-P) mod Q; [0]: (+P) mod Q
  // FP64 head of the v4 bootstrap kernel (head_stage1_f64): p, 1/p, the stage-1 twiddle psi^(m/2), its quotient w/p and
  // the conversion constant 1.5 2^52 + 2p, all as doubles
  double hp_p[MAXP], hp_pinv[MAXP], hp_w[MAXP], hp_wp[MAXP], hp_c[MAXP];
};

// ---- swizzled shared-memory index: keeps every radix-8 pass bank-conflict free ---------------------
// bits 3,4 ^= bits 6,7 (scalar accesses of the stride-8 / stride-64 passes); bit 2 ^= bit 5 (the two 16-byte
// halves a thread reads in the stride-1 pass land in different bank groups within a quarter warp).
__device__ __forceinline__ int swz(int i) { return i ^ (((i >> 6) & 3) << 3) ^ (((i >> 5) & 1) << 2); }

// ---- 32-bit modular primitives (p < 2^30) -----------------------------------------------------------
// Shoup: x any 32-bit value, w < p, wsh = floor(w 2^32 / p)  ->  x*w mod p in [0, 2p)
__device__ __forceinline__ uint32_t shoup_mul(uint32_t x, uint32_t w, uint32_t wsh, uint32_t p) {
  return x * w - __umulhi(x, wsh) * p;
}
__device__ __forceinline__ uint32_t csub(uint32_t x, uint32_t p) { return min(x, x - p); }   // [0,2p) -> [0,p)

// Harvey butterflies.  Forward (Cooley-Tukey): x in [0,4p), y any -> both in [0,4p).
// `z` is DevConst::zero: x + t + z compiles to IADD3 (ALU pipe) instead of IMAD.IADD (the saturated FMA pipe).
__device__ __forceinline__ void ct_bfly(uint32_t& x, uint32_t& y, uint2 w, uint32_t p, uint32_t p2, uint32_t z) {
  const uint32_t xr = min(x, x - p2);
  const uint32_t t = shoup_mul(y, w.x, w.y, p);
  x = xr + t + z;
  y = xr - t + p2;
}
// Inverse (Gentleman-Sande): x, y in [0,2p) -> both in [0,2p).
__device__ __forceinline__ void gs_bfly(uint32_t& x, uint32_t& y, uint2 w, uint32_t p, uint32_t p2, uint32_t z) {
  const uint32_t s = x + y + z, d = x - y + p2;
  x = min(s, s - p2);
  y = shoup_mul(d, w.x, w.y, p);
}
// The same with the NEGATED twiddle folded into the subtraction order: (x - y) (-w) = (y - x) w.  The v4 kernel reads the
// mirrored FORWARD table entry w = -psi^-k directly and never forms the inverse twiddle.
__device__ __forceinline__ void gs_bfly_negw(uint32_t& x, uint32_t& y, uint2 w, uint32_t p, uint32_t p2, uint32_t z) {
  const uint32_t s = x + y + z, d = y - x + p2;
  x = min(s, s - p2);
  y = shoup_mul(d, w.x, w.y, p);
}
// Montgomery reduction of T < p 2^32 -> T 2^-32 mod p in [0,2p)
__device__ __forceinline__ uint32_t redc(uint64_t T, uint32_t p, uint32_t pinv_neg) {
  const uint32_t mq = (uint32_t)T * pinv_neg;
  return (uint32_t)((T + (uint64_t)mq * p) >> 32);
}

// radix-2^LOGR register blocks; w[(1<<l)-1+g] is the twiddle of group g at level l
template <int LOGR>
__device__ __forceinline__ void fwd_block(uint32_t (&x)[1 << LOGR], const uint2* w, uint32_t p, uint32_t p2, uint32_t z) {
  constexpr int R = 1 << LOGR;
#pragma unroll
  for (int l = 0; l < LOGR; ++l) {
    const int half = R >> (l + 1);
#pragma unroll
    for (int g = 0; g < (1 << l); ++g)
#pragma unroll
      for (int k = 0; k < half; ++k) ct_bfly(x[g * 2 * half + k], x[g * 2 * half + k + half], w[(1 << l) - 1 + g], p, p2, z);
  }
}
template <int LOGR, bool NEGW = false>
__device__ __forceinline__ void inv_block(uint32_t (&x)[1 << LOGR], const uint2* w, uint32_t p, uint32_t p2, uint32_t z) {
  constexpr int R = 1 << LOGR;
#pragma unroll
  for (int l = LOGR - 1; l >= 0; --l) {
    const int half = R >> (l + 1);
#pragma unroll
    for (int g = 0; g < (1 << l); ++g)
#pragma unroll
      for (int k = 0; k < half; ++k) {
        if (NEGW) gs_bfly_negw(x[g * 2 * half + k], x[g * 2 * half + k + half], w[(1 << l) - 1 + g], p, p2, z);
        else gs_bfly(x[g * 2 * half + k], x[g * 2 * half + k + half], w[(1 << l) - 1 + g], p, p2, z);
      }
  }
}

// fwd_block without its first level (l = 0): the caller has already run it (head_stage1_f64)
template <int LOGR>
__device__ __forceinline__ void fwd_block_tail(uint32_t (&x)[1 << LOGR], const uint2* w, uint32_t p, uint32_t p2, uint32_t z) {
  constexpr int R = 1 << LOGR;
#pragma unroll
  for (int l = 1; l < LOGR; ++l) {
    const int half = R >> (l + 1);
#pragma unroll
    for (int g = 0; g < (1 << l); ++g)
#pragma unroll
      for (int k = 0; k < half; ++k) ct_bfly(x[g * 2 * half + k], x[g * 2 * half + k + half], w[(1 << l) - 1 + g], p, p2, z);
  }
}

// First forward stage of a digit polynomial on the FP64 pipe (idle otherwise; DFMA issues at the IMAD rate).
// d[k] are signed digits (|d| < 2^46) held exactly in doubles; pairs (k, k + R/2) with the single stage-1 twiddle w:
//   t  = d[k+R/2] w mod p, centred:  h + l = d w exactly (h = fl(d w), l = fma(d, w, -h)), q = rint(d (w/p)) by the
//        1.5 2^52 trick, t = fma(-q, p, h) + l  -- both exact (integers below 2^53), |t| <= (1/2 + 2^-7) p;
//   ra = d[k] mod p, centred, the same way;  x = ra + t + 2p, y = ra - t + 2p land in (p/2, 7p/2), inside [0, 4p), the
//   input range of the integer butterflies.  Adding 1.5 2^52 leaves the integer in the low word of the double.
// Replaces per pair 2 digit reductions + 1 Harvey butterfly (10.6 FMA-heavy slots) by 12 FP64 instructions.
template <int R>
__device__ __forceinline__ void head_stage1_f64(const double (&d)[R], uint32_t (&x)[R], double p, double pinv, double w,
                                                double wp, double cm) {
  constexpr double M = 6755399441055744.0;              // 1.5 2^52
#pragma unroll
  for (int k = 0; k < R / 2; ++k) {
    const double a = d[k], b = d[k + R / 2];
    const double h = __dmul_rn(b, w);
    const double l = __fma_rn(b, w, -h);
    const double q = __dadd_rn(__fma_rn(b, wp, M), -M);
    const double t = __dadd_rn(__fma_rn(-q, p, h), l);
    const double qa = __dadd_rn(__fma_rn(a, pinv, M), -M);
    const double rc = __dadd_rn(__fma_rn(-qa, p, a), cm);
    x[k] = (uint32_t)__double2loint(__dadd_rn(rc, t));
    x[k + R / 2] = (uint32_t)__double2loint(__dadd_rn(rc, -t));
  }
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP / SYNCS) ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// bounded spin (a lost copy traps instead of hanging the GPU)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 28); ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
// one elected thread: stage `bytes` (multiple of 16) of a twiddle table into shared memory
__device__ __forceinline__ void stage_table(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy reads of dst are done (barrier before)
  mbar_expect_tx(bar, bytes);
  constexpr uint32_t CH = 16384;
  for (uint32_t o = 0; o < bytes; o += CH)
    bulk_g2s(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, bytes - o < CH ? bytes - o : CH, bar);
}

// Ampere-style asynchronous 4-byte copies global -> shared (SASS LDGSTS): no destination registers, so data can be
// requested across a barrier without lengthening any register live range
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// inverse block without its last level (l = 0, pairs (k, k + R/2)): the caller folds a scaling into that level
template <int LOGR>
__device__ __forceinline__ void inv_block_upper(uint32_t (&x)[1 << LOGR], const uint2* w, uint32_t p, uint32_t p2, uint32_t z) {
  constexpr int R = 1 << LOGR;
#pragma unroll
  for (int l = LOGR - 1; l >= 1; --l) {
    const int half = R >> (l + 1);
#pragma unroll
    for (int g = 0; g < (1 << l); ++g)
#pragma unroll
      for (int k = 0; k < half; ++k) gs_bfly(x[g * 2 * half + k], x[g * 2 * half + k + half], w[(1 << l) - 1 + g], p, p2, z);
  }
}

// Compile-time shape of the on-chip transform of length M = 2^LOGM.
//   REM   = LOGM % 3 top stages are fused into the phases that fill / drain shared memory;
//   NPASS = LOGM / 3 radix-8 passes run over shared memory, active bits [3k, 3k+3);
//   G threads own one polynomial (one radix-8 block each), NG polynomials are processed side by side.
template <int LOGM>
struct Shape {
  static constexpr int M = 1 << LOGM;
  static constexpr int REM = LOGM % 3;
  static constexpr int NPASS = LOGM / 3;
  static constexpr int G = M / 8;
  static constexpr int T = (M / 2 > 1024) ? 1024 : M / 2;
  static constexpr int NG = T / G > 0 ? T / G : 1;
  static constexpr int LPT = G > T ? G / T : 1;   // radix-8 blocks per thread and polynomial when one polynomial has more blocks than the CTA threads (m = 16384)
  static constexpr int STR = M >> REM;       // stride of the fused top stages
  // RNS basis sizes implied by Params(n), m = 8n (checked against the host derivation in sgfhe_ctx_create)
  static constexpr int L = LOGM <= 10 ? 4 : (LOGM <= 13 ? 5 : 6);     // bootstrap external product
  static constexpr int LM = LOGM <= 10 ? 5 : (LOGM <= 12 ? 6 : 7);    // product of two full-size operands
  // m = 16384 (Params(2048), 93-bit Q): four transform buffers (256 KiB) do not fit an SM, so the "wide" kernels work on two
  // polynomials at a time and read their twiddles from global memory (L1/L2) instead of a staged table
  static constexpr bool WIDE = LOGM >= 14;
};

// One radix-8 pass (active bits [B, B+3)) over NPOLY polynomials in shared memory.
// Twiddles: tw[i] = (psi^bitrev(i), Shoup companion); the butterfly on bit b' of element idx uses
// tw[(M + idx) >> (b'+1)], so a block with base index `base` needs tw[t1], tw[2 t1 + {0,1}], tw[4 t1 + {0..3}].
template <int LOGM, int NPOLY, bool FWD, int B>
__device__ __forceinline__ void ntt_pass8(uint32_t* sm, const uint2* tw, uint32_t p, uint32_t z) {
  using S = Shape<LOGM>;
  constexpr int M = S::M;
  const int grp = S::LPT > 1 ? 0 : threadIdx.x / S::G;
  if (grp >= NPOLY) return;
  const uint32_t p2 = 2 * p;
#pragma unroll 1
  for (int lq = 0; lq < S::LPT; ++lq) {
  const int lane = S::LPT > 1 ? (int)threadIdx.x + lq * S::T : (int)threadIdx.x % S::G;
  const int base = ((lane >> B) << (B + 3)) | (lane & ((1 << B) - 1));
  const int t1 = (M >> (B + 3)) + (lane >> B);
  uint2 w[7];
  {
    w[0] = tw[t1];
    const uint4 a = *reinterpret_cast<const uint4*>(&tw[2 * t1]);
    w[1] = make_uint2(a.x, a.y); w[2] = make_uint2(a.z, a.w);
    const uint4 b0 = *reinterpret_cast<const uint4*>(&tw[4 * t1]);
    const uint4 b1 = *reinterpret_cast<const uint4*>(&tw[4 * t1 + 2]);
    w[3] = make_uint2(b0.x, b0.y); w[4] = make_uint2(b0.z, b0.w);
    w[5] = make_uint2(b1.x, b1.y); w[6] = make_uint2(b1.z, b1.w);
  }
  // swizzled addresses of the 8 elements, hoisted out of the polynomial loop
  int off[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) off[j] = swz(base + (j << B));
#pragma unroll
  for (int q = 0; q < (NPOLY + S::NG - 1) / S::NG; ++q) {
    const int poly = grp + q * S::NG;
    if (poly < NPOLY) {
      uint32_t* s = sm + poly * M;
      uint32_t x[8];
      if (B == 0) {
        const uint4 v0 = *reinterpret_cast<const uint4*>(s + off[0]);
        const uint4 v1 = *reinterpret_cast<const uint4*>(s + off[4]);
        x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = s[off[j]];
      }
      if (FWD) fwd_block<3>(x, w, p, p2, z); else inv_block<3>(x, w, p, p2, z);
      if (B == 0) {
        *reinterpret_cast<uint4*>(s + off[0]) = make_uint4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<uint4*>(s + off[4]) = make_uint4(x[4], x[5], x[6], x[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) s[off[j]] = x[j];
      }
    }
  }
  }
}

template <int LOGM, int NPOLY, bool FWD, int K>
struct PassLoop {
  static __device__ __forceinline__ void run(uint32_t* sm, const uint2* tw, uint32_t p, uint32_t z, const uint2* tw_low = nullptr) {
    constexpr int NP = Shape<LOGM>::NPASS;
    constexpr int B = FWD ? 3 * (NP - 1 - K) : 3 * K;
    // tw_low: a copy of table entries [0, M/8) (all the strided passes read: the pass with active bits [B, B+3) uses entries
    // below M >> B), e.g. in shared memory while the full table stays in global memory (wide kernels)
    ntt_pass8<LOGM, NPOLY, FWD, B>(sm, (B >= 3 && tw_low) ? tw_low : tw, p, z);
    // bits [0,6) of the index stay inside one group of 8 consecutive threads (one warp): the stride-8 and
    // stride-1 passes exchange data only within that group, so a warp-level barrier separates them.
    if ((FWD && B == 3) || (!FWD && B == 0 && NP > 1)) __syncwarp(); else __syncthreads();
    PassLoop<LOGM, NPOLY, FWD, K + 1>::run(sm, tw, p, z, tw_low);
  }
};
template <int LOGM, int NPOLY, bool FWD>
struct PassLoop<LOGM, NPOLY, FWD, Shape<LOGM>::NPASS> {
  static __device__ __forceinline__ void run(uint32_t*, const uint2*, uint32_t, uint32_t, const uint2* = nullptr) {}
};
// the NPASS radix-8 passes over bits [0, 3 NPASS); each pass ends with __syncthreads().
// Forward: high bits first (after the fused REM top stages); inverse: low bits first.
template <int LOGM, int NPOLY, bool FWD>
__device__ __forceinline__ void ntt_passes(uint32_t* sm, const uint2* tw, uint32_t p, uint32_t z, const uint2* tw_low = nullptr) {
  PassLoop<LOGM, NPOLY, FWD, 0>::run(sm, tw, p, z, tw_low);
}

// uniform twiddles of the fused top stages: tw[1 .. 2^REM - 1]
template <int REM>
__device__ __forceinline__ void top_twiddles(const uint2* __restrict__ tw, uint2* w) {
#pragma unroll
  for (int k = 1; k < (1 << REM); ++k) w[k - 1] = __ldg(&tw[k]);
}

// ---- 96-bit limb arithmetic (Q < 2^93) ---------------------------------------------------------------
struct u96 { uint32_t x0, x1, x2; };

__device__ __forceinline__ u96 add96(u96 a, u96 b) {
  u96 r;
  asm("add.cc.u32 %0, %3, %6;\n\taddc.cc.u32 %1, %4, %7;\n\taddc.u32 %2, %5, %8;"
      : "=r"(r.x0), "=r"(r.x1), "=r"(r.x2) : "r"(a.x0), "r"(a.x1), "r"(a.x2), "r"(b.x0), "r"(b.x1), "r"(b.x2));
  return r;
}
// r = a - b, borrow = 0xFFFFFFFF if a < b else 0
__device__ __forceinline__ u96 sub96(u96 a, u96 b, uint32_t& borrow) {
  u96 r;
  asm("sub.cc.u32 %0, %4, %7;\n\tsubc.cc.u32 %1, %5, %8;\n\tsubc.cc.u32 %2, %6, %9;\n\tsubc.u32 %3, 0, 0;"
      : "=r"(r.x0), "=r"(r.x1), "=r"(r.x2), "=r"(borrow)
      : "r"(a.x0), "r"(a.x1), "r"(a.x2), "r"(b.x0), "r"(b.x1), "r"(b.x2));
  return r;
}
__device__ __forceinline__ u96 sel96(bool c, u96 a, u96 b) { u96 r; r.x0 = c ? a.x0 : b.x0; r.x1 = c ? a.x1 : b.x1; r.x2 = c ? a.x2 : b.x2; return r; }
__device__ __forceinline__ u96 Q96(const DevConst& C) { u96 q; q.x0 = C.Ql[0]; q.x1 = C.Ql[1]; q.x2 = C.Ql[2]; return q; }
// x in [0, 2Q) -> [0, Q)
__device__ __forceinline__ u96 csubQ(u96 x, u96 Q) { uint32_t bw; const u96 t = sub96(x, Q, bw); return sel96(bw != 0, x, t); }
__device__ __forceinline__ u96 addmod96(u96 a, u96 b, u96 Q) { return csubQ(add96(a, b), Q); }
__device__ __forceinline__ u96 submod96(u96 a, u96 b, u96 Q) { uint32_t bw; const u96 t = sub96(a, b, bw); return sel96(bw != 0, add96(t, Q), t); }
__device__ __forceinline__ u96 negmod96(u96 a, u96 Q) { uint32_t bw; const u96 t = sub96(Q, a, bw); return sel96((a.x0 | a.x1 | a.x2) == 0, a, t); }
__device__ __forceinline__ u128 to128(u96 a) { return (u128)a.x0 | ((u128)a.x1 << 32) | ((u128)a.x2 << 64); }
__device__ __forceinline__ u96 from128(u128 v) { u96 r; r.x0 = (uint32_t)v; r.x1 = (uint32_t)(v >> 32); r.x2 = (uint32_t)(v >> 64); return r; }

__device__ __forceinline__ u96 ld96(const uint32_t* base, int stride, int idx) {
  u96 r; r.x0 = base[idx]; r.x1 = base[stride + idx]; r.x2 = base[2 * stride + idx]; return r;
}
__device__ __forceinline__ void st96(uint32_t* base, int stride, int idx, u96 v) {
  base[idx] = v.x0; base[stride + idx] = v.x1; base[2 * stride + idx] = v.x2;
}

// ---- wide (u128) helpers: seams and one-off kernels only ------------------------------------------------
__device__ __forceinline__ u128 addmodQ(u128 a, u128 b, u128 Q) { u128 s = a + b; return s >= Q ? s - Q : s; }
__device__ __forceinline__ u128 submodQ(u128 a, u128 b, u128 Q) { return a >= b ? a - b : a + Q - b; }
__device__ __forceinline__ u128 negmodQ(u128 a, u128 Q) { return a ? Q - a : (u128)0; }

// The accumulator is kept in OFFSET FORM a + offs mod Q (offs = s (1 + B), src/utils.jl:169,179), so the
// decomposition starts at the divrem.  flatten(nothing, a, Val(B), Val(2)) (src/utils.jl:155-189) on a_off = a + offs:
// returns the biased digit words dp_i = u_i - s + dig_bias (2^46; 2^48 at n = 2048).  KB = log2(B / 35) = 3 LOGM - 1 (B = 35 r^2 n, src/fhe.jl:87).
template <int KB>
__device__ __forceinline__ void decompose_off(const DevConst& C, u96 a, uint64_t& dp0, uint64_t& dp1) {
  // divrem(a, B), B = 35 2^KB   (src/utils.jl:172);  t = a >> KB < 35^2 2^KB < 2^49
  const uint64_t lo64 = (uint64_t)a.x0 | ((uint64_t)a.x1 << 32);
  const uint64_t hi64 = (uint64_t)a.x1 | ((uint64_t)a.x2 << 32);
  const uint64_t t = KB >= 32 ? (hi64 >> (KB - 32)) : ((lo64 >> KB) | ((uint64_t)a.x2 << (64 - KB)));
  const uint64_t lo = lo64 & ((1ull << KB) - 1);
  // u1 = floor(t / 35) on the idle FP64 pipe: (2^52 + t) - 2^52 = t exactly, then trunc(t * inv35 + 2^52) has
  // floor(t inv35) = floor(t / 35) in its mantissa (t inv35 - t/35 < 2^-8 < 1/35)
  const double td = __hiloint2double(0x43300000 | (int)(uint32_t)(t >> 32), (int)(uint32_t)t) - 4503599627370496.0;
  const double qd = __fma_rz(td, C.inv35, 4503599627370496.0);
  const uint32_t q_lo = (uint32_t)__double2loint(qd), q_hi = (uint32_t)__double2hiint(qd) & 0xFFFFFu;
  const uint32_t rem = (uint32_t)t - 35u * q_lo;                    // < 35
  const uint64_t u1 = (uint64_t)q_lo | ((uint64_t)q_hi << 32);
  const uint64_t u0 = ((uint64_t)rem << KB) | lo;
  dp0 = u0 + C.s46;                               // - s  (src/utils.jl:183-185), + storage bias
  dp1 = u1 + C.s46;
}

// flatten(rng, ...) (src/utils.jl:198-241) on the offset form: x0, x1 are the caller's draws
template <int KB>
__device__ __forceinline__ void decompose_off_rand(const DevConst& C, u96 a, int64_t x0, int64_t x1, uint64_t& dp0, uint64_t& dp1) {
  i128 X = (i128)x1 * (i128)C.B + (i128)x0;       // rand_a = a - x0 - x1 B   (src/utils.jl:222,232-233)
  // X mod Q without a 128-bit division: |x| <= xmax = 3B/2 and Q > 0.997 B^2 (Q in [1220, 1225] r^4 n^2, B^2 = 1225 r^4 n^2),
  // so |X| < 1.51 Q and two conditional corrections reach [0, Q)
  const i128 Qs = (i128)C.Q;
  if (X < 0) X += Qs;
  if (X < 0) X += Qs;
  if (X >= Qs) X -= Qs;
  a = from128(submodQ(to128(a), (u128)X, C.Q));
  decompose_off<KB>(C, a, dp0, dp1);
  dp0 += (uint64_t)x0;                            // + x                     (src/utils.jl:236-238)
  dp1 += (uint64_t)x1;
}
// canonical accumulator value <-> offset form
__device__ __forceinline__ u96 off96(const DevConst& C) { u96 o; o.x0 = C.offl[0]; o.x1 = C.offl[1]; o.x2 = C.offl[2]; return o; }
__device__ __forceinline__ u96 to_offset_form(const DevConst& C, u96 a) { return addmod96(a, off96(C), Q96(C)); }
__device__ __forceinline__ u96 from_offset_form(const DevConst& C, u96 a) { return submod96(a, off96(C), Q96(C)); }

// Signed digits (|d| < dig_bias) are stored biased, dp = d + dig_bias < 2^49, as two words: lo = dp mod 2^32, hi = dp >> 18.
__device__ __forceinline__ uint2 digit_words(uint64_t dp) { return make_uint2((uint32_t)dp, (uint32_t)(dp >> 18)); }
// the NEGATED digit as a double (exact): bits 0x433 | dp are 2^52 + dp, and (2^52 + 2^46) - (2^52 + dp) = -d
__device__ __forceinline__ double digit_f64(uint64_t dp) {
  return __dadd_rn(4573968371548160.0, -__hiloint2double((int)(0x43300000u | (uint32_t)(dp >> 32)), (int)(uint32_t)dp));
}
// residue of the NEGATED digit mod p in (p, 4p]: mu = floor(2^50 / p), negc4 = p - (dig_bias mod p) - 4p (mod 2^32).
// q p - (lo + negc4) is one IMAD with a negated addend; lo - q p + negc would need an extra register move to negate q.
// All four digit polynomials change sign together, which the CRT pre-scaling constants of the bootstrap basis undo
// (scale[0], scale_w are stored negated).
__device__ __forceinline__ uint32_t digit_mod(uint32_t lo, uint32_t hi, uint32_t mu, uint32_t negc4, uint32_t p) {
  return __umulhi(hi, mu) * p - (lo + negc4);
}

// canonical value c = lo + 2^64 hi of Z_Q, centred to (-Q/2, Q/2], as a residue mod p_i in [0,p).
// c = w0 + w1 2^32 + w2 2^64 with 32-bit words: two Shoup products by (2^32 mod p), (2^64 mod p), and w0 reduced by its
// top two bits (p > 2^30 - 2^27, checked on the host, keeps w0 - (w0 >> 30) p below 2p).
__device__ __forceinline__ uint32_t centred_mod(const DevConst& C, int i, uint64_t lo, uint64_t hi) {
  const uint32_t p = C.p[i], p2 = 2 * p;
  const uint32_t w0 = (uint32_t)lo, w1 = (uint32_t)(lo >> 32), w2 = (uint32_t)hi;
  uint32_t t = shoup_mul(w1, C.r32[i], C.r32_sh[i], p) + shoup_mul(w2, C.r64[i], C.r64_sh[i], p);   // [0, 4p)
  t = min(t, t - p2);
  t += w0 - (w0 >> 30) * p;                                                                          // [0, 4p)
  t = min(t, t - p2); t = min(t, t - p);
  const bool upper = hi > C.Qhalf[1] || (hi == C.Qhalf[1] && lo > C.Qhalf[0]);                      // c > Q/2: c - Q
  const uint32_t tn = t - C.qmodp[i];
  return upper ? min(tn, tn + p) : t;
}

// canonical value c = lo + 2^64 hi of Z_Q as a residue mod p_i in [0, 4p) (the input range of the forward butterflies),
// NOT centred: for products whose RNS basis has room for m Q^2 (DevConst::pm_uncentred)
__device__ __forceinline__ uint32_t plain_mod(const DevConst& C, int i, uint64_t lo, uint64_t hi) {
  const uint32_t p = C.p[i], p2 = 2 * p;
  const uint32_t w0 = (uint32_t)lo, w1 = (uint32_t)(lo >> 32), w2 = (uint32_t)hi;
  uint32_t t = shoup_mul(w1, C.r32[i], C.r32_sh[i], p) + shoup_mul(w2, C.r64[i], C.r64_sh[i], p);   // [0, 4p)
  t = min(t, t - p2);
  return t + (w0 - (w0 >> 30) * p);                                                                  // [0, 4p)
}

// 128-bit limb helpers for the unreduced CRT sums
__device__ __forceinline__ uint4 add128(uint4 a, uint4 b) {
  uint4 r;
  asm("add.cc.u32 %0, %4, %8;\n\taddc.cc.u32 %1, %5, %9;\n\taddc.cc.u32 %2, %6, %10;\n\taddc.u32 %3, %7, %11;"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w));
  return r;
}
__device__ __forceinline__ uint4 sub128(uint4 a, uint4 b) {
  uint4 r;
  asm("sub.cc.u32 %0, %4, %8;\n\tsubc.cc.u32 %1, %5, %9;\n\tsubc.cc.u32 %2, %6, %10;\n\tsubc.u32 %3, %7, %11;"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w));
  return r;
}

// CRT sum: residues y_i = z (P/p_i)^-1 mod p_i (canonical) of an integer |z| < P/32  ->  S = sum y_i c_i + v negP, an
// UNREDUCED representative of z mod Q with 0 <= S < (K 2^30 + K) Q, as four limbs.
//   v = round(sum y_i / p_i) from the top 14 bits of each residue (error < 2^-11, margin 0.47).
template <int BASIS, int K>
__device__ __forceinline__ uint4 crt_sum(const DevConst& C, const uint32_t* __restrict__ y, size_t stride) {
  uint64_t colA[3] = {0, 0, 0}, colB[3] = {0, 0, 0};
  uint32_t vs = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const uint32_t yi = y[i * stride];
    vs += (yi >> 16) * C.vk[i];                               // vk = round(2^44 / p): (y>>16) vk ~ (y/p) 2^28
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (i < 4) colA[k] += (uint64_t)yi * C.crt_c[BASIS][i][k];
      else colB[k] += (uint64_t)yi * C.crt_c[BASIS][i][k];
    }
  }
  const uint32_t v = (vs + (1u << 27)) >> 28;                 // round(sum y_i / p_i), 0 <= v <= K
#pragma unroll
  for (int k = 0; k < 3; ++k) colB[k] += (uint64_t)v * C.negP[BASIS][k];
  // S = sum_k (colA[k] + colB[k]) 2^(32k) < 2^126, as four limbs
  uint4 S;
  const uint64_t c0 = colA[0] + colB[0]; const uint32_t k0 = c0 < colA[0];
  const uint64_t c1 = colA[1] + colB[1]; const uint32_t k1 = c1 < colA[1];
  const uint64_t c2 = colA[2] + colB[2];                      // top column cannot overflow (c_i[2] < 2^29)
  S.x = (uint32_t)c0;
  const uint64_t t1 = (c0 >> 32) + (uint32_t)c1;
  S.y = (uint32_t)t1;
  const uint64_t t2 = (t1 >> 32) + (c1 >> 32) + k0 + (uint32_t)c2;
  S.z = (uint32_t)t2;
  S.w = (uint32_t)((t2 >> 32) + (c2 >> 32) + k1);
  return S;
}

// V mod Q for 0 <= V < 2^35 Q (four limbs); SB = bits(Q) - 1 = 6 LOGM + 8 (Q in [1220, 1225] r^4 n^2, src/fhe.jl:64-69).
// Quotient estimate on the otherwise idle FP64 pipe, without integer<->double conversion instructions:
//   T = floor(V / 2^(SB-16)) < 2^52 is placed in the mantissa of 2^52 + T, td = (2^52 + T) - 2^52 = T exactly, and
//   trunc(td * barrett_inv + 2^52) carries qh = floor(T * barrett_inv) in its mantissa.  barrett_inv = 2^(SB-16)/Q (1 - 2^-40)
//   never exceeds the true ratio, so floor(V/Q) - 1 <= qh <= floor(V/Q) and one conditional subtraction finishes.
template <int SB>
__device__ __forceinline__ u96 barrett96(const DevConst& C, uint4 V) {
  constexpr int SH = SB - 16;                                 // 46 <= SH <= 70
  uint32_t tl, th;
  if constexpr (SH >= 64) { tl = __funnelshift_r(V.z, V.w, SH - 64); th = V.w >> (SH - 64); }
  else { tl = __funnelshift_r(V.y, V.z, SH - 32); th = __funnelshift_r(V.z, V.w, SH - 32); }
  const double td = __hiloint2double((int)(0x43300000u | th), (int)tl) - 4503599627370496.0;
  const double qd = __fma_rz(td, C.barrett_inv, 4503599627370496.0);
  const uint32_t q0 = (uint32_t)__double2loint(qd), q1 = (uint32_t)__double2hiint(qd) & 0xFFFFFu;
  const uint64_t m0 = (uint64_t)q0 * C.Ql[0];
  const uint64_t m1 = (uint64_t)q0 * C.Ql[1] + (m0 >> 32);
  const uint64_t m1b = (uint64_t)q1 * C.Ql[0] + (uint32_t)m1;
  u96 prod; prod.x0 = (uint32_t)m0; prod.x1 = (uint32_t)m1b;
  prod.x2 = q0 * C.Ql[2] + q1 * C.Ql[1] + (uint32_t)(m1 >> 32) + (uint32_t)(m1b >> 32);
  u96 S; S.x0 = V.x; S.x1 = V.y; S.x2 = V.z;
  uint32_t bw;
  const u96 R = sub96(S, prod, bw);                           // exact mod 2^96; true value in [0, 2Q), 2Q < 2^96
  return csubQ(R, Q96(C));
}

// CRT lift to the canonical residue mod Q (standalone products; the bootstrap path keeps the sums unreduced)
template <int BASIS, int K, int SB>
__device__ __forceinline__ u96 crt_lift(const DevConst& C, const uint32_t* __restrict__ y, size_t stride) {
  return barrett96<SB>(C, crt_sum<BASIS, K>(C, y, stride));
}

// rescale(r, x, Q, round=true) with r = 2^logr (src/utils.jl:78-92 via reduce_modulus src/utils.jl:107-117)
__device__ __forceinline__ uint64_t modred(const DevConst& C, u128 x) {
  u128 rem = x; uint64_t q = 0;
  for (int i = 0; i < C.logr; ++i) {
    rem <<= 1; q <<= 1;
    if (rem >= C.Q) { rem -= C.Q; q |= 1; }
  }
  if (rem >= (C.Q >> 1) + (u128)((uint32_t)C.Q & 1)) { q += 1; if (q == (1ull << C.logr)) q = 0; }
  return q;
}

}  // namespace sgfhe
