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
  u128 Q, DQ, offs, B;          // offs = s * (1 + B) mod Q (src/utils.jl:169)
  uint64_t s;                   // digit offset (src/utils.jl:162-166)
  uint64_t barrett_mu;          // floor(2^(sbits+35) / Q)
  uint32_t p[MAXP], pinv_neg[MAXP], dig_mu[MAXP], vinv[MAXP];
  uint32_t r32[MAXP], r64[MAXP];            // 2^32 mod p, 2^64 mod p
  uint32_t qmodp[MAXP];                     // Q mod p
  uint32_t mont[MAXP], mont_sh[MAXP];       // 2^32 mod p as a Shoup constant
  uint64_t dig_bias[MAXP];                  // multiple of p, >= 2^47
  uint32_t scale[2][MAXP], scale_sh[2][MAXP];   // [0]: m^-1 (P_L/p)^-1 ; [1]: 2^32 m^-1 (P_LM/p)^-1  (mod p)
  uint32_t crt_c[2][MAXP][3];               // (P/p_i) mod Q, 32-bit limbs
  uint32_t negP[2][3];                      // (-P) mod Q
};

// ---- swizzled shared-memory index: keeps every radix-8 pass bank-conflict free ---------------------
__device__ __forceinline__ int swz(int i) { return i ^ (((i >> 6) & 3) << 3); }

// ---- 32-bit modular primitives (p < 2^30) -----------------------------------------------------------
// Shoup: x any 32-bit value, w < p, wsh = floor(w 2^32 / p)  ->  x*w mod p in [0, 2p)
__device__ __forceinline__ uint32_t shoup_mul(uint32_t x, uint32_t w, uint32_t wsh, uint32_t p) {
  return x * w - __umulhi(x, wsh) * p;
}
__device__ __forceinline__ uint32_t csub(uint32_t x, uint32_t p) { return min(x, x - p); }   // [0,2p) -> [0,p)

// Harvey butterflies.  Forward (Cooley-Tukey): x in [0,4p), y any -> both in [0,4p).
__device__ __forceinline__ void ct_bfly(uint32_t& x, uint32_t& y, uint2 w, uint32_t p, uint32_t p2) {
  const uint32_t xr = min(x, x - p2);
  const uint32_t t = shoup_mul(y, w.x, w.y, p);
  x = xr + t;
  y = xr - t + p2;
}
// Inverse (Gentleman-Sande): x, y in [0,2p) -> both in [0,2p).
__device__ __forceinline__ void gs_bfly(uint32_t& x, uint32_t& y, uint2 w, uint32_t p, uint32_t p2) {
  const uint32_t s = x + y, d = x - y + p2;
  x = min(s, s - p2);
  y = shoup_mul(d, w.x, w.y, p);
}
// Montgomery reduction of T < p 2^32 -> T 2^-32 mod p in [0,2p)
__device__ __forceinline__ uint32_t redc(uint64_t T, uint32_t p, uint32_t pinv_neg) {
  const uint32_t mq = (uint32_t)T * pinv_neg;
  return (uint32_t)((T + (uint64_t)mq * p) >> 32);
}

// ---- one radix-2^LOGR pass over `npoly` polynomials of N coefficients held in shared memory ---------
// Active index bits are [b, b+LOGR).  Twiddle table: tw[i] = (psi^bitrev(i), Shoup companion), i in [1,N);
// the butterfly on bit b' of element idx uses tw[(N + idx) >> (b'+1)].
template <int LOGR, bool FWD>
__device__ __forceinline__ void ntt_pass(uint32_t* sm, int npoly, int N, int b, const uint2* __restrict__ tw,
                                         uint32_t p) {
  constexpr int R = 1 << LOGR;
  const int nblk = N >> LOGR;
  const int nthr = blockDim.x;
  const int G = nblk < nthr ? nblk : nthr;
  const int ngrp = nthr / G;
  const int grp = threadIdx.x / G, lane = threadIdx.x - grp * G;
  const uint32_t p2 = 2 * p;
  if (grp >= ngrp) return;
  const int lowmask = (1 << b) - 1;
  for (int blk = lane; blk < nblk; blk += G) {
    const int base = ((blk & ~lowmask) << LOGR) | (blk & lowmask);
    const int t1 = (N + base) >> (b + LOGR);
    uint2 w[R - 1];
#pragma unroll
    for (int l = 0; l < LOGR; ++l)
#pragma unroll
      for (int j = 0; j < (1 << l); ++j) w[(1 << l) - 1 + j] = __ldg(&tw[(t1 << l) + j]);
    for (int poly = grp; poly < npoly; poly += ngrp) {
      uint32_t* s = sm + poly * N;
      uint32_t x[R];
      if (LOGR == 3 && b == 0) {
        const uint4* v = reinterpret_cast<const uint4*>(s + swz(base));
        const uint4 v0 = v[0], v1 = v[1];
        x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
        x[4 % R] = v1.x; x[5 % R] = v1.y; x[6 % R] = v1.z; x[7 % R] = v1.w;
      } else {
#pragma unroll
        for (int j = 0; j < R; ++j) x[j] = s[swz(base + (j << b))];
      }
      if (FWD) {
#pragma unroll
        for (int l = 0; l < LOGR; ++l) {
          const int half = R >> (l + 1);
#pragma unroll
          for (int g = 0; g < (1 << l); ++g)
#pragma unroll
            for (int k = 0; k < half; ++k) ct_bfly(x[g * 2 * half + k], x[g * 2 * half + k + half], w[(1 << l) - 1 + g], p, p2);
        }
      } else {
#pragma unroll
        for (int l = LOGR - 1; l >= 0; --l) {
          const int half = R >> (l + 1);
#pragma unroll
          for (int g = 0; g < (1 << l); ++g)
#pragma unroll
            for (int k = 0; k < half; ++k) gs_bfly(x[g * 2 * half + k], x[g * 2 * half + k + half], w[(1 << l) - 1 + g], p, p2);
        }
      }
      if (LOGR == 3 && b == 0) {
        uint4* v = reinterpret_cast<uint4*>(s + swz(base));
        v[0] = make_uint4(x[0], x[1], x[2], x[3]);
        v[1] = make_uint4(x[4 % R], x[5 % R], x[6 % R], x[7 % R]);
      } else {
#pragma unroll
        for (int j = 0; j < R; ++j) s[swz(base + (j << b))] = x[j];
      }
    }
  }
}

// Negacyclic forward NTT of `npoly` polys in shared memory: natural order in (values in [0,4p)),
// bit-reversed order out (values in [0,4p)).  Ends with __syncthreads().
__device__ __forceinline__ void ntt_forward(uint32_t* sm, int npoly, int N, int logN, const uint2* __restrict__ tw,
                                            uint32_t p) {
  const int rem = logN % 3;
  int b = logN - rem;
  if (rem == 1) { ntt_pass<1, true>(sm, npoly, N, b, tw, p); __syncthreads(); }
  if (rem == 2) { ntt_pass<2, true>(sm, npoly, N, b, tw, p); __syncthreads(); }
  while (b > 0) { b -= 3; ntt_pass<3, true>(sm, npoly, N, b, tw, p); __syncthreads(); }
}
// Inverse: bit-reversed in (values in [0,2p)), natural out ([0,2p)), WITHOUT the 1/N factor.
__device__ __forceinline__ void ntt_inverse(uint32_t* sm, int npoly, int N, int logN, const uint2* __restrict__ tw,
                                            uint32_t p) {
  const int rem = logN % 3;
  const int top = logN - rem;
  for (int b = 0; b < top; b += 3) { ntt_pass<3, false>(sm, npoly, N, b, tw, p); __syncthreads(); }
  if (rem == 1) { ntt_pass<1, false>(sm, npoly, N, top, tw, p); __syncthreads(); }
  if (rem == 2) { ntt_pass<2, false>(sm, npoly, N, top, tw, p); __syncthreads(); }
}

// ---- wide helpers --------------------------------------------------------------------------------------
__device__ __forceinline__ u128 addmodQ(u128 a, u128 b, u128 Q) { u128 s = a + b; return s >= Q ? s - Q : s; }
__device__ __forceinline__ u128 submodQ(u128 a, u128 b, u128 Q) { return a >= b ? a - b : a + Q - b; }
__device__ __forceinline__ u128 negmodQ(u128 a, u128 Q) { return a ? Q - a : (u128)0; }

__device__ __forceinline__ u128 load3(const uint32_t* base, int stride, int idx) {
  return (u128)base[idx] | ((u128)base[stride + idx] << 32) | ((u128)base[2 * stride + idx] << 64);
}
__device__ __forceinline__ void store3(uint32_t* base, int stride, int idx, u128 v) {
  base[idx] = (uint32_t)v; base[stride + idx] = (uint32_t)(v >> 32); base[2 * stride + idx] = (uint32_t)(v >> 64);
}

// flatten(rng|nothing, a, Val(B), Val(2)) as signed digits (src/utils.jl:155-189, 198-241).
// x0, x1 are the caller's draws (0, 0 for the deterministic form).
__device__ __forceinline__ void decompose(const DevConst& C, u128 a, int64_t x0, int64_t x1, bool random,
                                          int64_t& d0, int64_t& d1) {
  if (random) {                                   // rand_a = a - x0 - x1 B   (src/utils.jl:222,232-233)
    i128 X = (i128)x1 * (i128)C.B + (i128)x0;
    X %= (i128)C.Q;
    if (X < 0) X += (i128)C.Q;
    a = submodQ(a, (u128)X, C.Q);
  }
  a = addmodQ(a, C.offs, C.Q);                    // a += offset             (src/utils.jl:179)
  const uint64_t t = (uint64_t)(a >> C.kB);       // divrem(a, B), B = 35 2^kB (src/utils.jl:172)
  const uint64_t lo = (uint64_t)a & ((1ull << C.kB) - 1);
  const uint64_t u1 = t / 35u;
  const uint64_t u0 = ((t - u1 * 35u) << C.kB) | lo;
  d0 = (int64_t)(u0 - C.s) + x0;                  // - s (+ x)               (src/utils.jl:183-185, 236-238)
  d1 = (int64_t)(u1 - C.s) + x1;
}

// signed digit (|d| < 2^46) -> residue mod p_i in [0,3p)
__device__ __forceinline__ uint32_t digit_mod(const DevConst& C, int i, int64_t d) {
  const uint64_t dp = (uint64_t)(d + (int64_t)C.dig_bias[i]);
  const uint32_t q = __umulhi((uint32_t)(dp >> 18), C.dig_mu[i]);
  return (uint32_t)dp - q * C.p[i];
}

// canonical value of Z_Q, centred to (-Q/2, Q/2], as a residue mod p_i in [0,p)
__device__ __forceinline__ uint32_t centred_mod(const DevConst& C, int i, u128 c) {
  const uint32_t p = C.p[i];
  const uint64_t t = (uint64_t)(uint32_t)(c >> 64) * C.r64[i] + (uint64_t)(uint32_t)(c >> 32) * C.r32[i] + (uint32_t)c;
  uint32_t r = (uint32_t)(t % p);
  if (c > (C.Q >> 1)) r = r >= C.qmodp[i] ? r - C.qmodp[i] : r + p - C.qmodp[i];
  return r;
}

// CRT lift: residues y_i = z (P/p_i)^-1 mod p_i of an integer |z| < P/32  ->  z mod Q, canonical.
template <int BASIS>
__device__ __forceinline__ u128 crt_lift(const DevConst& C, int K, const uint32_t* y, size_t stride) {
  uint64_t colA[3] = {0, 0, 0}, colB[3] = {0, 0, 0}, vs = 0;
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    if (i < K) {
      const uint32_t yi = y[i * stride];
      vs += __umulhi(yi, C.vinv[i]);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (i < 4) colA[k] += (uint64_t)yi * C.crt_c[BASIS][i][k];
        else colB[k] += (uint64_t)yi * C.crt_c[BASIS][i][k];
      }
    }
  }
  const uint32_t v = (uint32_t)((vs + (1u << 28)) >> 29);   // round(sum y_i / p_i)
#pragma unroll
  for (int k = 0; k < 3; ++k) colB[k] += (uint64_t)v * C.negP[BASIS][k];
  const u128 S = (u128)colA[0] + colB[0] + (((u128)colA[1] + colB[1]) << 32) + (((u128)colA[2] + colB[2]) << 64);
  const uint64_t T = (uint64_t)(S >> (C.sbits - 29));
  const uint64_t qh = __umul64hi(T, C.barrett_mu);
  u128 R = S - (u128)qh * C.Q;
  if (R >= C.Q) R -= C.Q;
  if (R >= C.Q) R -= C.Q;
  return R;
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
