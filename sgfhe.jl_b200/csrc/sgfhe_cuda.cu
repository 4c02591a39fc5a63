// sgfhe_cuda.cu -- kernels and C ABI of libsgfhe_cuda.so (see include/sgfhe_cuda.h).
//
// Hot path = reference src/fhe.jl:559-595 (_bootstrap_internal) + src/fhe.jl:608-621 (bootstrap), in the
// algebraically identical form  (a,b) += (x^u - 1) * ([flatten(a); flatten(b)] . C^(k))  with the key C
// pre-transformed once.  One persistent CTA owns one gate for all n steps; only
// pre-transformed key tiles stream in (from L2: every resident CTA is on a nearby step).
#include "../../include/sgfhe_cuda.h"
#include "device_math.cuh"
#include "host_math.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace sgfhe;

// =========================================================================================================
// Kernels
// =========================================================================================================

enum : int { F_INIT = 1, F_FINAL = 2, F_RAW = 4, F_EXT = 8, F_DECOMP = 16, F_PACK = 32 };

struct GateArgs {
  const uint64_t* lwe1; const uint64_t* lwe2;   // [batch][n+1] over Z_r
  const int64_t* draws;                         // NULL or [batch][draw_steps][2][m][2]
  uint64_t* out_and; uint64_t* out_or; uint64_t* out_xor;   // [batch][n+1] (Z_r) or wide [batch][n+1][2] with F_RAW
  uint64_t* trace;                              // NULL or [2][m][2]: accumulator after the last step run
  const uint32_t* keyhat;                       // [rows][L][4][2][m] Montgomery form, NTT order
  const uint2* tw_f; const uint2* tw_i;         // [MAXP][m]
  uint8_t* scratch; size_t scratch_stride;      // per CTA: accumulator + digits (kept L2-persistent)
  uint8_t* zres; size_t zres_stride;            // per CTA: inverse-transform residues
  int batch, step_begin, step_end, draw_steps, flags;
  unsigned long long* timing;                   // NULL, or 8 phase-cycle accumulators written by CTA 0 (profiling aid)
  const uint64_t* pack_in;                      // F_PACK: [batch][m][2] wide polynomials, job g multiplies poly g with key row g
  const int64_t* pack_draws;                    // F_PACK: NULL or [batch][m][2]
  uint64_t rng_seed, rng_gate0;                 // draws == NULL and rng_seed != 0: flatten(rng, ...) with draws made on the device (DrawSrc)
  int* work_counter;                            // NULL (gate g = blockIdx.x + k gridDim.x) or a zeroed device counter: CTAs take the next gate when they finish one
  int stagger_cycles, stagger_slots;            // CTA b starts (b % slots) * cycles late: spreads the L2-bound phases of the CTAs in time
};

// acc: the accumulator (a, b) in OFFSET FORM x + offs mod Q (device_math.cuh), limb-major; dig: the biased digit words
// (dp mod 2^32, dp >> 18) of its gadget decomposition; zres: CRT-ready residues of the two product polynomials.
struct Scratch { uint32_t* acc; uint2* dig2; double* digd; uint32_t* zres; uint32_t* park; uint4* sums; };
__device__ __forceinline__ Scratch carve(uint8_t* base, uint8_t* zbase, int m) {
  Scratch s;
  s.acc = reinterpret_cast<uint32_t*>(base);                         // [2][3][m]
  s.dig2 = reinterpret_cast<uint2*>(base + (size_t)24 * m);          // [4][m] biased digit words (dp mod 2^32, dp >> 18) side by side: one 64-bit access per digit
  s.digd = reinterpret_cast<double*>(base + (size_t)24 * m);          // [4][m] negated digits as doubles (FP64 head; same bytes as dig2)
  s.zres = reinterpret_cast<uint32_t*>(zbase);                        // [L][2][m]
  s.park = reinterpret_cast<uint32_t*>(zbase + (size_t)64 * m);       // v5: [4][m/2] forward halves + [2][m/2] inverse halves (zres uses at most 8 * 8 m bytes)
  s.sums = reinterpret_cast<uint4*>(zbase + (size_t)76 * m);          // v5: [m] unreduced CRT sums
  return s;
}
static size_t scratch_bytes(int m) { return (size_t)56 * m; }

// Position of transform-domain point idx inside a pre-transformed key polynomial.  At m = 8192 every lane of the fused
// phase owns 16 consecutive points (64 bytes); the key is stored so that the four 16-byte loads of a warp are each
// 512 contiguous bytes: point e = 16 lane + 4 t + r of a 512-point slice sits at 128 t + 4 lane + r.
template <int LOGM>
__host__ __device__ __forceinline__ int key_pos(int idx) {
  if (LOGM != 13) return idx;
  const int e = idx & 511;
  return (idx & ~511) | (((e >> 2) & 3) << 7) | ((e >> 4) << 2) | (e & 3);
}
static size_t zres_bytes(int m, int L) { (void)L; return (size_t)(64 + 12 + 16) * m; }   // residues (up to 8 primes) + v5 parking + v5 sums

// accumulator init: a = 0, b = t(x) x^(-u_b) DQ   (src/fhe.jl:566-573, t(x) from src/fhe.jl:535-548)
template <int LOGM>
__device__ void gate_init(const DevConst& C, const Scratch& S, uint64_t ub) {
  constexpr int m = 1 << LOGM;
  u96 zero; zero.x0 = zero.x1 = zero.x2 = 0;
  zero = to_offset_form(C, zero);
  const u96 DQ = to_offset_form(C, from128(C.DQ)), nDQ = to_offset_form(C, from128(C.Q - C.DQ));
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    const int src = (int)((j + ub) & (uint64_t)(2 * m - 1));
    const int idx = src & (m - 1);
    int sgn = idx < m / 2 ? 1 : (idx == m / 2 ? 0 : -1);     // Dr = m/2: +1 on [0,m/2), 0, -1 on (m/2,m)
    if (src >= m) sgn = -sgn;
    st96(S.acc, m, j, zero);
    st96(S.acc + 3 * m, m, j, sgn == 0 ? zero : (sgn > 0 ? DQ : nDQ));
  }
}

// Where flatten(rng, ...) (src/utils.jl:198-241) takes its draws rand(rng, -xmax:xmax) from.  ptr: the caller's values for
// this step, [2][m][2] (polynomial a then b, coefficient, digit) -- the parity path, bit-exact with the reference for the
// caller's RNG.  ptr == NULL and seed != 0: made on the device by the counter-based generator Philox4x32-10 (Salmon et al.,
// SC'11), key = seed, counter = (coefficient, 2 step + polynomial, gate low word, gate high word ^ tag): no memory traffic, so the randomised mode is
// usable at paper size (host draws are 268 MB per gate there).  Valid randomised ciphertexts, but a different stream
// from any Julia RNG; the same (seed, gate, step) always gives the same draws.
struct DrawSrc {
  const int64_t* ptr; uint64_t seed, gate; uint32_t step;
  __device__ __forceinline__ bool on() const { return ptr != nullptr || seed != 0; }
};
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c.x), l0 = 0xD2511F53u * c.x;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c.z), l1 = 0xCD9E8D57u * c.z;
    c = make_uint4(h1 ^ c.y ^ k.x, l1, h0 ^ c.w ^ k.y, l0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}
// the two draws (digit 1, digit 2) of coefficient j of polynomial c: uniform on [-xmax, xmax] by multiply-shift of 64 random
// bits (bias below 2^-18 of one value's probability)
__device__ __forceinline__ void get_draws(const DevConst& C, const DrawSrc& d, int c, int j, int m, int64_t& x0, int64_t& x1) {
  if (d.ptr) { x0 = d.ptr[((size_t)c * m + j) * 2]; x1 = d.ptr[((size_t)c * m + j) * 2 + 1]; return; }
  const uint4 r = philox4x32_10(make_uint4((uint32_t)j, 2u * d.step + (uint32_t)c, (uint32_t)d.gate, 0x53474648u ^ (uint32_t)(d.gate >> 32)),
                                make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
  x0 = (int64_t)__umul64hi((uint64_t)r.x | ((uint64_t)r.y << 32), 2 * C.xmax + 1) - (int64_t)C.xmax;
  x1 = (int64_t)__umul64hi((uint64_t)r.z | ((uint64_t)r.w << 32), 2 * C.xmax + 1) - (int64_t)C.xmax;
}

// gadget decomposition of accumulator polynomial c (src/utils.jl:253-264) into S.dig[2c], S.dig[2c+1]
template <int LOGM, bool F64>
__device__ __forceinline__ void decompose_poly(const DevConst& C, const Scratch& S, int c, const DrawSrc draws) {
  constexpr int m = 1 << LOGM, KB = 3 * LOGM - 1;
  for (int idx = threadIdx.x; idx < m; idx += blockDim.x) {
    const u96 v = ld96(S.acc + c * 3 * m, m, idx);
    uint64_t dp0, dp1;
    if (draws.on()) { int64_t x0, x1; get_draws(C, draws, c, idx, m, x0, x1); decompose_off_rand<KB>(C, v, x0, x1, dp0, dp1); }
    else decompose_off<KB>(C, v, dp0, dp1);
    if (F64) { S.digd[(2 * c) * m + idx] = digit_f64(dp0); S.digd[(2 * c + 1) * m + idx] = digit_f64(dp1); }
    else {
      S.dig2[(2 * c) * m + idx] = digit_words(dp0);
      S.dig2[(2 * c + 1) * m + idx] = digit_words(dp1);
    }
  }
}

// Phase D for one accumulator polynomial, branch-free per coefficient so that the D coefficients in flight interleave:
//   EXT  (external-product seam): acc = z;   otherwise acc += x^u z - z  (mul_by_xj_minus_one, src/fhe.jl:554-556, applied
//   to the product), evaluated on the unreduced CRT sums as V = acc + (KQ - S[j]) + (S[j-u] or KQ - S[j-u]) with ONE Barrett
//   reduction; DEC: fused with the next step's gadget decomposition (RAND: with the caller's draws).
// GSUM: the staged CRT sums are in global memory (v5 kernel): read past L1 (other threads of the CTA wrote them)
template <bool GSUM>
__device__ __forceinline__ uint4 ld_sum(const uint4* p) { if (GSUM) return __ldcg(p); return *p; }

template <int LOGM, int T, int D, bool EXT, bool DEC, bool RAND, bool F64, bool GSUM = false>
__device__ __forceinline__ void update_poly(const DevConst& C, const Scratch& S, const uint4* sm4, int c,
                                            const DrawSrc draws_next, int u, u96 (&aq)[D]) {
  constexpr int m = 1 << LOGM, KB = 3 * LOGM - 1, SB = 6 * LOGM + 8;
  constexpr int NIT = m / T;
  const int tid = threadIdx.x;
  const uint4 KQ = make_uint4(C.KQ[0], C.KQ[1], C.KQ[2], C.KQ[3]);
  const uint4 OFF = make_uint4(C.offl[0], C.offl[1], C.offl[2], 0);
  uint32_t* acc = S.acc + c * 3 * m;
#pragma unroll 1
  for (int it0 = 0; it0 < NIT; it0 += D) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const int j = tid + (it0 + d) * T;
      uint4 V;
      // sm4 holds S' = -z (unreduced, 0 <= S' < KQ): the bootstrap basis' CRT constants are stored negated
      if (EXT) {
        V = add128(sub128(KQ, ld_sum<GSUM>(sm4 + j)), OFF);
      } else {
        const u96 a = aq[d];
        if (it0 + d + D < NIT) aq[d] = ld96(acc, m, j + D * T);
        const int src = (j - u) & (2 * m - 1);
        const uint4 nr = ld_sum<GSUM>(sm4 + (src & (m - 1)));         // -z[j-u]
        const uint4 pr = sub128(KQ, nr);                              // +z[j-u]
        const bool neg = src >= m;                                    // x^u z wraps with a sign flip
        const uint4 zs = make_uint4(neg ? nr.x : pr.x, neg ? nr.y : pr.y, neg ? nr.z : pr.z, neg ? nr.w : pr.w);
        V = add128(add128(make_uint4(a.x0, a.x1, a.x2, 0), ld_sum<GSUM>(sm4 + j)), zs);   // acc - z[j] +- z[j-u]  < Q + 2 KQ < 2^35 Q
      }
      const u96 res = barrett96<SB>(C, V);
      st96(acc, m, j, res);
      if (DEC) {
        uint64_t dp0, dp1;
        if (RAND) { int64_t x0, x1; get_draws(C, draws_next, c, j, m, x0, x1); decompose_off_rand<KB>(C, res, x0, x1, dp0, dp1); }
        else decompose_off<KB>(C, res, dp0, dp1);
        if (F64) { S.digd[(2 * c) * m + j] = digit_f64(dp0); S.digd[(2 * c + 1) * m + j] = digit_f64(dp1); }
        else {
          S.dig2[(2 * c) * m + j] = digit_words(dp0);
          S.dig2[(2 * c + 1) * m + j] = digit_words(dp1);
        }
      }
    }
  }
}

// Phase C/D per accumulator polynomial: the UNREDUCED CRT sums S (four limbs each, device_math.cuh crt_sum) go to
// shared memory, then update_poly.  Every loop reads data written a whole step ago (largely evicted to HBM), so D
// iterations of loads stay in flight and the first loads of each loop are issued one phase early, across the barriers.
// OWN: each thread reads only residues it stored itself (v4 kernel), so those loads may precede the entry barrier.
template <int LOGM, int T, bool OWN, bool F64, bool GSUM = false>
__device__ __forceinline__ void crt_update(const DevConst& C, const Scratch& S, uint4* sm4, uint32_t* stg,
                                           const DrawSrc draws_next, int u, bool ext, bool decompose_next,
                                           unsigned long long* timing, long long& tprev) {
  constexpr int m = 1 << LOGM, L = Shape<LOGM>::L;
  const int tid = threadIdx.x;
#define SGFHE_TICK(slot) do { if (timing && threadIdx.x == 0) { const long long tn_ = clock64(); timing[slot] += (unsigned long long)(tn_ - tprev); tprev = tn_; } } while (0)
  // sm4: [m] unreduced sums, 16 m bytes (the four transform buffers; global scratch in the v5 kernel); stg: staging area
  constexpr int NIT = m / T, D = (T <= 512 && NIT % 4 == 0) ? 4 : (NIT % 2 == 0 ? 2 : 1);   // coefficients in flight per thread (64-register kernels: two)
  uint32_t yq[D][L];
  u96 aq[D];
  if (OWN) {
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int i = 0; i < L; ++i) yq[d][i] = S.zres[(size_t)i * 2 * m + tid + d * T];
    __syncthreads();                                         // shared memory free for the CRT staging
  }
  // first residues of the SECOND polynomial: async copies into the idle twiddle-table region (behind the accumulator
  // staging), requested now, consumed when the second CRT loop starts
  constexpr bool RS = OWN && (size_t)D * (3 + L) * T * 4 <= (size_t)m * 8;
  uint32_t* stg_res = stg + D * 3 * T;                       // [D][L][T]
  if (RS) {
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int i = 0; i < L; ++i) cp_async4(stg_res + (d * L + i) * T + tid, S.zres + (size_t)m + (size_t)i * 2 * m + tid + d * T);
    cp_async_commit();
  }
  // the accumulator was last touched a whole step ago: pull it from HBM into L2 while the CRT sums run
  for (int line = tid; line < (2 * 3 * m * 4) / 128; line += T)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(S.acc) + (size_t)line * 128));
  for (int c = 0; c < 2; ++c) {
    const uint32_t* zr = S.zres + (size_t)c * m;
    if (RS && c == 1) {
      cp_async_wait_all();
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
        for (int i = 0; i < L; ++i) yq[d][i] = stg_res[(d * L + i) * T + tid];
    } else if (!OWN || c == 1) {
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
        for (int i = 0; i < L; ++i) yq[d][i] = zr[(size_t)i * 2 * m + tid + d * T];
    }
#pragma unroll 1
    for (int it0 = 0; it0 < NIT; it0 += D) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int idx = tid + (it0 + d) * T;
        sm4[idx] = crt_sum<0, L>(C, yq[d], 1);
        if (it0 + d + D < NIT) {
#pragma unroll
          for (int i = 0; i < L; ++i) yq[d][i] = zr[(size_t)i * 2 * m + idx + D * T];
        }
      }
    }
    constexpr int DS = OWN ? D : 0;                          // accumulator words of the first DS iterations: async copies
    // stg: [DS][3][T] in the (idle) twiddle-table region of the v4 kernel
    if (!ext && DS) {
#pragma unroll
      for (int d = 0; d < DS; ++d)
#pragma unroll
        for (int l = 0; l < 3; ++l) cp_async4(stg + (d * 3 + l) * T + tid, S.acc + (c * 3 + l) * m + tid + d * T);
      cp_async_commit();
    }
    __syncthreads();
    SGFHE_TICK(5);
    if (!ext) {
#pragma unroll
      for (int d = DS; d < D; ++d) aq[d] = ld96(S.acc + c * 3 * m, m, tid + d * T);
      if (DS) {
        cp_async_wait_all();
#pragma unroll
        for (int d = 0; d < DS; ++d) { aq[d].x0 = stg[(d * 3) * T + tid]; aq[d].x1 = stg[(d * 3 + 1) * T + tid]; aq[d].x2 = stg[(d * 3 + 2) * T + tid]; }
      }
    }
    if (ext) {
      if (!decompose_next) update_poly<LOGM, T, D, true, false, false, F64, GSUM>(C, S, sm4, c, draws_next, u, aq);
      else if (draws_next.on()) update_poly<LOGM, T, D, true, true, true, F64, GSUM>(C, S, sm4, c, draws_next, u, aq);
      else update_poly<LOGM, T, D, true, true, false, F64, GSUM>(C, S, sm4, c, draws_next, u, aq);
    } else {
      if (!decompose_next) update_poly<LOGM, T, D, false, false, false, F64, GSUM>(C, S, sm4, c, draws_next, u, aq);
      else if (draws_next.on()) update_poly<LOGM, T, D, false, true, true, F64, GSUM>(C, S, sm4, c, draws_next, u, aq);
      else update_poly<LOGM, T, D, false, true, false, F64, GSUM>(C, S, sm4, c, draws_next, u, aq);
    }
    __syncthreads();
    SGFHE_TICK(6);
  }
#undef SGFHE_TICK
}

// Tail of the v4 step with the two halves of the CTA on different work (T = 512 threads, 16 warps, four per scheduler).
// crt_update runs CRT(0), update(0), CRT(1), update(1) one after the other with all warps in the same phase: the CRT sums
// keep the FMA-heavy pipe busy (18 IMAD.WIDE per coefficient) while the ALU and FP64 pipes idle, the update / decomposition
// does the opposite.  Here warps 0-7 update polynomial 0 while warps 8-15 form the CRT sums of polynomial 1, so every
// scheduler holds two warps of each kind and interleaves them cycle by cycle:
//   A  all warps:  CRT sums of polynomial 0 -> shared memory (all 128 KiB of the transform buffers)
//   B  warps 0-7:  update + decomposition of polynomial 0 from shared memory
//      warps 8-15: CRT sums of polynomial 1 -> global scratch S.sums (shared memory is still occupied)
//   C  all warps:  update + decomposition of polynomial 1, sums read back from L2 (ld.cg)
// MEASURED (round 2, profiles/ab_r02_split_tail.txt): bit-exact, B takes 16.8k cycles against 21.0k for the two phases it replaces,
// but C takes 27.3k against 10.4k -- 256 KiB of sums per step come back through the 42 B/clk/SM L2 port, and shared memory
// cannot hold both polynomials' sums (2 x 128 KiB).  214.1k cycles/step against 198.4k: compiled only with -DSGFHE_TAIL_SPLIT.
#ifdef SGFHE_TAIL_SPLIT
template <int LOGM, int T, bool F64>
__device__ __forceinline__ void crt_update_split(const DevConst& C, const Scratch& S, uint4* sm4, uint32_t* stg,
                                                 const DrawSrc draws_next, int u, bool ext, bool decompose_next,
                                                 unsigned long long* timing, long long& tprev) {
  constexpr int m = 1 << LOGM, L = Shape<LOGM>::L, TH = T / 2, D = 4;
  static_assert((m / T) % D == 0, "tail pipelines are D deep");
  const int tid = threadIdx.x;
#define SGFHE_TICK(slot) do { if (timing && threadIdx.x == 0) { const long long tn_ = clock64(); timing[slot] += (unsigned long long)(tn_ - tprev); tprev = tn_; } } while (0)
#define SGFHE_UPDATE(TT, CC, GS, SRC) do { \
    if (ext) { \
      if (!decompose_next) update_poly<LOGM, TT, D, true, false, false, F64, GS>(C, S, SRC, CC, draws_next, u, aq); \
      else if (draws_next.on()) update_poly<LOGM, TT, D, true, true, true, F64, GS>(C, S, SRC, CC, draws_next, u, aq); \
      else update_poly<LOGM, TT, D, true, true, false, F64, GS>(C, S, SRC, CC, draws_next, u, aq); \
    } else { \
      if (!decompose_next) update_poly<LOGM, TT, D, false, false, false, F64, GS>(C, S, SRC, CC, draws_next, u, aq); \
      else if (draws_next.on()) update_poly<LOGM, TT, D, false, true, true, F64, GS>(C, S, SRC, CC, draws_next, u, aq); \
      else update_poly<LOGM, TT, D, false, true, false, F64, GS>(C, S, SRC, CC, draws_next, u, aq); \
    } } while (0)
  // ---- A: CRT sums of polynomial 0 (each thread reads only residues it stored itself: loads precede the barrier) ----
  {
    uint32_t yq[D][L];
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int i = 0; i < L; ++i) yq[d][i] = S.zres[(size_t)i * 2 * m + tid + d * T];
    __syncthreads();                                         // shared memory free for the CRT staging
    // the accumulator was last touched a whole step ago: pull it from HBM into L2 while the CRT sums run
    for (int line = tid; line < (2 * 3 * m * 4) / 128; line += T)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(S.acc) + (size_t)line * 128));
#pragma unroll 1
    for (int it0 = 0; it0 < m / T; it0 += D) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int idx = tid + (it0 + d) * T;
        sm4[idx] = crt_sum<0, L>(C, yq[d], 1);
        if (it0 + d + D < m / T) {
#pragma unroll
          for (int i = 0; i < L; ++i) yq[d][i] = S.zres[(size_t)i * 2 * m + idx + D * T];
        }
      }
    }
  }
  __syncthreads();
  SGFHE_TICK(5);
  // ---- B: update(0) on warps 0-7, CRT(1) on warps 8-15 ----
  if (tid < TH) {
    u96 aq[D];
    if (!ext) {
#pragma unroll
      for (int d = 0; d < D; ++d) aq[d] = ld96(S.acc, m, tid + d * TH);
    }
    SGFHE_UPDATE(TH, 0, false, sm4);
  } else {
    const int t2 = tid - TH;
    const uint32_t* zr = S.zres + (size_t)m;
    uint32_t yq[D][L];
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int i = 0; i < L; ++i) yq[d][i] = __ldcg(&zr[(size_t)i * 2 * m + t2 + d * TH]);
#pragma unroll 1
    for (int it0 = 0; it0 < m / TH; it0 += D) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int idx = t2 + (it0 + d) * TH;
        __stcg(&S.sums[idx], crt_sum<0, L>(C, yq[d], 1));
        if (it0 + d + D < m / TH) {
#pragma unroll
          for (int i = 0; i < L; ++i) yq[d][i] = __ldcg(&zr[(size_t)i * 2 * m + idx + D * TH]);
        }
      }
    }
  }
  // accumulator words of the first D iterations of phase C: async copies into the idle twiddle-table region
  if (!ext) {
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int l = 0; l < 3; ++l) cp_async4(stg + (d * 3 + l) * T + tid, S.acc + (3 + l) * m + tid + d * T);
    cp_async_commit();
  }
  __syncthreads();
  SGFHE_TICK(6);
  // ---- C: update(1), sums from L2 ----
  {
    u96 aq[D];
    if (!ext) {
      cp_async_wait_all();
#pragma unroll
      for (int d = 0; d < D; ++d) { aq[d].x0 = stg[(d * 3) * T + tid]; aq[d].x1 = stg[(d * 3 + 1) * T + tid]; aq[d].x2 = stg[(d * 3 + 2) * T + tid]; }
    }
    SGFHE_UPDATE(T, 1, true, S.sums);
  }
  __syncthreads();
  SGFHE_TICK(7);
#undef SGFHE_UPDATE
#undef SGFHE_TICK
}
#endif

// Last REM inverse stages of the two result polynomials (shared memory, swizzled) and store of the canonical residues, V
// neighbouring columns per thread so that the shared-memory reads and the global stores are 16 / 8 bytes wide when the CTA has
// fewer threads than columns (swz keeps an aligned group of four adjacent).  The CRT pre-scaling rides in the key words.
template <int LOGM>
__device__ __forceinline__ void store_residues(uint32_t* __restrict__ zres, const uint32_t* sm, const uint2* wt, uint32_t p, uint32_t z) {
  using SH = Shape<LOGM>;
  constexpr int m = SH::M, REM = SH::REM, R = 1 << REM, STR = SH::STR, T = SH::T;
  constexpr int V = 2 * STR / T >= 4 ? 4 : (2 * STR / T >= 2 ? 2 : 1);
  const uint32_t p2 = 2 * p;
#pragma unroll 2
  for (int e = (int)threadIdx.x * V; e < 2 * STR; e += T * V) {
    const int c = e / STR, idx = e % STR;
    uint32_t x[V][R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const uint32_t* src = sm + c * m + swz(idx + k * STR);
      if constexpr (V == 4) { const uint4 v = *reinterpret_cast<const uint4*>(src); x[0][k] = v.x; x[1][k] = v.y; x[2][k] = v.z; x[3][k] = v.w; }
      else if constexpr (V == 2) { const uint2 v = *reinterpret_cast<const uint2*>(src); x[0][k] = v.x; x[1][k] = v.y; }
      else x[0][k] = *src;
    }
#pragma unroll
    for (int v = 0; v < V; ++v) inv_block<REM>(x[v], wt, p, p2, z);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      uint32_t* dst = zres + (size_t)c * m + idx + k * STR;
      if constexpr (V == 4) *reinterpret_cast<uint4*>(dst) = make_uint4(csub(x[0][k], p), csub(x[1][k], p), csub(x[2][k], p), csub(x[3][k], p));
      else if constexpr (V == 2) *reinterpret_cast<uint2*>(dst) = make_uint2(csub(x[0][k], p), csub(x[1][k], p));
      else *dst = csub(x[0][k], p);
    }
  }
}

// One accumulation step (body of src/fhe.jl:579-582) on digits already in S.dig; leaves the new accumulator in
// S.acc and its decomposition (with `draws_next`, the following step's draws) in S.dig.  u = rotation in [0, 2m).
template <int LOGM>
__device__ void gate_step(const DevConst& C, const Scratch& S, uint32_t* sm, const uint32_t* __restrict__ keyrow,
                          const uint2* __restrict__ tw_f, const uint2* __restrict__ tw_i,
                          const DrawSrc draws_next, int u, bool ext, bool decompose_next,
                          uint2* tab, uint64_t* bar, uint32_t& parity, unsigned long long* timing) {
  using SH = Shape<LOGM>;
  long long tprev = timing ? clock64() : 0;
#define SGFHE_TICK(slot) do { if (timing && threadIdx.x == 0) { const long long tn_ = clock64(); timing[slot] += (unsigned long long)(tn_ - tprev); tprev = tn_; } } while (0)
  constexpr int m = SH::M, REM = SH::REM, R = 1 << REM, STR = SH::STR, T = SH::T, L = SH::L;
  const int tid = threadIdx.x;
  // Phase B: per RNS prime -- 4 forward NTTs, 8 MACs against the key tile, 2 inverse NTTs
#pragma unroll 1
  for (int i = 0; i < L; ++i) {
    const uint32_t p = C.p[i], p2 = 2 * p;
    if (tid == 0) stage_table(tab, tw_f + (size_t)i * m, m * 8, bar);      // TMA: forward twiddles of this prime
    {
      const uint32_t mu = C.dig_mu[i], negc = C.dig_negc[i];
      uint2 wt[R > 1 ? R - 1 : 1];
      top_twiddles<REM>(tw_f + (size_t)i * m, wt);
      // register double buffering: the loads of batch b+1 are in flight while batch b is reduced and stored
      constexpr int ITER = 4 * STR / T, U = ITER >= 4 ? 4 : ITER, NBATCH = ITER / U;
      uint32_t lo[2][U][R], hi[2][U][R];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int e = tid + q * T, j = e / STR, idx = e % STR;
#pragma unroll
        for (int k = 0; k < R; ++k) { const uint2 v = S.dig2[j * m + idx + k * STR]; lo[0][q][k] = v.x; hi[0][q][k] = v.y; }
      }
#pragma unroll
      for (int b = 0; b < NBATCH; ++b) {
        if (b + 1 < NBATCH) {
#pragma unroll
          for (int q = 0; q < U; ++q) {
            const int e = tid + ((b + 1) * U + q) * T, j = e / STR, idx = e % STR;
#pragma unroll
            for (int k = 0; k < R; ++k) { const uint2 v = S.dig2[j * m + idx + k * STR]; lo[(b + 1) & 1][q][k] = v.x; hi[(b + 1) & 1][q][k] = v.y; }
          }
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int e = tid + (b * U + q) * T, j = e / STR, idx = e % STR;
          uint32_t x[R];
#pragma unroll
          for (int k = 0; k < R; ++k) x[k] = digit_mod(lo[b & 1][q][k], hi[b & 1][q][k], mu, negc, p);
          fwd_block<REM>(x, wt, p, p2, C.zero);
#pragma unroll
          for (int k = 0; k < R; ++k) sm[j * m + swz(idx + k * STR)] = x[k];
        }
      }
    }
    __syncthreads();
    mbar_wait(bar, parity); parity ^= 1;
    SGFHE_TICK(0);
    ntt_passes<LOGM, 4, true>(sm, tab, p, C.zero);
    if (tid == 0) stage_table(tab, tw_i + (size_t)i * m, m * 8, bar);      // TMA: inverse twiddles, under the pointwise phase
    SGFHE_TICK(1);
    const uint32_t* K = keyrow + (size_t)i * 8 * m;      // [4][2][m] for this prime   (src/fhe.jl:527-528)
    const uint32_t pinv = C.pinv_neg[i];
    {
      // two neighbouring points per thread and iteration: 64-bit key loads and shared-memory accesses (swz and key_pos keep an
      // aligned pair adjacent); the key words of the next pair are in flight while one pair is multiplied
      constexpr int NPI = m / (2 * T);
      uint2 kq[2][8];
#pragma unroll
      for (int q = 0; q < 8; ++q) kq[0][q] = __ldg(reinterpret_cast<const uint2*>(&K[q * m + key_pos<LOGM>(2 * tid)]));
#pragma unroll
      for (int it = 0; it < NPI; ++it) {
        const int idx = 2 * (tid + it * T);
        if (it + 1 < NPI) {
#pragma unroll
          for (int q = 0; q < 8; ++q) kq[(it + 1) & 1][q] = __ldg(reinterpret_cast<const uint2*>(&K[q * m + key_pos<LOGM>(idx + 2 * T)]));
        }
        uint64_t sa0 = 0, sa1 = 0, sb0 = 0, sb1 = 0;
        const int si = swz(idx);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint2 d = *reinterpret_cast<const uint2*>(sm + j * m + si);
          d.x = min(d.x, d.x - p2); d.x = min(d.x, d.x - p); d.y = min(d.y, d.y - p2); d.y = min(d.y, d.y - p);
          const uint2 ka = kq[it & 1][2 * j], kb = kq[it & 1][2 * j + 1];
          sa0 += (uint64_t)d.x * ka.x; sa1 += (uint64_t)d.y * ka.y;
          sb0 += (uint64_t)d.x * kb.x; sb1 += (uint64_t)d.y * kb.y;
        }
        *reinterpret_cast<uint2*>(sm + si) = make_uint2(redc(sa0, p, pinv), redc(sa1, p, pinv));
        *reinterpret_cast<uint2*>(sm + m + si) = make_uint2(redc(sb0, p, pinv), redc(sb1, p, pinv));
      }
    }
    __syncthreads();
    mbar_wait(bar, parity); parity ^= 1;
    SGFHE_TICK(2);
    ntt_passes<LOGM, 2, false>(sm, tab, p, C.zero);
    SGFHE_TICK(3);
    {
      uint2 wt[R > 1 ? R - 1 : 1];
      top_twiddles<REM>(tw_i + (size_t)i * m, wt);
      store_residues<LOGM>(S.zres + (size_t)i * 2 * m, sm, wt, p, C.zero);
    }
    __syncthreads();
    SGFHE_TICK(4);
  }
  crt_update<LOGM, T, false, false>(C, S, reinterpret_cast<uint4*>(sm), sm + 4 * m, draws_next, u, ext, decompose_next, timing, tprev);
#undef SGFHE_TICK
}

// One accumulation step at m = 16384 (Params(2048): 93-bit Q, six primes).  Four transform buffers would be 256 KiB, so a
// prime is processed in two halves of two digit polynomials each (128 KiB of shared memory): half 0 transforms digits 0, 1
// and parks its share of the two key MAC sums (Montgomery-reduced words, thread-private) in global scratch, half 1 adds the
// share of digits 2, 3 and both result polynomials are transformed back.  Twiddles come from global memory (a table is
// 128 KiB per prime and direction); the unreduced CRT sums of the tail are staged in global scratch as well.
// Correctness-first: every stage is the generic building block of gate_step.
template <int LOGM>
__device__ void gate_step_wide(const DevConst& C, const Scratch& S, uint32_t* sm, const uint32_t* __restrict__ keyrow,
                               const uint2* __restrict__ tw_f, const uint2* __restrict__ tw_i,
                               const DrawSrc draws_next, int u, bool ext, bool decompose_next, unsigned long long* timing) {
  using SH = Shape<LOGM>;
  constexpr int m = SH::M, REM = SH::REM, R = 1 << REM, STR = SH::STR, T = SH::T, L = SH::L;
  const int tid = threadIdx.x;
  long long tprev = timing ? clock64() : 0;
#define SGFHE_TICK(slot) do { if (timing && threadIdx.x == 0) { const long long tn_ = clock64(); timing[slot] += (unsigned long long)(tn_ - tprev); tprev = tn_; } } while (0)
#pragma unroll 1
  for (int i = 0; i < L; ++i) {
    const uint32_t p = C.p[i], p2 = 2 * p, pinv = C.pinv_neg[i];
    const uint2* twf = tw_f + (size_t)i * m;
    const uint2* twi = tw_i + (size_t)i * m;
    const uint32_t* K = keyrow + (size_t)i * 8 * m;      // [4][2][m] for this prime   (src/fhe.jl:527-528)
    const uint32_t mu = C.dig_mu[i], negc = C.dig_negc[i];
    uint2 wt[R > 1 ? R - 1 : 1];
    top_twiddles<REM>(twf, wt);
    // entries [0, m/8) of both tables (all that the strided passes read: 2 x 16 KiB) go to shared memory behind the two transform
    // buffers; only the stride-1 passes read their twiddles from global memory.  Visible after the barrier that ends the digit load.
    uint2* twl_f = reinterpret_cast<uint2*>(sm + 2 * m + 64);
    uint2* twl_i = twl_f + m / 8;
    for (int e = tid; e < m / 16; e += T) {
      reinterpret_cast<uint4*>(twl_f)[e] = __ldg(reinterpret_cast<const uint4*>(twf) + e);
      reinterpret_cast<uint4*>(twl_i)[e] = __ldg(reinterpret_cast<const uint4*>(twi) + e);
    }
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
#pragma unroll 4
      for (int e = tid; e < 2 * STR; e += T) {
        const int jj = e / STR, idx = e % STR, j = 2 * h + jj;
        uint32_t x[R];
#pragma unroll
        for (int k = 0; k < R; ++k) { const uint2 v = S.dig2[j * m + idx + k * STR]; x[k] = digit_mod(v.x, v.y, mu, negc, p); }
        fwd_block<REM>(x, wt, p, p2, C.zero);
#pragma unroll
        for (int k = 0; k < R; ++k) sm[jj * m + swz(idx + k * STR)] = x[k];
      }
      __syncthreads();
      SGFHE_TICK(0);
      ntt_passes<LOGM, 2, true>(sm, twf, p, C.zero, twl_f);
      SGFHE_TICK(1);
      const uint32_t* Kh = K + (size_t)(4 * h) * m;      // key rows 2 j + c of digit polynomials j = 2h, 2h + 1
#pragma unroll 2
      for (int i4 = tid; i4 < m / 4; i4 += T) {           // four consecutive points per thread: 128-bit accesses throughout
        const int idx = 4 * i4, si = swz(idx), kp = key_pos<LOGM>(idx);      // swz and key_pos<14> keep aligned quads contiguous
        const uint4 k00 = __ldg(reinterpret_cast<const uint4*>(Kh + kp)), k01 = __ldg(reinterpret_cast<const uint4*>(Kh + m + kp));
        const uint4 k10 = __ldg(reinterpret_cast<const uint4*>(Kh + 2 * m + kp)), k11 = __ldg(reinterpret_cast<const uint4*>(Kh + 3 * m + kp));
        const uint4 v0 = *reinterpret_cast<const uint4*>(sm + si), v1 = *reinterpret_cast<const uint4*>(sm + m + si);
        uint4 pa = make_uint4(0, 0, 0, 0), pb = pa;
        if (h == 1) { pa = *reinterpret_cast<const uint4*>(S.park + idx); pb = *reinterpret_cast<const uint4*>(S.park + m + idx); }
        const uint32_t d0[4] = {v0.x, v0.y, v0.z, v0.w}, d1[4] = {v1.x, v1.y, v1.z, v1.w};
        const uint32_t a0[4] = {k00.x, k00.y, k00.z, k00.w}, b0[4] = {k01.x, k01.y, k01.z, k01.w};
        const uint32_t a1[4] = {k10.x, k10.y, k10.z, k10.w}, b1[4] = {k11.x, k11.y, k11.z, k11.w};
        const uint32_t qa[4] = {pa.x, pa.y, pa.z, pa.w}, qb[4] = {pb.x, pb.y, pb.z, pb.w};
        uint32_t ra[4], rb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint32_t x0 = d0[e], x1 = d1[e];
          x0 = min(x0, x0 - p2); x0 = min(x0, x0 - p); x1 = min(x1, x1 - p2); x1 = min(x1, x1 - p);
          ra[e] = redc((uint64_t)x0 * a0[e] + (uint64_t)x1 * a1[e], p, pinv);          // [0, 2p)
          rb[e] = redc((uint64_t)x0 * b0[e] + (uint64_t)x1 * b1[e], p, pinv);
          if (h == 1) { ra[e] += qa[e]; rb[e] += qb[e]; ra[e] = min(ra[e], ra[e] - p2); rb[e] = min(rb[e], rb[e] - p2); }   // [0, 2p): input range of the inverse butterflies
        }
        if (h == 0) {                                     // read back by this same thread
          *reinterpret_cast<uint4*>(S.park + idx) = make_uint4(ra[0], ra[1], ra[2], ra[3]);
          *reinterpret_cast<uint4*>(S.park + m + idx) = make_uint4(rb[0], rb[1], rb[2], rb[3]);
        } else {
          *reinterpret_cast<uint4*>(sm + si) = make_uint4(ra[0], ra[1], ra[2], ra[3]);
          *reinterpret_cast<uint4*>(sm + m + si) = make_uint4(rb[0], rb[1], rb[2], rb[3]);
        }
      }
      __syncthreads();
      SGFHE_TICK(2);
    }
    ntt_passes<LOGM, 2, false>(sm, twi, p, C.zero, twl_i);
    SGFHE_TICK(3);
    {
      uint2 wti[R > 1 ? R - 1 : 1];
      top_twiddles<REM>(twi, wti);
      store_residues<LOGM>(S.zres + (size_t)i * 2 * m, sm, wti, p, C.zero);
    }
    __syncthreads();
    SGFHE_TICK(4);
  }
#undef SGFHE_TICK
  crt_update<LOGM, T, false, false, true>(C, S, S.sums, nullptr, draws_next, u, ext, decompose_next, timing, tprev);
}

// extract + AND/OR/XOR assembly (src/fhe.jl:585-592) + reduce_modulus (src/fhe.jl:616-618)
__device__ void gate_final(const DevConst& C, const Scratch& S, uint64_t* out_and, uint64_t* out_or,
                           uint64_t* out_xor, bool raw) {
  const int m = C.m, n = C.n;
  for (int k = threadIdx.x; k <= n; k += blockDim.x) {
    u128 va, vo;
    if (k < n) {
      va = to128(from_offset_form(C, ld96(S.acc, m, 3 * m / 4 - k)));                           // extract(a, 3m/4+1, n)[k]
      vo = negmodQ(to128(from_offset_form(C, ld96(S.acc, m, m / 4 - k))), C.Q);                 // -extract(a, m/4+1, n)[k]
    } else {
      va = addmodQ(C.DQ, to128(from_offset_form(C, ld96(S.acc + 3 * m, m, 3 * m / 4))), C.Q);   // DQ + b[3m/4]
      vo = submodQ(C.DQ, to128(from_offset_form(C, ld96(S.acc + 3 * m, m, m / 4))), C.Q);       // DQ - b[m/4]
    }
    const u128 vx = submodQ(vo, va, C.Q);                                  // a_or - a_and
    if (raw) {
      out_and[2 * k] = (uint64_t)va; out_and[2 * k + 1] = (uint64_t)(va >> 64);
      out_or[2 * k] = (uint64_t)vo; out_or[2 * k + 1] = (uint64_t)(vo >> 64);
      out_xor[2 * k] = (uint64_t)vx; out_xor[2 * k + 1] = (uint64_t)(vx >> 64);
    } else {
      out_and[k] = modred(C, va); out_or[k] = modred(C, vo); out_xor[k] = modred(C, vx);
    }
  }
}

// =========================================================================================================
// v4: fused-pass step for m >= 4096 (512 threads, up to 128 registers per thread).
//   * the digit load runs the top 3+REM stages in registers (radix 8/16) before the first store to shared memory;
//   * the stride-1 forward pass, the 8 key MACs per point and the stride-1 inverse pass are one phase, so the
//     key tile streams from L2 underneath butterflies instead of in a phase of its own;
//   * the last inverse pass runs the top stages in registers and stores CRT-ready residues straight to HBM/L2;
//   * only the FORWARD twiddle table is staged (TMA, one prime ahead): psi^-bitrev(2^l+g) = -psi^bitrev(2^l+(g^(2^l-1))),
//     so the inverse passes read the mirrored forward entries and fold the sign into the butterfly; the twiddles of the
//     top stages (the same for every thread) are kernel parameters, i.e. constant-bank operands;
//   * at m = 8192 every warp owns one 512-point slice between the top stages (warp-level synchronisation only), a lane
//     works on two adjacent radix-8 blocks (64-bit shared-memory accesses);
//   * the CRT sums stay unreduced until the accumulator update (one Barrett step per coefficient, FP64-assisted), the first
//     loads of the tail arrive by cp.async in the then idle twiddle-table region;
//   * HF (round 2, the default): the digits are stored as doubles and the digit reduction plus the first forward stage run
//     on the FP64 pipe (head_stage1_f64), which the integer kernel left idle.
// =========================================================================================================
template <int LOGM, int TB_ = 1>
struct Shape4 {
  static constexpr int M = 1 << LOGM;
  static constexpr int REM = LOGM % 3;            // 0 or 1 supported
  static constexpr int LR0 = 3 + REM, R0 = 1 << LR0;
  static constexpr int TB = TB_;                  // top-stage blocks per thread and polynomial
  static constexpr int STR0 = M >> LR0;           // stride of the top-stage blocks (512)
  static constexpr int T = STR0 / TB;             // 512 threads (TB = 1), or 256 with two top-stage blocks per thread
  static constexpr int NB = (M / 8) / T;          // radix-8 blocks per thread and polynomial (1 or 2)
  static constexpr int L = Shape<LOGM>::L;
};
// GateTB<LOGM>::V = top-stage blocks per thread of the gate kernel.  V = 2 at m = 4096 gives 256-thread CTAs and TWO gates per SM
// (96 KiB of shared memory and 32 k registers each) with the per-warp 512-point slices of m = 8192 (64-bit shared-memory
// accesses, __syncwarp between the passes).  MEASURED (round 2, profiles/ab_r02_p512_two_ctas.txt): bit-exact; the strided
// passes get 20 % cheaper per gate, the key MACs and the update 15 % dearer (two gates on different key rows share the L2
// port), 107.0 k against 109.0 k cycles per gate-step, and with four instead of eight gates per CTA at 1 184 gates the batch
// ends on a longer tail: 4 600 - 4 780 against 5 117 gates/s.  V = 1 everywhere.
template <int LOGM> struct GateTB { static constexpr int V = 1; };
template <int LOGM> struct GateShape4 : Shape4<LOGM, GateTB<LOGM>::V> {};


// Between the top stages and the last inverse pass every element with index bits [9, LOGM) fixed is touched only by
// the 64 consecutive threads that own that 512-element slice, so the stride-64 <-> stride-8 hand-over needs a
// named barrier over those two warps only (ids 1..8), and warp pairs drift apart: some stream key words from L2 in
// the fused phase while others run butterflies.
__device__ __forceinline__ void group_bar64() {
  asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)(threadIdx.x >> 6)) : "memory");
}
// With two radix-8 blocks per thread and polynomial (m = 8192) each WARP owns one 512-element slice outright (blocks
// 2 lane and 2 lane + 1 of slice warp_id), so the three shared-memory passes and the fused phase between the top stages
// exchange data inside a warp only: __syncwarp replaces the block-level barriers and the 16 warps drift freely.
// The two blocks of a thread are ADJACENT: in the strided passes they share their twiddles and their elements are
// neighbours in memory, so every access is 64 bits wide (half the LDS/STS instructions; swz keeps bit 0).
template <int LOGM, int TB = 1>
__device__ __forceinline__ int block_of(int tid, int q) {
  using S4 = Shape4<LOGM, TB>;
  return S4::NB == 2 ? (((tid >> 5) << 6) | ((tid & 31) << 1) | q) : tid + q * S4::T;
}
template <int LOGM, int TB = 1>
__device__ __forceinline__ void slice_sync() {
  if (Shape4<LOGM, TB>::NB == 2) __syncwarp(); else group_bar64();
}

// twiddles of one radix-8 block from the staged forward table; the inverse passes take the MIRRORED forward entries
// (psi^-bitrev(2^l+g) = -psi^bitrev(2^l+(g^(2^l-1)))) and fold the sign into the butterfly (gs_bfly_negw)
template <bool FWD>
__device__ __forceinline__ void block_twiddles(const uint2* tab, int lvl, int g, uint32_t p, uint2 (&w)[7]) {
  const int t1 = lvl + (FWD ? g : (g ^ (lvl - 1)));
  const uint2 w0 = tab[t1];
  const uint4 a = *reinterpret_cast<const uint4*>(&tab[2 * t1]);
  const uint4 b0 = *reinterpret_cast<const uint4*>(&tab[4 * t1]);
  const uint4 b1 = *reinterpret_cast<const uint4*>(&tab[4 * t1 + 2]);
  w[0] = w0;
  if (FWD) {
    w[1] = make_uint2(a.x, a.y); w[2] = make_uint2(a.z, a.w);
    w[3] = make_uint2(b0.x, b0.y); w[4] = make_uint2(b0.z, b0.w); w[5] = make_uint2(b1.x, b1.y); w[6] = make_uint2(b1.z, b1.w);
  } else {
    w[1] = make_uint2(a.z, a.w); w[2] = make_uint2(a.x, a.y);
    w[3] = make_uint2(b1.z, b1.w); w[4] = make_uint2(b1.x, b1.y); w[5] = make_uint2(b0.z, b0.w); w[6] = make_uint2(b0.x, b0.y);
  }
}

template <int LOGM, int NPOLY, bool FWD, int B, int TB = 1>
__device__ __forceinline__ void pass8_v4(uint32_t* sm, const uint2* tab, uint32_t p, uint32_t z) {
  using S4 = Shape4<LOGM, TB>;
  constexpr int M = S4::M;
  const uint32_t p2 = 2 * p;
  if constexpr (S4::NB == 2 && B >= 1) {
    // blocks blk and blk + 1 (blk even): same twiddles, elements (e, e + 1) adjacent and 8-byte aligned
    const int blk = block_of<LOGM, TB>(threadIdx.x, 0);
    const int base = ((blk >> B) << (B + 3)) | (blk & ((1 << B) - 1));
    uint2 w[7];
    block_twiddles<FWD>(tab, M >> (B + 3), blk >> B, p, w);
    // swz(base + (j << B)) with the structure of the swizzle spelled out: only bits 2..4 depend on j through an XOR, the
    // rest is an immediate offset of the load / store
    int off[8];
    if constexpr (B == 6) {                               // bits 6,7 = j & 3 go to bits 3,4; bit 5 (of base) to bit 2
      const int b0 = swz(base);
#pragma unroll
      for (int j = 0; j < 8; ++j) off[j] = (b0 ^ ((j & 3) << 3)) + (j << 6);
    } else if constexpr (B == 3) {                        // bits 3,4 = j & 3 meet bits 6,7 of base; bit 5 = j >> 2 goes to bit 2
      const int c = (base >> 6) & 3;
#pragma unroll
      for (int j = 0; j < 8; ++j) off[j] = ((base + (((j & 3) ^ c) << 3)) ^ ((j >> 2) << 2)) + ((j >> 2) << 5);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) off[j] = swz(base + (j << B));
    }
#pragma unroll
    for (int poly = 0; poly < NPOLY; ++poly) {
      uint32_t* s = sm + poly * M;
      uint32_t x[8], y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { const uint2 v = *reinterpret_cast<const uint2*>(s + off[j]); x[j] = v.x; y[j] = v.y; }
      if (FWD) { fwd_block<3>(x, w, p, p2, z); fwd_block<3>(y, w, p, p2, z); }
      else { inv_block<3, true>(x, w, p, p2, z); inv_block<3, true>(y, w, p, p2, z); }
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<uint2*>(s + off[j]) = make_uint2(x[j], y[j]);
    }
  } else {
#pragma unroll
  for (int q = 0; q < S4::NB; ++q) {
    const int blk = block_of<LOGM, TB>(threadIdx.x, q);
    const int base = ((blk >> B) << (B + 3)) | (blk & ((1 << B) - 1));
    uint2 w[7];
    block_twiddles<FWD>(tab, M >> (B + 3), blk >> B, p, w);
    int off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) off[j] = swz(base + (j << B));
#pragma unroll
    for (int poly = 0; poly < NPOLY; ++poly) {
      uint32_t* s = sm + poly * M;
      uint32_t x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = s[off[j]];
      if (FWD) fwd_block<3>(x, w, p, p2, z); else inv_block<3, true>(x, w, p, p2, z);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[off[j]] = x[j];
    }
  }
  }
}

// Key tiles by TMA (-DSGFHE_KEY_TMA): the 2 x 1 KiB of key words a warp needs for one (radix-8 block, digit polynomial) --
// rows 2j and 2j+1 of the pre-transformed key, 8 transform points per lane, contiguous for the whole warp thanks to key_pos --
// arrive by two cp.async.bulk copies into a per-warp 2 KiB staging buffer, completing on a per-warp mbarrier; lane 0
// requests the next pair as soon as the warp has moved the current one into registers.
// MEASURED (round 2, profiles/ab_r02_key_tma.txt): bit-exact (78 GPU tests), but 227.8 k against 198.4 k cycles per step:
// the fused phase goes from 47.8 k to 72.3 k.  Shared memory has room for ONE pair per warp (32 KiB next to 192 KiB of
// transform buffers and twiddles), i.e. a lookahead of one (block, polynomial) of work, about 600 cycles, and a bulk copy
// takes longer than that to land; the LDG.128 loads with a register double buffer (the default) have the same lookahead
// and hide their latency.  Off by default.
#ifdef SGFHE_KEY_TMA
constexpr bool kKeyTma = true;
#else
constexpr bool kKeyTma = false;
#endif
constexpr int kKeyStageBytes = 16 * 2048 + 16 * 8;        // 16 warps x (2 rows x 1 KiB) + 16 mbarriers
template <int LOGM>
__device__ __forceinline__ uint8_t* key_stage_base(uint32_t* sm) { return reinterpret_cast<uint8_t*>(sm) + (size_t)24 * (1 << LOGM) + 128; }

template <int LOGM, bool HF>
__device__ void gate_step_v4(const DevConst& C, const Scratch& S, uint32_t* sm, const uint32_t* __restrict__ keyrow,
                             const uint2* __restrict__ tw_f, const DrawSrc draws_next, int u, bool ext,
                             bool decompose_next, uint2* tab, uint64_t* bar, uint32_t& parity,
                             unsigned long long* timing) {
  constexpr int TB = GateTB<LOGM>::V;
  using S4 = Shape4<LOGM, TB>;
  constexpr int m = S4::M, R0 = S4::R0, LR0 = S4::LR0, T = S4::T, NB = S4::NB, L = S4::L, STR0 = S4::STR0, NI = 4 * TB;
  const int tid = threadIdx.x;
  long long tprev = timing ? clock64() : 0;
#define SGFHE_TICK(slot) do { if (timing && threadIdx.x == 0) { const long long tn_ = clock64(); timing[slot] += (unsigned long long)(tn_ - tprev); tprev = tn_; } } while (0)
  // Digit words are prime independent.  Polynomials are processed in the order 2,3,0,1: buffers 2,3 are not read
  // by the previous prime's residue store, so that store overlaps the head of the next digit load.
  const int st = swz(tid);                               // swz(tid + k 256) = swz(tid) + k 256
  // head / top-inverse work items: (polynomial, top-stage block tb of this thread): elements tid + tb T + k STR0, k < R0
  // HF: the digits are doubles and the first stage runs on the FP64 pipe (head_stage1_f64); otherwise biased words
  uint32_t dl[HF ? 1 : 2][HF ? 1 : R0], dh[HF ? 1 : 2][HF ? 1 : R0];
  double dd[HF ? 2 : 1][HF ? R0 : 1];
#pragma unroll
  for (int k = 0; k < R0; ++k) {
    if constexpr (HF) dd[0][k] = S.digd[2 * m + tid + k * STR0];
    else { const uint2 v = S.dig2[2 * m + tid + k * STR0]; dl[0][k] = v.x; dh[0][k] = v.y; }
  }
#pragma unroll 1
  for (int i = 0; i < L; ++i) {
    const uint32_t p = C.p[i], p2 = 2 * p, z = C.zero;
    // ---- P0: digits -> residues -> top LR0 stages in registers -> shared memory (register double buffered) ----
    {
      const uint32_t mu = C.dig_mu[i], negc = C.dig_negc[i];
      const double fp = C.hp_p[i], fpinv = C.hp_pinv[i], fw = C.hp_w[i], fwp = C.hp_wp[i], fc = C.hp_c[i];
#pragma unroll
      for (int it = 0; it < NI; ++it) {
        const int j = (it / TB + 2) & 3, o = (it % TB) * T;
        if (it + 1 < NI) {
          const int jn = ((it + 1) / TB + 2) & 3, on = ((it + 1) % TB) * T;
#pragma unroll
          for (int k = 0; k < R0; ++k) {
            if constexpr (HF) dd[(it + 1) & 1][k] = S.digd[jn * m + tid + on + k * STR0];
            else { const uint2 v = S.dig2[jn * m + tid + on + k * STR0]; dl[(it + 1) & 1][k] = v.x; dh[(it + 1) & 1][k] = v.y; }
          }
        }
        // no barrier before buffers 0,1 are overwritten: their last reader (the previous prime's residue store) read, in
        // this same thread, exactly the addresses st + tb T + k STR0 that are written here
        uint32_t x[R0];
        if constexpr (HF) {
          head_stage1_f64<R0>(dd[it & 1], x, fp, fpinv, fw, fwp, fc);
          fwd_block_tail<LR0>(x, C.topf[i], p, p2, z);
        } else {
#pragma unroll
          for (int k = 0; k < R0; ++k) x[k] = digit_mod(dl[it & 1][k], dh[it & 1][k], mu, negc, p);
          fwd_block<LR0>(x, C.topf[i], p, p2, z);
        }
#pragma unroll
        for (int k = 0; k < R0; ++k) sm[j * m + st + o + k * STR0] = x[k];
      }
    }
    __syncthreads();
    mbar_wait(bar, parity); parity ^= 1;                 // forward table of this prime (staged one prime ago)
    SGFHE_TICK(0);
    pass8_v4<LOGM, 4, true, 6, TB>(sm, tab, p, z);
    slice_sync<LOGM, TB>();                                  // bits [0,9) stay inside one slice (one warp, or 64 consecutive threads)
    const uint32_t* K = keyrow + (size_t)i * 8 * m;      // [4][2][m] for this prime   (src/fhe.jl:527-528)
    uint4 kq[kKeyTma ? 1 : 2][4];                        // key words of (block, poly): rows 2j and 2j+1, 8 indices each;
    // TMA path: this warp's staging buffer [2][256] words, its mbarrier, and the request of item sidx = 4 q + j by lane 0
    const int warp = tid >> 5; [[maybe_unused]] const int lane = tid & 31;
    uint32_t* kbuf = reinterpret_cast<uint32_t*>(key_stage_base<LOGM>(sm)) + warp * 512;
    uint64_t* kbar = reinterpret_cast<uint64_t*>(key_stage_base<LOGM>(sm) + 16 * 2048) + warp;
    auto key_chunk0 = [&](int q) { return key_pos<LOGM>(8 * block_of<LOGM, TB>(warp * 32, q)); };   // first word of the warp's 1 KiB run in a key row
    auto key_request = [&](int sidx) {
      const int q = sidx / 4, j = sidx % 4;
      const uint32_t* r0 = K + (size_t)(2 * j) * m + key_chunk0(q);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the warp's reads of the buffer (before the __syncwarp) are done
      mbar_expect_tx(kbar, 2048);
      bulk_g2s(kbuf, r0, 1024, kbar);
      bulk_g2s(kbuf + 256, r0 + m, 1024, kbar);
    };
    if constexpr (kKeyTma) {
      if (lane == 0) key_request(0);                     // the first pair is requested before the stride-8 pass
    } else {                                             // the first set is requested before the stride-8 pass
      const int kb = key_pos<LOGM>(8 * block_of<LOGM, TB>(tid, 0)), kh = key_pos<LOGM>(8 * block_of<LOGM, TB>(tid, 0) + 4);
      kq[0][0] = __ldg(reinterpret_cast<const uint4*>(K + kb)); kq[0][1] = __ldg(reinterpret_cast<const uint4*>(K + kh));
      kq[0][2] = __ldg(reinterpret_cast<const uint4*>(K + m + kb)); kq[0][3] = __ldg(reinterpret_cast<const uint4*>(K + m + kh));
    }
    pass8_v4<LOGM, 4, true, 3, TB>(sm, tab, p, z);
    __syncwarp();                                        // bits [0,6) stay inside groups of 8 consecutive threads
    SGFHE_TICK(1);
    // ---- fused: stride-1 forward pass + 8 key MACs per point + stride-1 inverse pass ------------------------
    {
      const uint32_t pinv = C.pinv_neg[i];
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const int blk = block_of<LOGM, TB>(tid, q), base = 8 * blk;
        const int a0 = swz(base), a1 = a0 ^ 4;
        uint2 w[7];
        block_twiddles<true>(tab, m / 8, blk, p, w);
        uint64_t sa[8], sb[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { sa[e] = 0; sb[e] = 0; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int sidx = q * 4 + j;
          if constexpr (kKeyTma) {
            // items complete in order on the warp's mbarrier: phase parity = item parity (NB * 4 items per prime, an even number)
            mbar_wait(kbar, sidx & 1);
            const int o0 = key_pos<LOGM>(8 * blk) - key_chunk0(q), o1 = key_pos<LOGM>(8 * blk + 4) - key_chunk0(q);
            kq[0][0] = *reinterpret_cast<const uint4*>(kbuf + o0); kq[0][1] = *reinterpret_cast<const uint4*>(kbuf + o1);
            kq[0][2] = *reinterpret_cast<const uint4*>(kbuf + 256 + o0); kq[0][3] = *reinterpret_cast<const uint4*>(kbuf + 256 + o1);
            __syncwarp();
            if (sidx + 1 < NB * 4 && lane == 0) key_request(sidx + 1);
          } else if (sidx + 1 < NB * 4) {                // prefetch the next (block, poly) key words
            const int nq = (sidx + 1) / 4, nj = (sidx + 1) % 4;
            const int kb = key_pos<LOGM>(8 * block_of<LOGM, TB>(tid, nq)), kh = key_pos<LOGM>(8 * block_of<LOGM, TB>(tid, nq) + 4);
            const uint32_t* r0 = K + (size_t)(2 * nj) * m;
            const uint32_t* r1 = K + (size_t)(2 * nj + 1) * m;
            kq[(sidx + 1) & 1][0] = __ldg(reinterpret_cast<const uint4*>(r0 + kb)); kq[(sidx + 1) & 1][1] = __ldg(reinterpret_cast<const uint4*>(r0 + kh));
            kq[(sidx + 1) & 1][2] = __ldg(reinterpret_cast<const uint4*>(r1 + kb)); kq[(sidx + 1) & 1][3] = __ldg(reinterpret_cast<const uint4*>(r1 + kh));
          }
          uint32_t x[8];
          {
            const uint4 v0 = *reinterpret_cast<const uint4*>(sm + j * m + a0);
            const uint4 v1 = *reinterpret_cast<const uint4*>(sm + j * m + a1);
            x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
          }
          fwd_block<3>(x, w, p, p2, z);
          const uint4* kk = kq[kKeyTma ? 0 : (sidx & 1)];
          const uint32_t ka[8] = {kk[0].x, kk[0].y, kk[0].z, kk[0].w, kk[1].x, kk[1].y, kk[1].z, kk[1].w};
          const uint32_t kb[8] = {kk[2].x, kk[2].y, kk[2].z, kk[2].w, kk[3].x, kk[3].y, kk[3].z, kk[3].w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            // only two of the four transforms are corrected to [0, 2p): T < 2 (2p + 4p) p = 12 p^2, T + q p < 12 p^2 + 2^32 p < 2^64
            const uint32_t d = j < 2 ? min(x[e], x[e] - p2) : x[e];
            sa[e] += (uint64_t)d * ka[e];
            sb[e] += (uint64_t)d * kb[e];
          }
        }
        uint2 wi[7];
        block_twiddles<false>(tab, m / 8, blk, p, wi);
        uint32_t ya[8], yb[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {                                // redc of T < 12 p^2 lands in [0, 4p)
          ya[e] = redc(sa[e], p, pinv); ya[e] = min(ya[e], ya[e] - p2);
          yb[e] = redc(sb[e], p, pinv); yb[e] = min(yb[e], yb[e] - p2);
        }
        inv_block<3, true>(ya, wi, p, p2, z);
        inv_block<3, true>(yb, wi, p, p2, z);
        *reinterpret_cast<uint4*>(sm + a0) = make_uint4(ya[0], ya[1], ya[2], ya[3]);
        *reinterpret_cast<uint4*>(sm + a1) = make_uint4(ya[4], ya[5], ya[6], ya[7]);
        *reinterpret_cast<uint4*>(sm + m + a0) = make_uint4(yb[0], yb[1], yb[2], yb[3]);
        *reinterpret_cast<uint4*>(sm + m + a1) = make_uint4(yb[4], yb[5], yb[6], yb[7]);
      }
    }
    __syncwarp();
    SGFHE_TICK(2);
    pass8_v4<LOGM, 2, false, 3, TB>(sm, tab, p, z);
    slice_sync<LOGM, TB>();
    pass8_v4<LOGM, 2, false, 6, TB>(sm, tab, p, z);
    if (i + 1 < L) {                                     // next prime's first digit words: requested before the barrier, so a
#pragma unroll                                           // warp that arrives early waits with its loads already in flight
      for (int k = 0; k < R0; ++k) {
        if constexpr (HF) dd[0][k] = S.digd[2 * m + tid + k * STR0];
        else { const uint2 v = S.dig2[2 * m + tid + k * STR0]; dl[0][k] = v.x; dh[0][k] = v.y; }
      }
    }
    __syncthreads();                                     // last reader of `tab` for this prime is done
    {
      // TMA: next prime's forward table under the store phase.  After the last prime the table region serves the
      // CRT/update phases as a staging area; prime 0's table for the next step is requested after them.
      if (tid == 0 && i + 1 < L) stage_table(tab, tw_f + (size_t)(i + 1) * m, m * 8, bar);
    }
    SGFHE_TICK(3);
    // ---- top inverse stages in registers + CRT pre-scaling + store of the residues --------------------------
    // last level: x' = x + y, y' = (x - y) psi^(-m/2); the scaling s = m^-1 (P/p)^-1 rides in the key words (DevConst::keymul)
    {
      const uint32_t sw = C.lastw[i], sws = C.lastw_sh[i];
#pragma unroll
      for (int it = 0; it < 2 * TB; ++it) {
        const int c = it / TB, o = (it % TB) * T;
        uint32_t x[R0];
#pragma unroll
        for (int k = 0; k < R0; ++k) x[k] = sm[c * m + st + o + k * STR0];
        inv_block_upper<LR0>(x, C.topi[i], p, p2, z);
#pragma unroll
        for (int k = 0; k < R0 / 2; ++k) {
          const uint32_t s0 = x[k] + x[k + R0 / 2] + z, d0 = x[k] - x[k + R0 / 2] + p2;
          S.zres[((size_t)i * 2 + c) * m + tid + o + k * STR0] = csub(min(s0, s0 - p2), p);
          S.zres[((size_t)i * 2 + c) * m + tid + o + (k + R0 / 2) * STR0] = csub(shoup_mul(d0, sw, sws, p), p);
        }
      }
    }
    SGFHE_TICK(4);
  }
#ifdef SGFHE_TAIL_SPLIT                                   // measured slower: kept for A/B builds only
  crt_update_split<LOGM, T, HF>(C, S, reinterpret_cast<uint4*>(sm), sm + 4 * m, draws_next, u, ext, decompose_next, timing, tprev);
#else
  crt_update<LOGM, T, true, HF>(C, S, reinterpret_cast<uint4*>(sm), sm + 4 * m, draws_next, u, ext, decompose_next, timing, tprev);   // begins with the barrier that frees shared memory
#endif
  if (tid == 0) stage_table(tab, tw_f, m * 8, bar);      // ends with a barrier: the staging area is free again
#undef SGFHE_TICK
}

// =========================================================================================================
// v5: two gates per SM at m = 8192.  One gate per SM (v4) leaves every pipe idle part of the time: the transform phases
// saturate the FMA-heavy pipe while the CRT / update / decomposition tail is ALU- and latency-bound, and the phases of
// one gate cannot overlap because each depends on the previous one.  Two independent gates on one SM can.  A gate gets a
// 256-thread CTA (8 warps, 128 registers per thread) and 81 KiB of shared memory:
//   * after the top four stages the sixteen 512-point slices of a polynomial are independent, so a prime is processed in
//     two halves of eight slices (one slice per warp, exactly the per-warp structure of v4): 64 KiB for the four digit
//     polynomials of a half;
//   * the head runs the FP64 first stage once per (polynomial, radix-16 group) and parks the eight differences of the second
//     half in L2 (thread private); the inverse top stages of the first half park their eight outputs the same way and
//     the second half finishes the last stage against them;
//   * only twiddles 16..1023 (the two strided passes, 8 KiB per direction) are staged in shared memory by TMA; the
//     stride-1 level (used once per prime and thread) is read from L2 / L1;
//   * the unreduced CRT sums of the tail are staged in global scratch instead of shared memory.
// =========================================================================================================
struct Shape5 { static constexpr int LOGM = 13, M = 8192, T = 256, HALF = 4096, TABN = 1008; };

// twiddles of one radix-8 block, direct order, from a table that starts at entry `first`
__device__ __forceinline__ void block_twiddles_at(const uint2* tab, int first, int t1, uint2 (&w)[7]) {
  w[0] = tab[t1 - first];
  const uint4 a = *reinterpret_cast<const uint4*>(&tab[2 * t1 - first]);
  const uint4 b0 = *reinterpret_cast<const uint4*>(&tab[4 * t1 - first]);
  const uint4 b1 = *reinterpret_cast<const uint4*>(&tab[4 * t1 + 2 - first]);
  w[1] = make_uint2(a.x, a.y); w[2] = make_uint2(a.z, a.w);
  w[3] = make_uint2(b0.x, b0.y); w[4] = make_uint2(b0.z, b0.w); w[5] = make_uint2(b1.x, b1.y); w[6] = make_uint2(b1.z, b1.w);
}
__device__ __forceinline__ void block_twiddles_ldg(const uint2* __restrict__ tw, int t1, uint2 (&w)[7]) {
  const uint2 w0 = __ldg(&tw[t1]);
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(&tw[2 * t1]));
  const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(&tw[4 * t1]));
  const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(&tw[4 * t1 + 2]));
  w[0] = w0; w[1] = make_uint2(a.x, a.y); w[2] = make_uint2(a.z, a.w);
  w[3] = make_uint2(b0.x, b0.y); w[4] = make_uint2(b0.z, b0.w); w[5] = make_uint2(b1.x, b1.y); w[6] = make_uint2(b1.z, b1.w);
}

// one strided radix-8 pass (B = 6 or 3) over NPOLY half polynomials (4096 words each): the warp owns slice 8 h + warp, a lane
// two adjacent blocks (64-bit shared-memory accesses), twiddles from the staged table of entries 16..1023
template <int NPOLY, bool FWD, int B>
__device__ __forceinline__ void pass8_v5(uint32_t* sm, const uint2* tab, int h, uint32_t p, uint32_t z) {
  const uint32_t p2 = 2 * p;
  const int blk = ((threadIdx.x >> 5) << 6) | ((threadIdx.x & 31) << 1);
  const int base = ((blk >> B) << (B + 3)) | (blk & ((1 << B) - 1));
  uint2 w[7];
  block_twiddles_at(tab, 16, (Shape5::M >> (B + 3)) + ((blk + h * 512) >> B), w);
  int off[8];
  if constexpr (B == 6) {
    const int b0 = swz(base);
#pragma unroll
    for (int j = 0; j < 8; ++j) off[j] = (b0 ^ ((j & 3) << 3)) + (j << 6);
  } else {
    const int c = (base >> 6) & 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) off[j] = ((base + (((j & 3) ^ c) << 3)) ^ ((j >> 2) << 2)) + ((j >> 2) << 5);
  }
#pragma unroll
  for (int poly = 0; poly < NPOLY; ++poly) {
    uint32_t* s = sm + poly * Shape5::HALF;
    uint32_t x[8], y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { const uint2 v = *reinterpret_cast<const uint2*>(s + off[j]); x[j] = v.x; y[j] = v.y; }
    if (FWD) { fwd_block<3>(x, w, p, p2, z); fwd_block<3>(y, w, p, p2, z); }
    else { inv_block<3>(x, w, p, p2, z); inv_block<3>(y, w, p, p2, z); }
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<uint2*>(s + off[j]) = make_uint2(x[j], y[j]);
  }
}

// Tail of a v5 step.  96 KiB of shared memory hold (-z) mod Q of one result polynomial as three limb arrays (the unreduced
// four-limb sums of the v4 tail would need 128 KiB): CRT sum + Barrett per coefficient, barrier, then
// acc += x^u z - z as acc + r[j] + (r[j-u] or Q - r[j-u]) with two conditional subtractions, fused with the next step's
// decomposition.  Every thread handles the coefficients tid + 256 it, whose residues, accumulator words and digits are its own.
template <bool EXT, bool DEC, bool RAND>
__device__ __forceinline__ void update_v5(const DevConst& C, const Scratch& S, const uint32_t* sm, int c, const DrawSrc draws_next, int u) {
  constexpr int LOGM = Shape5::LOGM, m = Shape5::M, T = Shape5::T, KB = 3 * LOGM - 1, NIT = m / T, D = 4;
  const int tid = threadIdx.x;
  const u96 Q = Q96(C), OFF = off96(C);
  uint32_t* acc = S.acc + c * 3 * m;
  u96 aq[D];
  if (!EXT) {
#pragma unroll
    for (int d = 0; d < D; ++d) aq[d] = ld96(acc, m, tid + d * T);
  }
#pragma unroll 1
  for (int it0 = 0; it0 < NIT; it0 += D) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const int j = tid + (it0 + d) * T;
      const u96 rj = ld96(sm, m, j);                                  // (-z[j]) mod Q
      u96 res;
      if (EXT) res = addmod96(negmod96(rj, Q), OFF, Q);               // acc = z, in offset form
      else {
        const u96 a = aq[d];
        if (it0 + d + D < NIT) aq[d] = ld96(acc, m, j + D * T);
        const int src = (j - u) & (2 * m - 1);
        const u96 rs = ld96(sm, m, src & (m - 1));                    // (-z[j-u]) mod Q
        uint32_t bw;
        const u96 pos = sub96(Q, rs, bw);                             // +z[j-u] (= Q when z = 0: absorbed by the csubQ below)
        const u96 term = sel96(src >= m, rs, pos);                    // x^u z wraps with a sign flip
        res = csubQ(add96(csubQ(add96(a, rj), Q), term), Q);          // acc - z[j] +- z[j-u]
      }
      st96(acc, m, j, res);
      if (DEC) {
        uint64_t dp0, dp1;
        if (RAND) { int64_t x0, x1; get_draws(C, draws_next, c, j, m, x0, x1); decompose_off_rand<KB>(C, res, x0, x1, dp0, dp1); }
        else decompose_off<KB>(C, res, dp0, dp1);
        S.digd[(2 * c) * m + j] = digit_f64(dp0); S.digd[(2 * c + 1) * m + j] = digit_f64(dp1);
      }
    }
  }
}

__device__ __forceinline__ void tail_v5(const DevConst& C, const Scratch& S, uint32_t* sm, const DrawSrc draws_next, int u, bool ext,
                                        bool decompose_next, unsigned long long* timing, long long& tprev) {
  constexpr int LOGM = Shape5::LOGM, m = Shape5::M, T = Shape5::T, L = Shape<LOGM>::L, SB = 6 * LOGM + 8, NIT = m / T, D = 4;
  const int tid = threadIdx.x;
#define SGFHE_TICK(slot) do { if (timing && threadIdx.x == 0) { const long long tn_ = clock64(); timing[slot] += (unsigned long long)(tn_ - tprev); tprev = tn_; } } while (0)
  uint32_t yq[D][L];
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int i = 0; i < L; ++i) yq[d][i] = S.zres[(size_t)i * 2 * m + tid + d * T];
  __syncthreads();                                           // every warp has left the transform buffers
  for (int c = 0; c < 2; ++c) {
    const uint32_t* zr = S.zres + (size_t)c * m;
    if (c == 1) {
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
        for (int i = 0; i < L; ++i) yq[d][i] = zr[(size_t)i * 2 * m + tid + d * T];
    }
#pragma unroll 1
    for (int it0 = 0; it0 < NIT; it0 += D) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int idx = tid + (it0 + d) * T;
        const u96 r = barrett96<SB>(C, crt_sum<0, L>(C, yq[d], 1));
        if (it0 + d + D < NIT) {
#pragma unroll
          for (int i = 0; i < L; ++i) yq[d][i] = zr[(size_t)i * 2 * m + idx + D * T];
        }
        st96(sm, m, idx, r);
      }
    }
    __syncthreads();
    SGFHE_TICK(5);
    if (ext) {
      if (!decompose_next) update_v5<true, false, false>(C, S, sm, c, draws_next, u);
      else if (draws_next.on()) update_v5<true, true, true>(C, S, sm, c, draws_next, u);
      else update_v5<true, true, false>(C, S, sm, c, draws_next, u);
    } else {
      if (!decompose_next) update_v5<false, false, false>(C, S, sm, c, draws_next, u);
      else if (draws_next.on()) update_v5<false, true, true>(C, S, sm, c, draws_next, u);
      else update_v5<false, true, false>(C, S, sm, c, draws_next, u);
    }
    __syncthreads();
    SGFHE_TICK(6);
  }
#undef SGFHE_TICK
}

__device__ __forceinline__ void stage_tables_v5(uint2* tabf, uint2* tabi, const uint2* twf, const uint2* twi, uint64_t* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(bar, 2 * Shape5::TABN * 8);
  bulk_g2s(tabf, twf + 16, Shape5::TABN * 8, bar);
  bulk_g2s(tabi, twi + 16, Shape5::TABN * 8, bar);
}

__device__ void gate_step_v5(const DevConst& C, const Scratch& S, uint32_t* sm, const uint32_t* __restrict__ keyrow,
                             const uint2* __restrict__ tw_f, const uint2* __restrict__ tw_i, const DrawSrc draws_next, int u,
                             bool ext, bool decompose_next, uint2* tabf, uint2* tabi, uint64_t* bar, uint32_t& parity,
                             unsigned long long* timing) {
  constexpr int LOGM = Shape5::LOGM, m = Shape5::M, T = Shape5::T, HALF = Shape5::HALF, L = Shape<LOGM>::L;
  const int tid = threadIdx.x;
  long long tprev = timing ? clock64() : 0;
#define SGFHE_TICK(slot) do { if (timing && threadIdx.x == 0) { const long long tn_ = clock64(); timing[slot] += (unsigned long long)(tn_ - tprev); tprev = tn_; } } while (0)
  const int st = swz(tid);                               // swz(tid + 256 q + 512 k) = swz(tid) + 256 q + 512 k
  const int blk0 = ((tid >> 5) << 6) | ((tid & 31) << 1);     // the lane's first (even) radix-8 block inside the half
  double dd[2][16];
  auto load_digits = [&](const int b, const int item) {  // item = 2 poly + group: the 16 digits t + 512 k of radix-16 group t
#pragma unroll
    for (int k = 0; k < 16; ++k) dd[b][k] = S.digd[(item >> 1) * m + tid + (item & 1) * T + k * 512];
  };
  load_digits(0, 0);
#pragma unroll 1
  for (int i = 0; i < L; ++i) {
    const uint32_t p = C.p[i], p2 = 2 * p, z = C.zero;
    const uint32_t* K = keyrow + (size_t)i * 8 * m;      // [4][2][m] for this prime   (src/fhe.jl:527-528)
    const uint2* twf = tw_f + (size_t)i * m;
    const uint2* twi = tw_i + (size_t)i * m;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      // ---- head of this half: (h = 0) digits -> FP64 first stage -> sums go on, differences are parked;
      //      (h = 1) the parked differences come back.  Then three stages in registers -> shared memory. ----
      if (h == 0) {
        const double fp = C.hp_p[i], fpinv = C.hp_pinv[i], fw = C.hp_w[i], fwp = C.hp_wp[i], fc = C.hp_c[i];
#pragma unroll
        for (int item = 0; item < 8; ++item) {
          if (item + 1 < 8) load_digits((item + 1) & 1, item + 1);
          uint32_t x[16];
          head_stage1_f64<16>(dd[item & 1], x, fp, fpinv, fw, fwp, fc);
          uint32_t* pk = S.park + (size_t)item * 8 * T + tid;
#pragma unroll
          for (int k = 0; k < 8; ++k) pk[k * T] = x[8 + k];
          uint32_t lo[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) lo[k] = x[k];
          fwd_block<3>(lo, C.topf_h[i][0], p, p2, z);
          uint32_t* dst = sm + (item >> 1) * HALF + st + (item & 1) * T;
#pragma unroll
          for (int k = 0; k < 8; ++k) dst[k * 512] = lo[k];
        }
      } else {
        uint32_t pv[2][8];
#pragma unroll
        for (int k = 0; k < 8; ++k) pv[0][k] = S.park[(size_t)k * T + tid];
#pragma unroll
        for (int item = 0; item < 8; ++item) {
          if (item + 1 < 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) pv[(item + 1) & 1][k] = S.park[((size_t)(item + 1) * 8 + k) * T + tid];
          }
          uint32_t x[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) x[k] = pv[item & 1][k];
          fwd_block<3>(x, C.topf_h[i][1], p, p2, z);
          uint32_t* dst = sm + (item >> 1) * HALF + st + (item & 1) * T;
#pragma unroll
          for (int k = 0; k < 8; ++k) dst[k * 512] = x[k];
        }
      }
      __syncthreads();
      if (h == 0) { mbar_wait(bar, parity); parity ^= 1; }      // tables of this prime (staged one prime ago)
      SGFHE_TICK(0);
      pass8_v5<4, true, 6>(sm, tabf, h, p, z);
      __syncwarp();                                      // a warp owns its slice: bits [0,9) never leave it
      uint4 kq[2][4];
      {
        const int pt = 8 * (blk0 + h * 512);
        const int kb = key_pos<LOGM>(pt), kh = key_pos<LOGM>(pt + 4);
        kq[0][0] = __ldg(reinterpret_cast<const uint4*>(K + kb)); kq[0][1] = __ldg(reinterpret_cast<const uint4*>(K + kh));
        kq[0][2] = __ldg(reinterpret_cast<const uint4*>(K + m + kb)); kq[0][3] = __ldg(reinterpret_cast<const uint4*>(K + m + kh));
      }
      uint2 wnext[7];                                    // stride-1 twiddles come from L2: requested one block phase ahead
      block_twiddles_ldg(twf, m / 8 + blk0 + h * 512, wnext);
      pass8_v5<4, true, 3>(sm, tabf, h, p, z);
      __syncwarp();
      SGFHE_TICK(1);
      // ---- fused: stride-1 forward pass + 8 key MACs per point + stride-1 inverse pass ----
      {
        const uint32_t pinv = C.pinv_neg[i];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int blk = blk0 + q, base = 8 * blk, blkg = blk + h * 512;
          const int a0 = swz(base), a1 = a0 ^ 4;
          uint2 w[7], wi[7];
#pragma unroll
          for (int k = 0; k < 7; ++k) w[k] = wnext[k];
          block_twiddles_ldg(twi, m / 8 + blkg, wi);      // inverse twiddles of this block: in flight under the forward part
          uint64_t sa[8], sb[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { sa[e] = 0; sb[e] = 0; }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int sidx = q * 4 + j;
            if (sidx + 1 < 8) {                            // prefetch the next (block, poly) key words
              const int nq = (sidx + 1) / 4, nj = (sidx + 1) % 4;
              const int pt = 8 * (blk0 + nq + h * 512);
              const int kb = key_pos<LOGM>(pt), kh = key_pos<LOGM>(pt + 4);
              const uint32_t* r0 = K + (size_t)(2 * nj) * m;
              const uint32_t* r1 = K + (size_t)(2 * nj + 1) * m;
              kq[(sidx + 1) & 1][0] = __ldg(reinterpret_cast<const uint4*>(r0 + kb)); kq[(sidx + 1) & 1][1] = __ldg(reinterpret_cast<const uint4*>(r0 + kh));
              kq[(sidx + 1) & 1][2] = __ldg(reinterpret_cast<const uint4*>(r1 + kb)); kq[(sidx + 1) & 1][3] = __ldg(reinterpret_cast<const uint4*>(r1 + kh));
            }
            uint32_t x[8];
            {
              const uint4 v0 = *reinterpret_cast<const uint4*>(sm + j * HALF + a0);
              const uint4 v1 = *reinterpret_cast<const uint4*>(sm + j * HALF + a1);
              x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
            }
            fwd_block<3>(x, w, p, p2, z);
            const uint4* kk = kq[sidx & 1];
            const uint32_t ka[8] = {kk[0].x, kk[0].y, kk[0].z, kk[0].w, kk[1].x, kk[1].y, kk[1].z, kk[1].w};
            const uint32_t kb[8] = {kk[2].x, kk[2].y, kk[2].z, kk[2].w, kk[3].x, kk[3].y, kk[3].z, kk[3].w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint32_t d = min(x[e], x[e] - p2);                 // [0, 2p): four products < 8 p^2 < 2^63
              sa[e] += (uint64_t)d * ka[e];
              sb[e] += (uint64_t)d * kb[e];
            }
          }
          if (q == 0) block_twiddles_ldg(twf, m / 8 + blkg + 1, wnext);
          uint32_t ya[8], yb[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {                                // redc of T < 8 p^2 lands in [0, 3p)
            ya[e] = redc(sa[e], p, pinv); ya[e] = min(ya[e], ya[e] - p2);
            yb[e] = redc(sb[e], p, pinv); yb[e] = min(yb[e], yb[e] - p2);
          }
          inv_block<3>(ya, wi, p, p2, z);
          inv_block<3>(yb, wi, p, p2, z);
          *reinterpret_cast<uint4*>(sm + a0) = make_uint4(ya[0], ya[1], ya[2], ya[3]);
          *reinterpret_cast<uint4*>(sm + a1) = make_uint4(ya[4], ya[5], ya[6], ya[7]);
          *reinterpret_cast<uint4*>(sm + HALF + a0) = make_uint4(yb[0], yb[1], yb[2], yb[3]);
          *reinterpret_cast<uint4*>(sm + HALF + a1) = make_uint4(yb[4], yb[5], yb[6], yb[7]);
        }
      }
      __syncwarp();
      SGFHE_TICK(2);
      pass8_v5<2, false, 3>(sm, tabi, h, p, z);
      __syncwarp();
      pass8_v5<2, false, 6>(sm, tabi, h, p, z);
      if (h == 1 && i + 1 < L) load_digits(0, 0);        // next prime's first digits: in flight across the barrier
      __syncthreads();
      if (h == 1 && tid == 0) {                          // last reader of the tables is done: next prime's (or next step's prime 0)
        const int nx = i + 1 < L ? i + 1 : 0;
        stage_tables_v5(tabf, tabi, tw_f + (size_t)nx * m, tw_i + (size_t)nx * m, bar);
      }
      SGFHE_TICK(3);
      // ---- inverse top stages of this half in registers; the last stage joins the two halves ----
      {
        const uint32_t sw = C.lastw[i], sws = C.lastw_sh[i];
#pragma unroll
        for (int item = 0; item < 4; ++item) {             // item = 2 c + group
          const int c = item >> 1, t = tid + (item & 1) * T;
          uint32_t x[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) x[k] = sm[c * HALF + st + (item & 1) * T + k * 512];
          inv_block<3>(x, C.topi_h[i][h], p, p2, z);
          uint32_t* pk = sm + 4 * HALF + item * 8 * T + tid;   // thread-private words behind the four half buffers
          if (h == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) pk[k * T] = x[k];
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t y0 = pk[k * T];               // first half's value of the pair (k, k + 8)
              const uint32_t s0 = y0 + x[k] + z, d0 = y0 - x[k] + p2;
              S.zres[((size_t)i * 2 + c) * m + t + k * 512] = csub(min(s0, s0 - p2), p);
              S.zres[((size_t)i * 2 + c) * m + t + (k + 8) * 512] = csub(shoup_mul(d0, sw, sws, p), p);
            }
          }
        }
      }
      SGFHE_TICK(4);
    }
  }
  tail_v5(C, S, sm, draws_next, u, ext, decompose_next, timing, tprev);
#undef SGFHE_TICK
}

// Persistent gate loop shared by the two gate kernels (V4: fused-pass step for m >= 4096; HF: FP64 head, digits kept as
// doubles).  SMs do not all run at the same pace (two dies, different distances to L2), so with a work counter a CTA
// takes the next unprocessed gate when it finishes one instead of a fixed share of the batch.
// F_DECOMP: (re)compute S.dig from S.acc before the first step of this launch (set with F_INIT, and by the
// trace / external-product seams whose accumulator arrives from a previous launch or from the host).
template <int LOGM, int VER, bool HF>
__device__ __forceinline__ void run_gates(const DevConst& C, const GateArgs& A, uint32_t* sm, uint2* tab, uint64_t* bar,
                                          uint32_t& parity, uint32_t* slot) {
  constexpr int m = 1 << LOGM;
  const int n = C.n;
  const Scratch S = carve(A.scratch + (size_t)blockIdx.x * A.scratch_stride, A.zres + (size_t)blockIdx.x * A.zres_stride, m);
  const uint64_t rmask = (1ull << C.logr) - 1;
  if (A.stagger_cycles > 0) {                            // experiment knob: CTA b starts (b % slots) * cycles late
    const long long t0 = clock64(), wait = (long long)(blockIdx.x % A.stagger_slots) * A.stagger_cycles;
    while (clock64() - t0 < wait) __nanosleep(256);
  }
  for (int g = blockIdx.x;;) {
    if (A.work_counter) {
      if (threadIdx.x == 0) *slot = (uint32_t)atomicAdd(A.work_counter, 1);
      __syncthreads();
      g = (int)*slot;
    }
    if (g >= A.batch) break;
    const uint64_t* l1 = A.lwe1 + (size_t)g * (n + 1);
    const uint64_t* l2 = A.lwe2 + (size_t)g * (n + 1);
    const bool pack = (A.flags & F_PACK) != 0;           // shortened_external_product batch (src/fhe.jl:632-641, 683-684)
    const int64_t* dr = A.draws ? A.draws + (size_t)g * A.draw_steps * 4 * m : nullptr;
    const uint64_t seed = A.draws ? 0 : A.rng_seed;
    const uint64_t gid = A.rng_gate0 + (uint64_t)g;
    auto draws_of = [&](int k) {                          // the draws of accumulation step k of this gate (none: deterministic)
      DrawSrc d; d.ptr = dr ? dr + (size_t)(k - (pack ? 0 : A.step_begin)) * 4 * m : nullptr; d.seed = seed; d.gate = gid; d.step = (uint32_t)k;
      return d;
    };
    if (A.flags & F_INIT) gate_init<LOGM>(C, S, (l1[n] + l2[n]) & rmask);
    if (pack) {                                          // a = 0, b = input polynomial g: only the b-digit key rows contribute
      u96 zero; zero.x0 = zero.x1 = zero.x2 = 0;
      zero = to_offset_form(C, zero);
      for (int e = threadIdx.x; e < m; e += blockDim.x) {
        const uint64_t* src = A.pack_in + ((size_t)g * m + e) * 2;
        u96 v; v.x0 = (uint32_t)src[0]; v.x1 = (uint32_t)(src[0] >> 32); v.x2 = (uint32_t)src[1];
        st96(S.acc, m, e, zero); st96(S.acc + 3 * m, m, e, to_offset_form(C, v));
      }
    }
    __syncthreads();
    const int k_begin = pack ? g : A.step_begin, k_end = pack ? g + 1 : A.step_end;
    if ((A.flags & F_DECOMP) && k_begin < k_end) {
      DrawSrc d0 = draws_of(k_begin), d1 = d0;
      if (pack) {                                        // shortened product: a = 0 is flattened without draws, b with its own
        d0.ptr = nullptr; d0.seed = 0;
        d1.ptr = A.pack_draws ? A.pack_draws + (size_t)g * 2 * m - 2 * m : nullptr;      // indexed with c = 1
        d1.seed = A.pack_draws ? 0 : A.rng_seed; d1.step = (uint32_t)n;                 // a step index no gate uses
      }
      decompose_poly<LOGM, HF>(C, S, 0, d0);
      decompose_poly<LOGM, HF>(C, S, 1, d1);
    }
    __syncthreads();
    uint64_t usum = (A.flags & F_EXT) || k_begin >= k_end ? 0 : l1[k_begin] + l2[k_begin];
    for (int k = k_begin; k < k_end; ++k) {
      const int u = (int)(usum & rmask);                                          // u.a[k], src/fhe.jl:566,580
      const bool more = k + 1 < k_end;
      if (more && !(A.flags & F_EXT)) usum = l1[k + 1] + l2[k + 1];               // next step's rotation: in flight during this step
      const uint32_t* keyrow = A.keyhat + (size_t)k * C.L * 8 * m;
      DrawSrc dn = draws_of(k + 1);
      if (!more) { dn.ptr = nullptr; dn.seed = 0; }
      unsigned long long* tm = blockIdx.x == 0 ? A.timing : nullptr;
      if constexpr (VER == 5) gate_step_v5(C, S, sm, keyrow, A.tw_f, A.tw_i, dn, u, (A.flags & F_EXT) != 0, more, tab, tab + Shape5::TABN, bar, parity, tm);
      else if constexpr (VER == 4) gate_step_v4<LOGM, HF>(C, S, sm, keyrow, A.tw_f, dn, u, (A.flags & F_EXT) != 0, more, tab, bar, parity, tm);
      else if constexpr (Shape<LOGM>::WIDE) gate_step_wide<LOGM>(C, S, sm, keyrow, A.tw_f, A.tw_i, dn, u, (A.flags & F_EXT) != 0, more, tm);
      else gate_step<LOGM>(C, S, sm, keyrow, A.tw_f, A.tw_i, dn, u, (A.flags & F_EXT) != 0, more, tab, bar, parity, tm);
    }
    if (A.trace) {
      uint64_t* tr = A.trace + (pack ? (size_t)g * 4 * m : 0);
      for (int e = threadIdx.x; e < 2 * m; e += blockDim.x) {
        const u96 v = from_offset_form(C, ld96(S.acc + (e / m) * 3 * m, m, e % m));
        tr[2 * e] = (uint64_t)v.x0 | ((uint64_t)v.x1 << 32); tr[2 * e + 1] = v.x2;
      }
    }
    if (A.flags & F_FINAL) {
      const size_t w = (A.flags & F_RAW) ? 2 : 1;
      gate_final(C, S, A.out_and + (size_t)g * (n + 1) * w, A.out_or + (size_t)g * (n + 1) * w,
                 A.out_xor + (size_t)g * (n + 1) * w, (A.flags & F_RAW) != 0);
    }
    __syncthreads();
    if (!A.work_counter) g += gridDim.x;
  }
}

template <int LOGM>
__global__ void __launch_bounds__(Shape<LOGM>::T, 1024 / Shape<LOGM>::T)
bootstrap_kernel(const __grid_constant__ DevConst C, const __grid_constant__ GateArgs A) {
  extern __shared__ __align__(16) uint32_t sm[];
  constexpr int m = 1 << LOGM;
  constexpr int NW = Shape<LOGM>::WIDE ? 2 * m : 6 * m;               // wide: two transform buffers, no staged table
  uint2* tab = reinterpret_cast<uint2*>(sm + 4 * m);                 // staged twiddle table of the current (prime, direction)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + NW);
  uint32_t parity = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  run_gates<LOGM, 3, false>(C, A, sm, tab, bar, parity, sm + NW + 2);
}

template <int LOGM, bool HF>
__global__ void __launch_bounds__(GateShape4<LOGM>::T, GateTB<LOGM>::V)
bootstrap_kernel_v4(const __grid_constant__ DevConst C, const __grid_constant__ GateArgs A) {
  extern __shared__ __align__(16) uint32_t sm[];
  constexpr int m = 1 << LOGM;
  uint2* tab = reinterpret_cast<uint2*>(sm + 4 * m);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 6 * m);
  uint32_t parity = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    if (kKeyTma) for (int w = 0; w < 16; ++w) mbar_init(reinterpret_cast<uint64_t*>(key_stage_base<LOGM>(sm) + 16 * 2048) + w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stage_table(tab, A.tw_f, m * 8, bar);
  }
  __syncthreads();
  run_gates<LOGM, 4, HF>(C, A, sm, tab, bar, parity, sm + 6 * m + 2);
  mbar_wait(bar, parity);                                // the table staged for a step that never runs
}

// two CTAs (gates) per SM; the CTAs of the second wave start half a step late so that the two gates of an SM are in
// different phases from the start
__global__ void __launch_bounds__(Shape5::T, 2)
bootstrap_kernel_v5(const __grid_constant__ DevConst C, const __grid_constant__ GateArgs A) {
  extern __shared__ __align__(16) uint32_t sm[];
  uint2* tabf = reinterpret_cast<uint2*>(sm + 6 * Shape5::HALF);    // 96 KiB: four half buffers + inverse parking; three limb arrays in the tail
  uint64_t* bar = reinterpret_cast<uint64_t*>(tabf + 2 * Shape5::TABN);
  uint32_t parity = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stage_tables_v5(tabf, tabf + Shape5::TABN, A.tw_f, A.tw_i, bar);
  }
  __syncthreads();
  if (2 * blockIdx.x >= gridDim.x && A.stagger_cycles >= 0) {
    const long long t0 = clock64(), wait = A.stagger_cycles > 0 ? A.stagger_cycles : 100000;
    while (clock64() - t0 < wait) __nanosleep(1024);
  }
  GateArgs B = A; B.stagger_cycles = 0;
  run_gates<Shape5::LOGM, 5, true>(C, B, sm, tabf, bar, parity, reinterpret_cast<uint32_t*>(bar + 1));
  mbar_wait(bar, parity);                                // the tables staged for a step that never runs
}
static constexpr size_t kSmemV5 = (size_t)6 * Shape5::HALF * 4 + (size_t)2 * Shape5::TABN * 8 + 32;

// Key pre-transform (K10): coefficient-form wide polys -> per-prime NTT domain, Montgomery form.
// grid = (npolys, L).  coef: [npolys][m][2];  out: poly P of row k=P/8, slot jc=P%8 -> keyhat[((k L + i) 8 + jc) m ..]
template <int LOGM>
__global__ void __launch_bounds__(Shape<LOGM>::T, 1024 / Shape<LOGM>::T)
key_transform_kernel(const __grid_constant__ DevConst C, const uint64_t* __restrict__ coef, uint32_t* __restrict__ keyhat,
                     const uint2* __restrict__ tw_f, int poly0) {
  extern __shared__ __align__(16) uint32_t sm[];
  using SH = Shape<LOGM>;
  constexpr int m = SH::M, REM = SH::REM, R = 1 << REM, STR = SH::STR;
  const int i = blockIdx.y;
  const uint32_t p = C.p[i], p2 = 2 * p;
  uint2* tab = reinterpret_cast<uint2*>(sm + m);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 3 * m);
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) stage_table(tab, tw_f + (size_t)i * m, m * 8, bar);
  const uint64_t* src = coef + (size_t)blockIdx.x * m * 2;
  uint2 wt[R > 1 ? R - 1 : 1];
  top_twiddles<REM>(tw_f + (size_t)i * m, wt);
  for (int idx = threadIdx.x; idx < STR; idx += blockDim.x) {
    uint32_t x[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int e = idx + k * STR;
      x[k] = centred_mod(C, i, src[2 * e], src[2 * e + 1]);
    }
    fwd_block<REM>(x, wt, p, p2, C.zero);
#pragma unroll
    for (int k = 0; k < R; ++k) sm[swz(idx + k * STR)] = x[k];
  }
  __syncthreads();
  mbar_wait(bar, 0);
  ntt_passes<LOGM, 1, true>(sm, tab, p, C.zero);
  const int P = poly0 + blockIdx.x, k = P >> 3, jc = P & 7;
  uint32_t* dst = keyhat + (((size_t)k * C.L + i) * 8 + jc) * m;
  for (int idx = threadIdx.x; idx < m; idx += blockDim.x) {
    uint32_t v = sm[swz(idx)];
    v = min(v, v - p2); v = min(v, v - p);
    dst[key_pos<LOGM>(idx)] = csub(shoup_mul(v, C.keymul[i], C.keymul_sh[i], p), p);   // 2^32 (Montgomery) and the CRT pre-scaling
  }
}

// Standalone negacyclic product of two full-size operands (seam for DarkIntegers `Polynomial *`).
// grid = batch CTAs; scratch per CTA: [LM][m] u32.
template <int LOGM>
__global__ void __launch_bounds__(Shape<LOGM>::T, 1024 / Shape<LOGM>::T)
polymul_kernel(const __grid_constant__ DevConst C, const uint64_t* __restrict__ a, const uint64_t* __restrict__ b,
               uint64_t* __restrict__ out, const uint2* __restrict__ tw_f, const uint2* __restrict__ tw_i,
               uint32_t* scratch, int batch, int b_bcast) {
  extern __shared__ __align__(16) uint32_t sm[];
  using SH = Shape<LOGM>;
  constexpr int m = SH::M, REM = SH::REM, R = 1 << REM, STR = SH::STR;
  uint32_t* zres = scratch + (size_t)blockIdx.x * SH::LM * m;
  constexpr bool GTW = SH::WIDE;                          // m = 16384: two 64 KiB operands leave no room for a 128 KiB table
  uint2* tab = reinterpret_cast<uint2*>(sm + 2 * m);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + (GTW ? 2 * m : 4 * m));
  uint32_t parity = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  for (int g = blockIdx.x; g < batch; g += gridDim.x) {
#pragma unroll 1
    for (int i = 0; i < SH::LM; ++i) {
      const uint32_t p = C.p[i], p2 = 2 * p;
      if (!GTW && threadIdx.x == 0) stage_table(tab, tw_f + (size_t)i * m, m * 8, bar);
      uint2 wt[R > 1 ? R - 1 : 1];
      top_twiddles<REM>(tw_f + (size_t)i * m, wt);
      {
        // operand coefficients (16 bytes each) one iteration ahead of their reduction
        constexpr int NITER = 2 * STR / SH::T;
        ulonglong2 cur[R], nxt[R];
        auto fetch = [&](ulonglong2 (&dst)[R], int e) {
          const int c = e / STR, idx = e % STR;
          const ulonglong2* src = reinterpret_cast<const ulonglong2*>(c ? b + (b_bcast ? 0 : (size_t)g * m * 2) : a + (size_t)g * m * 2);
#pragma unroll
          for (int k = 0; k < R; ++k) dst[k] = src[idx + k * STR];
        };
        fetch(cur, threadIdx.x);
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
          const int e = threadIdx.x + it * SH::T, c = e / STR, idx = e % STR;
          if (it + 1 < NITER) fetch(nxt, e + SH::T);
          uint32_t x[R];
#pragma unroll
          for (int k = 0; k < R; ++k) x[k] = centred_mod(C, i, cur[k].x, cur[k].y);
          fwd_block<REM>(x, wt, p, p2, C.zero);
#pragma unroll
          for (int k = 0; k < R; ++k) sm[c * m + swz(idx + k * STR)] = x[k];
#pragma unroll
          for (int k = 0; k < R; ++k) cur[k] = nxt[k];
        }
      }
      __syncthreads();
      if (!GTW) { mbar_wait(bar, parity); parity ^= 1; }
      ntt_passes<LOGM, 2, true>(sm, GTW ? tw_f + (size_t)i * m : tab, p, C.zero);
      if (!GTW && threadIdx.x == 0) stage_table(tab, tw_i + (size_t)i * m, m * 8, bar);
      for (int idx = threadIdx.x; idx < m; idx += blockDim.x) {
        uint32_t x = sm[swz(idx)], y = sm[m + swz(idx)];
        x = min(x, x - p2); x = min(x, x - p); y = min(y, y - p2); y = min(y, y - p);
        sm[swz(idx)] = redc((uint64_t)x * y, p, C.pinv_neg[i]);
      }
      __syncthreads();
      if (!GTW) { mbar_wait(bar, parity); parity ^= 1; }
      ntt_passes<LOGM, 1, false>(sm, GTW ? tw_i + (size_t)i * m : tab, p, C.zero);
      top_twiddles<REM>(tw_i + (size_t)i * m, wt);
      for (int idx = threadIdx.x; idx < STR; idx += blockDim.x) {
        uint32_t x[R];
#pragma unroll
        for (int k = 0; k < R; ++k) x[k] = sm[swz(idx + k * STR)];
        inv_block<REM>(x, wt, p, p2, C.zero);
#pragma unroll
        for (int k = 0; k < R; ++k)
          zres[(size_t)i * m + idx + k * STR] = csub(shoup_mul(x[k], C.scale[1][i], C.scale_sh[1][i], p), p);
      }
      __syncthreads();
    }
    for (int idx = threadIdx.x; idx < m; idx += blockDim.x) {
      const u96 z = crt_lift<1, SH::LM, 6 * LOGM + 8>(C, zres + idx, (size_t)m);
      out[((size_t)g * m + idx) * 2] = (uint64_t)z.x0 | ((uint64_t)z.x1 << 32); out[((size_t)g * m + idx) * 2 + 1] = z.x2;
    }
    __syncthreads();
  }
}

// Standalone products with the building blocks of the v4 bootstrap kernel (m >= 4096): one CTA multiplies TWO pairs of
// operands at a time, so a prime is 4 forward + 2 inverse transforms, the same shape as a bootstrap step: top stages
// in registers straight from the operands, warp-local shared-memory passes, the pointwise products fused between the
// stride-1 forward and inverse passes, residues to scratch [LM][2][m], CRT lift at the end.
template <int LOGM>
__global__ void __launch_bounds__(Shape4<LOGM>::T, 1)
polymul_kernel_v4(const __grid_constant__ DevConst C, const uint64_t* __restrict__ a, const uint64_t* __restrict__ b,
                  uint64_t* __restrict__ out, const uint2* __restrict__ tw_f, uint32_t* scratch, int batch, int b_bcast) {
  extern __shared__ __align__(16) uint32_t sm[];
  using S4 = Shape4<LOGM>;
  constexpr int m = S4::M, R0 = S4::R0, LR0 = S4::LR0, T = S4::T, NB = S4::NB, LM = Shape<LOGM>::LM, SB = 6 * LOGM + 8;
  const int tid = threadIdx.x;
  uint2* tab = reinterpret_cast<uint2*>(sm + 4 * m);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 6 * m);
  uint32_t* zres = scratch + (size_t)blockIdx.x * LM * 2 * m;
  uint32_t parity = 0;
  if (tid == 0) {
    mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stage_table(tab, tw_f, m * 8, bar);
  }
  __syncthreads();
  const int st = swz(tid);
  const bool unc = C.pm_uncentred != 0;
  const int npairs = (batch + 1) / 2;
  for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
    const int g0 = 2 * pr, g1 = (2 * pr + 1 < batch) ? 2 * pr + 1 : g0;     // an odd batch repeats its last product
    // operand polynomial of buffer j: (a, b) of the first product, then of the second
    auto operand = [&](int j) {
      return reinterpret_cast<const ulonglong2*>(j & 1 ? b : a) + ((j & 1) && b_bcast ? (size_t)0 : (size_t)(j & 2 ? g1 : g0) * m);
    };
    // Half a polynomial's operand words (8 coefficients per thread, 32 registers) are kept in flight: the second half is
    // requested while the first is reduced, the next polynomial's first half while this one runs its top stages.
    constexpr int H = R0 / 2;
    ulonglong2 raw[H];
    {
      const ulonglong2* s0 = operand(2);
#pragma unroll
      for (int k = 0; k < H; ++k) raw[k] = s0[tid + k * T];
    }
#pragma unroll 1
    for (int i = 0; i < LM; ++i) {
      const uint32_t p = C.p[i], p2 = 2 * p, z = C.zero;
      // ---- operands -> centred residues -> top LR0 stages in registers -> shared memory (order 2,3,0,1, see gate_step_v4)
#pragma unroll 1
      for (int jj = 0; jj < 4; ++jj) {
        const int j = (jj + 2) & 3;
        const ulonglong2* sj = operand(j);
        const ulonglong2* sn = operand((jj + 3) & 3);        // next polynomial (after the fourth: the first one again, next prime)
        if (jj == 2) __syncthreads();                      // previous prime's residue store has left buffers 0,1
        uint32_t x[R0];
#pragma unroll
        for (int k = 0; k < H; ++k) { x[k] = unc ? plain_mod(C, i, raw[k].x, raw[k].y) : centred_mod(C, i, raw[k].x, raw[k].y); raw[k] = sj[tid + (k + H) * T]; }
#pragma unroll
        for (int k = 0; k < H; ++k) { x[k + H] = unc ? plain_mod(C, i, raw[k].x, raw[k].y) : centred_mod(C, i, raw[k].x, raw[k].y); raw[k] = sn[tid + k * T]; }
        fwd_block<LR0>(x, C.topf[i], p, p2, z);
#pragma unroll
        for (int k = 0; k < R0; ++k) sm[j * m + st + k * T] = x[k];
      }
      __syncthreads();
      mbar_wait(bar, parity); parity ^= 1;
      pass8_v4<LOGM, 4, true, 6>(sm, tab, p, z);
      slice_sync<LOGM>();
      pass8_v4<LOGM, 4, true, 3>(sm, tab, p, z);
      __syncwarp();
      // ---- fused: stride-1 forward pass of the four operands, two pointwise products, stride-1 inverse pass
      {
        const uint32_t pinv = C.pinv_neg[i];
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          const int blk = block_of<LOGM>(tid, q), base = 8 * blk;
          const int a0 = swz(base), a1 = a0 ^ 4;
          uint2 w[7];
          block_twiddles<true>(tab, m / 8, blk, p, w);
          uint32_t y[2][8];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t u[8], v[8];
            {
              const uint4 v0 = *reinterpret_cast<const uint4*>(sm + (2 * c) * m + a0), v1 = *reinterpret_cast<const uint4*>(sm + (2 * c) * m + a1);
              u[0] = v0.x; u[1] = v0.y; u[2] = v0.z; u[3] = v0.w; u[4] = v1.x; u[5] = v1.y; u[6] = v1.z; u[7] = v1.w;
              const uint4 w0 = *reinterpret_cast<const uint4*>(sm + (2 * c + 1) * m + a0), w1 = *reinterpret_cast<const uint4*>(sm + (2 * c + 1) * m + a1);
              v[0] = w0.x; v[1] = w0.y; v[2] = w0.z; v[3] = w0.w; v[4] = w1.x; v[5] = w1.y; v[6] = w1.z; v[7] = w1.w;
            }
            fwd_block<3>(u, w, p, p2, z);
            fwd_block<3>(v, w, p, p2, z);
#pragma unroll
            for (int e = 0; e < 8; ++e) {                   // operands corrected to [0, 2p): product < 4 p^2, redc in [0, 2p)
              const uint32_t ue = min(u[e], u[e] - p2), ve = min(v[e], v[e] - p2);
              y[c][e] = redc((uint64_t)ue * ve, p, pinv);
            }
          }
          uint2 wi[7];
          block_twiddles<false>(tab, m / 8, blk, p, wi);
          inv_block<3, true>(y[0], wi, p, p2, z);
          inv_block<3, true>(y[1], wi, p, p2, z);
          *reinterpret_cast<uint4*>(sm + a0) = make_uint4(y[0][0], y[0][1], y[0][2], y[0][3]);
          *reinterpret_cast<uint4*>(sm + a1) = make_uint4(y[0][4], y[0][5], y[0][6], y[0][7]);
          *reinterpret_cast<uint4*>(sm + m + a0) = make_uint4(y[1][0], y[1][1], y[1][2], y[1][3]);
          *reinterpret_cast<uint4*>(sm + m + a1) = make_uint4(y[1][4], y[1][5], y[1][6], y[1][7]);
        }
      }
      __syncwarp();
      pass8_v4<LOGM, 2, false, 3>(sm, tab, p, z);
      slice_sync<LOGM>();
      pass8_v4<LOGM, 2, false, 6>(sm, tab, p, z);
      __syncthreads();                                     // last reader of `tab` for this prime is done
      if (tid == 0) stage_table(tab, tw_f + (size_t)((i + 1 == LM) ? 0 : i + 1) * m, m * 8, bar);
      // ---- top inverse stages in registers + CRT pre-scaling (with the 2^32 of the Montgomery product) + store
      {
        const uint32_t sc = C.scale[1][i], scs = C.scale_sh[1][i], sw = C.scale_w1[i], sws = C.scale_w1_sh[i];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t x[R0];
#pragma unroll
          for (int k = 0; k < R0; ++k) x[k] = sm[c * m + st + k * T];
          inv_block_upper<LR0>(x, C.topi[i], p, p2, z);
#pragma unroll
          for (int k = 0; k < R0 / 2; ++k) {
            const uint32_t s0 = x[k] + x[k + R0 / 2] + z, d0 = x[k] - x[k + R0 / 2] + p2;
            zres[((size_t)i * 2 + c) * m + tid + k * T] = csub(shoup_mul(s0, sc, scs, p), p);
            zres[((size_t)i * 2 + c) * m + tid + (k + R0 / 2) * T] = csub(shoup_mul(d0, sw, sws, p), p);
          }
        }
      }
    }
    // ---- CRT lift (each thread reads only residues it stored itself)
    for (int c = 0; c < (g1 != g0 ? 2 : 1); ++c) {
      uint64_t* o = out + (size_t)(c ? g1 : g0) * m * 2;
      const uint32_t* zr = zres + (size_t)c * m;
      constexpr int NIT = m / T, DP = 2;                    // residues of DP coefficients in flight
      uint32_t yq[DP][LM];
#pragma unroll
      for (int d = 0; d < DP; ++d)
#pragma unroll
        for (int q = 0; q < LM; ++q) yq[d][q] = zr[(size_t)q * 2 * m + tid + d * T];
#pragma unroll 1
      for (int it0 = 0; it0 < NIT; it0 += DP) {
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          const int idx = tid + (it0 + d) * T;
          const u96 zv = crt_lift<1, LM, SB>(C, yq[d], 1);
          if (it0 + d + DP < NIT) {
#pragma unroll
            for (int q = 0; q < LM; ++q) yq[d][q] = zr[(size_t)q * 2 * m + idx + DP * T];
          }
          reinterpret_cast<ulonglong2*>(o)[idx] = make_ulonglong2((uint64_t)zv.x0 | ((uint64_t)zv.x1 << 32), zv.x2);
        }
      }
    }
    __syncthreads();                                       // buffers 2,3 of the next pair are written before its first barrier
  }
  mbar_wait(bar, parity);                                  // the table staged for a prime that never runs
}

// flatten_poly seam (src/utils.jl:253-264): a [m] wide -> out [2][m] wide residues mod Q
__global__ void flatten_kernel(const __grid_constant__ DevConst C, const uint64_t* __restrict__ a,
                               const int64_t* __restrict__ draws, uint64_t* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C.m) return;
  const u128 v = (u128)a[2 * idx] | ((u128)a[2 * idx + 1] << 64);
  int64_t d[2];
  {
    uint64_t dp0 = 0, dp1 = 0;
    const u96 vo = to_offset_form(C, from128(v));
#define SGFHE_FLATTEN_CASE(LOGM_) case LOGM_: if (draws) decompose_off_rand<3 * LOGM_ - 1>(C, vo, draws[2 * idx], draws[2 * idx + 1], dp0, dp1); else decompose_off<3 * LOGM_ - 1>(C, vo, dp0, dp1); break;
    switch (C.logm) { SGFHE_FLATTEN_CASE(9) SGFHE_FLATTEN_CASE(10) SGFHE_FLATTEN_CASE(11) SGFHE_FLATTEN_CASE(12) SGFHE_FLATTEN_CASE(14) default: SGFHE_FLATTEN_CASE(13) }
#undef SGFHE_FLATTEN_CASE
    d[0] = (int64_t)dp0 - (int64_t)C.dig_bias; d[1] = (int64_t)dp1 - (int64_t)C.dig_bias;
  }
  for (int i = 0; i < 2; ++i) {
    const u128 r = d[i] >= 0 ? (u128)d[i] : C.Q - (u128)(-d[i]);
    out[((size_t)i * C.m + idx) * 2] = (uint64_t)r; out[((size_t)i * C.m + idx) * 2 + 1] = (uint64_t)(r >> 64);
  }
}


// ---- Scheme 2 element type (src/rns.jl:8-60): limb-wise arithmetic on (v mod M1, v mod M2), M < 2^48 ------------
// Barrett: mu = floor(2^96 / M), 2^32 < M < 2^48;  a, b < M
__device__ __forceinline__ uint64_t mulmod48(uint64_t a, uint64_t b, uint64_t M, uint64_t mu) {
  const uint64_t lo = a * b, hi = __umul64hi(a, b);
  const uint64_t q = __umul64hi((hi << 32) | (lo >> 32), mu);          // in [ab/M - 2, ab/M]
  uint64_t r = lo - q * M;                                             // exact modulo 2^64, true value in [0, 3M)
  r = r >= M ? r - M : r;
  r = r >= M ? r - M : r;
  return r;
}
__global__ void rns2_kernel(int op, size_t count, const uint64_t* __restrict__ a1, const uint64_t* __restrict__ a2,
                            const uint64_t* __restrict__ b1, const uint64_t* __restrict__ b2, uint64_t M1, uint64_t M2,
                            uint64_t i1, uint64_t i2, uint64_t* __restrict__ o1, uint64_t* __restrict__ o2) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const uint64_t x1 = a1[i], x2 = a2[i], y1 = b1[i], y2 = b2[i];
    uint64_t r1, r2;
    if (op == 0) { r1 = mulmod48(x1, y1, M1, i1); r2 = mulmod48(x2, y2, M2, i2); }                          // rns.jl:51-52
    else if (op == 1) { r1 = x1 + y1; r1 = r1 >= M1 ? r1 - M1 : r1; r2 = x2 + y2; r2 = r2 >= M2 ? r2 - M2 : r2; }   // rns.jl:55-56
    else { r1 = x1 >= y1 ? x1 - y1 : x1 + M1 - y1; r2 = x2 >= y2 ? x2 - y2 : x2 + M2 - y2; }              // rns.jl:59-60
    o1[i] = r1; o2[i] = r2;
  }
}

// split_ciphertext (src/fhe.jl:287-290): one CTA row per output LWE; i = blockIdx.x % n + 1 is the reference's 1-based index.
// extract(a, i, n) (src/fhe.jl:237-244): element k (1-based) is a[i-k+1] while k <= i, else -a[N+i+1-k].
__global__ void split_kernel(int n, int N, uint64_t r, const uint64_t* __restrict__ a, const uint64_t* __restrict__ b,
                             uint64_t* __restrict__ lwes) {
  const int ct = blockIdx.x / n, i = blockIdx.x % n + 1;
  const uint64_t* ap = a + (size_t)ct * N;
  uint64_t* out = lwes + (size_t)blockIdx.x * (n + 1);
  for (int k = threadIdx.x + 1; k <= n; k += blockDim.x) {
    uint64_t v;
    if (k <= i) v = ap[i - k];                                   // a.coeffs[i-k+1]
    else { const uint64_t t = ap[N + i - k]; v = t ? r - t : 0; }     // -a.coeffs[N+i+1-k]
    out[k - 1] = v;
  }
  if (threadIdx.x == 0) out[n] = b[(size_t)ct * N + i - 1];
}

// decrypt(key, ::EncryptedBit) (src/fhe.jl:504-507): one warp per LWE
__global__ void decrypt_kernel(int n, int count, uint64_t rmask, uint64_t Dr, const uint64_t* __restrict__ lwes,
                               const uint8_t* __restrict__ sk, uint8_t* __restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= count) return;
  const uint64_t* l = lwes + (size_t)w * (n + 1);
  uint64_t acc = 0;
  for (int k = lane; k < n; k += 32) acc += sk[k] ? l[k] : 0;
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[w] = (uint8_t)((((l[n] - acc) + Dr / 2) & rmask) / Dr);
}

// BootstrapKey rows from the products a_j * ext_key (src/fhe.jl:192-198): key[i][j] = (a_j, a_j s + e_j) + s_i G[j,:],
// G = [1 0; B 0; 0 1; 0 B] (src/fhe.jl:119-122).  a, prod: [rows][4][m][2] wide; e: [rows][4][m]; out: [rows][4][2][m][2].
__global__ void keygen_assemble_kernel(const __grid_constant__ DevConst C, const uint64_t* __restrict__ a,
                                       const uint64_t* __restrict__ prod, const int64_t* __restrict__ e,
                                       const uint8_t* __restrict__ sk, int row0, int rows, uint64_t* __restrict__ out) {
  const int m = C.m;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * 4 * m) return;
  const int k = (int)(idx % m), j = (int)((idx / m) & 3), i = (int)(idx / ((size_t)4 * m));
  u128 av = (u128)a[2 * idx] | ((u128)a[2 * idx + 1] << 64);
  u128 bv = (u128)prod[2 * idx] | ((u128)prod[2 * idx + 1] << 64);
  const int64_t ev = e[idx];
  bv = ev >= 0 ? addmodQ(bv, (u128)ev, C.Q) : submodQ(bv, (u128)(-ev), C.Q);          // b_j = a_j * s + e_j   (src/fhe.jl:195)
  if (k == 0 && sk[row0 + i]) {                                                      // + s_i G                 (src/fhe.jl:196)
    const u128 g = (j & 1) ? C.B : (u128)1;
    if (j < 2) av = addmodQ(av, g, C.Q); else bv = addmodQ(bv, g, C.Q);
  }
  uint64_t* o = out + (((size_t)i * 4 + j) * 2 * m + k) * 2;
  o[0] = (uint64_t)av; o[1] = (uint64_t)(av >> 64);
  o[(size_t)2 * m] = (uint64_t)bv; o[(size_t)2 * m + 1] = (uint64_t)(bv >> 64);
}

// pack_encrypted_bits (src/fhe.jl:675-678): as[i].coeffs[j] = new_lwes[j].a[i] for j < n, 0 above (resize to m)
__global__ void pack_transpose_kernel(int n, int m, const uint64_t* __restrict__ lwes /* [n][n+1][2] */, uint64_t* __restrict__ polys /* [n][m][2] */) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * m) return;
  const int j = (int)(idx % m), i = (int)(idx / m);
  uint64_t lo = 0, hi = 0;
  if (j < n) { lo = lwes[((size_t)j * (n + 1) + i) * 2]; hi = lwes[((size_t)j * (n + 1) + i) * 2 + 1]; }
  polys[2 * idx] = lo; polys[2 * idx + 1] = hi;
}

// pack_encrypted_bits tail (src/fhe.jl:686-693): w~ = sum_i w_i, v~ = sum_i v_i; w = ModRed(-w~), v = ModRed(b - v~) with
// b.coeffs[j] = new_lwes[j].b for j < n, 0 above.  wv: [n][2][m][2] wide; out_w, out_v: [m] over Z_r.
__global__ void pack_tail_kernel(const __grid_constant__ DevConst C, const uint64_t* __restrict__ wv,
                                 const uint64_t* __restrict__ lwes, uint64_t* __restrict__ out_w, uint64_t* __restrict__ out_v) {
  const int m = C.m, n = C.n;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * m) return;
  const int c = idx / m, k = idx % m;
  u128 sum = 0;                                               // n values below Q < 2^93: no overflow
  for (int i = 0; i < n; ++i) {
    const uint64_t* t = wv + (((size_t)i * 2 + c) * m + k) * 2;
    sum += (u128)t[0] | ((u128)t[1] << 64);
  }
  sum %= C.Q;
  if (c == 0) out_w[k] = modred(C, negmodQ(sum, C.Q));
  else {
    u128 b = 0;
    if (k < n) { const uint64_t* t = lwes + ((size_t)k * (n + 1) + n) * 2; b = (u128)t[0] | ((u128)t[1] << 64); }
    out_v[k] = modred(C, submodQ(b, sum, C.Q));
  }
}

// test seam: the draws DrawSrc makes on the device for (seed, gate, steps step0..), as the host would have to supply them
__global__ void device_draws_kernel(const __grid_constant__ DevConst C, uint64_t seed, uint64_t gate, int step0, int steps,
                                    int64_t* __restrict__ out) {
  const int m = C.m;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)steps * 2 * m) return;
  const int j = (int)(idx % m), c = (int)((idx / m) & 1), k = (int)(idx / ((size_t)2 * m));
  DrawSrc d; d.ptr = nullptr; d.seed = seed; d.gate = gate; d.step = (uint32_t)(step0 + k);
  int64_t x0, x1;
  get_draws(C, d, c, j, m, x0, x1);
  out[2 * idx] = x0; out[2 * idx + 1] = x1;
}

// wide [2][m][2] -> accumulator scratch (SoA limbs) and back
__global__ void acc_load_kernel(const __grid_constant__ DevConst C, const uint64_t* __restrict__ ab, uint32_t* acc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x, m = C.m;
  if (e >= 2 * m) return;
  const u128 v = (u128)ab[2 * e] | ((u128)ab[2 * e + 1] << 64);
  st96(acc + (e / m) * 3 * m, m, e % m, to_offset_form(C, from128(v)));
}

// =========================================================================================================
// Host side
// =========================================================================================================

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct sgfhe_ctx {
  int device = 0;
  HostParams hp;
  DevConst dc;
  int num_sms = 0, threads = 0, boot_threads = 0, boot_threads_gate = 0, max_ctas = 0;
  bool use_v4 = false;
  bool head_f64 = true;                            // v4 gate kernel: digits as doubles, first forward stage on the FP64 pipe
  bool use_v5 = false;                             // m = 8192: two gates per SM (bootstrap_kernel_v5)
  size_t smem_bytes = 0, smem_gate_v4 = 0, scratch_stride = 0, zres_stride = 0;
  int persist_l2 = 0;                              // pin accumulator + digit scratch in L2 (access policy window)
  uint2* d_tw_f = nullptr; uint2* d_tw_i = nullptr;
  uint32_t* d_keyhat = nullptr; int key_rows = 0; size_t keyhat_capacity_rows = 0;
  uint8_t* d_scratch = nullptr; int scratch_ctas = 0;
  uint32_t* d_pm_scratch = nullptr; int pm_ctas = 0;
  uint64_t* d_io = nullptr; size_t io_capacity = 0;
  int* d_counter = nullptr;                        // work counter of the persistent gate kernels
  uint64_t key_token = 0;                          // identifies the key content of d_keyhat (0 = none); see sgfhe_bkey_token
  uint8_t* d_arena = nullptr; size_t arena_bytes = 0;   // grow-only staging for the host-buffer entry points
};
static std::atomic<uint64_t> g_token{0};
static void new_key_token(sgfhe_ctx* c) { c->key_token = ++g_token; }

static void to_limbs(u128 v, uint32_t out[3]) { out[0] = (uint32_t)v; out[1] = (uint32_t)(v >> 32); out[2] = (uint32_t)(v >> 64); }

static int choose_primes(double need_bits, const std::vector<uint32_t>& primes) {
  double have = 0;
  for (size_t k = 0; k < primes.size(); ++k) {
    have += log2((double)primes[k]);
    if (have >= need_bits) return (int)k + 1;
  }
  return -1;
}

static int build_consts(const HostParams& hp, DevConst* dc, std::vector<uint2>* twf, std::vector<uint2>* twi) {
  memset(dc, 0, sizeof *dc);
  bool dc_uncentred_ok = false;
  const std::vector<uint32_t> primes = h_rns_primes(MAXP);
  const double lq = log2((double)hp.Q), lB = log2((double)hp.B), lm = hp.logm;
  // |sum of 4 m products digit * centred key| < 4 m (2B+1) Q/2 ; +1 sign bit, +4 bits for the CRT rounding margin
  const int L = choose_primes(2 + lm + (lB + 1.001) + (lq - 1) + 1 + 4, primes);
  const int LM = choose_primes(lm + 2 * (lq - 1) + 1 + 4, primes);
  if (L < 0 || LM < 0) return -1;
  {
    double have = 0;
    for (int k = 0; k < LM; ++k) have += log2((double)primes[k]);
    dc_uncentred_ok = have >= lm + 2 * lq + 5;             // m Q^2 < P / 32 without centring the operands
  }
  dc->pm_uncentred = dc_uncentred_ok ? 1 : 0;
  dc->n = hp.n; dc->m = hp.m; dc->logm = hp.logm; dc->logr = hp.logr; dc->kB = hp.kB; dc->L = L; dc->LM = LM;
  dc->sbits = h_bits(hp.Q) - 1;
  dc->Q = hp.Q; dc->DQ = hp.DQ; dc->B = hp.B;
  dc->Qhalf[0] = (uint64_t)(hp.Q >> 1); dc->Qhalf[1] = (uint64_t)(hp.Q >> 65);
  const u128 s = hp.B / 2 - 1;                       // B is even (src/utils.jl:162-166)
  dc->s = (uint64_t)s;
  dc->offs = h_mulmod(s, (1 + hp.B) % hp.Q, hp.Q);
  to_limbs(hp.Q, dc->Ql); to_limbs(dc->offs, dc->offl);
  if (dc->sbits != 6 * hp.logm + 8) return -1;                             // compile-time shift of barrett96
  dc->barrett_inv = ldexp(1.0, dc->sbits - 16) / (double)hp.Q * (1.0 - ldexp(1.0, -40));
  dc->inv35 = nextafter(nextafter(1.0 / 35.0, 1.0), 1.0);
  dc->dig_bias = (uint64_t)1 << (hp.logm >= 14 ? 48 : 46);                // |digit| <= 2B with the largest draws: 2^44.1 at n = 1024, 2^47.2 at n = 2048
  dc->s46 = dc->dig_bias - dc->s;
  dc->xmax = (uint64_t)(hp.B / 2 * 3);
  {
    const u128 K = ((u128)L << 30) + L + 1, KQ = K * hp.Q;                 // < 2^34 Q < 2^124
    dc->KQ[0] = (uint32_t)KQ; dc->KQ[1] = (uint32_t)(KQ >> 32); dc->KQ[2] = (uint32_t)(KQ >> 64); dc->KQ[3] = (uint32_t)(KQ >> 96);
  }
  const int NP = L > LM ? L : LM;
  for (int i = 0; i < NP; ++i) {
    const uint32_t p = primes[i];
    dc->p[i] = p;
    uint32_t inv = p; for (int k = 0; k < 5; ++k) inv *= 2 - p * inv;      // p^-1 mod 2^32
    dc->pinv_neg[i] = 0u - inv;
    dc->dig_mu[i] = (uint32_t)(((uint64_t)1 << 50) / p);
    dc->vinv[i] = (uint32_t)(((uint64_t)1 << 61) / p);
    dc->vk[i] = (uint32_t)((((uint64_t)1 << 44) + p / 2) / p);
    dc->r32[i] = (uint32_t)(((uint64_t)1 << 32) % p);
    dc->r64[i] = (uint32_t)((((u128)1) << 64) % p);
    dc->qmodp[i] = (uint32_t)(hp.Q % p);
    dc->r32_sh[i] = (uint32_t)(((uint64_t)dc->r32[i] << 32) / p);
    dc->r64_sh[i] = (uint32_t)(((uint64_t)dc->r64[i] << 32) / p);
    if (5ull * ((1u << 30) - p) >= (1u << 30)) return -1;                  // centred_mod's reduction of the low word
    dc->mont[i] = dc->r32[i];
    dc->mont_sh[i] = (uint32_t)(((uint64_t)dc->mont[i] << 32) / p);
    dc->dig_negc[i] = p - (uint32_t)(dc->dig_bias % p) - 4u * p;            // wraps mod 2^32 on purpose
    dc->hp_p[i] = (double)p; dc->hp_pinv[i] = 1.0 / (double)p; dc->hp_c[i] = 6755399441055744.0 + 2.0 * (double)p;
  }
  for (int basis = 0; basis < 2; ++basis) {
    const int K = basis == 0 ? L : LM;
    u128 Pm = 1 % hp.Q;
    for (int i = 0; i < K; ++i) Pm = h_mulmod(Pm, primes[i], hp.Q);
    // basis 0 (bootstrap): the sums represent -z, so its constants are negated: (+P) mod Q and -(P/p_i) mod Q
    to_limbs(basis == 0 ? Pm : (hp.Q - Pm) % hp.Q, dc->negP[basis]);
    for (int i = 0; i < K; ++i) {
      const uint32_t p = primes[i];
      u128 c = 1; uint64_t cp = 1;
      for (int j = 0; j < K; ++j) if (j != i) { c = h_mulmod(c, primes[j], hp.Q); cp = h_mulmod64(cp, primes[j] % p, p); }
      to_limbs(basis == 0 ? (hp.Q - c) % hp.Q : c, dc->crt_c[basis][i]);
      uint64_t sc = h_mulmod64(h_powmod64(cp, p - 2, p), h_powmod64((uint64_t)hp.m % p, p - 2, p), p);
      if (basis == 1) sc = h_mulmod64(sc, dc->r32[i], p);
      else sc = (p - sc) % p;                              // digit_mod yields the negated digits
      dc->scale[basis][i] = (uint32_t)sc;
      dc->scale_sh[basis][i] = (uint32_t)((sc << 32) / p);
      if (basis == 0) {                                   // psi^(-m/2) = (psi^-1)^(m/2)
        const uint64_t psi_inv = h_powmod64(h_root_2m(p, hp.m), p - 2, p);
        const uint64_t lw = h_powmod64(psi_inv, hp.m / 2, p), km = h_mulmod64(sc, dc->r32[i], p);
        dc->lastw[i] = (uint32_t)lw; dc->lastw_sh[i] = (uint32_t)((lw << 32) / p);
        dc->keymul[i] = (uint32_t)km; dc->keymul_sh[i] = (uint32_t)((km << 32) / p);
      } else {
        const uint64_t psi_inv = h_powmod64(h_root_2m(p, hp.m), p - 2, p);
        const uint64_t sw = h_mulmod64(sc, h_powmod64(psi_inv, hp.m / 2, p), p);
        dc->scale_w1[i] = (uint32_t)sw; dc->scale_w1_sh[i] = (uint32_t)((sw << 32) / p);
      }
    }
  }
  const int m = hp.m;
  twf->assign((size_t)MAXP * m, make_uint2(0, 0)); twi->assign((size_t)MAXP * m, make_uint2(0, 0));
  std::vector<uint64_t> pf(m), pi(m);
  for (int i = 0; i < NP; ++i) {
    const uint64_t p = primes[i];
    const uint64_t psi = h_root_2m((uint32_t)p, m), psi_inv = h_powmod64(psi, p - 2, p);
    pf[0] = pi[0] = 1;
    for (int k = 1; k < m; ++k) { pf[k] = pf[k - 1] * psi % p; pi[k] = pi[k - 1] * psi_inv % p; }
    for (int k = 0; k < m; ++k) {
      const int r = h_bitrev(k, hp.logm);
      (*twf)[(size_t)i * m + k] = make_uint2((uint32_t)pf[r], (uint32_t)((pf[r] << 32) / p));
      (*twi)[(size_t)i * m + k] = make_uint2((uint32_t)pi[r], (uint32_t)((pi[r] << 32) / p));
    }
    // top-stage twiddles of the v4 kernels (constant bank): tw[k] and its inverse, the negated mirrored forward entry
    const int r0 = hp.logm % 3 == 1 ? 16 : 8;
    for (int k = 1; k < r0 && hp.logm >= 12; ++k) {
      int lvl = 1; while (2 * lvl <= k) lvl *= 2;
      const uint2 f = (*twf)[(size_t)i * m + k], mir = (*twf)[(size_t)i * m + lvl + ((k - lvl) ^ (lvl - 1))];
      dc->topf[i][k - 1] = f;
      dc->topi[i][k - 1] = make_uint2((uint32_t)p - mir.x, ~mir.y);
    }
    for (int h = 0; h < 2 && hp.logm >= 12; ++h) {       // v5: the radix-8 that follows the first stage inside half h of a radix-16 group
      const int idx[7] = {2 + h, 4 + 2 * h, 5 + 2 * h, 8 + 4 * h, 9 + 4 * h, 10 + 4 * h, 11 + 4 * h};
      for (int k = 0; k < 7; ++k) { dc->topf_h[i][h][k] = (*twf)[(size_t)i * m + idx[k]]; dc->topi_h[i][h][k] = (*twi)[(size_t)i * m + idx[k]]; }
    }
    dc->hp_w[i] = (double)(*twf)[(size_t)i * m + 1].x;                       // stage-1 twiddle psi^(m/2)
    dc->hp_wp[i] = dc->hp_w[i] / (double)p;
  }
  return 0;
}


// ---- LOGM dispatch: every kernel is compiled for m = 512 .. 16384 (v4 kernels: 4096, 8192) -------------------------------------------
#define SGFHE_DISPATCH(logm, STMT)                         \
  switch (logm) {                                          \
    case 9:  { constexpr int LOGM_ = 9;  STMT; } break;    \
    case 10: { constexpr int LOGM_ = 10; STMT; } break;    \
    case 11: { constexpr int LOGM_ = 11; STMT; } break;    \
    case 12: { constexpr int LOGM_ = 12; STMT; } break;    \
    case 14: { constexpr int LOGM_ = 14; STMT; } break;    \
    default: { constexpr int LOGM_ = 13; STMT; } break;    \
  }

#define SGFHE_DISPATCH_V4(logm, STMT)                      \
  switch (logm) {                                          \
    case 12: { constexpr int LOGM_ = 12; STMT; } break;    \
    default: { constexpr int LOGM_ = 13; STMT; } break;    \
  }

static cudaError_t configure_kernels(sgfhe_ctx* c, int* occ) {
  cudaError_t e = cudaSuccess;
  c->use_v4 = c->hp.logm >= 12 && c->hp.logm <= 13 && !getenv("SGFHE_FORCE_V3");
  c->head_f64 = !getenv("SGFHE_HEAD_INT");           // A/B knob: the all-integer head of round 1
  // experimental, off by default: two gates per SM.  Measured slower than v4 (231 k against 201 k cycles per gate-step): 296
  // gates in flight double the scratch working set to 246 MB, the L2 hit rate falls from 78 % to 48 % and the DRAM traffic per
  // gate rises from 1.0 to 2.5 GB (profiles/ncu_r02_v5_summary.txt); kept selectable for that comparison and parity-tested
  c->use_v5 = c->use_v4 && c->hp.logm == 13 && c->head_f64 && getenv("SGFHE_V5");
  if (c->use_v4) {
    SGFHE_DISPATCH_V4(c->hp.logm, {
      c->boot_threads = Shape4<LOGM_>::T;                  // standalone products
      c->boot_threads_gate = GateShape4<LOGM_>::T;
      e = cudaFuncSetAttribute(polymul_kernel_v4<LOGM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(bootstrap_kernel_v4<LOGM_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_gate_v4);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(bootstrap_kernel_v4<LOGM_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_gate_v4);
      if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, bootstrap_kernel_v4<LOGM_, true>, c->boot_threads_gate, c->smem_gate_v4);
    });
    if (e != cudaSuccess) return e;
    if (c->use_v5) {
      e = cudaFuncSetAttribute(bootstrap_kernel_v5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemV5);
      if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, bootstrap_kernel_v5, Shape5::T, kSmemV5);
      if (e != cudaSuccess) return e;
      c->boot_threads_gate = Shape5::T;
    }
  }
  SGFHE_DISPATCH(c->hp.logm, {
    c->threads = Shape<LOGM_>::T;
    if (!c->use_v4) c->boot_threads = c->threads;
    e = cudaFuncSetAttribute(bootstrap_kernel<LOGM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(key_transform_kernel<LOGM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(c->smem_bytes, (size_t)c->hp.m * 12 + 16));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(polymul_kernel<LOGM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes);
    if (e == cudaSuccess && !c->use_v4) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, bootstrap_kernel<LOGM_>, c->threads, c->smem_bytes);
  });
  return e;
}
static void launch_bootstrap(const sgfhe_ctx* c, int grid, cudaStream_t st, const GateArgs& A) {
  if (c->use_v5) {
    bootstrap_kernel_v5<<<grid, Shape5::T, kSmemV5, st>>>(c->dc, A);
    ++g_launches;
    return;
  }
  if (c->use_v4) {
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(c->boot_threads_gate); cfg.dynamicSmemBytes = c->smem_gate_v4; cfg.stream = st;
    cudaLaunchAttribute attr[1]; int nattr = 0;
    if (c->persist_l2) {                               // accumulator + digits of the resident CTAs stay in L2 across steps
      attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
      attr[0].val.accessPolicyWindow.base_ptr = A.scratch;
      attr[0].val.accessPolicyWindow.num_bytes = (size_t)grid * A.scratch_stride;
      attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
      attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      nattr = 1;
    }
    cfg.attrs = attr; cfg.numAttrs = nattr;
    if (c->head_f64) { SGFHE_DISPATCH_V4(c->hp.logm, (cudaLaunchKernelEx(&cfg, bootstrap_kernel_v4<LOGM_, true>, c->dc, A))); }
    else { SGFHE_DISPATCH_V4(c->hp.logm, (cudaLaunchKernelEx(&cfg, bootstrap_kernel_v4<LOGM_, false>, c->dc, A))); }
  } else {
    SGFHE_DISPATCH(c->hp.logm, (bootstrap_kernel<LOGM_><<<grid, c->threads, c->smem_bytes, st>>>(c->dc, A)));
  }
  ++g_launches;
}
static void launch_key_transform(const sgfhe_ctx* c, int npolys, const uint64_t* d_coef, uint32_t* d_keyhat, int poly0) {
  SGFHE_DISPATCH(c->hp.logm, (key_transform_kernel<LOGM_><<<dim3(npolys, c->dc.L), c->threads, (size_t)c->hp.m * 12 + 16>>>(
                                  c->dc, d_coef, d_keyhat, c->d_tw_f, poly0)));
  ++g_launches;
}
static void launch_polymul(const sgfhe_ctx* c, int grid, cudaStream_t st, const uint64_t* a, const uint64_t* b, uint64_t* out,
                           int batch, int b_bcast = 0) {
  if (c->use_v4 && !getenv("SGFHE_POLYMUL_V3")) {
    SGFHE_DISPATCH_V4(c->hp.logm, (polymul_kernel_v4<LOGM_><<<grid, c->boot_threads, c->smem_bytes, st>>>(
                                       c->dc, a, b, out, c->d_tw_f, c->d_pm_scratch, batch, b_bcast)));
    ++g_launches;
    return;
  }
  SGFHE_DISPATCH(c->hp.logm, (polymul_kernel<LOGM_><<<grid, c->threads, (size_t)c->hp.m * (Shape<LOGM_>::WIDE ? 8 : 16) + 16, st>>>(
                                  c->dc, a, b, out, c->d_tw_f, c->d_tw_i, c->d_pm_scratch, batch, b_bcast)));
  ++g_launches;
}

extern "C" const char* sgfhe_last_error(void) { return g_err.c_str(); }
// shared with scheme2.cu (not part of the public header)
extern "C" void sgfhe_set_error_(const char* msg) { g_err = msg; }
extern "C" void sgfhe_count_launch_(void) { ++g_launches; }
extern "C" uint64_t sgfhe_launch_count(void) { return g_launches.load(); }

static void fill_params(const HostParams& hp, int L, sgfhe_params* out) {
  out->n = hp.n; out->t = hp.t; out->m = hp.m; out->rns_primes = L;
  out->r = hp.r; out->q = hp.q; out->Dr = hp.Dr; out->Dq = hp.Dq;
  out->Q[0] = (uint64_t)hp.Q; out->Q[1] = (uint64_t)(hp.Q >> 64);
  out->B[0] = (uint64_t)hp.B; out->B[1] = (uint64_t)(hp.B >> 64);
  out->DQ_tilde[0] = (uint64_t)hp.DQ; out->DQ_tilde[1] = (uint64_t)(hp.DQ >> 64);
}

extern "C" int sgfhe_params_derive(int32_t n, sgfhe_params* out) {
  if (!out) return fail(SGFHE_ERR_ARG, "out is NULL");
  HostParams hp;
  const int rc = h_params(n, &hp);
  if (rc == -1) return fail(SGFHE_ERR_ARG, "n must be a power of two >= 64 (src/fhe.jl:45-46)");
  if (rc) return fail(SGFHE_ERR_MODULUS, "could not find a modulus / n is too large (src/utils.jl:26, src/fhe.jl:77)");
  fill_params(hp, 0, out);
  return SGFHE_OK;
}

extern "C" int sgfhe_ctx_create(int32_t n, int32_t device, sgfhe_ctx** out) {
  if (!out) return fail(SGFHE_ERR_ARG, "out is NULL");
  *out = nullptr;
  HostParams hp;
  const int rc = h_params(n, &hp);
  if (rc == -1) return fail(SGFHE_ERR_ARG, "n must be a power of two >= 64 (src/fhe.jl:45-46)");
  if (rc) return fail(SGFHE_ERR_MODULUS, "could not find a modulus / n is too large (src/utils.jl:26, src/fhe.jl:77)");
  if (n > 2048) return fail(SGFHE_ERR_ARG, "this backend supports n <= 2048 (Q < 2^96; src/fhe.jl:71-78 itself stops at 128 bits)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(SGFHE_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(SGFHE_ERR_ARG, "bad device ordinal");
  CK(cudaSetDevice(device));
  sgfhe_ctx* c = new sgfhe_ctx();
  c->device = device; c->hp = hp;
  std::vector<uint2> twf, twi;
  if (build_consts(hp, &c->dc, &twf, &twi)) { delete c; return fail(SGFHE_ERR_MODULUS, "RNS basis too small for these parameters"); }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->num_sms = prop.multiProcessorCount;
  if (getenv("SGFHE_L2_PERSIST")) {
    fprintf(stderr, "[sgfhe] L2 %d MB, persisting max %d MB, window max %d MB\n", prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20, prop.accessPolicyMaxWindowSize >> 20);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize);
  }
  const int m = hp.m;
  c->smem_bytes = (size_t)(hp.logm >= 14 ? 10 : 24) * m + 16 + 1024;   // 4 NTT buffers + staged twiddle table + mbarrier, work-counter word (+ spare); m = 16384: two buffers, twiddles from global
  c->smem_gate_v4 = (size_t)24 * m + 128 + (kKeyTma ? kKeyStageBytes : 1024);          // + per-warp key staging buffers and their mbarriers
  int occ = 0;
  CK(configure_kernels(c, &occ));
  {
    int wantL = 0, wantLM = 0;
    SGFHE_DISPATCH(hp.logm, { wantL = Shape<LOGM_>::L; wantLM = Shape<LOGM_>::LM; });
    if (wantL != c->dc.L || wantLM != c->dc.LM) { delete c; return fail(SGFHE_ERR_MODULUS, "compiled RNS basis size does not match Params"); }
  }
  if (occ < 1) { delete c; return fail(SGFHE_ERR_CUDA, "bootstrap kernel does not fit on an SM"); }
  c->max_ctas = occ * c->num_sms;
  c->scratch_stride = (scratch_bytes(m) + 255) & ~(size_t)255;
  c->zres_stride = (zres_bytes(m, c->dc.L) + 255) & ~(size_t)255;
  c->persist_l2 = getenv("SGFHE_L2_PERSIST") ? atoi(getenv("SGFHE_L2_PERSIST")) : 0;
  CK(cudaMalloc(&c->d_tw_f, twf.size() * sizeof(uint2)));
  CK(cudaMalloc(&c->d_tw_i, twi.size() * sizeof(uint2)));
  CK(cudaMemcpy(c->d_tw_f, twf.data(), twf.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_tw_i, twi.data(), twi.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  *out = c;
  return SGFHE_OK;
}

extern "C" int sgfhe_ctx_destroy(sgfhe_ctx* c) {
  if (!c) return SGFHE_OK;
  cudaSetDevice(c->device);
  cudaFree(c->d_tw_f); cudaFree(c->d_tw_i); cudaFree(c->d_keyhat); cudaFree(c->d_scratch); cudaFree(c->d_pm_scratch); cudaFree(c->d_io); cudaFree(c->d_counter); cudaFree(c->d_arena);
  delete c;
  return SGFHE_OK;
}

extern "C" int sgfhe_params_get(const sgfhe_ctx* c, sgfhe_params* out) {
  if (!c || !out) return fail(SGFHE_ERR_ARG, "NULL argument");
  fill_params(c->hp, c->dc.L, out);
  return SGFHE_OK;
}

static size_t keyhat_row_words(const sgfhe_ctx* c) { return (size_t)c->dc.L * 8 * c->hp.m; }

static int ensure_keyhat(sgfhe_ctx* c, int rows) {
  if ((size_t)rows <= c->keyhat_capacity_rows) return SGFHE_OK;
  if (c->d_keyhat) { cudaFree(c->d_keyhat); c->d_keyhat = nullptr; c->keyhat_capacity_rows = 0; c->key_rows = 0; }
  if (cudaMalloc(&c->d_keyhat, (size_t)rows * keyhat_row_words(c) * sizeof(uint32_t)) != cudaSuccess)
    return fail(SGFHE_ERR_NOMEM, "cudaMalloc of the pre-transformed key failed");
  c->keyhat_capacity_rows = rows;
  return SGFHE_OK;
}

static int ensure_scratch(sgfhe_ctx* c, int ctas) {
  if (ctas <= c->scratch_ctas) return SGFHE_OK;
  if (c->d_scratch) { cudaFree(c->d_scratch); c->d_scratch = nullptr; c->scratch_ctas = 0; }
  if (cudaMalloc(&c->d_scratch, (size_t)ctas * (c->scratch_stride + c->zres_stride)) != cudaSuccess)
    return fail(SGFHE_ERR_NOMEM, "cudaMalloc of the gate scratch failed");
  c->scratch_ctas = ctas;
  return SGFHE_OK;
}

// grow-only device staging shared by the host-buffer entry points (no cudaMalloc / cudaFree per call)
static int ensure_arena(sgfhe_ctx* c, size_t bytes) {
  if (bytes <= c->arena_bytes) return SGFHE_OK;
  cudaFree(c->d_arena); c->d_arena = nullptr; c->arena_bytes = 0;
  bytes = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
  if (cudaMalloc(&c->d_arena, bytes) != cudaSuccess) return fail(SGFHE_ERR_NOMEM, "cudaMalloc of the staging arena failed");
  c->arena_bytes = bytes;
  return SGFHE_OK;
}

// transform `npolys` coefficient-form polys (host) into keyhat starting at poly index poly0
static int transform_polys(sgfhe_ctx* c, const uint64_t* h_coef, int poly0, int npolys, uint32_t* d_keyhat) {
  const int m = c->hp.m;
  const int chunk = 512;                                   // polys per staging chunk
  uint64_t* d_stage = nullptr;
  const size_t poly_bytes = (size_t)m * 2 * sizeof(uint64_t);
  if (cudaMalloc(&d_stage, (size_t)(npolys < chunk ? npolys : chunk) * poly_bytes) != cudaSuccess)
    return fail(SGFHE_ERR_NOMEM, "cudaMalloc of the key staging buffer failed");
  for (int done = 0; done < npolys; done += chunk) {
    const int cnt = npolys - done < chunk ? npolys - done : chunk;
    cudaError_t e = cudaMemcpy(d_stage, h_coef + (size_t)done * m * 2, (size_t)cnt * poly_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      launch_key_transform(c, cnt, d_stage, d_keyhat, poly0 + done);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d_stage); return fail(SGFHE_ERR_CUDA, std::string("key transform: ") + cudaGetErrorString(e)); }
  }
  cudaFree(d_stage);
  return SGFHE_OK;
}

extern "C" int sgfhe_bkey_upload(sgfhe_ctx* c, const uint64_t* key, int32_t rows) {
  if (!c || !key) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (rows < 1 || rows > c->hp.n) return fail(SGFHE_ERR_ARG, "rows must be in [1, n]");
  CK(cudaSetDevice(c->device));
  int rc = ensure_keyhat(c, rows); if (rc) return rc;
  c->key_rows = 0; c->key_token = 0;                   // a failed transform leaves no usable key behind
  rc = transform_polys(c, key, 0, rows * 8, c->d_keyhat); if (rc) return rc;
  c->key_rows = rows; new_key_token(c);
  return SGFHE_OK;
}

extern "C" int sgfhe_bkey_token(const sgfhe_ctx* c, uint64_t* token, int32_t* rows) {
  if (!c || !token) return fail(SGFHE_ERR_ARG, "NULL argument");
  *token = c->key_rows > 0 ? c->key_token : 0;
  if (rows) *rows = c->key_rows;
  return SGFHE_OK;
}

extern "C" int sgfhe_bkey_device_buffer(sgfhe_ctx* c, int32_t rows, void** d_ptr, uint64_t* bytes) {
  if (!c || !d_ptr || !bytes) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (rows < 1 || rows > c->hp.n) return fail(SGFHE_ERR_ARG, "rows must be in [1, n]");
  CK(cudaSetDevice(c->device));
  int rc = ensure_keyhat(c, rows); if (rc) return rc;
  *d_ptr = c->d_keyhat; *bytes = (uint64_t)rows * keyhat_row_words(c) * sizeof(uint32_t);
  return SGFHE_OK;                                     // the caller may read (broadcast root) or overwrite + sgfhe_bkey_adopt
}

extern "C" int sgfhe_bkey_adopt(sgfhe_ctx* c, int32_t rows) {
  if (!c) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (rows < 1 || (size_t)rows > c->keyhat_capacity_rows) return fail(SGFHE_ERR_ARG, "rows exceeds the key buffer");
  c->key_rows = rows; new_key_token(c);
  return SGFHE_OK;
}


// Serialised form of the pre-transformed key: header + raw words, so a key is transformed once and reloaded later.
struct KeyBlobHeader { uint32_t magic, version; int32_t n, m, L, rows; uint32_t p[MAXP]; uint64_t Q[2]; };
static const uint32_t KEY_MAGIC = 0x53474B31u;   // "SGK1"

extern "C" int sgfhe_bkey_export_size(sgfhe_ctx* c, int32_t rows, uint64_t* bytes) {
  if (!c || !bytes) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (rows < 1 || rows > c->key_rows) return fail(SGFHE_ERR_STATE, "rows exceeds the uploaded key rows");
  *bytes = sizeof(KeyBlobHeader) + (uint64_t)rows * keyhat_row_words(c) * sizeof(uint32_t);
  return SGFHE_OK;
}

extern "C" int sgfhe_bkey_export(sgfhe_ctx* c, int32_t rows, void* blob, uint64_t bytes) {
  uint64_t need = 0;
  int rc = sgfhe_bkey_export_size(c, rows, &need); if (rc) return rc;
  if (!blob || bytes < need) return fail(SGFHE_ERR_ARG, "blob buffer too small");
  CK(cudaSetDevice(c->device));
  KeyBlobHeader h; memset(&h, 0, sizeof h);
  h.magic = KEY_MAGIC; h.version = 3; h.n = c->hp.n; h.m = c->hp.m; h.L = c->dc.L; h.rows = rows;
  for (int i = 0; i < MAXP; ++i) h.p[i] = c->dc.p[i];
  h.Q[0] = (uint64_t)c->hp.Q; h.Q[1] = (uint64_t)(c->hp.Q >> 64);
  memcpy(blob, &h, sizeof h);
  CK(cudaMemcpy(static_cast<char*>(blob) + sizeof h, c->d_keyhat, need - sizeof h, cudaMemcpyDeviceToHost));
  return SGFHE_OK;
}

extern "C" int sgfhe_bkey_import(sgfhe_ctx* c, const void* blob, uint64_t bytes) {
  if (!c || !blob || bytes < sizeof(KeyBlobHeader)) return fail(SGFHE_ERR_ARG, "bad blob");
  KeyBlobHeader h; memcpy(&h, blob, sizeof h);
  if (h.magic != KEY_MAGIC || h.version != 3) return fail(SGFHE_ERR_ARG, "not a serialised sgfhe key");
  if (h.n != c->hp.n || h.m != c->hp.m || h.L != c->dc.L || h.Q[0] != (uint64_t)c->hp.Q || h.Q[1] != (uint64_t)(c->hp.Q >> 64))
    return fail(SGFHE_ERR_ARG, "serialised key belongs to other parameters");
  for (int i = 0; i < h.L; ++i) if (h.p[i] != c->dc.p[i]) return fail(SGFHE_ERR_ARG, "serialised key uses another RNS basis");
  if (h.rows < 1 || h.rows > c->hp.n) return fail(SGFHE_ERR_ARG, "bad row count");
  const uint64_t need = sizeof h + (uint64_t)h.rows * keyhat_row_words(c) * sizeof(uint32_t);
  if (bytes < need) return fail(SGFHE_ERR_ARG, "truncated blob");
  CK(cudaSetDevice(c->device));
  int rc = ensure_keyhat(c, h.rows); if (rc) return rc;
  c->key_rows = 0; c->key_token = 0;
  CK(cudaMemcpy(c->d_keyhat, static_cast<const char*>(blob) + sizeof h, need - sizeof h, cudaMemcpyHostToDevice));
  c->key_rows = h.rows; new_key_token(c);
  return SGFHE_OK;
}

static int launch_gates(sgfhe_ctx* c, GateArgs& A, cudaStream_t st) {
  const int grid = A.batch < c->max_ctas ? A.batch : c->max_ctas;
  int rc = ensure_scratch(c, grid); if (rc) return rc;
  A.keyhat = c->d_keyhat; A.tw_f = c->d_tw_f; A.tw_i = c->d_tw_i;
  A.scratch = c->d_scratch; A.scratch_stride = c->scratch_stride;
  A.zres = c->d_scratch + (size_t)c->scratch_ctas * c->scratch_stride; A.zres_stride = c->zres_stride;
  if (A.batch > grid) {                              // more gates than CTAs: dynamic distribution
    if (!c->d_counter && cudaMalloc(&c->d_counter, sizeof(int)) != cudaSuccess) return fail(SGFHE_ERR_NOMEM, "cudaMalloc of the work counter failed");
    CK(cudaMemsetAsync(c->d_counter, 0, sizeof(int), st));
    A.work_counter = c->d_counter;
  }
  launch_bootstrap(c, grid, st, A);
  CK(cudaGetLastError());
  return SGFHE_OK;
}

extern "C" int sgfhe_bootstrap_batch_device(sgfhe_ctx* c, int32_t batch, const uint64_t* d_lwe1, const uint64_t* d_lwe2,
                                            const int64_t* d_draws, uint64_t* d_and, uint64_t* d_or, uint64_t* d_xor,
                                            void* stream) {
  if (!c || !d_lwe1 || !d_lwe2 || !d_and || !d_or || !d_xor) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (batch < 0) return fail(SGFHE_ERR_ARG, "negative batch");
  if (c->key_rows != c->hp.n) return fail(SGFHE_ERR_STATE, "no complete bootstrap key uploaded");
  if (batch == 0) return SGFHE_OK;
  CK(cudaSetDevice(c->device));
  GateArgs A; memset(&A, 0, sizeof A);
  A.lwe1 = d_lwe1; A.lwe2 = d_lwe2; A.draws = d_draws; A.out_and = d_and; A.out_or = d_or; A.out_xor = d_xor;
  A.batch = batch; A.step_begin = 0; A.step_end = c->hp.n; A.draw_steps = c->hp.n; A.flags = F_INIT | F_DECOMP | F_FINAL;
  if (const char* sc = getenv("SGFHE_STAGGER")) { A.stagger_cycles = atoi(sc); A.stagger_slots = getenv("SGFHE_STAGGER_SLOTS") ? atoi(getenv("SGFHE_STAGGER_SLOTS")) : 16; }
  if (getenv("SGFHE_PHASE_TIMING")) {            // profiling aid: per-phase cycles of CTA 0, printed to stderr
    unsigned long long* d_t = nullptr; unsigned long long h_t[8] = {0};
    CK(cudaMalloc(&d_t, sizeof h_t)); CK(cudaMemset(d_t, 0, sizeof h_t));
    A.timing = d_t;
    int rc = launch_gates(c, A, (cudaStream_t)stream);
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h_t, d_t, sizeof h_t, cudaMemcpyDeviceToHost)); cudaFree(d_t);
    static const char* names[8] = {"digit load+top stage", "forward passes", "pointwise", "inverse passes", "top stage+store", "crt", "update+decompose", "tail C (split tail)"};
    unsigned long long tot = 0; for (int i = 0; i < 8; ++i) tot += h_t[i];
    for (int i = 0; i < 8; ++i) if (i < 7 || h_t[i]) fprintf(stderr, "[sgfhe phase] %-22s %12llu cycles  %5.1f%%\n", names[i], h_t[i], 100.0 * h_t[i] / (tot ? tot : 1));
    fprintf(stderr, "[sgfhe phase] total %llu cycles over %d steps = %.0f cycles/step\n", tot, c->hp.n, (double)tot / c->hp.n);
    return rc;
  }
  return launch_gates(c, A, (cudaStream_t)stream);
}

// bootstrap(bkey, rng, ...) with the flatten draws made on the device (DrawSrc): gate g of the batch uses the stream
// (seed, gate0 + g).  seed != 0.
extern "C" int sgfhe_bootstrap_batch_rng_device(sgfhe_ctx* c, int32_t batch, const uint64_t* d_lwe1, const uint64_t* d_lwe2,
                                                uint64_t seed, uint64_t gate0, uint64_t* d_and, uint64_t* d_or, uint64_t* d_xor,
                                                void* stream) {
  if (!c || !d_lwe1 || !d_lwe2 || !d_and || !d_or || !d_xor) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (batch < 0) return fail(SGFHE_ERR_ARG, "negative batch");
  if (seed == 0) return fail(SGFHE_ERR_ARG, "seed must be non-zero (0 selects the deterministic flatten)");
  if (c->key_rows != c->hp.n) return fail(SGFHE_ERR_STATE, "no complete bootstrap key uploaded");
  if (batch == 0) return SGFHE_OK;
  CK(cudaSetDevice(c->device));
  GateArgs A; memset(&A, 0, sizeof A);
  A.lwe1 = d_lwe1; A.lwe2 = d_lwe2; A.out_and = d_and; A.out_or = d_or; A.out_xor = d_xor; A.rng_seed = seed; A.rng_gate0 = gate0;
  A.batch = batch; A.step_begin = 0; A.step_end = c->hp.n; A.draw_steps = c->hp.n; A.flags = F_INIT | F_DECOMP | F_FINAL;
  return launch_gates(c, A, (cudaStream_t)stream);
}

extern "C" int sgfhe_bootstrap_batch_rng(sgfhe_ctx* c, int32_t batch, const uint64_t* lwe1, const uint64_t* lwe2, uint64_t seed,
                                         uint64_t gate0, uint64_t* out_and, uint64_t* out_or, uint64_t* out_xor) {
  if (!c || !lwe1 || !lwe2 || !out_and || !out_or || !out_xor) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (batch < 0) return fail(SGFHE_ERR_ARG, "negative batch");
  if (batch == 0) return SGFHE_OK;
  CK(cudaSetDevice(c->device));
  const size_t w = (size_t)batch * (c->hp.n + 1);
  int rc = ensure_arena(c, 5 * w * 8); if (rc) return rc;
  uint64_t* d = reinterpret_cast<uint64_t*>(c->d_arena);
  cudaError_t e = cudaMemcpyAsync(d, lwe1, w * 8, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + w, lwe2, w * 8, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) rc = sgfhe_bootstrap_batch_rng_device(c, batch, d, d + w, seed, gate0, d + 2 * w, d + 3 * w, d + 4 * w, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_and, d + 2 * w, w * 8, cudaMemcpyDeviceToHost, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_or, d + 3 * w, w * 8, cudaMemcpyDeviceToHost, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_xor, d + 4 * w, w * 8, cudaMemcpyDeviceToHost, nullptr);
  const cudaError_t es = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = es;
  if (rc) return rc;
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("bootstrap_batch_rng: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}

// test seam: the draws the device generator makes for accumulation steps step0 .. step0+steps-1 of gate `gate`
extern "C" int sgfhe_device_draws(sgfhe_ctx* c, uint64_t seed, uint64_t gate, int32_t step0, int32_t steps, int64_t* out) {
  if (!c || !out) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (seed == 0 || step0 < 0 || steps < 1) return fail(SGFHE_ERR_ARG, "bad seed / step range");
  CK(cudaSetDevice(c->device));
  const size_t count = (size_t)steps * 2 * c->hp.m;
  int rc = ensure_arena(c, count * 16); if (rc) return rc;
  device_draws_kernel<<<(unsigned)((count + 255) / 256), 256>>>(c->dc, seed, gate, step0, steps, reinterpret_cast<int64_t*>(c->d_arena));
  ++g_launches;
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, c->d_arena, count * 16, cudaMemcpyDeviceToHost));
  return SGFHE_OK;
}

extern "C" int sgfhe_bootstrap_batch(sgfhe_ctx* c, int32_t batch, const uint64_t* lwe1, const uint64_t* lwe2,
                                     const int64_t* draws, uint64_t* out_and, uint64_t* out_or, uint64_t* out_xor) {
  if (!c || !lwe1 || !lwe2 || !out_and || !out_or || !out_xor) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (batch < 0) return fail(SGFHE_ERR_ARG, "negative batch");
  if (c->key_rows != c->hp.n) return fail(SGFHE_ERR_STATE, "no complete bootstrap key uploaded");
  if (batch == 0) return SGFHE_OK;
  CK(cudaSetDevice(c->device));
  const size_t lwe_bytes = (size_t)batch * (c->hp.n + 1) * sizeof(uint64_t);
  const size_t draw_bytes = draws ? (size_t)batch * c->hp.n * 4 * c->hp.m * sizeof(int64_t) : 0;
  if (5 * lwe_bytes > c->io_capacity) {                 // LWE staging buffers are kept between calls (grow only)
    cudaFree(c->d_io); c->d_io = nullptr; c->io_capacity = 0;
    if (cudaMalloc(&c->d_io, 5 * lwe_bytes) != cudaSuccess) return fail(SGFHE_ERR_NOMEM, "cudaMalloc of LWE buffers failed");
    c->io_capacity = 5 * lwe_bytes;
  }
  uint64_t* d_io = c->d_io; int64_t* d_draws = nullptr;
  if (draws) { const int rca = ensure_arena(c, draw_bytes); if (rca) return rca; d_draws = reinterpret_cast<int64_t*>(c->d_arena); }
  const size_t w = lwe_bytes / sizeof(uint64_t);
  int rc = SGFHE_OK;
  cudaError_t e = cudaMemcpyAsync(d_io, lwe1, lwe_bytes, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_io + w, lwe2, lwe_bytes, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess && draws) e = cudaMemcpyAsync(d_draws, draws, draw_bytes, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) rc = sgfhe_bootstrap_batch_device(c, batch, d_io, d_io + w, d_draws, d_io + 2 * w, d_io + 3 * w, d_io + 4 * w, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_and, d_io + 2 * w, lwe_bytes, cudaMemcpyDeviceToHost, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_or, d_io + 3 * w, lwe_bytes, cudaMemcpyDeviceToHost, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_xor, d_io + 4 * w, lwe_bytes, cudaMemcpyDeviceToHost, nullptr);
  const cudaError_t es = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = es;
  if (rc) return rc;
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("bootstrap_batch: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}


// _bootstrap_internal for a batch with device buffers: out_* [batch][n+1][2] wide over Z_Q
static int internal_batch_device(sgfhe_ctx* c, int batch, const uint64_t* d_l1, const uint64_t* d_l2, const int64_t* d_draws,
                                 uint64_t* d_and, uint64_t* d_or, uint64_t* d_xor, cudaStream_t st) {
  GateArgs A; memset(&A, 0, sizeof A);
  A.lwe1 = d_l1; A.lwe2 = d_l2; A.draws = d_draws; A.out_and = d_and; A.out_or = d_or; A.out_xor = d_xor;
  A.batch = batch; A.step_begin = 0; A.step_end = c->hp.n; A.draw_steps = c->hp.n; A.flags = F_INIT | F_DECOMP | F_FINAL | F_RAW;
  return launch_gates(c, A, st);
}

// shortened_external_product batch with device buffers: d_polys [count][m][2], d_out [count][2][m][2]
static int shortened_device(sgfhe_ctx* c, int count, const uint64_t* d_polys, const int64_t* d_draws, uint64_t* d_out,
                            uint64_t* d_dummy, cudaStream_t st) {
  GateArgs A; memset(&A, 0, sizeof A);
  A.lwe1 = d_dummy; A.lwe2 = d_dummy; A.out_and = A.out_or = A.out_xor = d_dummy;
  A.batch = count; A.flags = F_EXT | F_DECOMP | F_PACK;
  A.pack_in = d_polys; A.pack_draws = d_draws; A.trace = d_out;
  return launch_gates(c, A, st);
}

extern "C" int sgfhe_bootstrap_internal_batch(sgfhe_ctx* c, int32_t batch, const uint64_t* lwe1, const uint64_t* lwe2,
                                              const int64_t* draws, uint64_t* out_and, uint64_t* out_or, uint64_t* out_xor) {
  if (!c || !lwe1 || !lwe2 || !out_and || !out_or || !out_xor) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (batch < 0) return fail(SGFHE_ERR_ARG, "negative batch");
  if (c->key_rows != c->hp.n) return fail(SGFHE_ERR_STATE, "no complete bootstrap key uploaded");
  if (batch == 0) return SGFHE_OK;
  CK(cudaSetDevice(c->device));
  const size_t lwe_w = (size_t)batch * (c->hp.n + 1);
  const size_t draw_bytes = draws ? (size_t)batch * c->hp.n * 4 * c->hp.m * sizeof(int64_t) : 0;
  int rc = ensure_arena(c, 8 * lwe_w * 8 + draw_bytes); if (rc) return rc;
  uint64_t* d_io = reinterpret_cast<uint64_t*>(c->d_arena);
  int64_t* d_draws = draws ? reinterpret_cast<int64_t*>(c->d_arena + 8 * lwe_w * 8) : nullptr;
  cudaError_t e = cudaMemcpyAsync(d_io, lwe1, lwe_w * 8, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_io + lwe_w, lwe2, lwe_w * 8, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess && draws) e = cudaMemcpyAsync(d_draws, draws, draw_bytes, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) rc = internal_batch_device(c, batch, d_io, d_io + lwe_w, d_draws, d_io + 2 * lwe_w, d_io + 4 * lwe_w, d_io + 6 * lwe_w, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_and, d_io + 2 * lwe_w, lwe_w * 16, cudaMemcpyDeviceToHost, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_or, d_io + 4 * lwe_w, lwe_w * 16, cudaMemcpyDeviceToHost, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_xor, d_io + 6 * lwe_w, lwe_w * 16, cudaMemcpyDeviceToHost, nullptr);
  const cudaError_t es = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = es;
  if (rc) return rc;
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("bootstrap_internal_batch: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}

extern "C" int sgfhe_shortened_products(sgfhe_ctx* c, int32_t count, const uint64_t* polys, const int64_t* draws,
                                        uint64_t* out) {
  if (!c || !polys || !out) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (count < 0 || count > c->key_rows) return fail(count < 0 ? SGFHE_ERR_ARG : SGFHE_ERR_STATE, "count exceeds the uploaded key rows");
  if (count == 0) return SGFHE_OK;
  CK(cudaSetDevice(c->device));
  const size_t m = c->hp.m, in_w = (size_t)count * m * 2, out_w = (size_t)count * 4 * m;
  int rc = ensure_arena(c, (in_w + out_w + 8 + (draws ? in_w : 0)) * 8); if (rc) return rc;
  uint64_t* d = reinterpret_cast<uint64_t*>(c->d_arena);
  int64_t* d_draws = draws ? reinterpret_cast<int64_t*>(d + in_w + out_w + 8) : nullptr;
  cudaError_t e = cudaMemcpyAsync(d, polys, in_w * 8, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess && draws) e = cudaMemcpyAsync(d_draws, draws, in_w * 8, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) rc = shortened_device(c, count, d, d_draws, d + in_w, d + in_w + out_w, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out, d + in_w, out_w * 8, cudaMemcpyDeviceToHost, nullptr);
  const cudaError_t es = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = es;
  if (rc) return rc;
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("shortened_products: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}

// pack_encrypted_bits (src/fhe.jl:660-696) on the device.  Arena layout shared by the two entry points below, in uint64
// words: triv | bits | and, or, xor (wide, [n][n+1][2] each) | polys [n][m][2] | wv [n][2][m][2] | w, v [m] | dummy | draws
struct PackLayout { size_t triv, bits, o_and, o_or, o_xor, polys, wv, w, v, dummy, db, ds, total; };
static PackLayout pack_layout(size_t n, size_t m, size_t db_bytes, size_t ds_bytes) {
  PackLayout L; const size_t lwe_w = n * (n + 1);
  L.triv = 0; L.bits = lwe_w; L.o_and = 2 * lwe_w; L.o_or = 4 * lwe_w; L.o_xor = 6 * lwe_w; L.polys = 8 * lwe_w;
  L.wv = L.polys + n * m * 2; L.w = L.wv + n * 4 * m; L.v = L.w + m; L.dummy = L.v + m; L.db = L.dummy + 8;
  L.ds = L.db + db_bytes / 8; L.total = L.ds + ds_bytes / 8;
  return L;
}
// stages after the n internal bootstraps: transposition (src/fhe.jl:675-678), n shortened external products (:683-684),
// sums, negate / subtract (:686-690), ModRed (:692-693).  d + PL.o_and holds the pre-ModRed LWEs.
static int pack_tail_device(sgfhe_ctx* c, uint64_t* d, const PackLayout& PL, bool have_ds, cudaError_t* e) {
  const size_t n = c->hp.n, m = c->hp.m;
  pack_transpose_kernel<<<(unsigned)((n * m + 255) / 256), 256>>>((int)n, (int)m, d + PL.o_and, d + PL.polys);
  ++g_launches;
  *e = cudaGetLastError();
  if (*e != cudaSuccess) return SGFHE_OK;
  int rc = shortened_device(c, (int)n, d + PL.polys, have_ds ? reinterpret_cast<int64_t*>(d + PL.ds) : nullptr, d + PL.wv, d + PL.dummy, nullptr);
  if (rc) return rc;
  pack_tail_kernel<<<(unsigned)((2 * m + 127) / 128), 128>>>(c->dc, d + PL.wv, d + PL.o_and, d + PL.w, d + PL.v);
  ++g_launches;
  *e = cudaGetLastError();
  return SGFHE_OK;
}

extern "C" int sgfhe_pack_encrypted_bits(sgfhe_ctx* c, const uint64_t* enc_bits, const int64_t* draws_boot,
                                         const int64_t* draws_short, uint64_t* out_w, uint64_t* out_v) {
  if (!c || !enc_bits || !out_w || !out_v) return fail(SGFHE_ERR_ARG, "NULL argument");
  if ((draws_boot == nullptr) != (draws_short == nullptr)) return fail(SGFHE_ERR_ARG, "draws_boot and draws_short: both or neither");
  if (c->key_rows != c->hp.n) return fail(SGFHE_ERR_STATE, "no complete bootstrap key uploaded");
  CK(cudaSetDevice(c->device));
  const size_t n = c->hp.n, m = c->hp.m, lwe_w = n * (n + 1);
  const size_t db_bytes = draws_boot ? n * n * 4 * m * 8 : 0, ds_bytes = draws_short ? n * m * 2 * 8 : 0;
  const PackLayout PL = pack_layout(n, m, db_bytes, ds_bytes);
  int rc = ensure_arena(c, PL.total * 8); if (rc) return rc;
  uint64_t* d = reinterpret_cast<uint64_t*>(c->d_arena);
  std::vector<uint64_t> triv(lwe_w, 0);
  for (size_t i = 0; i < n; ++i) triv[i * (n + 1) + n] = c->hp.Dr;       // trivial LWE encrypting 1 (src/fhe.jl:670-671)
  cudaError_t e = cudaMemcpyAsync(d + PL.triv, triv.data(), lwe_w * 8, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + PL.bits, enc_bits, lwe_w * 8, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess && draws_boot) e = cudaMemcpyAsync(d + PL.db, draws_boot, db_bytes, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess && draws_short) e = cudaMemcpyAsync(d + PL.ds, draws_short, ds_bytes, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);              // `triv` is pageable: copied before it goes out of scope
  if (e == cudaSuccess) rc = internal_batch_device(c, (int)n, d + PL.triv, d + PL.bits, draws_boot ? reinterpret_cast<int64_t*>(d + PL.db) : nullptr,
                                                   d + PL.o_and, d + PL.o_or, d + PL.o_xor, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) rc = pack_tail_device(c, d, PL, draws_short != nullptr, &e);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_w, d + PL.w, m * 8, cudaMemcpyDeviceToHost, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_v, d + PL.v, m * 8, cudaMemcpyDeviceToHost, nullptr);
  const cudaError_t es = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = es;
  if (rc) return rc;
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("pack_encrypted_bits: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}

// Test seam: the stages of pack_encrypted_bits after the n bootstraps, from given pre-ModRed LWEs (src/fhe.jl:675-693).
extern "C" int sgfhe_pack_from_lwes(sgfhe_ctx* c, const uint64_t* new_lwes, const int64_t* draws_short, uint64_t* out_w, uint64_t* out_v) {
  if (!c || !new_lwes || !out_w || !out_v) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (c->key_rows != c->hp.n) return fail(SGFHE_ERR_STATE, "no complete bootstrap key uploaded");
  CK(cudaSetDevice(c->device));
  const size_t n = c->hp.n, m = c->hp.m, lwe_w = n * (n + 1);
  const size_t ds_bytes = draws_short ? n * m * 2 * 8 : 0;
  const PackLayout PL = pack_layout(n, m, 0, ds_bytes);
  int rc = ensure_arena(c, PL.total * 8); if (rc) return rc;
  uint64_t* d = reinterpret_cast<uint64_t*>(c->d_arena);
  cudaError_t e = cudaMemcpyAsync(d + PL.o_and, new_lwes, lwe_w * 16, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess && draws_short) e = cudaMemcpyAsync(d + PL.ds, draws_short, ds_bytes, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) rc = pack_tail_device(c, d, PL, draws_short != nullptr, &e);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_w, d + PL.w, m * 8, cudaMemcpyDeviceToHost, nullptr);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpyAsync(out_v, d + PL.v, m * 8, cudaMemcpyDeviceToHost, nullptr);
  const cudaError_t es = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = es;
  if (rc) return rc;
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("pack_from_lwes: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}

// BootstrapKey(rng, sk) (src/fhe.jl:181-201) on the device, output already in the transform domain.
//   sk: n bytes (0/1).  a_rand: [rows][4][m][2] wide, the uniform polynomials a_1..a_4 of rows row0.. (src/fhe.jl:193);
//   e_rand: [rows][4][m] int64 in [-n, n] (src/fhe.jl:194) -- the caller's RNG draws them in the reference's order.
//   The 4 rows products a_j * ext_key (src/fhe.jl:195), + e_j, + s_i G (src/fhe.jl:196) and the key pre-transform run on the
//   device; key_out (NULL or host [rows][4][2][m][2] wide) receives the coefficient form only on request.
extern "C" int sgfhe_bkey_generate(sgfhe_ctx* c, const uint8_t* sk, const uint64_t* a_rand, const int64_t* e_rand,
                                   int32_t row0, int32_t rows, uint64_t* key_out) {
  if (!c || !sk || !a_rand || !e_rand) return fail(SGFHE_ERR_ARG, "NULL argument");
  const int n = c->hp.n; const size_t m = c->hp.m;
  if (row0 < 0 || rows < 1 || row0 + rows > n) return fail(SGFHE_ERR_ARG, "rows out of range");
  if (row0 > c->key_rows) return fail(SGFHE_ERR_STATE, "key rows must be generated in order (row0 exceeds the rows present)");
  CK(cudaSetDevice(c->device));
  int rc = SGFHE_OK;
  if ((size_t)n > c->keyhat_capacity_rows) {           // growing the buffer drops its content: only legal at the start
    if (row0 != 0) return fail(SGFHE_ERR_STATE, "key buffer smaller than n rows: start at row 0");
    rc = ensure_keyhat(c, n); if (rc) return rc;
  }
  const int chunk = (int)std::max<size_t>(1, ((size_t)1 << 25) / (4 * m * 16) * 4);    // rows per pass: <= 128 MiB of a_rand
  // arena: sk | ext (m wide) | a | prod | e | coef
  const size_t crow = (size_t)(rows < chunk ? rows : chunk);
  const size_t o_sk = 0, o_ext = 4096, o_a = o_ext + m * 16, o_prod = o_a + crow * 4 * m * 16, o_e = o_prod + crow * 4 * m * 16,
               o_coef = o_e + crow * 4 * m * 8, total = o_coef + crow * 8 * m * 16;
  if (n > 4096) return fail(SGFHE_ERR_ARG, "n too large");          // the secret key occupies the first o_ext bytes of the arena
  rc = ensure_arena(c, total); if (rc) return rc;
  uint8_t* d = c->d_arena;
  std::vector<uint64_t> ext(m * 2, 0);
  for (int i = 0; i < n; ++i) ext[2 * (size_t)i] = sk[i] ? 1 : 0;           // resize(polynomial_Q(sk.key), m)  (src/fhe.jl:185)
  CK(cudaMemcpy(d + o_sk, sk, n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d + o_ext, ext.data(), m * 16, cudaMemcpyHostToDevice));
  const int old_rows = c->key_rows;
  c->key_rows = row0 < old_rows ? row0 : old_rows; c->key_token = 0;        // rows from row0 on are being rewritten
  for (int done = 0; done < rows; done += (int)crow) {
    const int cnt = rows - done < (int)crow ? rows - done : (int)crow;
    const size_t polys = (size_t)cnt * 4;
    CK(cudaMemcpyAsync(d + o_a, a_rand + (size_t)done * 4 * m * 2, polys * m * 16, cudaMemcpyHostToDevice, nullptr));
    CK(cudaMemcpyAsync(d + o_e, e_rand + (size_t)done * 4 * m, polys * m * 8, cudaMemcpyHostToDevice, nullptr));
    {                                                   // products a_j * ext_key, second operand broadcast
      const bool v4 = c->use_v4 && !getenv("SGFHE_POLYMUL_V3");
      const int cap = v4 ? c->num_sms : c->num_sms * std::max(2, 1024 / c->threads), units = v4 ? ((int)polys + 1) / 2 : (int)polys;
      const int grid = units < cap ? units : cap;
      if (grid > c->pm_ctas) {
        cudaFree(c->d_pm_scratch); c->d_pm_scratch = nullptr; c->pm_ctas = 0;
        if (cudaMalloc(&c->d_pm_scratch, (size_t)grid * c->dc.LM * 2 * m * sizeof(uint32_t)) != cudaSuccess)
          return fail(SGFHE_ERR_NOMEM, "cudaMalloc of polymul scratch failed");
        c->pm_ctas = grid;
      }
      launch_polymul(c, grid, nullptr, reinterpret_cast<uint64_t*>(d + o_a), reinterpret_cast<uint64_t*>(d + o_ext),
                     reinterpret_cast<uint64_t*>(d + o_prod), (int)polys, 1);
      CK(cudaGetLastError());
    }
    keygen_assemble_kernel<<<(unsigned)((polys * m + 255) / 256), 256>>>(c->dc, reinterpret_cast<uint64_t*>(d + o_a),
        reinterpret_cast<uint64_t*>(d + o_prod), reinterpret_cast<int64_t*>(d + o_e), d + o_sk, row0 + done, cnt,
        reinterpret_cast<uint64_t*>(d + o_coef));
    ++g_launches;
    CK(cudaGetLastError());
    launch_key_transform(c, cnt * 8, reinterpret_cast<uint64_t*>(d + o_coef), c->d_keyhat, (row0 + done) * 8);
    CK(cudaGetLastError());
    if (key_out) CK(cudaMemcpyAsync(key_out + (size_t)done * 8 * m * 2, d + o_coef, (size_t)cnt * 8 * m * 16, cudaMemcpyDeviceToHost, nullptr));
    CK(cudaDeviceSynchronize());
  }
  c->key_rows = std::max(old_rows, row0 + rows); new_key_token(c);
  return SGFHE_OK;
}

extern "C" int sgfhe_bootstrap_trace(sgfhe_ctx* c, const uint64_t* lwe1, const uint64_t* lwe2, const int64_t* draws,
                                     int32_t n_steps, uint64_t* trace, uint64_t* out_and, uint64_t* out_or,
                                     uint64_t* out_xor) {
  if (!c || !lwe1 || !lwe2 || !out_and || !out_or || !out_xor) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (n_steps < 0 || n_steps > c->key_rows) return fail(n_steps < 0 ? SGFHE_ERR_ARG : SGFHE_ERR_STATE, "n_steps exceeds the uploaded key rows");
  CK(cudaSetDevice(c->device));
  const int n = c->hp.n, m = c->hp.m;
  const size_t lwe_w = n + 1, tr_w = (size_t)2 * m * 2, dr_w = (size_t)4 * m;
  uint64_t* d_buf = nullptr; int64_t* d_draws = nullptr;
  const size_t words = 2 * lwe_w + 3 * lwe_w * 2 + (trace ? (size_t)n_steps * tr_w : 0);
  if (cudaMalloc(&d_buf, words * sizeof(uint64_t)) != cudaSuccess) return fail(SGFHE_ERR_NOMEM, "cudaMalloc failed");
  if (draws && n_steps && cudaMalloc(&d_draws, (size_t)n_steps * dr_w * sizeof(int64_t)) != cudaSuccess) { cudaFree(d_buf); return fail(SGFHE_ERR_NOMEM, "cudaMalloc failed"); }
  uint64_t* d_l1 = d_buf; uint64_t* d_l2 = d_buf + lwe_w; uint64_t* d_out = d_buf + 2 * lwe_w; uint64_t* d_tr = d_out + 6 * lwe_w;
  int rc = SGFHE_OK;
  cudaError_t e = cudaMemcpy(d_l1, lwe1, lwe_w * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_l2, lwe2, lwe_w * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && d_draws) e = cudaMemcpy(d_draws, draws, (size_t)n_steps * dr_w * 8, cudaMemcpyHostToDevice);
  // one launch per step so the accumulator can be copied out after each (test seam; the product path
  // runs all n steps in one launch)
  for (int k = -1; k < n_steps && rc == SGFHE_OK && e == cudaSuccess; ++k) {
    GateArgs A; memset(&A, 0, sizeof A);
    A.lwe1 = d_l1; A.lwe2 = d_l2; A.batch = 1;
    A.out_and = d_out; A.out_or = d_out + 2 * lwe_w; A.out_xor = d_out + 4 * lwe_w;
    if (k < 0) { A.flags = F_INIT; A.step_begin = A.step_end = 0; }
    else {
      A.step_begin = k; A.step_end = k + 1; A.draw_steps = 1; A.flags = F_DECOMP;
      A.draws = d_draws ? d_draws + (size_t)k * dr_w : nullptr;
      A.trace = trace ? d_tr + (size_t)k * tr_w : nullptr;
    }
    if (k == n_steps - 1) A.flags |= F_FINAL | F_RAW;
    rc = launch_gates(c, A, nullptr);
  }
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaDeviceSynchronize();
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpy(out_and, d_out, lwe_w * 16, cudaMemcpyDeviceToHost);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpy(out_or, d_out + 2 * lwe_w, lwe_w * 16, cudaMemcpyDeviceToHost);
  if (rc == SGFHE_OK && e == cudaSuccess) e = cudaMemcpy(out_xor, d_out + 4 * lwe_w, lwe_w * 16, cudaMemcpyDeviceToHost);
  if (rc == SGFHE_OK && e == cudaSuccess && trace && n_steps) e = cudaMemcpy(trace, d_tr, (size_t)n_steps * tr_w * 8, cudaMemcpyDeviceToHost);
  cudaFree(d_buf); cudaFree(d_draws);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("bootstrap_trace: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}


extern "C" int sgfhe_split_ciphertext_device(sgfhe_ctx* c, int32_t count, int32_t N, const uint64_t* d_a, const uint64_t* d_b,
                                             uint64_t* d_lwes, void* stream) {
  if (!c || !d_a || !d_b || !d_lwes) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (count < 0) return fail(SGFHE_ERR_ARG, "negative count");
  if (N < c->hp.n) return fail(SGFHE_ERR_ARG, "polynomial shorter than n (src/fhe.jl:238)");
  if (count == 0) return SGFHE_OK;
  CK(cudaSetDevice(c->device));
  split_kernel<<<(unsigned)count * c->hp.n, 256, 0, (cudaStream_t)stream>>>(c->hp.n, N, c->hp.r, d_a, d_b, d_lwes);
  ++g_launches;
  CK(cudaGetLastError());
  return SGFHE_OK;
}

extern "C" int sgfhe_split_ciphertext(sgfhe_ctx* c, int32_t count, int32_t N, const uint64_t* a, const uint64_t* b, uint64_t* lwes) {
  if (!c || !a || !b || !lwes) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (count < 0) return fail(SGFHE_ERR_ARG, "negative count");
  if (N < c->hp.n) return fail(SGFHE_ERR_ARG, "polynomial shorter than n (src/fhe.jl:238)");
  if (count == 0) return SGFHE_OK;
  CK(cudaSetDevice(c->device));
  const size_t in_w = (size_t)count * N, out_w = (size_t)count * c->hp.n * (c->hp.n + 1);
  int rc = ensure_arena(c, (2 * in_w + out_w) * 8); if (rc) return rc;
  uint64_t* d = reinterpret_cast<uint64_t*>(c->d_arena);
  cudaError_t e = cudaMemcpyAsync(d, a, in_w * 8, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + in_w, b, in_w * 8, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) rc = sgfhe_split_ciphertext_device(c, count, N, d, d + in_w, d + 2 * in_w, nullptr);
  if (!rc && e == cudaSuccess) e = cudaMemcpy(lwes, d + 2 * in_w, out_w * 8, cudaMemcpyDeviceToHost);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("split_ciphertext: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}

extern "C" int sgfhe_decrypt_bits_device(sgfhe_ctx* c, int32_t count, const uint64_t* d_lwes, const uint8_t* d_sk, uint8_t* d_out,
                                         void* stream) {
  if (!c || !d_lwes || !d_sk || !d_out) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (count < 0) return fail(SGFHE_ERR_ARG, "negative count");
  if (count == 0) return SGFHE_OK;
  CK(cudaSetDevice(c->device));
  const int threads = 256, blocks = (int)(((size_t)count * 32 + threads - 1) / threads);
  decrypt_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(c->hp.n, count, c->hp.r - 1, c->hp.Dr, d_lwes, d_sk, d_out);
  ++g_launches;
  CK(cudaGetLastError());
  return SGFHE_OK;
}

extern "C" int sgfhe_decrypt_bits(sgfhe_ctx* c, int32_t count, const uint64_t* lwes, const uint8_t* sk, uint8_t* out) {
  if (!c || !lwes || !sk || !out) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (count < 0) return fail(SGFHE_ERR_ARG, "negative count");
  if (count == 0) return SGFHE_OK;
  CK(cudaSetDevice(c->device));
  const size_t lw = (size_t)count * (c->hp.n + 1) * 8;
  int rc = ensure_arena(c, lw + c->hp.n + count); if (rc) return rc;
  uint8_t* d = c->d_arena;
  cudaError_t e = cudaMemcpyAsync(d, lwes, lw, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + lw, sk, c->hp.n, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) rc = sgfhe_decrypt_bits_device(c, count, reinterpret_cast<uint64_t*>(d), d + lw, d + lw + c->hp.n, nullptr);
  if (!rc && e == cudaSuccess) e = cudaMemcpy(out, d + lw + c->hp.n, count, cudaMemcpyDeviceToHost);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("decrypt_bits: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}

static uint64_t h_isqrt(uint64_t x) { uint64_t r = 0; while ((r + 1) * (r + 1) <= x) ++r; return r; }

extern "C" int sgfhe_scheme2_params_derive(int32_t k, sgfhe_scheme2_params* out) {
  if (!out) return fail(SGFHE_ERR_ARG, "out is NULL");
  if (k < 1 || k > 5) return fail(SGFHE_ERR_ARG, "k must be in 1..5 (src/fhe2.jl:39)");
  const uint64_t n = 1024, r = ((uint64_t)1 << (k + 6)) * h_isqrt(n), m = r / 2, l = 2, tau = 2 * h_isqrt(n);
  int t = 0; while (((uint64_t)1 << t) < r) ++t;
  const u128 q = h_find_modulus(2 * n, (u128)128 * r * n, 0);
  const u128 Bp = h_find_modulus(r, (u128)15 * ((u128)1 << (2 * k + 2)) * r * tau * h_isqrt(2 * l * m), 0);
  const u128 B = h_find_modulus(r, Bp + 1, 0);
  out->n = (int32_t)n; out->k = k; out->t = t - 1; out->pad = 0; out->r = r; out->m = m; out->q = (uint64_t)q; out->tau = tau;
  out->B = (uint64_t)B; out->Bp = (uint64_t)Bp; out->Dr = r >> (k + 2); out->Dq = (uint64_t)q >> (k + 2);
  return SGFHE_OK;
}

extern "C" int sgfhe_rns2_op_device(int32_t device, int32_t op, uint64_t count, const uint64_t* d_a1, const uint64_t* d_a2,
                                    const uint64_t* d_b1, const uint64_t* d_b2, uint64_t M1, uint64_t M2, uint64_t* d_o1,
                                    uint64_t* d_o2, void* stream) {
  if (!d_a1 || !d_a2 || !d_b1 || !d_b2 || !d_o1 || !d_o2) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (op < 0 || op > 2) return fail(SGFHE_ERR_ARG, "op must be 0 (*), 1 (+) or 2 (-)");
  if (M1 <= ((uint64_t)1 << 32) || M2 <= ((uint64_t)1 << 32) || M1 >= ((uint64_t)1 << 48) || M2 >= ((uint64_t)1 << 48))
    return fail(SGFHE_ERR_ARG, "moduli must be in (2^32, 2^48) (the range of Scheme2.Params(1..5))");
  if (count == 0) return SGFHE_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(SGFHE_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
  CK(cudaSetDevice(device));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int threads = 256;
  size_t blocks = (count + threads - 1) / threads;
  const size_t cap = (size_t)sms * 8;
  if (blocks > cap) blocks = cap;
  const uint64_t mu1 = (uint64_t)((((u128)1) << 96) / M1), mu2 = (uint64_t)((((u128)1) << 96) / M2);
  rns2_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(op, count, d_a1, d_a2, d_b1, d_b2, M1, M2, mu1, mu2, d_o1, d_o2);
  ++g_launches;
  CK(cudaGetLastError());
  return SGFHE_OK;
}

extern "C" int sgfhe_rns2_op(int32_t device, int32_t op, uint64_t count, const uint64_t* a1, const uint64_t* a2,
                             const uint64_t* b1, const uint64_t* b2, uint64_t M1, uint64_t M2, uint64_t* o1, uint64_t* o2) {
  if (!a1 || !a2 || !b1 || !b2 || !o1 || !o2) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (count == 0) return SGFHE_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(SGFHE_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
  CK(cudaSetDevice(device));
  uint64_t* d = nullptr;
  const size_t w = count;
  if (cudaMalloc(&d, 6 * w * 8) != cudaSuccess) return fail(SGFHE_ERR_NOMEM, "cudaMalloc failed");
  const uint64_t* src[4] = {a1, a2, b1, b2};
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaMemcpy(d + i * w, src[i], w * 8, cudaMemcpyHostToDevice);
  int rc = SGFHE_OK;
  if (e == cudaSuccess) { rc = sgfhe_rns2_op_device(device, op, count, d, d + w, d + 2 * w, d + 3 * w, M1, M2, d + 4 * w, d + 5 * w, nullptr); if (!rc) e = cudaDeviceSynchronize(); }
  if (!rc && e == cudaSuccess) e = cudaMemcpy(o1, d + 4 * w, w * 8, cudaMemcpyDeviceToHost);
  if (!rc && e == cudaSuccess) e = cudaMemcpy(o2, d + 5 * w, w * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("rns2_op: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}

extern "C" int sgfhe_polymul_device(sgfhe_ctx* c, int32_t batch, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out,
                                    void* stream) {
  if (!c || !d_a || !d_b || !d_out) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (batch < 0) return fail(SGFHE_ERR_ARG, "negative batch");
  if (batch == 0) return SGFHE_OK;
  CK(cudaSetDevice(c->device));
  const bool v4 = c->use_v4 && !getenv("SGFHE_POLYMUL_V3");
  // v4: one CTA per SM, two products at a time; small-m kernel: 64 registers per thread, as many CTAs as 1024 threads allow (at least two)
  const int cap = v4 ? c->num_sms : c->num_sms * std::max(2, 1024 / c->threads), units = v4 ? (batch + 1) / 2 : batch;
  const int grid = units < cap ? units : cap;
  if (grid > c->pm_ctas) {
    cudaFree(c->d_pm_scratch); c->d_pm_scratch = nullptr; c->pm_ctas = 0;
    if (cudaMalloc(&c->d_pm_scratch, (size_t)grid * c->dc.LM * 2 * c->hp.m * sizeof(uint32_t)) != cudaSuccess)
      return fail(SGFHE_ERR_NOMEM, "cudaMalloc of polymul scratch failed");
    c->pm_ctas = grid;
  }
  launch_polymul(c, grid, (cudaStream_t)stream, d_a, d_b, d_out, batch);
  CK(cudaGetLastError());
  return SGFHE_OK;
}

extern "C" int sgfhe_polymul(sgfhe_ctx* c, int32_t batch, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  if (!c || !a || !b || !out) return fail(SGFHE_ERR_ARG, "NULL argument");
  if (batch < 0) return fail(SGFHE_ERR_ARG, "negative batch");
  if (batch == 0) return SGFHE_OK;
  CK(cudaSetDevice(c->device));
  const size_t bytes = (size_t)batch * c->hp.m * 16;
  int rc = ensure_arena(c, 3 * bytes); if (rc) return rc;
  uint64_t* d = reinterpret_cast<uint64_t*>(c->d_arena);
  uint64_t* da = d; uint64_t* db = d + bytes / 8; uint64_t* dout = d + 2 * (bytes / 8);
  cudaError_t e = cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) e = cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, nullptr);
  if (e == cudaSuccess) rc = sgfhe_polymul_device(c, batch, da, db, dout, nullptr);
  if (!rc && e == cudaSuccess) e = cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("polymul: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}

extern "C" int sgfhe_flatten_poly(sgfhe_ctx* c, const uint64_t* a, const int64_t* draws, uint64_t* out) {
  if (!c || !a || !out) return fail(SGFHE_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(c->device));
  const int m = c->hp.m;
  uint64_t* d = nullptr;
  if (cudaMalloc(&d, (size_t)m * (16 + 32 + 16)) != cudaSuccess) return fail(SGFHE_ERR_NOMEM, "cudaMalloc failed");
  uint64_t* da = d; uint64_t* dout = d + 2 * (size_t)m; int64_t* dd = reinterpret_cast<int64_t*>(d + 6 * (size_t)m);
  cudaError_t e = cudaMemcpy(da, a, (size_t)m * 16, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && draws) e = cudaMemcpy(dd, draws, (size_t)m * 16, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    flatten_kernel<<<(m + 255) / 256, 256>>>(c->dc, da, draws ? dd : nullptr, dout);
    ++g_launches;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, dout, (size_t)m * 32, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("flatten_poly: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}

extern "C" int sgfhe_external_product(sgfhe_ctx* c, const uint64_t* a, const uint64_t* b, const uint64_t* Amat,
                                      const int64_t* draws, uint64_t* a_out, uint64_t* b_out) {
  if (!c || !a || !b || !Amat || !a_out || !b_out) return fail(SGFHE_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(c->device));
  const int m = c->hp.m;
  uint32_t* d_khat = nullptr; uint64_t* d_ab = nullptr; int64_t* d_draws = nullptr;
  int rc = ensure_scratch(c, 1); if (rc) return rc;
  if (cudaMalloc(&d_khat, keyhat_row_words(c) * sizeof(uint32_t)) != cudaSuccess) return fail(SGFHE_ERR_NOMEM, "cudaMalloc failed");
  if (cudaMalloc(&d_ab, (size_t)2 * m * 16 + 64) != cudaSuccess) { cudaFree(d_khat); return fail(SGFHE_ERR_NOMEM, "cudaMalloc failed"); }
  if (draws && cudaMalloc(&d_draws, (size_t)4 * m * 8) != cudaSuccess) { cudaFree(d_khat); cudaFree(d_ab); return fail(SGFHE_ERR_NOMEM, "cudaMalloc failed"); }
  rc = transform_polys(c, Amat, 0, 8, d_khat);
  cudaError_t e = cudaSuccess;
  if (!rc) {
    e = cudaMemcpy(d_ab, a, (size_t)m * 16, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_ab + 2 * (size_t)m, b, (size_t)m * 16, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && draws) e = cudaMemcpy(d_draws, draws, (size_t)4 * m * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      acc_load_kernel<<<(2 * m + 255) / 256, 256>>>(c->dc, d_ab, reinterpret_cast<uint32_t*>(c->d_scratch));
      ++g_launches;
      uint64_t* dummy = d_ab + 4 * (size_t)m;     // lwe pointers are not dereferenced for u when F_EXT is set ... but
      GateArgs A; memset(&A, 0, sizeof A);        // ... they are indexed for the pointer arithmetic only
      A.lwe1 = dummy; A.lwe2 = dummy; A.batch = 1; A.step_begin = 0; A.step_end = 1; A.draw_steps = 1;
      A.draws = d_draws; A.flags = F_EXT | F_DECOMP; A.trace = d_ab;
      A.out_and = A.out_or = A.out_xor = dummy;
      A.keyhat = d_khat; A.tw_f = c->d_tw_f; A.tw_i = c->d_tw_i; A.scratch = c->d_scratch; A.scratch_stride = c->scratch_stride;
      A.zres = c->d_scratch + (size_t)c->scratch_ctas * c->scratch_stride; A.zres_stride = c->zres_stride;
      launch_bootstrap(c, 1, nullptr, A);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(a_out, d_ab, (size_t)m * 16, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(b_out, d_ab + 2 * (size_t)m, (size_t)m * 16, cudaMemcpyDeviceToHost);
  }
  cudaFree(d_khat); cudaFree(d_ab); cudaFree(d_draws);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(SGFHE_ERR_CUDA, std::string("external_product: ") + cudaGetErrorString(e));
  return SGFHE_OK;
}
