// scheme2.cu -- the Scheme-2 variant (reference src/fhe2.jl, src/rns.jl): two-limb arithmetic on the GPU.
//
// Upstream Scheme 2 is "experimental, not finished" (src/fhe2.jl:1-7): it defines Params, PrivateKey, BootstrapKey,
// PublicKey, encrypt/decrypt and NO flatten, external product or bootstrap.  What exists and is mirrored here:
//   * the ring element RNS2Number{UInt64, B, B'} = (v mod B, v mod B')  (src/rns.jl:8-24) with limb-wise * + -
//     (src/rns.jl:51-60), as polynomial arithmetic in (Z_B x Z_B')[x]/(x^m+1): B and B' are NTT primes, B-1 and B'-1
//     divisible by r = 2m (src/fhe2.jl:57-60), so each limb is one negacyclic NTT product;
//   * BootstrapKey (src/fhe2.jl:104-131): n = 1024 matrices 4x2 of such polynomials, (a_j, a_j s + e_j) + s_i G.
// Sizes: k = 1..5 -> m = 2048..32768, B, B' in (2^32, 2^47).  A transform longer than 4096 points does not fit one CTA's
// shared memory as 64-bit words next to its twiddles, so it runs in two passes: the top log2(m) - 12 stages in registers
// straight from / to global memory (columns of stride m / 2^T, coalesced), then independent 4096-point blocks in shared
// memory.  Arithmetic: Shoup / Harvey lazy butterflies on 64-bit words (values below 4p < 2^50).
#include "../../include/sgfhe_cuda.h"
#include "host_math.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

using namespace sgfhe;

extern "C" void sgfhe_set_error_(const char* msg);
extern "C" void sgfhe_count_launch_(void);

namespace {

constexpr int S2_LOCAL_LOG = 12;                 // points per shared-memory block: 4096 x 8 B = 32 KiB
struct Tw64 { uint64_t w, wsh; };                // twiddle and floor(w 2^64 / p)

__device__ __forceinline__ uint64_t shoup64(uint64_t x, Tw64 t, uint64_t p) { return x * t.w - __umul64hi(x, t.wsh) * p; }   // [0, 2p)
__device__ __forceinline__ uint64_t csub64(uint64_t x, uint64_t p) { return x >= p ? x - p : x; }
// Barrett for two variable operands: mu = floor(2^96 / p), 2^32 < p < 2^48, a, b < p
__device__ __forceinline__ uint64_t mulmod48(uint64_t a, uint64_t b, uint64_t p, uint64_t mu) {
  const uint64_t lo = a * b, hi = __umul64hi(a, b);
  const uint64_t q = __umul64hi((hi << 32) | (lo >> 32), mu);
  uint64_t r = lo - q * p;
  r = csub64(r, p);
  return csub64(r, p);
}
__device__ __forceinline__ void ct64(uint64_t& x, uint64_t& y, Tw64 w, uint64_t p) {       // x, y in [0, 4p) -> [0, 4p)
  const uint64_t p2 = 2 * p, xr = x >= p2 ? x - p2 : x, t = shoup64(y, w, p);
  x = xr + t; y = xr - t + p2;
}
__device__ __forceinline__ void gs64(uint64_t& x, uint64_t& y, Tw64 w, uint64_t p) {       // x, y in [0, 2p) -> [0, 2p)
  const uint64_t p2 = 2 * p, s = x + y, d = x - y + p2;
  x = s >= p2 ? s - p2 : s; y = shoup64(d, w, p);
}

struct PrimeC { uint64_t p, mu; Tw64 minv; const Tw64* twf; const Tw64* twi; };   // minv = m^-1 mod p

// Top T forward stages on columns: element c + k (m >> T), k < 2^T, of polynomial `poly`; twiddles tw[2^l + g].
template <int T>
__global__ void s2_fwd_top(uint64_t* data, PrimeC P, int logm, size_t npoly) {
  const size_t cols = (size_t)1 << (logm - T), idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npoly * cols) return;
  const size_t poly = idx / cols, c = idx % cols;
  uint64_t* src = data + (poly << logm) + c;
  uint64_t* dst = src;                                               // in place: every column is private to its thread
  uint64_t x[1 << T];
#pragma unroll
  for (int k = 0; k < (1 << T); ++k) x[k] = src[(size_t)k << (logm - T)];
#pragma unroll
  for (int l = 0; l < T; ++l) {
    const int half = (1 << T) >> (l + 1);
#pragma unroll
    for (int g = 0; g < (1 << l); ++g)
#pragma unroll
      for (int k = 0; k < half; ++k) ct64(x[g * 2 * half + k], x[g * 2 * half + k + half], P.twf[(1 << l) + g], P.p);
  }
#pragma unroll
  for (int k = 0; k < (1 << T); ++k) dst[(size_t)k << (logm - T)] = x[k];
}

// Top T inverse stages (mirror of s2_fwd_top) with the scaling by m^-1 folded behind the last one; output canonical.
template <int T>
__global__ void s2_inv_top(uint64_t* __restrict__ data, PrimeC P, int logm, size_t npoly) {
  const size_t cols = (size_t)1 << (logm - T), idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npoly * cols) return;
  const size_t poly = idx / cols, c = idx % cols;
  uint64_t* d = data + (poly << logm) + c;
  uint64_t x[1 << T];
#pragma unroll
  for (int k = 0; k < (1 << T); ++k) x[k] = d[(size_t)k << (logm - T)];
#pragma unroll
  for (int l = T - 1; l >= 0; --l) {
    const int half = (1 << T) >> (l + 1);
#pragma unroll
    for (int g = 0; g < (1 << l); ++g)
#pragma unroll
      for (int k = 0; k < half; ++k) gs64(x[g * 2 * half + k], x[g * 2 * half + k + half], P.twi[(1 << l) + g], P.p);
  }
#pragma unroll
  for (int k = 0; k < (1 << T); ++k) d[(size_t)k << (logm - T)] = csub64(shoup64(x[k], P.minv, P.p), P.p);
}

// 64-bit words in shared memory: 16 banks of 8 bytes.  XOR-ing index bits 4..6 into bits 0..2 keeps the radix-8 passes with
// strides 1, 64 and 512 conflict free and the stride-8 pass two-way.
__device__ __forceinline__ int swz64(int i) { return i ^ ((i >> 4) & 7); }

// radix-2^LOGR register block over elements base + j 2^B of the local array; twiddle of group g at level l is
// tw[(t1 << l) + g] with t1 = 2^s + (global block index of the first level), s the stage of that level
template <int LOGR, bool FWD>
__device__ __forceinline__ void local_pass(uint64_t* sm, int B, int t, const Tw64* tw, size_t t1, uint64_t p) {
  constexpr int R = 1 << LOGR;
  const int base = ((t >> B) << (B + LOGR)) | (t & ((1 << B) - 1));
  uint64_t x[R];
#pragma unroll
  for (int j = 0; j < R; ++j) x[j] = sm[swz64(base + (j << B))];
  if (FWD) {
#pragma unroll
    for (int l = 0; l < LOGR; ++l) {
      const int half = R >> (l + 1);
#pragma unroll
      for (int g = 0; g < (1 << l); ++g) {
        const Tw64 w = tw[(t1 << l) + g];
#pragma unroll
        for (int k = 0; k < half; ++k) ct64(x[g * 2 * half + k], x[g * 2 * half + k + half], w, p);
      }
    }
  } else {
#pragma unroll
    for (int l = LOGR - 1; l >= 0; --l) {
      const int half = R >> (l + 1);
#pragma unroll
      for (int g = 0; g < (1 << l); ++g) {
        const Tw64 w = tw[(t1 << l) + g];
#pragma unroll
        for (int k = 0; k < half; ++k) gs64(x[g * 2 * half + k], x[g * 2 * half + k + half], w, p);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < R; ++j) sm[swz64(base + (j << B))] = x[j];
}

// stages s0 .. logm-1 of the forward transform on block `blk` (of 2^s0) of a polynomial held (swizzled) in shared memory:
// the (logm - s0) % 3 top stages as one small pass, then radix-8 passes from the high bits down
__device__ void local_fwd(uint64_t* sm, int logm, int s0, int blk, PrimeC P) {
  const int ll = logm - s0, rem = ll % 3;
  int s = s0, B = ll;                                    // next stage, bits [0, B) still to do
  if (rem) {
    B -= rem;
    for (int t = threadIdx.x; t < (1 << (ll - rem)); t += blockDim.x) {
      const size_t t1 = ((size_t)1 << s) + ((size_t)blk << (s - s0)) + (t >> B);
      if (rem == 1) local_pass<1, true>(sm, B, t, P.twf, t1, P.p); else local_pass<2, true>(sm, B, t, P.twf, t1, P.p);
    }
    s += rem;
    __syncthreads();
  }
  while (B > 0) {
    B -= 3;
    for (int t = threadIdx.x; t < (1 << (ll - 3)); t += blockDim.x)
      local_pass<3, true>(sm, B, t, P.twf, ((size_t)1 << s) + ((size_t)blk << (s - s0)) + (t >> B), P.p);
    s += 3;
    __syncthreads();
  }
}
__device__ void local_inv(uint64_t* sm, int logm, int s0, int blk, PrimeC P) {
  const int ll = logm - s0, rem = ll % 3;
  int s = logm, B = 0;                                   // stages [s, logm) are undone, bits [0, B) are done
  while (B + 3 <= ll - rem) {
    s -= 3;
    for (int t = threadIdx.x; t < (1 << (ll - 3)); t += blockDim.x)
      local_pass<3, false>(sm, B, t, P.twi, ((size_t)1 << s) + ((size_t)blk << (s - s0)) + (t >> B), P.p);
    B += 3;
    __syncthreads();
  }
  if (rem) {
    s -= rem;
    for (int t = threadIdx.x; t < (1 << (ll - rem)); t += blockDim.x) {
      const size_t t1 = ((size_t)1 << s) + ((size_t)blk << (s - s0)) + (t >> B);
      if (rem == 1) local_pass<1, false>(sm, B, t, P.twi, t1, P.p); else local_pass<2, false>(sm, B, t, P.twi, t1, P.p);
    }
    __syncthreads();
  }
}

// grid = npoly * 2^s0 blocks: forward stages s0.. in shared memory, in place in global memory; output in [0, p)
__global__ void s2_fwd_local(uint64_t* __restrict__ data, PrimeC P, int logm, int s0) {
  extern __shared__ __align__(16) uint64_t sm64[];
  const int nloc = 1 << (logm - s0), blk = blockIdx.x & ((1 << s0) - 1);
  uint64_t* d = data + (size_t)blockIdx.x * nloc;                      // blocks of one polynomial are consecutive
  for (int i = threadIdx.x; i < nloc; i += blockDim.x) sm64[swz64(i)] = d[i];
  __syncthreads();
  local_fwd(sm64, logm, s0, blk, P);
  for (int i = threadIdx.x; i < nloc; i += blockDim.x) { uint64_t v = sm64[swz64(i)]; v = v >= 2 * P.p ? v - 2 * P.p : v; d[i] = csub64(v, P.p); }
}

// pointwise product of two transformed polynomials (b broadcast when b_stride == 0) + inverse stages logm-1 .. s0;
// with s0 == 0 also the scaling by m^-1 (otherwise s2_inv_top finishes)
__global__ void s2_mulinv_local(const uint64_t* ah, const uint64_t* bh, size_t b_stride, uint64_t* out /* may alias ah */,
                                PrimeC P, int logm, int s0) {
  extern __shared__ __align__(16) uint64_t sm64[];
  const int nloc = 1 << (logm - s0), blk = blockIdx.x & ((1 << s0) - 1);
  const size_t poly = blockIdx.x >> s0, off = (size_t)blockIdx.x * nloc;
  const uint64_t* b = bh + poly * b_stride + (size_t)blk * nloc;
  for (int i = threadIdx.x; i < nloc; i += blockDim.x) sm64[swz64(i)] = mulmod48(ah[off + i], b[i], P.p, P.mu);
  __syncthreads();
  local_inv(sm64, logm, s0, blk, P);
  for (int i = threadIdx.x; i < nloc; i += blockDim.x) {
    uint64_t v = sm64[swz64(i)];
    if (s0 == 0) v = csub64(shoup64(v, P.minv, P.p), P.p);
    out[off + i] = v;                                                  // [0, 2p) when s2_inv_top follows
  }
}

// inverse transform of data already in the transform domain (no product): local part
__global__ void s2_inv_local(uint64_t* __restrict__ data, PrimeC P, int logm, int s0) {
  extern __shared__ __align__(16) uint64_t sm64[];
  const int nloc = 1 << (logm - s0), blk = blockIdx.x & ((1 << s0) - 1);
  uint64_t* d = data + (size_t)blockIdx.x * nloc;
  for (int i = threadIdx.x; i < nloc; i += blockDim.x) sm64[swz64(i)] = d[i];
  __syncthreads();
  local_inv(sm64, logm, s0, blk, P);
  for (int i = threadIdx.x; i < nloc; i += blockDim.x) {
    uint64_t v = sm64[swz64(i)];
    if (s0 == 0) v = csub64(shoup64(v, P.minv, P.p), P.p);
    d[i] = v;
  }
}

// The multiply-accumulate of an external product in the transform domain, one limb: out[c] = sum_j d[j] . K[j][c]
// (the shape of src/fhe.jl:527-528 with RNS2Number elements, src/rns.jl:51-56).  d: [count][4][m], K: [count][4][2][m].
__global__ void s2_mac8(const uint64_t* __restrict__ d, const uint64_t* __restrict__ K, uint64_t* __restrict__ out,
                        uint64_t p, uint64_t mu, int logm, size_t count) {
  const size_t m = (size_t)1 << logm;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count * m; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t g = idx >> logm, i = idx & (m - 1);
    uint64_t sa = 0, sb = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t dj = d[(g * 4 + j) * m + i];
      sa += mulmod48(dj, K[((g * 4 + j) * 2) * m + i], p, mu);          // four terms below p: no overflow
      sb += mulmod48(dj, K[((g * 4 + j) * 2 + 1) * m + i], p, mu);
    }
    out[(g * 2) * m + i] = sa % p; out[(g * 2 + 1) * m + i] = sb % p;
  }
}

// wide integers (lo, hi) below Q = B B' -> the two residues (RNS2Number(x, m1, m2), src/rns.jl:16-18)
__global__ void s2_split_wide(const uint64_t* __restrict__ wide, uint64_t* __restrict__ v1, uint64_t* __restrict__ v2,
                              uint64_t B, uint64_t Bp, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const u128 x = (u128)wide[2 * i] | ((u128)wide[2 * i + 1] << 64);
  v1[i] = (uint64_t)(x % B); v2[i] = (uint64_t)(x % Bp);
}

// BootstrapKey rows (src/fhe2.jl:122-126): key[i][j] = (a_j, a_j s + e_j) + s_i G[j,:], G = [1 0; B 0; 0 1; 0 B] as
// RNS2Numbers, i.e. B -> (0, B mod B').  a, prod: planar [2][rows*4][m]; out: [rows][4][2][m][2] = (v1, v2) pairs.
__global__ void s2_key_assemble(const uint64_t* __restrict__ a, const uint64_t* __restrict__ prod, const int64_t* __restrict__ e,
                                const uint8_t* __restrict__ sk, uint64_t B, uint64_t Bp, int logm, int row0, size_t rows,
                                uint64_t* __restrict__ out) {
  const size_t m = (size_t)1 << logm, idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x, total = rows * 4 * m;
  if (idx >= total) return;
  const size_t k = idx & (m - 1), j = (idx >> logm) & 3, i = idx >> (logm + 2);
  const uint64_t mod[2] = {B, Bp};
  const int64_t ev = e[idx];
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const uint64_t p = mod[l];
    uint64_t av = a[(size_t)l * total + idx], bv = prod[(size_t)l * total + idx];
    bv = ev >= 0 ? csub64(bv + (uint64_t)ev, p) : (bv >= (uint64_t)(-ev) ? bv - (uint64_t)(-ev) : bv + p - (uint64_t)(-ev));
    if (k == 0 && sk[row0 + i]) {
      const uint64_t g = (j & 1) ? B % p : 1;
      if (j < 2) av = csub64(av + g, p); else bv = csub64(bv + g, p);
    }
    uint64_t* o = out + ((((i * 4 + j) * 2) * m + k) * 2) + l;
    o[0] = av; o[2 * m] = bv;
  }
}

}  // namespace

// =========================================================================================================
struct sgfhe_s2_ctx {
  int device = 0, k = 0, logm = 0;
  sgfhe_scheme2_params prm;
  PrimeC pc[2];
  Tw64* d_tw = nullptr;                 // [2 primes][2 directions][m]
  uint8_t* d_arena = nullptr; size_t arena_bytes = 0;
};

static int s2_fail(int code, const std::string& msg) { sgfhe_set_error_(msg.c_str()); return code; }
#define S2CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return s2_fail(SGFHE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

static uint64_t h_root64(uint64_t p, uint64_t m) {            // psi with psi^m = -1 mod p
  for (uint64_t g = 2;; ++g) {
    const uint64_t w = h_powmod64(g, (p - 1) / (2 * m), p);
    if (h_powmod64(w, m, p) == p - 1) return w;
  }
}
static Tw64 h_tw(uint64_t w, uint64_t p) { Tw64 t; t.w = w; t.wsh = (uint64_t)(((u128)w << 64) / p); return t; }

static int s2_arena(sgfhe_s2_ctx* c, size_t bytes) {
  if (bytes <= c->arena_bytes) return SGFHE_OK;
  cudaFree(c->d_arena); c->d_arena = nullptr; c->arena_bytes = 0;
  if (cudaMalloc(&c->d_arena, bytes) != cudaSuccess) return s2_fail(SGFHE_ERR_NOMEM, "cudaMalloc of the scheme-2 staging arena failed");
  c->arena_bytes = bytes;
  return SGFHE_OK;
}

extern "C" int sgfhe_s2_ctx_create(int32_t k, int32_t device, sgfhe_s2_ctx** out) {
  if (!out) return s2_fail(SGFHE_ERR_ARG, "out is NULL");
  *out = nullptr;
  sgfhe_scheme2_params prm;
  int rc = sgfhe_scheme2_params_derive(k, &prm); if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return s2_fail(SGFHE_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return s2_fail(SGFHE_ERR_ARG, "bad device ordinal");
  S2CK(cudaSetDevice(device));
  sgfhe_s2_ctx* c = new sgfhe_s2_ctx();
  c->device = device; c->k = k; c->prm = prm;
  const uint64_t m = prm.m;
  while (((uint64_t)1 << c->logm) < m) ++c->logm;
  std::vector<Tw64> tw((size_t)4 * m);
  const uint64_t mod[2] = {prm.B, prm.Bp};
  if (cudaMalloc(&c->d_tw, tw.size() * sizeof(Tw64)) != cudaSuccess) { delete c; return s2_fail(SGFHE_ERR_NOMEM, "cudaMalloc of twiddles failed"); }
  for (int l = 0; l < 2; ++l) {
    const uint64_t p = mod[l];
    if (p <= ((uint64_t)1 << 32) || p >= ((uint64_t)1 << 48) || (p - 1) % (2 * m)) { cudaFree(c->d_tw); delete c; return s2_fail(SGFHE_ERR_MODULUS, "scheme-2 modulus outside (2^32, 2^48) or not = 1 mod 2m"); }
    const uint64_t psi = h_root64(p, m), psi_inv = h_powmod64(psi, p - 2, p);
    std::vector<uint64_t> pf(m), pi(m);
    pf[0] = pi[0] = 1;
    for (uint64_t i = 1; i < m; ++i) { pf[i] = h_mulmod64(pf[i - 1], psi, p); pi[i] = h_mulmod64(pi[i - 1], psi_inv, p); }
    for (uint64_t i = 0; i < m; ++i) {
      const int r = h_bitrev((int)i, c->logm);
      tw[((size_t)l * 2) * m + i] = h_tw(pf[r], p);
      tw[((size_t)l * 2 + 1) * m + i] = h_tw(pi[r], p);
    }
    c->pc[l].p = p; c->pc[l].mu = (uint64_t)((((u128)1) << 96) / p);
    c->pc[l].minv = h_tw(h_powmod64(m % p, p - 2, p), p);
    c->pc[l].twf = c->d_tw + ((size_t)l * 2) * m; c->pc[l].twi = c->d_tw + ((size_t)l * 2 + 1) * m;
  }
  if (cudaMemcpy(c->d_tw, tw.data(), tw.size() * sizeof(Tw64), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(c->d_tw); delete c; return s2_fail(SGFHE_ERR_CUDA, "twiddle upload failed"); }
  *out = c;
  return SGFHE_OK;
}

extern "C" int sgfhe_s2_ctx_destroy(sgfhe_s2_ctx* c) {
  if (!c) return SGFHE_OK;
  cudaSetDevice(c->device);
  cudaFree(c->d_tw); cudaFree(c->d_arena);
  delete c;
  return SGFHE_OK;
}

// ---- transforms of `npoly` polynomials of one limb, in place -------------------------------------------------------
static void s2_forward(const sgfhe_s2_ctx* c, int limb, uint64_t* d, size_t npoly, cudaStream_t st) {
  const int logm = c->logm, T = logm > S2_LOCAL_LOG ? logm - S2_LOCAL_LOG : 0;
  const PrimeC& P = c->pc[limb];
  const size_t cols = npoly << (logm - T);
  const unsigned tb = 256, gb = (unsigned)((cols + tb - 1) / tb);
  if (T == 1) s2_fwd_top<1><<<gb, tb, 0, st>>>(d, P, logm, npoly);
  else if (T == 2) s2_fwd_top<2><<<gb, tb, 0, st>>>(d, P, logm, npoly);
  else if (T == 3) s2_fwd_top<3><<<gb, tb, 0, st>>>(d, P, logm, npoly);
  if (T) sgfhe_count_launch_();
  const int nloc = 1 << (logm - T);
  s2_fwd_local<<<(unsigned)(npoly << T), nloc / 2 > 512 ? 512 : nloc / 2, (size_t)nloc * 8, st>>>(d, P, logm, T);
  sgfhe_count_launch_();
}
static void s2_inverse_tail(const sgfhe_s2_ctx* c, int limb, uint64_t* d, size_t npoly, cudaStream_t st) {
  const int logm = c->logm, T = logm > S2_LOCAL_LOG ? logm - S2_LOCAL_LOG : 0;
  if (!T) return;
  const PrimeC& P = c->pc[limb];
  const size_t cols = npoly << (logm - T);
  const unsigned tb = 256, gb = (unsigned)((cols + tb - 1) / tb);
  if (T == 1) s2_inv_top<1><<<gb, tb, 0, st>>>(d, P, logm, npoly);
  else if (T == 2) s2_inv_top<2><<<gb, tb, 0, st>>>(d, P, logm, npoly);
  else s2_inv_top<3><<<gb, tb, 0, st>>>(d, P, logm, npoly);
  sgfhe_count_launch_();
}
// out = a * b (b_stride 0: one b for every product); a and b are overwritten by their transforms
static void s2_product(const sgfhe_s2_ctx* c, int limb, uint64_t* a, uint64_t* b, size_t b_stride, bool b_transformed,
                       uint64_t* out, size_t npoly, cudaStream_t st) {
  const int logm = c->logm, T = logm > S2_LOCAL_LOG ? logm - S2_LOCAL_LOG : 0, nloc = 1 << (logm - T);
  s2_forward(c, limb, a, npoly, st);
  if (!b_transformed) s2_forward(c, limb, b, b_stride ? npoly : 1, st);
  s2_mulinv_local<<<(unsigned)(npoly << T), nloc / 2 > 512 ? 512 : nloc / 2, (size_t)nloc * 8, st>>>(a, b, b_stride, out, c->pc[limb], logm, T);
  sgfhe_count_launch_();
  s2_inverse_tail(c, limb, out, npoly, st);
}

extern "C" int sgfhe_s2_params_get(const sgfhe_s2_ctx* c, sgfhe_scheme2_params* out) {
  if (!c || !out) return s2_fail(SGFHE_ERR_ARG, "NULL argument");
  *out = c->prm;
  return SGFHE_OK;
}

extern "C" int sgfhe_s2_polymul_device(sgfhe_s2_ctx* c, int32_t batch, uint64_t* d_a1, uint64_t* d_a2, uint64_t* d_b1,
                                       uint64_t* d_b2, int32_t b_broadcast, uint64_t* d_o1, uint64_t* d_o2, void* stream) {
  if (!c || !d_a1 || !d_a2 || !d_b1 || !d_b2 || !d_o1 || !d_o2) return s2_fail(SGFHE_ERR_ARG, "NULL argument");
  if (batch < 0) return s2_fail(SGFHE_ERR_ARG, "negative batch");
  if (batch == 0) return SGFHE_OK;
  S2CK(cudaSetDevice(c->device));
  const size_t bs = b_broadcast ? 0 : (size_t)1 << c->logm;
  s2_product(c, 0, d_a1, d_b1, bs, false, d_o1, batch, (cudaStream_t)stream);
  s2_product(c, 1, d_a2, d_b2, bs, false, d_o2, batch, (cudaStream_t)stream);
  S2CK(cudaGetLastError());
  return SGFHE_OK;
}

extern "C" int sgfhe_s2_polymul(sgfhe_s2_ctx* c, int32_t batch, const uint64_t* a1, const uint64_t* a2, const uint64_t* b1,
                                const uint64_t* b2, uint64_t* o1, uint64_t* o2) {
  if (!c || !a1 || !a2 || !b1 || !b2 || !o1 || !o2) return s2_fail(SGFHE_ERR_ARG, "NULL argument");
  if (batch < 0) return s2_fail(SGFHE_ERR_ARG, "negative batch");
  if (batch == 0) return SGFHE_OK;
  S2CK(cudaSetDevice(c->device));
  const size_t w = (size_t)batch << c->logm;
  int rc = s2_arena(c, 6 * w * 8); if (rc) return rc;
  uint64_t* d = reinterpret_cast<uint64_t*>(c->d_arena);
  const uint64_t* src[4] = {a1, a2, b1, b2};
  for (int i = 0; i < 4; ++i) S2CK(cudaMemcpyAsync(d + i * w, src[i], w * 8, cudaMemcpyHostToDevice, nullptr));
  rc = sgfhe_s2_polymul_device(c, batch, d, d + w, d + 2 * w, d + 3 * w, 0, d + 4 * w, d + 5 * w, nullptr); if (rc) return rc;
  S2CK(cudaMemcpyAsync(o1, d + 4 * w, w * 8, cudaMemcpyDeviceToHost, nullptr));
  S2CK(cudaMemcpyAsync(o2, d + 5 * w, w * 8, cudaMemcpyDeviceToHost, nullptr));
  S2CK(cudaDeviceSynchronize());
  return SGFHE_OK;
}

extern "C" int sgfhe_s2_ntt_device(sgfhe_s2_ctx* c, int32_t inverse, int32_t count, uint64_t* d_x1, uint64_t* d_x2, void* stream) {
  if (!c || !d_x1 || !d_x2) return s2_fail(SGFHE_ERR_ARG, "NULL argument");
  if (count < 0) return s2_fail(SGFHE_ERR_ARG, "negative count");
  if (count == 0) return SGFHE_OK;
  S2CK(cudaSetDevice(c->device));
  uint64_t* x[2] = {d_x1, d_x2};
  const int logm = c->logm, T = logm > S2_LOCAL_LOG ? logm - S2_LOCAL_LOG : 0, nloc = 1 << (logm - T);
  for (int l = 0; l < 2; ++l) {
    if (!inverse) s2_forward(c, l, x[l], count, (cudaStream_t)stream);
    else {
      s2_inv_local<<<(unsigned)((size_t)count << T), nloc / 2 > 512 ? 512 : nloc / 2, (size_t)nloc * 8, (cudaStream_t)stream>>>(x[l], c->pc[l], logm, T);
      sgfhe_count_launch_();
      s2_inverse_tail(c, l, x[l], count, (cudaStream_t)stream);
    }
  }
  S2CK(cudaGetLastError());
  return SGFHE_OK;
}

extern "C" int sgfhe_s2_mac8_device(sgfhe_s2_ctx* c, int32_t count, const uint64_t* d_d1, const uint64_t* d_d2, const uint64_t* d_k1,
                                    const uint64_t* d_k2, uint64_t* d_o1, uint64_t* d_o2, void* stream) {
  if (!c || !d_d1 || !d_d2 || !d_k1 || !d_k2 || !d_o1 || !d_o2) return s2_fail(SGFHE_ERR_ARG, "NULL argument");
  if (count < 0) return s2_fail(SGFHE_ERR_ARG, "negative count");
  if (count == 0) return SGFHE_OK;
  S2CK(cudaSetDevice(c->device));
  int sms = 0; S2CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  const size_t total = (size_t)count << c->logm;
  size_t blocks = (total + 255) / 256; if (blocks > (size_t)sms * 16) blocks = (size_t)sms * 16;
  s2_mac8<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_d1, d_k1, d_o1, c->pc[0].p, c->pc[0].mu, c->logm, count);
  s2_mac8<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_d2, d_k2, d_o2, c->pc[1].p, c->pc[1].mu, c->logm, count);
  sgfhe_count_launch_(); sgfhe_count_launch_();
  S2CK(cudaGetLastError());
  return SGFHE_OK;
}

extern "C" int sgfhe_s2_bkey_generate(sgfhe_s2_ctx* c, const uint8_t* sk, const uint64_t* a_rand, const int64_t* e_rand,
                                      int32_t row0, int32_t rows, uint64_t* key_out) {
  if (!c || !sk || !a_rand || !e_rand || !key_out) return s2_fail(SGFHE_ERR_ARG, "NULL argument");
  const int n = c->prm.n; const size_t m = c->prm.m;
  if (row0 < 0 || rows < 1 || row0 + rows > n) return s2_fail(SGFHE_ERR_ARG, "rows out of range");
  S2CK(cudaSetDevice(c->device));
  const size_t crow = std::max<size_t>(1, std::min<size_t>(rows, ((size_t)1 << 26) / (4 * m * 16)));   // <= 64 MiB of a_rand per pass
  const size_t np = crow * 4, pw = np * m;                             // polynomials / words per limb and pass
  // arena: sk | ext [2][m] | wide a | a [2][pw] | prod [2][pw] | e | key
  const size_t o_sk = 0, o_ext = 4096, o_wide = o_ext + 2 * m * 8, o_a = o_wide + pw * 16, o_prod = o_a + 2 * pw * 8,
               o_e = o_prod + 2 * pw * 8, o_key = o_e + pw * 8, total = o_key + crow * 8 * m * 16;
  int rc = s2_arena(c, total); if (rc) return rc;
  uint8_t* d = c->d_arena;
  uint64_t* ext = reinterpret_cast<uint64_t*>(d + o_ext);
  uint64_t* a = reinterpret_cast<uint64_t*>(d + o_a);
  uint64_t* prod = reinterpret_cast<uint64_t*>(d + o_prod);
  std::vector<uint64_t> hext(2 * m, 0);
  for (int i = 0; i < n; ++i) hext[i] = hext[m + i] = sk[i] ? 1 : 0;   // resize(polynomial_Q(params, sk.key), m)  (src/fhe2.jl:113)
  S2CK(cudaMemcpy(d + o_sk, sk, n, cudaMemcpyHostToDevice));
  S2CK(cudaMemcpy(ext, hext.data(), 2 * m * 8, cudaMemcpyHostToDevice));
  s2_forward(c, 0, ext, 1, nullptr);
  s2_forward(c, 1, ext + m, 1, nullptr);
  for (int done = 0; done < rows; done += (int)crow) {
    const size_t cnt = std::min<size_t>(crow, rows - done), cp = cnt * 4, cw = cp * m;
    S2CK(cudaMemcpyAsync(d + o_wide, a_rand + (size_t)done * 4 * m * 2, cw * 16, cudaMemcpyHostToDevice, nullptr));
    S2CK(cudaMemcpyAsync(d + o_e, e_rand + (size_t)done * 4 * m, cw * 8, cudaMemcpyHostToDevice, nullptr));
    s2_split_wide<<<(unsigned)((cw + 255) / 256), 256>>>(reinterpret_cast<uint64_t*>(d + o_wide), a, a + pw, c->prm.B, c->prm.Bp, cw);
    sgfhe_count_launch_();
    // the transforms overwrite their input: keep a_j (it is half of the key) and transform a copy in `prod`
    S2CK(cudaMemcpyAsync(prod, a, cw * 8, cudaMemcpyDeviceToDevice, nullptr));
    S2CK(cudaMemcpyAsync(prod + pw, a + pw, cw * 8, cudaMemcpyDeviceToDevice, nullptr));
    s2_product(c, 0, prod, ext, 0, true, prod, cp, nullptr);
    s2_product(c, 1, prod + pw, ext + m, 0, true, prod + pw, cp, nullptr);
    // s2_key_assemble indexes limb l at l * (cnt*4*m): compact the second limb when the pass is partial
    if (cnt != crow) {
      S2CK(cudaMemcpyAsync(a + cw, a + pw, cw * 8, cudaMemcpyDeviceToDevice, nullptr));
      S2CK(cudaMemcpyAsync(prod + cw, prod + pw, cw * 8, cudaMemcpyDeviceToDevice, nullptr));
    }
    s2_key_assemble<<<(unsigned)((cw + 255) / 256), 256>>>(a, prod, reinterpret_cast<int64_t*>(d + o_e), d + o_sk, c->prm.B, c->prm.Bp,
                                                         c->logm, row0 + done, cnt, reinterpret_cast<uint64_t*>(d + o_key));
    sgfhe_count_launch_();
    S2CK(cudaGetLastError());
    S2CK(cudaMemcpyAsync(key_out + (size_t)done * 8 * m * 2, d + o_key, cnt * 8 * m * 16, cudaMemcpyDeviceToHost, nullptr));
    S2CK(cudaDeviceSynchronize());
  }
  return SGFHE_OK;
}
