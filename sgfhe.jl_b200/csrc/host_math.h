// host_math.h -- host-side integer helpers for libsgfhe_cuda: Params derivation (reference
// src/fhe.jl:43-97, src/utils.jl:7-28), the RNS basis of 30-bit NTT primes, CRT constants.
// Product code: does not use anything under oracle/.
#pragma once
#include <cstdint>
#include <vector>

namespace sgfhe {

typedef unsigned __int128 u128;

inline u128 h_mulmod(u128 a, u128 b, u128 m) {   // shift-add; setup only (m < 2^127)
  u128 r = 0; a %= m; b %= m;
  while (b) { if (b & 1) { r += a; if (r >= m) r -= m; } a <<= 1; if (a >= m) a -= m; b >>= 1; }
  return r;
}
inline u128 h_powmod(u128 a, u128 e, u128 m) {
  u128 r = 1 % m; a %= m;
  while (e) { if (e & 1) r = h_mulmod(r, a, m); a = h_mulmod(a, a, m); e >>= 1; }
  return r;
}
inline uint64_t h_mulmod64(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)((u128)a * b % m); }
inline uint64_t h_powmod64(uint64_t a, uint64_t e, uint64_t m) {
  uint64_t r = 1 % m; a %= m;
  while (e) { if (e & 1) r = h_mulmod64(r, a, m); a = h_mulmod64(a, a, m); e >>= 1; }
  return r;
}

// Primes.isprime stand-in (src/utils.jl:19): strong-probable-prime test to 24 prime bases.
inline bool h_isprime(u128 x) {
  static const int bs[] = {2,3,5,7,11,13,17,19,23,29,31,37,41,43,47,53,59,61,67,71,73,79,83,89};
  if (x < 2) return false;
  for (int b : bs) { if (x == (u128)b) return true; if (x % b == 0) return false; }
  u128 d = x - 1; int s = 0;
  while (!(d & 1)) { d >>= 1; ++s; }
  for (int b : bs) {
    u128 y = h_powmod(b, d, x);
    if (y == 1 || y == x - 1) continue;
    bool comp = true;
    for (int k = 1; k < s && comp; ++k) { y = h_mulmod(y, y, x); if (y == x - 1) comp = false; }
    if (comp) return false;
  }
  return true;
}

// src/utils.jl:7-28; qmax == 0 means `nothing`.  Returns 0 when nothing is found.
inline u128 h_find_modulus(u128 n, u128 qmin, u128 qmax) {
  u128 j = (qmin - 1 + n - 1) / n;
  for (;;) {
    u128 q = j * n + 1;
    if (qmax != 0 && q > qmax) return 0;
    if (h_isprime(q)) return q;
    ++j;
  }
}

struct HostParams {          // src/fhe.jl:27-99
  int n, t, m, logm, logr, kB;
  uint64_t r, q, Dr, Dq;
  u128 Q, B, DQ;
};

// returns 0 ok, -1 bad n (src/fhe.jl:45-46), -2 no modulus / too large (src/utils.jl:26, src/fhe.jl:77)
inline int h_params(int n, HostParams* P) {
  if (n < 64 || (n & (n - 1))) return -1;
  if (n > 2048) return -2;
  u128 bn = n, r = bn * 16;
  u128 q = h_find_modulus(2 * bn, r * bn, 0);
  int logr = 0; while (((u128)1 << logr) < r) ++logr;
  u128 m = r / 2, r4n2 = r * r * r * r * bn * bn;
  u128 Q = h_find_modulus(2 * m, r4n2 * 1220, r4n2 * 1225);
  if (!Q) return -2;
  P->n = n; P->t = logr - 1; P->m = (int)m; P->logr = logr; P->logm = logr - 1;
  P->r = (uint64_t)r; P->q = (uint64_t)q; P->Dr = (uint64_t)(r / 4); P->Dq = (uint64_t)(q / 4);
  P->Q = Q; P->B = r * r * bn * 35; P->DQ = Q / 8;
  int logn = 0; while ((1 << logn) < n) ++logn;
  P->kB = 2 * logr + logn;           // B = 35 * 2^kB
  return 0;
}

// 30-bit primes p = k*2^15 + 1 (so that 2m | p-1 for every m <= 2^14), descending from 2^30.
inline std::vector<uint32_t> h_rns_primes(int count) {
  std::vector<uint32_t> out;
  for (uint64_t k = (1u << 15) - 1; k > 0 && (int)out.size() < count; --k) {
    uint64_t p = (k << 15) + 1;
    bool prime = true;
    for (uint64_t d = 3; d * d <= p; d += 2) if (p % d == 0) { prime = false; break; }
    if (prime) out.push_back((uint32_t)p);
  }
  return out;
}

inline uint32_t h_root_2m(uint32_t p, int m) {   // psi with psi^m = -1 mod p
  for (uint64_t g = 2;; ++g) {
    uint64_t w = h_powmod64(g, (p - 1) / (2 * (uint64_t)m), p);
    if (h_powmod64(w, m, p) == p - 1) return (uint32_t)w;
  }
}

inline int h_bitrev(int x, int bits) { int r = 0; for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; } return r; }

// number of bits of x
inline int h_bits(u128 x) { int b = 0; while (x) { ++b; x >>= 1; } return b; }

}  // namespace sgfhe
