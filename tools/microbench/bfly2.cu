// Second butterfly-block microbenchmark (where do the cycles go: pipes, shared memory or barriers?)
// Butterfly-block microbenchmark for B200 (sm_100a): one radix-8 register block per thread over polynomials in
// shared memory, exactly the shape of pass8_v4 in sgfhe_cuda.cu (512 threads, one CTA per SM, 128 KiB of data).
// Compares the Shoup quotient floor(y wsh / 2^32) computed by IMAD.HI with the bit-identical
//     lo32( fma.rz.f64( (2^52 + y), double(wsh), 2^84 - 2^52 wsh ) )
// on the otherwise idle FP64 pipe, selectable per butterfly level (bit l of MASK -> level l uses DFMA).
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o bfly bfly.cu
// Run:   ./bfly > bfly.json
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int M = 8192, T = 512, ITERS = 256;
constexpr uint32_t P = 1073479681u;   // any odd modulus < 2^30; arithmetic identity only, not a transform

struct Tw { uint32_t w, wsh; double C, K; };

__device__ __forceinline__ uint32_t hi_dfma(uint32_t y, double C, double K) {
  double d, r; uint32_t lo, hi;
  asm("mov.b64 %0, {%1, %2};" : "=d"(d) : "r"(y), "r"(0x43300000u));
  asm("fma.rz.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(d), "d"(C), "d"(K));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "d"(r));
  return lo;
}
template <bool D>
__device__ __forceinline__ uint32_t shoup(uint32_t x, const Tw& t, uint32_t p) {
  if (D) return x * t.w - hi_dfma(x, t.C, t.K) * p;
  return x * t.w - __umulhi(x, t.wsh) * p;
}
template <bool D>
__device__ __forceinline__ void ct(uint32_t& x, uint32_t& y, const Tw& w, uint32_t p, uint32_t p2, uint32_t z) {
  const uint32_t xr = min(x, x - p2);
  const uint32_t t = shoup<D>(y, w, p);
  x = xr + t + z; y = xr - t + p2;
}
template <bool D>
__device__ __forceinline__ void gs(uint32_t& x, uint32_t& y, const Tw& w, uint32_t p, uint32_t p2, uint32_t z) {
  const uint32_t s = x + y + z, d = x - y + p2;
  x = min(s, s - p2);
  y = shoup<D>(d, w, p);
}
template <bool D>
__device__ __forceinline__ void cvt(Tw& t) {
  if (D) {
    double dd; asm("mov.b64 %0, {%1, %2};" : "=d"(dd) : "r"(t.wsh), "r"(0x43300000u));
    t.C = dd - 4503599627370496.0;                                   // double(wsh), exact
    t.K = fma(-4503599627370496.0, t.C, 19342813113834066795298816.0);   // 2^84 - 2^52 wsh, exact
  }
}


template <int MASK, bool FWD>
__device__ __forceinline__ void block8(uint32_t (&x)[8], const Tw (&w)[7], uint32_t p, uint32_t p2, uint32_t z) {
  if (FWD) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ct<(MASK & 1) != 0>(x[kk], x[kk + 4], w[0], p, p2, z);
#pragma unroll
    for (int gg = 0; gg < 2; ++gg)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) ct<(MASK & 2) != 0>(x[gg * 4 + kk], x[gg * 4 + kk + 2], w[1 + gg], p, p2, z);
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) ct<(MASK & 4) != 0>(x[gg * 2], x[gg * 2 + 1], w[3 + gg], p, p2, z);
  } else {
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) gs<(MASK & 4) != 0>(x[gg * 2], x[gg * 2 + 1], w[3 + gg], p, p2, z);
#pragma unroll
    for (int gg = 0; gg < 2; ++gg)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) gs<(MASK & 2) != 0>(x[gg * 4 + kk], x[gg * 4 + kk + 2], w[1 + gg], p, p2, z);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) gs<(MASK & 1) != 0>(x[kk], x[kk + 4], w[0], p, p2, z);
  }
}

// MODE 0: data stays in registers (pipe limit as scheduled by ptxas); 1: LDS/STS per block, no barrier;
// 2: + __syncthreads per iteration; 3: + 64-thread named barrier per iteration.
template <int MASK, int NPOLY, bool FWD, int MODE, int TT>
__global__ void __launch_bounds__(TT, 1) k(uint32_t* g, const uint2* gtab, uint32_t z, long long* cyc) {
  extern __shared__ __align__(16) uint32_t sm[];
  uint2* tab = reinterpret_cast<uint2*>(sm + 4 * M);
  for (int i = threadIdx.x; i < 4 * M; i += TT) sm[i] = g[(size_t)blockIdx.x * 4 * M + i];
  for (int i = threadIdx.x; i < M; i += TT) tab[i] = gtab[i];
  __syncthreads();
  const uint32_t p = P, p2 = 2 * P;
  constexpr int STRIDE = M / 8;                      // 1024: element j of the block lives at tid' + j * STRIDE
  uint32_t xr[MODE == 0 ? NPOLY : 1][8];
  if (MODE == 0) {
#pragma unroll
    for (int poly = 0; poly < NPOLY; ++poly)
#pragma unroll
      for (int j = 0; j < 8; ++j) xr[poly][j] = sm[poly * M + (threadIdx.x % STRIDE) + j * STRIDE];
  }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    const int t1 = 1024 + ((threadIdx.x + it) & 1023);
    Tw w[7];
    {
      const uint2 w0 = tab[t1];
      const uint4 a = *reinterpret_cast<const uint4*>(&tab[2 * t1]);
      const uint4 b0 = *reinterpret_cast<const uint4*>(&tab[4 * t1]);
      const uint4 b1 = *reinterpret_cast<const uint4*>(&tab[4 * t1 + 2]);
      w[0].w = w0.x; w[0].wsh = w0.y; w[1].w = a.x; w[1].wsh = a.y; w[2].w = a.z; w[2].wsh = a.w;
      w[3].w = b0.x; w[3].wsh = b0.y; w[4].w = b0.z; w[4].wsh = b0.w; w[5].w = b1.x; w[5].wsh = b1.y; w[6].w = b1.z; w[6].wsh = b1.w;
      cvt<(MASK & 1) != 0>(w[0]);
      cvt<(MASK & 2) != 0>(w[1]); cvt<(MASK & 2) != 0>(w[2]);
      cvt<(MASK & 4) != 0>(w[3]); cvt<(MASK & 4) != 0>(w[4]); cvt<(MASK & 4) != 0>(w[5]); cvt<(MASK & 4) != 0>(w[6]);
    }
    const int e0 = (threadIdx.x + (MODE >= 2 ? it : 0)) % STRIDE;   // with a barrier the blocks may move between threads
#pragma unroll
    for (int poly = 0; poly < NPOLY; ++poly) {
      if (MODE == 0) {
        block8<MASK, FWD>(xr[poly], w, p, p2, z);
      } else {
        uint32_t x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = sm[poly * M + e0 + j * STRIDE];
        block8<MASK, FWD>(x, w, p, p2, z);
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[poly * M + e0 + j * STRIDE] = x[j];
      }
    }
    if (MODE == 2) __syncthreads();
    if (MODE == 3) asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)(threadIdx.x >> 6)) : "memory");
  }
  const long long t1c = clock64();
  if (MODE == 0) {
#pragma unroll
    for (int poly = 0; poly < NPOLY; ++poly)
#pragma unroll
      for (int j = 0; j < 8; ++j) sm[poly * M + (threadIdx.x % STRIDE) + j * STRIDE] = xr[poly][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * M; i += TT) g[(size_t)blockIdx.x * 4 * M + i] = sm[i];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1c - t0;
}

static uint32_t* d_g; static uint32_t* d_g0; static uint2* d_tab; static long long* d_cyc; static int nsm;

template <int MASK, int NPOLY, bool FWD, int MODE, int TT>
int run(const char* name, bool last) {
  const size_t smem = 4 * M * 4 + M * 8;
  CK(cudaFuncSetAttribute(k<MASK, NPOLY, FWD, MODE, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  double best = 1e30;
  for (int r = 0; r < 4; ++r) {
    CK(cudaMemcpy(d_g, d_g0, (size_t)nsm * 4 * M * 4, cudaMemcpyDeviceToDevice));
    k<MASK, NPOLY, FWD, MODE, TT><<<nsm, TT, smem>>>(d_g, d_tab, 0u, d_cyc);
    CK(cudaDeviceSynchronize());
    long long h[1024]; CK(cudaMemcpy(h, d_cyc, 8 * nsm, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += (double)h[i]; avg /= nsm;
    if (avg < best) best = avg;
  }
  // per iteration and SMSP: (TT / 128) warps x NPOLY x 12 warp-butterflies
  const double per_it = best / ITERS;
  printf("  \"%s\": {\"clk_per_iter\": %.1f, \"clk_per_warp_bfly_per_smsp\": %.3f}%s\n", name, per_it,
         per_it / (12.0 * NPOLY * (TT / 128)), last ? "" : ",");
  return 0;
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0)); nsm = pr.multiProcessorCount;
  CK(cudaMalloc(&d_g, (size_t)nsm * 4 * M * 4)); CK(cudaMalloc(&d_g0, (size_t)nsm * 4 * M * 4));
  CK(cudaMalloc(&d_tab, M * 8)); CK(cudaMalloc(&d_cyc, 8 * 1024));
  {
    uint32_t* h = new uint32_t[(size_t)nsm * 4 * M]; uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < (size_t)nsm * 4 * M; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)(s % P); }
    CK(cudaMemcpy(d_g0, h, (size_t)nsm * 4 * M * 4, cudaMemcpyHostToDevice));
    uint2* t = new uint2[M];
    for (int i = 0; i < M; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; const uint32_t w = (uint32_t)(s % P); t[i] = make_uint2(w, (uint32_t)(((uint64_t)w << 32) / P)); }
    CK(cudaMemcpy(d_tab, t, M * 8, cudaMemcpyHostToDevice));
  }
  printf("{\n  \"gpu\": \"%s\", \"note\": \"name = dir npoly _ mask (DFMA levels) _ mode (0 regs, 1 smem, 2 smem+syncthreads, 3 smem+bar64) _ threads\",\n", pr.name);
#define R(MASK, NP, FWD, MODE, TT, LAST) if (run<MASK, NP, FWD, MODE, TT>(#FWD "_np" #NP "_mask" #MASK "_mode" #MODE "_t" #TT, LAST)) return 1;
  R(0, 1, true, 0, 512, false) R(0, 2, true, 0, 512, false) R(0, 4, true, 0, 512, false)
  R(7, 1, true, 0, 512, false) R(7, 2, true, 0, 512, false) R(7, 4, true, 0, 512, false)
  R(1, 2, true, 0, 512, false) R(3, 2, true, 0, 512, false)
  R(0, 2, true, 0, 1024, false) R(7, 2, true, 0, 1024, false) R(0, 1, true, 0, 1024, false)
  R(0, 4, true, 1, 512, false) R(0, 4, true, 2, 512, false) R(0, 4, true, 3, 512, false)
  R(7, 4, true, 1, 512, false) R(3, 4, true, 1, 512, false)
  R(0, 4, true, 1, 1024, false) R(0, 4, true, 2, 1024, false) R(7, 4, true, 1, 1024, false)
  R(0, 2, false, 0, 512, false) R(7, 2, false, 0, 512, false) R(0, 2, false, 1, 512, false) R(0, 2, false, 2, 512, false)
  R(0, 2, false, 2, 1024, true)
  printf("}\n");
  return 0;
}
