// Third pipe microbenchmark (round 2): does FP64 work overlap with the integer butterfly instruction mix when both sit in
// ONE instruction stream?  Same harness as intpipe2.cu, 512 threads per SM (the occupancy of the gate kernel) and 1024.
//   units_per_slot: 1.0 = one instruction at 64 lanes/clk/SM.  Perfect overlap of a DFMA with integer work adds 0, none adds 1.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
constexpr int ITERS = 2048, CHAINS = 8;
enum Op { DFMA_ONLY, LO_ONLY, HI_ONLY, DFMA_LO, DFMA_HI, DFMA_LO_IADD, BFLY, BFLY_DFMA1, BFLY_DFMA2, BFLY_DFMA4 };

#define ADD3(x, y, z) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x) : "r"(y), "r"(z))
#define LO(x, y, z) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define HI(x, y, z) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define DF(x) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(1.0000001), "d"(1e-12))
// Harvey butterfly as in device_math.cuh: VIADDMNMX, IMAD.HI, 2 IMAD, 2 IADD3
__device__ __forceinline__ void bfly(uint32_t& x, uint32_t& y, uint32_t w, uint32_t wsh, uint32_t p, uint32_t z) {
  const uint32_t p2 = 2 * p, xr = min(x, x - p2);
  const uint32_t t = y * w - __umulhi(y, wsh) * p;
  x = xr + t + z; y = xr - t + p2;
}

template <int OP, int TT>
__global__ void __launch_bounds__(TT, 1) k(uint32_t* out, uint32_t seed, uint32_t z, long long* cyc) {
  uint32_t a[CHAINS], b[CHAINS]; double d[CHAINS];
  uint32_t m0 = seed * 2654435761u + threadIdx.x, m1 = seed ^ 0x9e3779b9u;
  const uint32_t p = 1073479681u;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { a[i] = (m0 + i * 77u) % p; b[i] = (m1 + i * 13u) % p; d[i] = 1.0 + 1e-9 * (double)(i + threadIdx.x); }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) {
        if (OP == DFMA_ONLY) { DF(d[i]); }
        else if (OP == LO_ONLY) { LO(a[i], m0, m1); }
        else if (OP == HI_ONLY) { HI(a[i], m0, m1); }
        else if (OP == DFMA_LO) { DF(d[i]); LO(a[i], m0, m1); }
        else if (OP == DFMA_HI) { DF(d[i]); HI(a[i], m0, m1); }
        else if (OP == DFMA_LO_IADD) { DF(d[i]); LO(a[i], m0, m1); ADD3(b[i], a[i], m0); }
        else if (OP == BFLY) { bfly(a[i], b[i], m0 | 1u, m1, p, z); }
        else if (OP == BFLY_DFMA1) { bfly(a[i], b[i], m0 | 1u, m1, p, z); DF(d[i]); }
        else if (OP == BFLY_DFMA2) { bfly(a[i], b[i], m0 | 1u, m1, p, z); DF(d[i]); DF(d[(i + 1) % CHAINS]); }
        else if (OP == BFLY_DFMA4) { bfly(a[i], b[i], m0 | 1u, m1, p, z); DF(d[i]); DF(d[(i + 1) % CHAINS]); DF(d[(i + 2) % CHAINS]); DF(d[(i + 3) % CHAINS]); }
      }
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc ^= a[i] ^ b[i] ^ (uint32_t)__double2ll_rz(d[i] * 1e6);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP, int TT>
int run(const char* name, int nsm, uint32_t* out, long long* cyc, bool last) {
  for (int i = 0; i < 3; ++i) k<OP, TT><<<nsm, TT>>>(out, 1234u + i, 0u, cyc);
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < 5; ++r) {
    k<OP, TT><<<nsm, TT>>>(out, 99u + r, 0u, cyc);
    CK(cudaDeviceSynchronize());
    long long h[1024]; CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += (double)h[i]; avg /= nsm;
    if (avg < best) best = avg;
  }
  // SM clocks per slot of all TT threads; one full-rate (64 lanes/clk/SM) instruction for TT threads takes TT/64 clocks
  const double clk_per_slot = best / ((double)ITERS * 4 * CHAINS);
  printf("  \"%s@%d\": {\"sm_clk_per_slot\": %.3f, \"units_per_slot\": %.3f}%s\n", name, TT, clk_per_slot, clk_per_slot / (TT / 64.0), last ? "" : ",");
  return 0;
}
template <int TT>
int all(int nsm, uint32_t* out, long long* cyc, bool last) {
  run<DFMA_ONLY, TT>("dfma", nsm, out, cyc, false);
  run<LO_ONLY, TT>("lo", nsm, out, cyc, false);
  run<HI_ONLY, TT>("hi", nsm, out, cyc, false);
  run<DFMA_LO, TT>("dfma+lo", nsm, out, cyc, false);
  run<DFMA_HI, TT>("dfma+hi", nsm, out, cyc, false);
  run<DFMA_LO_IADD, TT>("dfma+lo+iadd3", nsm, out, cyc, false);
  run<BFLY, TT>("bfly", nsm, out, cyc, false);
  run<BFLY_DFMA1, TT>("bfly+1dfma", nsm, out, cyc, false);
  run<BFLY_DFMA2, TT>("bfly+2dfma", nsm, out, cyc, false);
  run<BFLY_DFMA4, TT>("bfly+4dfma", nsm, out, cyc, last);
  return 0;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int nsm = p.multiProcessorCount;
  uint32_t* out; long long* cyc; CK(cudaMalloc(&out, 4 * nsm * 1024)); CK(cudaMalloc(&cyc, 8 * 1024));
  printf("{\n  \"note\": \"units_per_slot: 1.0 = one instruction at 64 lanes/clk/SM; a DFMA that overlaps integer work completely adds 0\",\n");
  all<512>(nsm, out, cyc, false);
  all<1024>(nsm, out, cyc, true);
  printf("}\n");
  return 0;
}
