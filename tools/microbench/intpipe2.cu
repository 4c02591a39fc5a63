// Second integer-pipe microbenchmark: pure IMAD.WIDE (no addend), IMAD+IADD3 overlap,
// IMAD.WIDE+IADD3 ratios.  Same harness as intpipe.cu.  Loop-only SASS mix is checked
// with loopmix.py before trusting a line.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
constexpr int ITERS = 2048, CHAINS = 8;
enum Op { WIDE_RZ, LO_IADD_1_1, LO_IADD_1_2, WIDE_IADD_1_1, WIDE_IADD_1_2, WIDE_IADD_1_3, WIDE_LO_1_1, HI_IADD_1_1, DFMA_WIDE_1_1, DFMA_IADD_1_1, DFMA_WIDE_IADD, LO_LOP_SHF };

#define ADD3(x, y, z) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x) : "r"(y), "r"(z))
#define WIDE(wv) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(wv) : "r"((uint32_t)(wv)), "r"((uint32_t)((wv) >> 32)))
#define LO(x, y, z) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define HI(x, y, z) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define DF(x) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(1.0000001), "d"(1e-12))

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t a[CHAINS], b[CHAINS], c[CHAINS]; uint64_t w[CHAINS]; double d[CHAINS];
  uint32_t m0 = seed * 2654435761u + threadIdx.x, m1 = seed ^ 0x9e3779b9u;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { a[i] = m0 + i * 77u; b[i] = m1 + i * 13u; c[i] = a[i] ^ b[i]; w[i] = ((uint64_t)a[i] << 32) | (b[i] | 1u); d[i] = 1.0 + 1e-9 * (double)(i + threadIdx.x); }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) {
        if (OP == WIDE_RZ) { WIDE(w[i]); }
        else if (OP == LO_IADD_1_1) { LO(a[i], m0, m1); ADD3(b[i], c[i], m0); }
        else if (OP == LO_IADD_1_2) { LO(a[i], m0, m1); ADD3(b[i], c[i], m0); ADD3(c[i], b[i], m1); }
        else if (OP == WIDE_IADD_1_1) { WIDE(w[i]); ADD3(b[i], c[i], m0); }
        else if (OP == WIDE_IADD_1_2) { WIDE(w[i]); ADD3(b[i], c[i], m0); ADD3(c[i], b[i], m1); }
        else if (OP == WIDE_IADD_1_3) { WIDE(w[i]); ADD3(b[i], c[i], m0); ADD3(c[i], b[i], m1); ADD3(a[i], a[i], m1); }
        else if (OP == WIDE_LO_1_1) { WIDE(w[i]); LO(a[i], m0, m1); }
        else if (OP == HI_IADD_1_1) { HI(a[i], m0, m1); ADD3(b[i], c[i], m0); }
        else if (OP == DFMA_WIDE_1_1) { DF(d[i]); WIDE(w[i]); }
        else if (OP == DFMA_IADD_1_1) { DF(d[i]); ADD3(b[i], c[i], m0); }
        else if (OP == DFMA_WIDE_IADD) { DF(d[i]); WIDE(w[i]); ADD3(b[i], c[i], m0); ADD3(c[i], b[i], m1); }
        else if (OP == LO_LOP_SHF) { LO(a[i], m0, m1); asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(b[i]) : "r"(c[i]), "r"(m0)); asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(c[i]) : "r"(b[i])); }
      }
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc ^= a[i] ^ b[i] ^ c[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)__double2ll_rz(d[i] * 1e6);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(const char* name, int nsm, uint32_t* out, long long* cyc, bool last) {
  for (int i = 0; i < 3; ++i) k<OP><<<nsm, 1024>>>(out, 1234u + i, cyc);
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < 5; ++r) {
    k<OP><<<nsm, 1024>>>(out, 99u + r, cyc);
    CK(cudaDeviceSynchronize());
    long long h[1024]; CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += (double)h[i]; avg /= nsm;
    if (avg < best) best = avg;
  }
  // SM clocks per (one slot of all 1024 threads); 1 "unit" = 16 clk = one full-rate (64 lanes/clk/SM) instruction
  double clk_per_slot = best / ((double)ITERS * 4 * CHAINS);
  printf("  \"%s\": {\"sm_clk_per_slot_1024thr\": %.3f, \"units_per_slot\": %.3f}%s\n", name, clk_per_slot, clk_per_slot / 16.0, last ? "" : ",");
  return 0;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int nsm = p.multiProcessorCount;
  uint32_t* out; long long* cyc; CK(cudaMalloc(&out, 4 * nsm * 1024)); CK(cudaMalloc(&cyc, 8 * 1024));
  printf("{\n  \"note\": \"units_per_slot: 1.0 = one instruction at 64 lanes/clk/SM\",\n");
  run<WIDE_RZ>("wide_rz", nsm, out, cyc, false);
  run<LO_IADD_1_1>("lo+iadd3", nsm, out, cyc, false);
  run<LO_IADD_1_2>("lo+2iadd3", nsm, out, cyc, false);
  run<WIDE_IADD_1_1>("wide+iadd3", nsm, out, cyc, false);
  run<WIDE_IADD_1_2>("wide+2iadd3", nsm, out, cyc, false);
  run<WIDE_IADD_1_3>("wide+3iadd3", nsm, out, cyc, false);
  run<WIDE_LO_1_1>("wide+lo", nsm, out, cyc, false);
  run<HI_IADD_1_1>("hi+iadd3", nsm, out, cyc, false);
  run<DFMA_WIDE_1_1>("dfma+wide", nsm, out, cyc, false);
  run<DFMA_IADD_1_1>("dfma+iadd3", nsm, out, cyc, false);
  run<DFMA_WIDE_IADD>("dfma+wide+2iadd3", nsm, out, cyc, false);
  run<LO_LOP_SHF>("lo+lop3+shf", nsm, out, cyc, true);
  printf("}\n");
  return 0;
}
