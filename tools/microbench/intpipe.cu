// Integer-pipe microbenchmark for B200 (sm_100a).
//
// the binding roofline of the bootstrap path is the integer
// pipe, and MEASURED_PEAKS.json carries no integer peak.  This program measures
// warp-instruction throughput per SM per clock for the instruction classes the
// multi-limb Montgomery kernels are built from (IMAD, IMAD.WIDE, IMAD.HI,
// IADD3, LOP3, SHF, mixes of them, DFMA) and prints one JSON object.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o intpipe intpipe.cu
// Run:    ./intpipe > int_peaks.json

#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 2048;   // loop trips
constexpr int CHAINS = 8;     // independent dependency chains per thread

// Each body executes CHAINS independent ops of one class per unrolled slot.
enum Op { IMAD_LO, IMAD_WIDE, IMAD_HI, IADD3, LOP3, SHF, MIX_WIDE_IADD, MIX_WIDE_2IADD, DFMA, MIX_LO_HI, FFMA, MIX_WIDE_FFMA, ACC_WIDE, ACC_WIDE2 };

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t a[CHAINS], b[CHAINS];
  uint64_t w[CHAINS], cw[CHAINS];
  double d[CHAINS];
  float f[CHAINS];
  uint32_t m0 = seed * 2654435761u + threadIdx.x, m1 = seed ^ 0x9e3779b9u;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    a[i] = m0 + i * 77u; b[i] = m1 + i * 13u; w[i] = ((uint64_t)a[i] << 32) | b[i]; cw[i] = w[i] * 0x9e3779b97f4a7c15ull + seed;
    d[i] = 1.0 + 1e-9 * (double)(a[i] & 1023); f[i] = 1.0f + 1e-6f * (float)(b[i] & 1023);
  }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) {
        if (OP == IMAD_LO) {
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m0), "r"(m1));
        } else if (OP == IMAD_WIDE) {
          asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)), "l"(cw[i]));
        } else if (OP == IMAD_HI) {
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m0), "r"(m1));
        } else if (OP == IADD3) {
          asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(b[i]), "r"(m0));
          asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(b[i]) : "r"(a[i]), "r"(m1));
        } else if (OP == LOP3) {
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(m1));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(b[i]) : "r"(a[i]), "r"(m0));
        } else if (OP == SHF) {
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(b[i]) : "r"(a[i]));
        } else if (OP == MIX_WIDE_IADD) {
          asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)), "l"(cw[i]));
          asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(b[i]), "r"((uint32_t)(w[i]>>32)));
        } else if (OP == MIX_WIDE_2IADD) {
          asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)), "l"(cw[i]));
          asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(b[i]), "r"((uint32_t)(w[i]>>32)));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(b[i]) : "r"(a[i]), "r"(m1));
        } else if (OP == ACC_WIDE) {
          // accumulate a varying product into a 64-bit accumulator: 1 IADD3 feeds 1 IMAD.WIDE
          asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(b[i]), "r"(m0));
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));
        } else if (OP == ACC_WIDE2) {
          // two accumulating wide MADs per IADD3
          asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(b[i]), "r"(m0));
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(cw[i]) : "r"(a[i]), "r"(m1));
        } else if (OP == DFMA) {
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(1.0000001), "d"(1e-12));
        } else if (OP == MIX_LO_HI) {
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m0), "r"(m1));
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m0), "r"(m1));
        } else if (OP == FFMA) {
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(1.0000001f), "f"(1e-7f));
        } else if (OP == MIX_WIDE_FFMA) {
          asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)), "l"(cw[i]));
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(1.0000001f), "f"(1e-7f));
        }
      }
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc ^= a[i] ^ b[i] ^ (uint32_t)cw[i] ^ (uint32_t)(cw[i] >> 32) ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)d[i] ^ (uint32_t)f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Res { double ms; double inst_per_clk_sm_cyc; double inst_per_s; };

template <int OP>
int run(const char* name, int ops_per_slot, int nsm, uint32_t* out, long long* cyc, bool last) {
  int blocks = nsm, threads = 1024;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) k<OP><<<blocks, threads>>>(out, 1234u + i, cyc);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    k<OP><<<blocks, threads>>>(out, 99u + r, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  long long h[1024];
  CK(cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += (double)h[i]; avg /= blocks;
  // thread-level instructions per block
  double inst_thread = (double)ITERS * 4 * CHAINS * ops_per_slot;
  double lanes_per_clk_sm = inst_thread * threads / avg;   // thread-ops per SM clock (in-kernel clock64)
  double total = inst_thread * threads * blocks;
  double per_s = total / (best * 1e-3);
  printf("  \"%s\": {\"ops_per_slot\": %d, \"ms\": %.4f, \"thread_ops_per_clk_per_sm\": %.2f, \"thread_ops_per_s\": %.4e, \"implied_mhz\": %.0f}%s\n",
         name, ops_per_slot, best, lanes_per_clk_sm, per_s, avg / (best * 1e-3) / 1e6, last ? "" : ",");
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(uint32_t) * nsm * 1024));
  CK(cudaMalloc(&cyc, sizeof(long long) * 1024));
  printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\",\n", p.name, nsm, p.major, p.minor);
  if (run<IMAD_LO>("imad_lo", 1, nsm, out, cyc, false)) return 1;
  if (run<IMAD_WIDE>("imad_wide", 1, nsm, out, cyc, false)) return 1;
  if (run<IMAD_HI>("imad_hi", 1, nsm, out, cyc, false)) return 1;
  if (run<MIX_LO_HI>("imad_lo+imad_hi", 2, nsm, out, cyc, false)) return 1;
  if (run<IADD3>("iadd3", 2, nsm, out, cyc, false)) return 1;
  if (run<LOP3>("lop3", 2, nsm, out, cyc, false)) return 1;
  if (run<SHF>("shf", 2, nsm, out, cyc, false)) return 1;
  if (run<MIX_WIDE_IADD>("imad_wide+iadd", 2, nsm, out, cyc, false)) return 1;
  if (run<MIX_WIDE_2IADD>("imad_wide+iadd+lop3", 3, nsm, out, cyc, false)) return 1;
  if (run<FFMA>("ffma", 1, nsm, out, cyc, false)) return 1;
  if (run<MIX_WIDE_FFMA>("imad_wide+ffma", 2, nsm, out, cyc, false)) return 1;
  if (run<ACC_WIDE>("iadd3+imad_wide_acc", 2, nsm, out, cyc, false)) return 1;
  if (run<ACC_WIDE2>("iadd3+2imad_wide_acc", 3, nsm, out, cyc, false)) return 1;
  if (run<DFMA>("dfma", 1, nsm, out, cyc, true)) return 1;
  printf("}\n");
  return 0;
}
