// Developer tool: how many issue slots do the sm_100a packed-FP32 instructions take, and which pipes overlap?
// Each kernel runs 8 independent chains per thread of a fixed instruction mix (asm volatile keeps every
// instruction); 8 warps per SM sub-partition.  Output: cycles per warp-level "group" per sub-partition, where a
// group is the mix listed.  If the slots simply add, cycles/group = number of instructions in the group.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)

typedef unsigned long long u64;
#define FFMA2(v, a, b) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(a), "l"(b))
#define FFMA(x, a, b) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(a), "f"(b))
#define LOP(i, a) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(i) : "r"(a))
#define I2F(f, i) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(i))
#define RCP(f) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f))
#define SHFL(f) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+f"(f) : "r"(ln))

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
  u64 v[8];
  float x[8], y[8];
  int n[8];
  const int ln = (threadIdx.x + 1) & 31;
  u64 ca, cb;
  asm volatile("mov.b64 %0, {%1,%1};" : "=l"(ca) : "f"(seed * 1.0001f));
  asm volatile("mov.b64 %0, {%1,%1};" : "=l"(cb) : "f"(seed * 0.0001f));
  const float fa = seed * 1.0001f, fb = seed * 0.0001f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    x[c] = seed + c + threadIdx.x; y[c] = seed - c; n[c] = threadIdx.x * 9 + c;
    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v[c]) : "f"(x[c]), "f"(y[c]));
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (MODE == 0) { FFMA2(v[c], ca, cb); }                                   // 1 FFMA2
      if (MODE == 1) { FFMA(x[c], fa, fb); }                                    // 1 FFMA
      if (MODE == 2) { LOP(n[c], it); }                                         // 1 LOP3 (ALU)
      if (MODE == 3) { FFMA2(v[c], ca, cb); LOP(n[c], it); }                    // FFMA2 + LOP3
      if (MODE == 4) { FFMA(x[c], fa, fb); LOP(n[c], it); }                     // FFMA + LOP3
      if (MODE == 5) { FFMA2(v[c], ca, cb); FFMA(x[c], fa, fb); }               // FFMA2 + FFMA
      if (MODE == 6) { I2F(y[c], n[c]); LOP(n[c], it); }                        // I2FP + LOP3 (both ALU?)
      if (MODE == 7) { FFMA2(v[c], ca, cb); I2F(y[c], n[c]); n[c] += it; }      // FFMA2 + I2FP + IADD
      if (MODE == 8) { RCP(x[c]); }                                             // MUFU.RCP
      if (MODE == 9) { FFMA2(v[c], ca, cb); FFMA2(v[c], cb, ca); FFMA2(v[c], ca, ca); RCP(x[c]); }  // 3 FFMA2 + MUFU
      if (MODE == 10) { SHFL(x[c]); }                                           // SHFL
      if (MODE == 11) { FFMA2(v[c], ca, cb); SHFL(x[c]); }                      // FFMA2 + SHFL
      if (MODE == 12) { FFMA2(v[c], ca, cb); LOP(n[c], it); LOP(n[c], c); }     // FFMA2 + 2 LOP3
      if (MODE == 13) { FFMA(x[c], fa, fb); FFMA(y[c], fa, fb); LOP(n[c], it); LOP(n[c], c); }  // 2 FFMA + 2 LOP3
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float lo, hi;
    asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[c]));
    s += lo + hi + x[c] + y[c] + (float)n[c];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, int sms, float* d_out, double clock_ghz) {
  const int iters = 2048, blocks = sms * 4, threads = 256;   // 8 warps per sub-partition
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<MODE><<<blocks, threads>>>(d_out, 32, 1.0f);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k<MODE><<<blocks, threads>>>(d_out, iters, 1.0f);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  // groups per sub-partition = iters * 8 chains * 8 warps
  const double cyc = ms * 1e-3 * clock_ghz * 1e9 / ((double)iters * 8 * 8);
  printf("%-34s %8.3f ms  %6.2f cycles per group per sub-partition\n", name, ms, cyc);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  const double ghz = p.clockRate / 1e6;
  printf("%s, %d SMs, %.3f GHz (nominal; cycles assume this clock)\n", p.name, sms, ghz);
  float* d;
  CK(cudaMalloc(&d, (size_t)sms * 4 * 256 * 4));
  run<0>("FFMA2", sms, d, ghz);
  run<1>("FFMA", sms, d, ghz);
  run<2>("LOP3", sms, d, ghz);
  run<3>("FFMA2 + LOP3", sms, d, ghz);
  run<4>("FFMA + LOP3", sms, d, ghz);
  run<5>("FFMA2 + FFMA", sms, d, ghz);
  run<6>("I2FP + LOP3", sms, d, ghz);
  run<7>("FFMA2 + I2FP + IADD", sms, d, ghz);
  run<8>("MUFU.RCP", sms, d, ghz);
  run<9>("3 FFMA2 + MUFU.RCP", sms, d, ghz);
  run<10>("SHFL", sms, d, ghz);
  run<11>("FFMA2 + SHFL", sms, d, ghz);
  run<12>("FFMA2 + 2 LOP3", sms, d, ghz);
  run<13>("2 FFMA + 2 LOP3", sms, d, ghz);
  return 0;
}
