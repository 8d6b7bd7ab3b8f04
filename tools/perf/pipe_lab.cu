// Throughput of the integer multiply-add flavours on sm_100a (which ones share the "heavy" FMA pipe, how many passes).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
typedef uint64_t u64; typedef uint32_t u32;
template <int KIND>
__global__ void k(u32* out, u32 a, u32 b, int iters) {
  u32 x[8]; u64 y[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 3 + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (KIND == 0) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b));
        if (KIND == 1) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(x[i]), "+r"(*(u32*)&y[i]) : "r"(a), "r"(b));
        if (KIND == 2) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b));
        if (KIND == 3) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b));
        if (KIND == 4) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b));
        if (KIND == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
        if (KIND == 6) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
        if (KIND == 7) { asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(*(u32*)&y[i]) : "r"(a)); }
        if (KIND == 8) { asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(x[i]), "+r"(*((u32*)&y[i] + 1)) : "r"(a), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(*(u32*)&y[i]) : "r"(a)); }
      }
    }
  }
  u32 acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc ^= x[i] ^ (u32)y[i] ^ (u32)(y[i] >> 32);
  if (acc == 0xdeadbeef) out[0] = acc;
}
template <int KIND> double run(const char* name, int sms) {
  u32* d; cudaMalloc(&d, 64);
  const int iters = 2048, block = 256, grid = sms * 8;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(a); k<KIND><<<grid, block>>>(d, 3u, 0x01020305u, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms;
  }
  double ops = (double)grid * block * iters * 64.0;
  double g = ops / (best * 1e-3) / 1e9;
  printf("%-34s %9.1f Ginstr/s  = %.1f lanes/clk/SM\n", name, g, g * 1e9 / (sms * 1.965e9));
  cudaFree(d);
  return g;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  run<0>("mad.lo.u32 (IMAD)", sms);
  run<1>("mad.lo.cc+madc.hi (IMAD.WIDE)", sms);
  run<2>("dp2a.lo.u32.u32 (IDP.2A)", sms);
  run<3>("dp4a.u32.u32 (IDP.4A)", sms);
  run<4>("mad.hi.u32 (IMAD.HI)", sms);
  run<5>("add.u32 (IADD3)", sms);
  run<6>("prmt.b32 (PRMT)", sms);
  run<7>("dp2a + add pairs (each counted 1)", sms);
  run<8>("IMAD.WIDE + add pairs", sms);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
