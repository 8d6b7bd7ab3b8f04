// Microbenchmarks of intra-CTA hand-off latency on sm_100a: named barriers vs shared-memory flags vs __syncthreads.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
typedef uint64_t u64;
__device__ __forceinline__ void nb_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// two warps ping-pong a value through shared memory with named barriers
__global__ void pingpong_bar(long long* out, int iters, int wb) {
  __shared__ volatile u64 slot[2][32];
  unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  u64 v = lane;
  long long t0 = clock64();
  if (wid == 0) {
    for (int i = 0; i < iters; i++) {
      slot[0][lane] = v; nb_arrive(1, 64); nb_sync(2, 64); v = slot[1][lane] + 1;
    }
  } else if (wid == (unsigned)wb) {
    for (int i = 0; i < iters; i++) {
      nb_sync(1, 64); u64 x = slot[0][lane]; slot[1][lane] = x + 1; nb_arrive(2, 64);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)v; }
}
// same with spin flags
__global__ void pingpong_flag(long long* out, int iters, int wb) {
  __shared__ volatile u64 slot[2][32];
  __shared__ volatile int flag[2];
  unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) { flag[0] = 0; flag[1] = 0; }
  __syncthreads();
  u64 v = lane;
  long long t0 = clock64();
  if (wid == 0) {
    for (int i = 1; i <= iters; i++) {
      slot[0][lane] = v; __syncwarp(); if (lane == 0) { __threadfence_block(); flag[0] = i; }
      while (flag[1] != i) {}
      v = slot[1][lane] + 1;
    }
  } else if (wid == (unsigned)wb) {
    for (int i = 1; i <= iters; i++) {
      while (flag[0] != i) {}
      u64 x = slot[0][lane]; slot[1][lane] = x + 1; __syncwarp(); if (lane == 0) { __threadfence_block(); flag[1] = i; }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)v; }
}
// flag-free: the payload itself carries a sequence tag (value never repeats), consumer spins on the payload
__global__ void pingpong_tag(long long* out, int iters, int wb) {
  __shared__ volatile u64 slot[2][32];
  unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  slot[0][lane] = 0; slot[1][lane] = 0;
  __syncthreads();
  u64 v = 1;
  long long t0 = clock64();
  if (wid == 0) {
    for (int i = 1; i <= iters; i++) {
      slot[0][lane] = v;
      u64 x; do { x = slot[1][lane]; } while (x != v + 1);
      v = x + 1;
    }
  } else if (wid == (unsigned)wb) {
    u64 expect = 1;
    for (int i = 1; i <= iters; i++) {
      u64 x; do { x = slot[0][lane]; } while (x != expect);
      slot[1][lane] = x + 1; expect += 2;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)v; }
}
// N warps, one __syncthreads per iteration with a shared write + read
__global__ void sync_all(long long* out, int iters) {
  __shared__ volatile u64 slot[2][16][32];
  unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  u64 v = lane;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    slot[i & 1][wid][lane] = v;
    __syncthreads();
    v = slot[i & 1][(wid + 1) % nw][lane] + 1;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)v; }
}
// dependent chain latencies of single instructions, one warp
__global__ void lat_lds(long long* out, int iters) {
  __shared__ volatile unsigned idx[32];
  idx[threadIdx.x] = threadIdx.x;
  __syncwarp();
  unsigned j = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) j = idx[j];
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = j; }
}
int main() {
  long long* d; cudaMalloc(&d, 64); long long h[2];
  const int iters = 10000;
  for (int wb : {1, 4, 3}) {
    pingpong_bar<<<1, 32 * (wb + 1)>>>(d, iters, wb); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("named-barrier ping-pong warp0<->warp%d: %.1f cycles per round trip (%.1f per hand-off)  err=%s\n", wb, (double)h[0] / iters, (double)h[0] / iters / 2, cudaGetErrorString(cudaGetLastError()));
    pingpong_flag<<<1, 32 * (wb + 1)>>>(d, iters, wb); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("flag-spin     ping-pong warp0<->warp%d: %.1f cycles per round trip (%.1f per hand-off)\n", wb, (double)h[0] / iters, (double)h[0] / iters / 2);
    pingpong_tag<<<1, 32 * (wb + 1)>>>(d, iters, wb); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("payload-spin  ping-pong warp0<->warp%d: %.1f cycles per round trip (%.1f per hand-off) v=%lld\n", wb, (double)h[0] / iters, (double)h[0] / iters / 2, h[1]);
  }
  for (int nw : {2, 4, 12, 16}) {
    sync_all<<<1, 32 * nw>>>(d, iters); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("__syncthreads exchange, %2d warps: %.1f cycles per iteration\n", nw, (double)h[0] / iters);
  }
  lat_lds<<<1, 32>>>(d, iters); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("LDS dependent latency: %.1f cycles\n", (double)h[0] / iters);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
