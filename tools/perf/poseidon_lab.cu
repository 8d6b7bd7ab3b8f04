// Standalone lab for the Poseidon leaf sponge (K2): times kernel variants on synthetic column-major data and checks
// every variant against a one-thread-per-leaf kernel and a host permutation.  Not part of the product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I starky_bls12_381_b200/csrc \
//        tools/perf/poseidon_lab.cu -o gpurun_out/poseidon_lab
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#ifdef LAB_TRACE
#define SP_TRACE 1
#endif
#include "leafhash.cuh"
#include "leafhash_mm.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__host__ __device__ inline u64 splitmix(u64 x) {
  x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
__host__ __device__ inline u64 cell(u64 idx) { u64 v = splitmix(idx); return v >= GL_P ? v - GL_P : v; }
__global__ void fill_kernel(u64* d, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = cell(i);
}

// one thread per leaf, whole state in registers
__global__ void __launch_bounds__(128) ref_leaf_kernel(const u64* __restrict__ cols, uint32_t leaf_len, uint32_t n_leaves, u64* digests) {
  uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n_leaves) return;
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = 0;
  uint32_t chunks = (leaf_len + 7) / 8;
  for (uint32_t c = 0; c < chunks; c++) {
#pragma unroll
    for (int i = 0; i < 8; i++) if (c * 8 + i < leaf_len) s[i] = cols[(size_t)(c * 8 + i) * n_leaves + pos];
    poseidon_permute(s);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) digests[4ull * pos + i] = s[i];
}

__global__ void mul_check_kernel(const u64* a, const u64* b, u64* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = gl_mul(a[i], b[i]);
}
static int mul_check() {
  std::vector<u64> edge = {0, 1, 2, 7, 0xFFFFFFFFULL, 0x100000000ULL, 0x100000001ULL, 0xFFFFFFFEULL, GL_P - 1, GL_P, GL_P + 1, GL_P - 2,
                           0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFF00000000ULL, 0xFFFFFFFEFFFFFFFFULL, 0x8000000000000000ULL,
                           0x7FFFFFFFFFFFFFFFULL, 0x00000000FFFFFFFEULL, 0xFFFFFFFE00000001ULL, 0xFFFFFFFE00000002ULL, 0x0000000100000000ULL - 2,
                           0xFFFF0000FFFF0000ULL, 0x0000FFFF0000FFFFULL, 0x00000001FFFFFFFFULL, 0xFFFFFFFF00000002ULL};
  std::vector<u64> a, b;
  for (u64 x : edge) for (u64 y : edge) { a.push_back(x); b.push_back(y); }
  for (int i = 0; i < 2000000; i++) {
    u64 x = splitmix(2 * i + 11), y = splitmix(2 * i + 12);
    if (i % 7 == 0) x &= 0xFFFFFFFFULL; if (i % 11 == 0) y |= 0xFFFFFFFF00000000ULL; if (i % 13 == 0) x |= 0xFFFFFFFF00000000ULL;
    if (i % 17 == 0) y &= 0xFFFFFFFF00000000ULL; if (i % 19 == 0) x = edge[i % edge.size()];
    a.push_back(x); b.push_back(y);
  }
  int n = (int)a.size();
  u64 *da, *db, *dc;
  CK(cudaMalloc(&da, 8ull * n)); CK(cudaMalloc(&db, 8ull * n)); CK(cudaMalloc(&dc, 8ull * n));
  CK(cudaMemcpy(da, a.data(), 8ull * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b.data(), 8ull * n, cudaMemcpyHostToDevice));
  mul_check_kernel<<<(n + 255) / 256, 256>>>(da, db, dc, n);
  std::vector<u64> c(n);
  CK(cudaMemcpy(c.data(), dc, 8ull * n, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < n; i++) {
    unsigned __int128 m = (unsigned __int128)a[i] * b[i];
    u64 want = (u64)(m % GL_P);
    if (c[i] != want) { if (bad < 5) printf("  mul mismatch: %016llx * %016llx = %016llx want %016llx\n", (unsigned long long)a[i], (unsigned long long)b[i], (unsigned long long)c[i], (unsigned long long)want); bad++; }
  }
  printf("gl_mul check: %d pairs, %d mismatches\n", n, bad);
  cudaFree(da); cudaFree(db); cudaFree(dc);
  return bad;
}

// latency probes: one warp, dependent chains
__global__ void mul_latency_kernel(u64* out, u64 a, int iters, long long* cycles) {
  u64 x = a + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) x = gl_mul_lazy(x, x);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}
template <int ILP>
__global__ void mul_tput_kernel(u64* out, u64 a, int iters) {
  u64 x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; j++) x[j] = a + threadIdx.x * 7 + j + blockIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
      for (int j = 0; j < ILP; j++) x[j] = gl_mul_lazy(x[j], x[j]);
  }
  u64 acc = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) acc ^= x[j];
  if (acc == 0x1234567) out[0] = acc;
}
// ---- alternative Goldilocks multiply formulations (lab only) ----
// v2: the fold's x2 * eps + (x1:x0) as ONE multiply-add with carry out (mad.lo.cc / madc.hi.cc fuse into IMAD.WIDE),
//     then - x3 and the two eps repairs.
__device__ __forceinline__ u64 gl_mul_lazy_v2(u64 a, u64 b) {
  const u64 lo = a * b, hi = __umul64hi(a, b);
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 c1, b1;\n\t"
      "mad.lo.cc.u32 %0, %4, 0xFFFFFFFF, %2;\n\t"     // (x1:x0) + x2 * eps, carry c1
      "madc.hi.cc.u32 %1, %4, 0xFFFFFFFF, %3;\n\t"
      "addc.u32 c1, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, %5;\n\t"                     // - x3, borrow b1
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.u32 b1, 0, 0;\n\t"                         // 0 or 0xFFFFFFFF
      "neg.s32 c1, c1;\n\t"                            // 0 or 0xFFFFFFFF = c1 * eps
      "add.cc.u32 %0, %0, c1;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "sub.cc.u32 %0, %0, b1;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"((u32)lo), "r"((u32)(lo >> 32)), "r"((u32)hi), "r"((u32)(hi >> 32)));
  return ((u64)r1 << 32) | r0;
}
template <int V> __device__ __forceinline__ u64 mulv(u64 a, u64 b) { return V == 2 ? gl_mul_lazy_v2(a, b) : gl_mul_lazy(a, b); }
template <int V> __device__ __forceinline__ u64 sboxv(u64 x) {
  const u64 x2 = mulv<V>(x, x), x4 = mulv<V>(x2, x2), x3 = mulv<V>(x2, x);
  return mulv<V>(x3, x4);
}
template <int V>
__global__ void sbox_variant_kernel(u64* out, u64 a, int iters, int busy, long long* cycles) {
  const unsigned warp = threadIdx.x >> 5;
  u64 x = a + threadIdx.x;
  if (warp == 0 || busy) {
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int k = 0; k < 4; k++) x = gl_add_lazy_canon(sboxv<V>(x), 0x123456789ULL);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// S-box latency in the shape of the sp kernel: `warps` warps per block, warp 0 runs a dependent chain of S-boxes
// (x <- x^7 + c), the other warps either wait at a barrier (busy == 0) or run their own chains (busy == 1)
__global__ void sbox_latency_kernel(u64* out, u64 a, int iters, int busy, long long* cycles) {
  const unsigned warp = threadIdx.x >> 5;
  u64 x = a + threadIdx.x;
  if (warp == 0 || busy) {
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int k = 0; k < 4; k++) x = gl_add_lazy_canon(poseidon_sbox(x), 0x123456789ULL);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void perm_latency_kernel(u64* out, int iters, long long* cycles) {
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) poseidon_permute(s);
  long long t1 = clock64();
#pragma unroll
  for (int i = 0; i < 12; i++) out[threadIdx.x * 12 + i] = s[i];
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <class F>
static float time_ms(F&& launch, int reps = 3) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(a); launch(); cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  return best;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s, %d SMs, %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);

#ifdef LAB_TRACE
  if (argc > 1 && !strcmp(argv[1], "trace")) {   // clock64 timeline of the word-0 warp of the sp kernel (one permutation)
    uint32_t nl = 4096, ll = 256;
    size_t cells = (size_t)nl * ll;
    u64 *d_cols, *d_dig; long long* d_tr;
    CK(cudaMalloc(&d_cols, 8 * cells)); CK(cudaMalloc(&d_dig, 32ull * nl)); CK(cudaMalloc(&d_tr, 64 * 8));
    CK(cudaMemset(d_tr, 0, 64 * 8));
    CK(cudaMemcpyToSymbol(g_sp_trace, &d_tr, sizeof(d_tr)));
    fill_kernel<<<(unsigned)((cells + 255) / 256), 256>>>(d_cols, cells);
    for (int rep = 0; rep < 2; rep++) leaf_sponge_sp_kernel<0><<<(nl + 31) / 32, 416>>>(d_cols, ll, nl, 0, d_dig);
    CK(cudaDeviceSynchronize());
    long long h[64]; CK(cudaMemcpy(h, d_tr, sizeof(h), cudaMemcpyDeviceToHost));
    printf("perm start %lld\n", 0ll);
    printf("last full round of the first four (stamps of the crit warp): sbox+store issued at +%lld, barrier passed +%lld (wait %lld), mds+recombine issued +%lld; previous round ended +%lld\n",
           h[50] - h[0], h[51] - h[0], h[51] - h[50], h[52] - h[0], h[3] - h[0]);
    for (int i = 1; i <= 4; i++) printf("full round %d done  +%lld (d %lld)\n", i - 1, h[i] - h[0], h[i] - h[i - 1]);
    for (int r = 0; r < 8; r++)
      printf("partial %d: sbox issued +%lld | barrier passed +%lld (wait %lld) | x0 issued +%lld   round total %lld\n", r, h[8 + 4 * r] - h[0],
             h[9 + 4 * r] - h[0], h[9 + 4 * r] - h[8 + 4 * r], h[10 + 4 * r] - h[0], r ? h[10 + 4 * r] - h[10 + 4 * (r - 1)] : h[10] - h[4]);
    printf("partial rounds done +%lld  (22 rounds: %lld, avg %lld)\n", h[5] - h[0], h[5] - h[4], (h[5] - h[4]) / 22);
    for (int i = 0; i < 4; i++) printf("full round %d done +%lld (d %lld)\n", 26 + i, h[40 + i] - h[0], h[40 + i] - (i ? h[40 + i - 1] : h[5]));
    return 0;
  }
#endif
  if (argc > 4 && !strcmp(argv[1], "one")) {   // one variant, one shape (for ncu): one <wps> <n_leaves> <leaf_len>
    int wps = atoi(argv[2]); uint32_t nl = atoi(argv[3]), ll = atoi(argv[4]);
    size_t cells = (size_t)nl * ll;
    u64 *d_cols, *d_dig;
    CK(cudaMalloc(&d_cols, 8 * cells)); CK(cudaMalloc(&d_dig, 32ull * nl));
    fill_kernel<<<(unsigned)((cells + 255) / 256), 256>>>(d_cols, cells);
    CK(cudaDeviceSynchronize());
    unsigned g32 = (nl + 31) / 32;
    for (int rep = 0; rep < 2; rep++) {
      float ms = time_ms([&] {
        if (wps == 201) leaf_sponge_mm_kernel<1><<<(nl + 7) / 8, 32>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 202) leaf_sponge_mm_kernel<2><<<g32, 64>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 204) leaf_sponge_mm_kernel<4><<<(nl + 63) / 64, 64>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 43) leaf_sponge_dp_kernel<0><<<g32, 128>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 44) leaf_sponge_dp_kernel<2><<<g32, 128>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 45) leaf_sponge_ds_kernel<<<g32, 128>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 101) leaf_sponge_st_kernel<<<g32, 32>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 130) leaf_sponge_sp_kernel<0><<<g32, 416>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 131) leaf_sponge_sp_kernel<1><<<g32, 512>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 137) leaf_sponge_sp_kernel<7><<<g32, 512>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 122) leaf_sponge_w12_kernel<2><<<g32, 512>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 120) leaf_sponge_w12_kernel<0><<<g32, 384>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 12) leaf_sponge_ws_kernel<12><<<g32, 384>>>(d_cols, ll, nl, 0, d_dig);
        else if (wps == 4) leaf_sponge_ws_kernel<4><<<g32, 128>>>(d_cols, ll, nl, 0, d_dig);
        else leaf_sponge_ws_kernel<1><<<g32, 32>>>(d_cols, ll, nl, 0, d_dig);
      }, 1);
      printf("ws<%d> N=%u C=%u: %.3f ms\n", wps, nl, ll, ms);
    }
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "mm")) {   // dense MDS on IMMA (leafhash_mm.cuh) against the dp / sp kernels
    struct Shape { const char* name; uint32_t n_leaves, leaf_len; };
    for (const Shape& sh : {Shape{"ragged", 1000, 77}, Shape{"ragged2", 1013, 131}, Shape{"ML-like", 2048, 8003}, Shape{"PP-like", 4096, 8003},
                            Shape{"ECC-like", 32768, 3339}, Shape{"FE-like", 32768, 8003}, Shape{"FE 1/2 box", 16384, 8003}, Shape{"FE 1/4 box", 8192, 8003}}) {
      const size_t cells = (size_t)sh.n_leaves * sh.leaf_len;
      u64 *d_cols, *d_dig;
      CK(cudaMalloc(&d_cols, 8 * cells)); CK(cudaMalloc(&d_dig, 32ull * sh.n_leaves));
      fill_kernel<<<(unsigned)((cells + 255) / 256), 256>>>(d_cols, cells);
      CK(cudaDeviceSynchronize());
      const double perms = (double)((sh.leaf_len + 7) / 8) * sh.n_leaves;
      const unsigned g32 = (sh.n_leaves + 31) / 32;
      std::vector<u64> ref(4ull * sh.n_leaves), got(4ull * sh.n_leaves);
      printf("%-10s N=%6u C=%6u\n", sh.name, sh.n_leaves, sh.leaf_len);
      auto run = [&](const char* name, auto launch, bool is_ref) {
        CK(cudaMemset(d_dig, 0, 32ull * sh.n_leaves));
        float t = time_ms(launch);
        CK(cudaMemcpy(is_ref ? ref.data() : got.data(), d_dig, 32ull * sh.n_leaves, cudaMemcpyDeviceToHost));
        bool ok = is_ref || !memcmp(got.data(), ref.data(), 32ull * sh.n_leaves);
        printf("    %-46s %9.3f ms  %7.1f Mperm/s  %s\n", name, t, perms / t / 1e3, ok ? "ok" : "MISMATCH");
      };
      run("ref (1 thread per leaf)", [&] { ref_leaf_kernel<<<g32, 32>>>(d_cols, sh.leaf_len, sh.n_leaves, d_dig); }, true);
      run("dp<0>", [&] { leaf_sponge_dp_kernel<0><<<g32, 128>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); }, false);
      run("sp<0>", [&] { leaf_sponge_sp_kernel<0><<<g32, 416>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); }, false);
      run("mm<1>  8 leaves per warp, block 32", [&] { leaf_sponge_mm_kernel<1><<<(sh.n_leaves + 7) / 8, 32>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); }, false);
      run("mm<1> 4 leaves per warp + helper lanes", [&] { leaf_sponge_mm_kernel<1, 256><<<(sh.n_leaves + 3) / 4, 32>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); }, false);
      run("mm<2> 16 leaves per warp, block 64", [&] { leaf_sponge_mm_kernel<2><<<(sh.n_leaves + 31) / 32, 64>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); }, false);
      run("mm<4> 32 leaves per warp, block 64", [&] { leaf_sponge_mm_kernel<4><<<(sh.n_leaves + 63) / 64, 64>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); }, false);
#define HET(N4, N2, N1, label) do { using Hh = MmHet<N4, N2, N1>; \
      CK(cudaFuncSetAttribute(leaf_sponge_mm_het_kernel<N4, N2, N1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 << 10)); \
      run(label, [&] { leaf_sponge_mm_het_kernel<N4, N2, N1, 0><<<(sh.n_leaves + Hh::LEAVES - 1) / Hh::LEAVES, Hh::THREADS, 120 << 10>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); }, false); } while (0)
      HET(0, 3, 1, "mm het: 3 x 16 + 8 leaves per scheduler");
      HET(1, 1, 1, "mm het: 32 + 16 + 8 leaves per scheduler");
      HET(1, 0, 3, "mm het: 32 + 3 x 8 leaves per scheduler");
      HET(0, 2, 3, "mm het: 2 x 16 + 3 x 8 leaves per scheduler");
      HET(0, 0, 7, "mm het: 7 x 8 leaves per scheduler");
      cudaFree(d_cols); cudaFree(d_dig);
    }
    printf("lab done\n");
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "mmx")) {   // where does the time of the matrix-instruction sponge go: timing by elimination (wrong digests)
    struct Shape { const char* name; uint32_t n_leaves, leaf_len; };
    for (const Shape& sh : {Shape{"ML-like", 2048, 8003}, Shape{"FE-like", 32768, 8003}}) {
      const size_t cells = (size_t)sh.n_leaves * sh.leaf_len;
      u64 *d_cols, *d_dig;
      CK(cudaMalloc(&d_cols, 8 * cells)); CK(cudaMalloc(&d_dig, 32ull * sh.n_leaves));
      fill_kernel<<<(unsigned)((cells + 255) / 256), 256>>>(d_cols, cells);
      CK(cudaDeviceSynchronize());
      const unsigned g8 = (sh.n_leaves + 7) / 8, g64 = (sh.n_leaves + 63) / 64;
      const double chain = (sh.leaf_len + 7) / 8;
      printf("%-10s N=%6u C=%6u\n", sh.name, sh.n_leaves, sh.leaf_len);
      auto run = [&](const char* name, auto launch) {
        float t = time_ms(launch);
        printf("    %-58s %9.3f ms  %8.0f cycles per permutation\n", name, t, t * 1e-3 / chain * 1.965e9);
      };
#define MMX(NLV, DBGV, G, B, label) run(label, [&] { leaf_sponge_mm_kernel<NLV, DBGV><<<G, B>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); })
      MMX(1, 0, g8, 32, "mm<1> complete");
      MMX(1, 1, g8, 32, "mm<1> no partial-round S-box");
      MMX(1, 2, g8, 32, "mm<1> no matrix instruction");
      MMX(1, 4, g8, 32, "mm<1> no recombination");
      MMX(1, 8, g8, 32, "mm<1> no full-round S-boxes");
      MMX(1, 9, g8, 32, "mm<1> no S-box at all");
      MMX(1, 15, g8, 32, "mm<1> pack + constants + loop only");
      MMX(4, 0, g64, 64, "mm<4> complete");
      MMX(4, 1, g64, 64, "mm<4> no partial-round S-box");
      MMX(4, 2, g64, 64, "mm<4> no matrix instruction");
      MMX(4, 4, g64, 64, "mm<4> no recombination");
      MMX(4, 8, g64, 64, "mm<4> no full-round S-boxes");
      MMX(4, 9, g64, 64, "mm<4> no S-box at all");
      MMX(4, 64, g64, 64, "mm<4> (correct) full-round S-boxes in lock step, groups of 3");
      MMX(4, 128, g64, 64, "mm<4> (correct) full-round S-boxes in lock step, groups of 6");
      MMX(4, 192, g64, 64, "mm<4> (correct) full-round S-boxes in lock step, groups of 12");
      MMX(2, 0, (sh.n_leaves + 31) / 32, 64, "mm<2> complete");
      MMX(2, 64, (sh.n_leaves + 31) / 32, 64, "mm<2> (correct) full-round S-boxes in lock step, groups of 3");
      MMX(2, 128, (sh.n_leaves + 31) / 32, 64, "mm<2> (correct) full-round S-boxes in lock step, groups of 6");
      MMX(1, 64, g8, 32, "mm<1> (correct) full-round S-boxes in lock step, groups of 3");
      MMX(4, 16, g64, 64, "mm<4> (correct) fold on the FMA pipe");
      MMX(4, 32, g64, 64, "mm<4> (correct) limb pairs on the ALU pipe");
      MMX(4, 48, g64, 64, "mm<4> (correct) both");
      MMX(1, 16, g8, 32, "mm<1> (correct) fold on the FMA pipe");
      MMX(1, 32, g8, 32, "mm<1> (correct) limb pairs on the ALU pipe");
      cudaFree(d_cols); cudaFree(d_dig);
    }
    printf("lab done\n");
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "dp")) {   // throughput-bound shapes: which warp owns the partial-round S-box
    struct Shape { const char* name; uint32_t n_leaves, leaf_len; };
    for (const Shape& sh : {Shape{"ECC-like", 32768, 3339}, Shape{"FE-like", 32768, 8003}, Shape{"FE/4", 32768, 18382}, Shape{"FE 1/2 box", 16384, 18382}}) {
      const size_t cells = (size_t)sh.n_leaves * sh.leaf_len;
      u64 *d_cols, *d_ref, *d_dig;
      CK(cudaMalloc(&d_cols, 8 * cells)); CK(cudaMalloc(&d_ref, 32ull * sh.n_leaves)); CK(cudaMalloc(&d_dig, 32ull * sh.n_leaves));
      fill_kernel<<<(unsigned)((cells + 255) / 256), 256>>>(d_cols, cells);
      CK(cudaDeviceSynchronize());
      const double perms = (double)((sh.leaf_len + 7) / 8) * sh.n_leaves;
      const unsigned g32 = (sh.n_leaves + 31) / 32;
      std::vector<u64> ref(4ull * sh.n_leaves), got(4ull * sh.n_leaves);
      printf("%-10s N=%6u C=%6u\n", sh.name, sh.n_leaves, sh.leaf_len);
      auto run = [&](const char* name, auto launch, bool is_ref) {
        CK(cudaMemset(d_dig, 0, 32ull * sh.n_leaves));
        float t = time_ms(launch);
        CK(cudaMemcpy(is_ref ? ref.data() : got.data(), d_dig, 32ull * sh.n_leaves, cudaMemcpyDeviceToHost));
        bool ok = is_ref || !memcmp(got.data(), ref.data(), 32ull * sh.n_leaves);
        printf("    %-46s %9.3f ms  %7.1f Mperm/s  %s\n", name, t, perms / t / 1e3, ok ? "ok" : "MISMATCH");
      };
      for (int rep = 0; rep < 2; rep++) {
        run("dp<0> words 0..2 always on warp 0", [&] { leaf_sponge_dp_kernel<0><<<g32, 128>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); }, true);
        run("dp<1> S-box warp rotates by block index", [&] { leaf_sponge_dp_kernel<1><<<g32, 128>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); }, false);
        run("dp<2> S-box warp rotates by arrival on the SM", [&] { leaf_sponge_dp_kernel<2><<<g32, 128>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); }, false);
      }
      cudaFree(d_cols); cudaFree(d_ref); cudaFree(d_dig);
    }
    printf("lab done\n");
    return 0;
  }
  const bool only_sp = argc > 1 && !strcmp(argv[1], "sp");     // the latency-bound shapes and the sp variants only
  // ---- latency / throughput probes ----
  if (!only_sp) {
  if (mul_check()) return 1;
  u64* d_out; long long* d_cyc; CK(cudaMalloc(&d_out, 1 << 20)); CK(cudaMalloc(&d_cyc, 64));
  long long cyc;
  mul_latency_kernel<<<1, 32>>>(d_out, 12345, 64, d_cyc); mul_latency_kernel<<<1, 32>>>(d_out, 12345, 64, d_cyc);
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  printf("gl_mul_lazy dependent latency: %.1f cycles\n", (double)cyc / (64 * 16));
  perm_latency_kernel<<<1, 32>>>(d_out, 16, d_cyc); perm_latency_kernel<<<1, 32>>>(d_out, 16, d_cyc);
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  printf("poseidon_permute one-warp latency: %.0f cycles\n", (double)cyc / 16);
  {
    const int iters = 2048, grid = prop.multiProcessorCount * 8, block = 256;
    float ms1 = time_ms([&] { mul_tput_kernel<1><<<grid, block>>>(d_out, 3, iters); });
    float ms4 = time_ms([&] { mul_tput_kernel<4><<<grid, block>>>(d_out, 3, iters / 4); });
    float ms8 = time_ms([&] { mul_tput_kernel<8><<<grid, block>>>(d_out, 3, iters / 8); });
    double muls = (double)grid * block * iters * 4;
    printf("gl_mul_lazy throughput: ILP1 %.1f  ILP4 %.1f  ILP8 %.1f Gmul/s\n", muls / ms1 / 1e6, muls / ms4 / 1e6, muls / ms8 / 1e6);
  }

  {  // S-box latency in situ
    u64* d_out; long long* d_cyc; CK(cudaMalloc(&d_out, 8 * 416 * 148)); CK(cudaMalloc(&d_cyc, 8 * 148));
    const int iters = 256;
    for (int warps : {1, 4, 13}) for (int busy : {0, 1}) for (int blocks : {1, 128}) {
      sbox_latency_kernel<<<blocks, 32 * warps>>>(d_out, 3, iters, busy, d_cyc);
      CK(cudaDeviceSynchronize());
      long long c; CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
      printf("S-box chain latency: %2d warps/block, others %s, %3d blocks: %.1f cycles per S-box\n", warps, busy ? "busy" : "idle", blocks,
             (double)c / (iters * 4));
    }
    u64 h1[32], h2[32];
    for (int busy : {0, 1}) {
      sbox_variant_kernel<1><<<1, 416>>>(d_out, 3, iters, busy, d_cyc); CK(cudaDeviceSynchronize());
      long long c1; CK(cudaMemcpy(&c1, d_cyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h1, d_out, sizeof(h1), cudaMemcpyDeviceToHost));
      sbox_variant_kernel<2><<<1, 416>>>(d_out, 3, iters, busy, d_cyc); CK(cudaDeviceSynchronize());
      long long c2; CK(cudaMemcpy(&c2, d_cyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2, d_out, sizeof(h2), cudaMemcpyDeviceToHost));
      bool same = true;
      for (int i = 0; i < 32; i++) same = same && ((h1[i] % GL_P) == (h2[i] % GL_P));
      printf("S-box variants, 13 warps, others %s: current fold %.1f, multiply-add fold %.1f cycles per S-box  (%s)\n", busy ? "busy" : "idle",
             (double)c1 / (iters * 4), (double)c2 / (iters * 4), same ? "same values" : "VALUES DIFFER");
    }
    cudaFree(d_out); cudaFree(d_cyc);
  }
  }

  // ---- leaf sponge variants ----
  struct Shape { const char* name; uint32_t n_leaves, leaf_len; };
  std::vector<Shape> shapes = {{"ML-like", 2048, 8003}, {"PP-like", 4096, 8003}, {"ECC-like", 32768, 3339}, {"FE-like", 32768, 8003},
                               {"ragged", 1000, 77}};
  if (only_sp) shapes = {{"ML-like", 2048, 8003}, {"PP-like", 4096, 8003}, {"ragged", 1000, 77}};
  if (argc > 1 && !strcmp(argv[1], "full")) {
    shapes.push_back({"PP", 4096, 29376}); shapes.push_back({"ML", 2048, 97330}); shapes.push_back({"FE", 32768, 73527});
  }
  for (const Shape& sh : shapes) {
    const size_t cells = (size_t)sh.n_leaves * sh.leaf_len;
    u64 *d_cols, *d_ref, *d_dig;
    CK(cudaMalloc(&d_cols, 8 * cells)); CK(cudaMalloc(&d_ref, 32ull * sh.n_leaves)); CK(cudaMalloc(&d_dig, 32ull * sh.n_leaves));
    fill_kernel<<<(unsigned)((cells + 255) / 256), 256>>>(d_cols, cells);
    CK(cudaDeviceSynchronize());
    const double perms = (double)((sh.leaf_len + 7) / 8) * sh.n_leaves;
    float ms = time_ms([&] { ref_leaf_kernel<<<(sh.n_leaves + 31) / 32, 32>>>(d_cols, sh.leaf_len, sh.n_leaves, d_ref); }, 2);
    printf("%-8s N=%6u C=%6u  ref(1 thread/leaf, block 32): %9.3f ms  %7.1f Mperm/s\n", sh.name, sh.n_leaves, sh.leaf_len, ms, perms / ms / 1e3);
    std::vector<u64> ref(4ull * sh.n_leaves), got(4ull * sh.n_leaves);
    CK(cudaMemcpy(ref.data(), d_ref, 32ull * sh.n_leaves, cudaMemcpyDeviceToHost));
    // host check of 3 leaves
    for (uint32_t leaf : {0u, sh.n_leaves / 2 + 1, sh.n_leaves - 1}) {
      u64 s[12] = {0};
      for (uint32_t c = 0; c < (sh.leaf_len + 7) / 8; c++) {
        for (int i = 0; i < 8; i++) if (c * 8 + i < sh.leaf_len) s[i] = cell((size_t)(c * 8 + i) * sh.n_leaves + leaf);
        poseidon_permute(s);
      }
      if (memcmp(s, &ref[4ull * leaf], 32)) { printf("  HOST MISMATCH at leaf %u\n", leaf); return 1; }
    }
    auto run = [&](const char* name, auto launch) {
      CK(cudaMemset(d_dig, 0, 32ull * sh.n_leaves));
      float t = time_ms(launch);
      CK(cudaMemcpy(got.data(), d_dig, 32ull * sh.n_leaves, cudaMemcpyDeviceToHost));
      bool ok = !memcmp(got.data(), ref.data(), 32ull * sh.n_leaves);
      printf("    %-28s %9.3f ms  %7.1f Mperm/s  %s\n", name, t, perms / t / 1e3, ok ? "ok" : "MISMATCH");
    };
    const unsigned g32 = (sh.n_leaves + 31) / 32;
    if (!only_sp) {
    run("ws<1>  (12 words/thread)", [&] { leaf_sponge_ws_kernel<1><<<g32, 32>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("ws<2>  (6 words/thread)", [&] { leaf_sponge_ws_kernel<2><<<g32, 64>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("ws<3>  (4 words/thread)", [&] { leaf_sponge_ws_kernel<3><<<g32, 96>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("ws<4>  (3 words/thread)", [&] { leaf_sponge_ws_kernel<4><<<g32, 128>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("ws<6>  (2 words/thread)", [&] { leaf_sponge_ws_kernel<6><<<g32, 192>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("ws<12> (1 word/thread)", [&] { leaf_sponge_ws_kernel<12><<<g32, 384>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("w12<0> (two-barrier partial)", [&] { leaf_sponge_w12_kernel<0><<<g32, 384>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("dp (3 words/thread, dp2a MDS)", [&] { leaf_sponge_dp_kernel<0><<<g32, 128>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("dp<1> S-box warp rotates by block index", [&] { leaf_sponge_dp_kernel<1><<<g32, 128>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("dp<2> S-box warp rotates by arrival on the SM", [&] { leaf_sponge_dp_kernel<2><<<g32, 128>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("ds (dp2a full + sparse partial)", [&] { leaf_sponge_ds_kernel<<<g32, 128>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("st (sparse, one thread per leaf, block 32)", [&] { leaf_sponge_st_kernel<<<g32, 32>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("st (sparse, one thread per leaf, block 64)", [&] { leaf_sponge_st_kernel<<<(sh.n_leaves + 63) / 64, 64>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    }
    run("sp<0> (sparse partial rounds, 13 warps)", [&] { leaf_sponge_sp_kernel<0><<<g32, 416>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("sp<1> word-0 warp alone on its scheduler", [&] { leaf_sponge_sp_kernel<1><<<g32, 512>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("sp<2> round-3 row in registers", [&] { leaf_sponge_sp_kernel<2><<<g32, 416>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("sp<4> full-round rows on dp2a", [&] { leaf_sponge_sp_kernel<4><<<g32, 416>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("sp<3> = 1 + 2", [&] { leaf_sponge_sp_kernel<3><<<g32, 512>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("sp<6> = 2 + 4", [&] { leaf_sponge_sp_kernel<6><<<g32, 416>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    run("sp<7> = 1 + 2 + 4", [&] { leaf_sponge_sp_kernel<7><<<g32, 512>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    if (!only_sp)
    run("w12<0,2> (2 chains/half)", [&] { leaf_sponge_w12_kernel<0, 2><<<g32, 384>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    if (!only_sp)
    run("w12<0,3> (3 chains/half)", [&] { leaf_sponge_w12_kernel<0, 3><<<g32, 384>>>(d_cols, sh.leaf_len, sh.n_leaves, 0, d_dig); });
    cudaFree(d_cols); cudaFree(d_ref); cudaFree(d_dig);
  }
  printf("lab done\n");
  return 0;
}
