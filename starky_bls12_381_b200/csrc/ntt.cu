// K1 / K5: column-batched Goldilocks NTTs for sm_100a.
// Replaces PolynomialValues::ifft, PolynomialCoeffs::lde + coset_fft(7) and the transpose +
// reverse_index_bits of PolynomialBatch::from_values / from_coeffs (plonky2; reached from
// starky::prover::prove, reference call sites /root/reference/src/aggregate_proof.rs:59,105,138,169,212).
//
// Data layout (DESIGN.md "HBM layout"):
//   values [C][n]  column-major, natural order           (what Vec<PolynomialValues<F>> holds)
//   coeffs [C][n]  position p holds coefficient c_{bitrev(p)}  ("bit-reversed coefficient order")
//   lde    [C][N]  position J*n + k holds P(7 * w_N^(j + 2^r k)), j = bitrev_r(J)   ("coset-major order")
// With that choice a column needs NO permutation pass: a decimation-in-frequency inverse transform takes natural
// values to bit-reversed coefficients, and per coset a decimation-in-time forward transform takes bit-reversed
// (scaled) coefficients to natural-order coset values.  plonky2's leaf index of position (J,k) is J*n + bitrev_n(k);
// only the 32-byte digests are scattered to it (merkle.cu), the 8*C*N bytes of LDE never are.
#include <stdlib.h>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// twiddles and coset scale tables (computed once per size, cached in the ctx)
// ---------------------------------------------------------------------------------------------------------
__global__ void twiddle_kernel(u64* fwd, u64* inv, u64 w, u64 w_inv, uint32_t half) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= half) return;
  fwd[k] = gl_pow(w, k);
  inv[k] = gl_pow(w_inv, k);
}

const Twiddles& sb_twiddles(sb_ctx* ctx, unsigned log_size) {
  auto it = ctx->tw.find(log_size);
  if (it != ctx->tw.end()) return it->second;
  Twiddles& t = ctx->tw[log_size];
  t.log_size = log_size;
  uint32_t half = log_size ? (1u << (log_size - 1)) : 1;
  t.fwd.ensure(8ull * half);
  t.inv.ensure(8ull * half);
  u64 w = gl_root(log_size), wi = gl_inv(w);
  LAUNCH(ctx, twiddle_kernel, (half + 255) / 256, 256, 0, t.fwd.as<u64>(), t.inv.as<u64>(), w, wi, half);
  return t;
}

// scale[J][p] = n^-1 * (7 * w_N^j)^{bitrev_n(p)},  j = bitrev_r(J)
__global__ void coset_scale_kernel(u64* scale, unsigned log_n, unsigned rate_bits, u64 w_N, u64 n_inv) {
  uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t n = 1u << log_n;
  if (idx >= (n << rate_bits)) return;
  uint32_t J = idx >> log_n, p = idx & (n - 1);
  uint32_t j = bitrev32(J, rate_bits), k = bitrev32(p, log_n);
  u64 base = gl_mul(7, gl_pow(w_N, j));
  scale[idx] = gl_mul(n_inv, gl_pow(base, k));
}

static const u64* coset_scale(sb_ctx* ctx, unsigned log_n, unsigned rate_bits) {
  uint64_t key = ((uint64_t)log_n << 8) | rate_bits;
  auto it = ctx->coset_scale.find(key);
  if (it != ctx->coset_scale.end()) return it->second.as<u64>();
  DevBuf& b = ctx->coset_scale[key];
  uint32_t total = 1u << (log_n + rate_bits);
  b.ensure(8ull * total);
  LAUNCH(ctx, coset_scale_kernel, (total + 255) / 256, 256, 0, b.as<u64>(), log_n, rate_bits,
         gl_root(log_n + rate_bits), gl_inv(1ull << log_n));
  return b.as<u64>();
}

// ---------------------------------------------------------------------------------------------------------
// shared-memory butterflies.  `buf` holds `cpb` vectors of n = 2^log_n elements; tw is the table of the
// enclosing transform of size 2^log_tw (tw[k] = w^k), so stage `half` uses tw[j << (log_tw - 1 - log2(half))].
// ---------------------------------------------------------------------------------------------------------
template <int EPT>
__device__ __forceinline__ void dif_stages(u64* buf, unsigned log_n, unsigned log_tw, const u64* __restrict__ tw,
                                           unsigned T) {
  // decimation in frequency: natural order in -> bit-reversed order out
  const unsigned half_n = 1u << (log_n - 1);
  for (int s = (int)log_n - 1; s >= 0; s--) {
    const unsigned half = 1u << s;
#pragma unroll
    for (int m = 0; m < EPT / 2; m++) {
      unsigned w = threadIdx.x + m * T;
      unsigned col = w >> (log_n - 1), b = w & (half_n - 1);
      unsigned j = b & (half - 1);
      unsigned i0 = ((b >> s) << (s + 1)) + j + (col << log_n);
      u64 u = buf[i0], v = buf[i0 + half];
      buf[i0] = gl_add(u, v);
      buf[i0 + half] = gl_mul(gl_sub(u, v), __ldg(tw + ((size_t)j << (log_tw - 1 - s))));
    }
    __syncthreads();
  }
}
template <int EPT>
__device__ __forceinline__ void dit_stages(u64* buf, unsigned log_n, unsigned log_tw, const u64* __restrict__ tw,
                                           unsigned T) {
  // decimation in time: bit-reversed order in -> natural order out
  const unsigned half_n = 1u << (log_n - 1);
  for (unsigned s = 0; s < log_n; s++) {
    const unsigned half = 1u << s;
#pragma unroll
    for (int m = 0; m < EPT / 2; m++) {
      unsigned w = threadIdx.x + m * T;
      unsigned col = w >> (log_n - 1), b = w & (half_n - 1);
      unsigned j = b & (half - 1);
      unsigned i0 = ((b >> s) << (s + 1)) + j + (col << log_n);
      u64 u = buf[i0], t = gl_mul(buf[i0 + half], __ldg(tw + ((size_t)j << (log_tw - 1 - s))));
      buf[i0] = gl_add(u, t);
      buf[i0 + half] = gl_sub(u, t);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// K1: fused iNTT -> (n^-1, coset shift) scaling -> 2^r coset NTTs, one pass over the trace.
// Algorithmic HBM bytes per column: 8n read + 8n (coefficients, kept for openings/FRI) + 8N written.
// ---------------------------------------------------------------------------------------------------------
// Where LDE value (column col, position pos) goes.  Local output: [row block][column][position in block] (one block =
// the plain [column][N] layout).  Peer output ("P2P stores fused into the K1 epilogue"): dst_tab[b] is the row
// buffer [C_total][2^log_rb] of the rank that owns row block b -- this GPU's or, over NVLink, a peer's -- so the
// column-sharded LDE lands row-sharded without a separate all-to-all pass.
__device__ __forceinline__ u64* lde_slot(u64* lde, u64* const* dst_tab, uint32_t col0, uint32_t n_cols, uint32_t col, size_t pos,
                                         unsigned log_rb) {
  const size_t blk = pos >> log_rb, r = pos & (((size_t)1 << log_rb) - 1);
  if (dst_tab) return dst_tab[blk] + (((size_t)(col0 + col)) << log_rb) + r;
  return lde + (((blk * n_cols) + col) << log_rb) + r;
}

template <int EPT>
__global__ void __launch_bounds__(EPT == 8 ? 1024 : 512) lde_kernel(const u64* __restrict__ values, u64* __restrict__ coeffs, u64* __restrict__ lde,
                           uint32_t n_cols, unsigned log_n, unsigned rate_bits, unsigned cpb,
                           const u64* __restrict__ tw_fwd, const u64* __restrict__ tw_inv,
                           const u64* __restrict__ scale, u64 n_inv, unsigned log_rb, u64* const* __restrict__ dst_tab, uint32_t dst_col0) {
  extern __shared__ u64 buf[];
  const unsigned T = blockDim.x;
  const uint32_t n = 1u << log_n;
  const uint32_t col0 = blockIdx.x * cpb;
  u64 c[EPT];
#pragma unroll
  for (int m = 0; m < EPT; m++) {
    unsigned e = threadIdx.x + m * T;
    uint32_t col = col0 + (e >> log_n);
    buf[e] = col < n_cols ? gl_canon(values[(size_t)col * n + (e & (n - 1))]) : 0;
  }
  __syncthreads();
  dif_stages<EPT>(buf, log_n, log_n, tw_inv, T);
  // buf[p] = n * c_{bitrev(p)}; the 1/n is folded into the coset scale table
#pragma unroll
  for (int m = 0; m < EPT; m++) c[m] = buf[threadIdx.x + m * T];
  for (unsigned J = 0; J < (1u << rate_bits); J++) {
    __syncthreads();
#pragma unroll
    for (int m = 0; m < EPT; m++) {
      unsigned e = threadIdx.x + m * T;
      buf[e] = gl_mul(c[m], __ldg(scale + ((size_t)J << log_n) + (e & (n - 1))));
    }
    __syncthreads();
    dit_stages<EPT>(buf, log_n, log_n, tw_fwd, T);
#pragma unroll
    for (int m = 0; m < EPT; m++) {
      unsigned e = threadIdx.x + m * T;
      uint32_t col = col0 + (e >> log_n);
      // output order [row block][column][position in block], blocks of 2^log_rb positions: one block (log_rb = log N) is
      // the plain column-major [C][N]; G blocks make every destination rank's slab contiguous for the all-to-all (8e)
      const size_t pos = ((size_t)J << log_n) + (e & (n - 1));
      if (col < n_cols) *lde_slot(lde, dst_tab, dst_col0, n_cols, col, pos, log_rb) = buf[e];
    }
  }
  if (coeffs) {
    const u64 ninv = n_inv;
#pragma unroll
    for (int m = 0; m < EPT; m++) {
      unsigned e = threadIdx.x + m * T;
      uint32_t col = col0 + (e >> log_n);
      if (col < n_cols) coeffs[(size_t)col * n + (e & (n - 1))] = gl_mul(c[m], ninv);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K1, radix-8 form (n >= 64).  The radix-2 kernel above spends its time on shared-memory round trips and barriers
// (log n per transform) with the ALU pipe at 73 %; here a thread keeps eight elements in registers for three stages:
//   inverse (DIF): group (log n - 1 .. log n - 3) straight from global memory (stride n/8: coalesced), middle groups
//                  through shared memory, last group (2, 1, 0) leaves thread t with positions 8t .. 8t+7 = exactly the
//                  eight it needs for the first forward group, so the coefficients never leave its registers;
//   forward (DIT), per coset: scale, group (0, 1, 2) in registers, middle groups, last group (log n - 3 .. log n - 1)
//                  straight to global memory (stride n/8: coalesced).
// Barriers per column: (1 + 2^r) x (1 + middle groups) instead of (1 + 2^r) x log n.  Shared rows are padded by one
// word per 32 so that the "eight consecutive words per thread" accesses stay at two wavefronts per warp.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned padi(unsigned i) { return i + (i >> 5); }

// R stages of decimation in frequency on 2^R registers; element m sits at index i0 + m * unit, `low` = i0 mod unit,
// the top stage is s = log2(unit) + R - 1
template <int R>
__device__ __forceinline__ void dif_group(u64 (&x)[1 << R], unsigned low, unsigned log_unit, unsigned log_tw, const u64* __restrict__ tw) {
#pragma unroll
  for (int st = 0; st < R; st++) {
    constexpr int E = 1 << R;
    const int half_m = E >> (st + 1);
    const unsigned cur = log_unit + R - 1 - st;           // stage index: half = 2^cur
#pragma unroll
    for (int m = 0; m < E; m++) {
      if (!(m & half_m)) {
        const unsigned j = low + ((unsigned)(m & (half_m - 1)) << log_unit);
        const u64 w = __ldg(tw + ((size_t)j << (log_tw - 1 - cur)));
        const u64 u = x[m], v = x[m + half_m];
        x[m] = gl_add(u, v);
        x[m + half_m] = gl_mul(gl_sub(u, v), w);
      }
    }
  }
}
// R stages of decimation in time; the bottom stage is s = log2(unit)
template <int R>
__device__ __forceinline__ void dit_group(u64 (&x)[1 << R], unsigned low, unsigned log_unit, unsigned log_tw, const u64* __restrict__ tw) {
#pragma unroll
  for (int st = 0; st < R; st++) {
    constexpr int E = 1 << R;
    const int half_m = 1 << st;
    const unsigned cur = log_unit + st;
#pragma unroll
    for (int m = 0; m < E; m++) {
      if (!(m & half_m)) {
        const unsigned j = low + ((unsigned)(m & (half_m - 1)) << log_unit);
        const u64 w = __ldg(tw + ((size_t)j << (log_tw - 1 - cur)));
        const u64 u = x[m], t = gl_mul(x[m + half_m], w);
        x[m] = gl_add(u, t);
        x[m + half_m] = gl_sub(u, t);
      }
    }
  }
}
// one middle pass over a column held in (padded) shared memory: groups of R stages starting at stage `base` (DIT:
// bottom stage, DIF: the group covers base + R - 1 .. base); every thread does 8 / 2^R work items
template <int R, bool DIF>
__device__ __forceinline__ void smem_pass(u64* b, unsigned t, unsigned log_n, unsigned base, const u64* __restrict__ tw) {
  constexpr int E = 1 << R, ITEMS = 8 / E;
  const unsigned per_col = 1u << (log_n - 3);
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const unsigned wi = t + it * per_col;                              // work item in [0, n / 2^R)
    const unsigned low = wi & ((1u << base) - 1), high = wi >> base;
    const unsigned i0 = (high << (base + R)) | low;
    u64 x[E];
#pragma unroll
    for (int m = 0; m < E; m++) x[m] = b[padi(i0 + ((unsigned)m << base))];
    if (DIF) dif_group<R>(x, low, base, log_n, tw); else dit_group<R>(x, low, base, log_n, tw);
#pragma unroll
    for (int m = 0; m < E; m++) b[padi(i0 + ((unsigned)m << base))] = x[m];
  }
}

__global__ void __launch_bounds__(1024) lde8_kernel(const u64* __restrict__ values, u64* __restrict__ coeffs, u64* __restrict__ lde,
                                                   uint32_t n_cols, unsigned log_n, unsigned rate_bits, unsigned log_cpb,
                                                   const u64* __restrict__ tw_fwd, const u64* __restrict__ tw_inv,
                                                   const u64* __restrict__ scale, u64 n_inv, unsigned log_rb, u64* const* __restrict__ dst_tab, uint32_t dst_col0) {
  extern __shared__ u64 buf[];
  const uint32_t n = 1u << log_n, unit = n >> 3;
  const unsigned lc = threadIdx.x >> (log_n - 3), t = threadIdx.x & (unit - 1);
  const uint32_t col = (blockIdx.x << log_cpb) + lc;
  const bool live = col < n_cols;
  u64* b = buf + (size_t)lc * (n + (n >> 5));
  const unsigned mid = log_n - 6, r = mid % 3;          // middle stages 3 .. log n - 4: mid / 3 full groups + one of r
  u64 x[8], c[8];

  // ---- inverse transform, decimation in frequency ----
  const u64* vin = values + (size_t)col * n + t;
#pragma unroll
  for (int m = 0; m < 8; m++) x[m] = live ? gl_canon(vin[(size_t)m * unit]) : 0;   // plonky2 fields may hold [p, 2^64)
  dif_group<3>(x, t, log_n - 3, log_n, tw_inv);
#pragma unroll
  for (int m = 0; m < 8; m++) b[padi(t + m * unit)] = x[m];
  __syncthreads();
  for (int base = (int)log_n - 6; base >= 3 + (int)r; base -= 3) { smem_pass<3, true>(b, t, log_n, base, tw_inv); __syncthreads(); }
  if (r == 2) { smem_pass<2, true>(b, t, log_n, 3, tw_inv); __syncthreads(); }
  if (r == 1) { smem_pass<1, true>(b, t, log_n, 3, tw_inv); __syncthreads(); }
#pragma unroll
  for (int m = 0; m < 8; m++) x[m] = b[padi(8 * t + m)];
  dif_group<3>(x, 0, 0, log_n, tw_inv);
#pragma unroll
  for (int m = 0; m < 8; m++) c[m] = x[m];              // n * c_{bitrev(8t + m)}
  if (coeffs && live) {
    u64* co = coeffs + (size_t)col * n + 8 * t;
#pragma unroll
    for (int m = 0; m < 8; m += 2) *(ulonglong2*)(co + m) = make_ulonglong2(gl_mul(c[m], n_inv), gl_mul(c[m + 1], n_inv));
  }

  // ---- forward transforms, decimation in time, one per coset ----
  for (unsigned J = 0; J < (1u << rate_bits); J++) {
    const u64* sc = scale + ((size_t)J << log_n) + 8 * t;
#pragma unroll
    for (int m = 0; m < 8; m++) x[m] = gl_mul(c[m], __ldg(sc + m));
    dit_group<3>(x, 0, 0, log_n, tw_fwd);
    __syncthreads();                                      // everyone is done reading the previous contents of b
#pragma unroll
    for (int m = 0; m < 8; m++) b[padi(8 * t + m)] = x[m];
    __syncthreads();
    if (r == 1) { smem_pass<1, false>(b, t, log_n, 3, tw_fwd); __syncthreads(); }
    if (r == 2) { smem_pass<2, false>(b, t, log_n, 3, tw_fwd); __syncthreads(); }
    for (unsigned base = 3 + r; base + 3 <= log_n - 3; base += 3) { smem_pass<3, false>(b, t, log_n, base, tw_fwd); __syncthreads(); }
#pragma unroll
    for (int m = 0; m < 8; m++) x[m] = b[padi(t + m * unit)];
    dit_group<3>(x, t, log_n - 3, log_n, tw_fwd);
    if (live) {
#pragma unroll
      for (int m = 0; m < 8; m++) {
        const size_t pos = ((size_t)J << log_n) + t + (size_t)m * unit;
        *lde_slot(lde, dst_tab, dst_col0, n_cols, col, pos, log_rb) = x[m];
      }
    }
  }
}

void sb_lde_trace(sb_ctx* ctx, const u64* d_values, u64* d_coeffs, u64* d_lde, uint32_t n_cols, unsigned log_n,
                  unsigned rate_bits, unsigned log_row_blocks, u64* const* d_dst_tab, uint32_t col0) {
  const unsigned log_rb = log_n + rate_bits - log_row_blocks;
  if (log_n < 1 || log_n > 13) SB_THROW(SB_EINVAL, "trace height 2^%u unsupported (1 <= log_n <= 13)", log_n);
  const Twiddles& tw = sb_twiddles(ctx, log_n);
  const u64* scale = coset_scale(ctx, log_n, rate_bits);
  const uint32_t n = 1u << log_n;
  if (log_n >= 6 && !getenv("SB_LDE_RADIX2")) {
    // radix-8 kernel: n / 8 threads per column, at least 256 threads per block
    unsigned log_cpb = 0;
    while (((n >> 3) << log_cpb) < 256) log_cpb++;
    const unsigned T = (n >> 3) << log_cpb;
    const size_t smem8 = 8ull * ((size_t)(n + (n >> 5)) << log_cpb);
    // per device (function attributes are per context): a cheap host call, so no process-wide "done" flag
    CUDA_CHECK(cudaFuncSetAttribute(lde8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (8192 + 256)));
    LAUNCH(ctx, lde8_kernel, (n_cols + (1u << log_cpb) - 1) >> log_cpb, T, smem8, d_values, d_coeffs, d_lde, n_cols, log_n, rate_bits,
           log_cpb, tw.fwd.as<u64>(), tw.inv.as<u64>(), scale, gl_inv((u64)n), log_rb, d_dst_tab, col0);
    return;
  }
  // block shape: EPT elements per thread, cpb columns per block so that a block has >= 128 threads
  int ept = n >= 4096 ? 8 : 4;
  unsigned cpb = 1;
  while (cpb * n / ept < 128) cpb *= 2;
  unsigned T = cpb * n / ept;
  size_t smem = 8ull * cpb * n;
  uint32_t grid = (n_cols + cpb - 1) / cpb;
  if (ept == 8) {
    CUDA_CHECK(cudaFuncSetAttribute(lde_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(ctx, lde_kernel<8>, grid, T, smem, d_values, d_coeffs, d_lde, n_cols, log_n, rate_bits, cpb,
           tw.fwd.as<u64>(), tw.inv.as<u64>(), scale, gl_inv((u64)n), log_rb, d_dst_tab, col0);
  } else {
    CUDA_CHECK(cudaFuncSetAttribute(lde_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(ctx, lde_kernel<4>, grid, T, smem, d_values, d_coeffs, d_lde, n_cols, log_n, rate_bits, cpb,
           tw.fwd.as<u64>(), tw.inv.as<u64>(), scale, gl_inv((u64)n), log_rb, d_dst_tab, col0);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Generic batched transform of `count` vectors of size 2^log_size (quotient polys, FRI polynomials, tests).
//   dif = true : natural in  -> bit-reversed out      dif = false: bit-reversed in -> natural out
//   inverse    : use w^-1 and scale by size^-1
// Stages with half >= 2^LOG_BLOCK run as one global pass each; the rest run in shared memory per 2^LOG_BLOCK block.
// ---------------------------------------------------------------------------------------------------------
__global__ void ntt_global_stage_kernel(u64* data, unsigned log_size, unsigned s, const u64* __restrict__ tw,
                                        bool dif, uint64_t total_butterflies) {
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= total_butterflies) return;
  const uint64_t half_size = 1ull << (log_size - 1);
  uint64_t vec = w >> (log_size - 1), b = w & (half_size - 1);
  uint64_t half = 1ull << s, j = b & (half - 1);
  u64* d = data + (vec << log_size);
  uint64_t i0 = ((b >> s) << (s + 1)) + j;
  u64 t = __ldg(tw + (j << (log_size - 1 - s)));
  u64 u = d[i0], v = d[i0 + half];
  if (dif) {
    d[i0] = gl_add(u, v);
    d[i0 + half] = gl_mul(gl_sub(u, v), t);
  } else {
    v = gl_mul(v, t);
    d[i0] = gl_add(u, v);
    d[i0 + half] = gl_sub(u, v);
  }
}

template <int EPT>
__global__ void __launch_bounds__(1024) ntt_block_kernel(u64* data, unsigned log_size, unsigned log_block, const u64* __restrict__ tw,
                                 bool dif, u64 post_scale) {
  extern __shared__ u64 buf[];
  const unsigned T = blockDim.x;
  u64* d = data + ((size_t)blockIdx.x << log_block);
#pragma unroll
  for (int m = 0; m < EPT; m++) buf[threadIdx.x + m * T] = d[threadIdx.x + m * T];
  __syncthreads();
  if (dif) dif_stages<EPT>(buf, log_block, log_size, tw, T);
  else dit_stages<EPT>(buf, log_block, log_size, tw, T);
#pragma unroll
  for (int m = 0; m < EPT; m++) {
    u64 v = buf[threadIdx.x + m * T];
    d[threadIdx.x + m * T] = post_scale == 1 ? v : gl_mul(v, post_scale);
  }
}

__global__ void scale_kernel(u64* d, uint64_t n, u64 s) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = gl_mul(d[i], s);
}
void sb_scale_device(sb_ctx* ctx, u64* d, uint64_t n, u64 s) {
  LAUNCH(ctx, scale_kernel, (unsigned)((n + 255) / 256), 256, 0, d, n, s);
}

void sb_ntt_device(sb_ctx* ctx, u64* d_data, unsigned log_size, uint32_t count, bool inverse, bool dif) {
  if (log_size == 0 || count == 0) return;
  const Twiddles& tw = sb_twiddles(ctx, log_size);
  const u64* t = inverse ? tw.inv.as<u64>() : tw.fwd.as<u64>();
  unsigned log_block = log_size < 12 ? log_size : 12;
  u64 post = inverse ? gl_inv(1ull << log_size) : 1;
  uint64_t total_bf = (uint64_t)count << (log_size - 1);
  auto global_stages = [&](bool descending) {
    if (descending) for (int s = (int)log_size - 1; s >= (int)log_block; s--)
      LAUNCH(ctx, ntt_global_stage_kernel, (unsigned)((total_bf + 255) / 256), 256, 0, d_data, log_size, (unsigned)s, t, true, total_bf);
    else for (unsigned s = log_block; s < log_size; s++)
      LAUNCH(ctx, ntt_global_stage_kernel, (unsigned)((total_bf + 255) / 256), 256, 0, d_data, log_size, s, t, false, total_bf);
  };
  auto block_pass = [&](u64 scale) {
    uint32_t blocks = count << (log_size - log_block);
    uint32_t nb = 1u << log_block;
    size_t smem = 8ull * nb;
    if (nb >= 1024) {
      unsigned T = nb / 4;
      LAUNCH(ctx, ntt_block_kernel<4>, blocks, T, smem, d_data, log_size, log_block, t, dif, scale);
    } else if (nb >= 64) {
      unsigned T = nb / 2;
      LAUNCH(ctx, ntt_block_kernel<2>, blocks, T, smem, d_data, log_size, log_block, t, dif, scale);
    } else {
      // tiny transforms: one thread per butterfly, at least one
      unsigned T = nb / 2;
      LAUNCH(ctx, ntt_block_kernel<2>, blocks, T, smem, d_data, log_size, log_block, t, dif, scale);
    }
  };
  if (dif) { global_stages(true); block_pass(post); }
  else {
    if (log_block == log_size) block_pass(post);
    else {
      block_pass(1); global_stages(false);
      if (post != 1) sb_scale_device(ctx, d_data, (uint64_t)count << log_size, post);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// row-major [n][C] (u64 or u32) -> column-major [C][n] u64: the device replacement of
// starky::util::trace_rows_to_poly_values (aggregate_proof.rs:57,104,137,168,211).  32x32 smem tiles.
// ---------------------------------------------------------------------------------------------------------
template <class T>
__global__ void transpose_kernel(const T* __restrict__ rows, u64* __restrict__ cols, uint32_t n_rows, uint32_t n_cols) {
  __shared__ u64 tile[32][33];
  uint32_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    uint32_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < n_rows && c < n_cols) ? gl_canon((u64)rows[(size_t)r * n_cols + c]) : 0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    uint32_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < n_rows && c < n_cols) cols[(size_t)c * n_rows + r] = tile[threadIdx.x][i];
  }
}
void sb_transpose_rows_to_cols(sb_ctx* ctx, const void* d_rows, u64* d_cols, uint32_t n_rows, uint32_t n_cols, bool is_u32) {
  dim3 grid((n_cols + 31) / 32, (n_rows + 31) / 32), block(32, 8);
  if (is_u32) { LAUNCH(ctx, transpose_kernel<uint32_t>, grid, block, 0, (const uint32_t*)d_rows, d_cols, n_rows, n_cols); }
  else { LAUNCH(ctx, transpose_kernel<u64>, grid, block, 0, (const u64*)d_rows, d_cols, n_rows, n_cols); }
}
