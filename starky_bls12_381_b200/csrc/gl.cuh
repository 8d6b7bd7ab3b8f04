// Goldilocks field p = 2^64 - 2^32 + 1 and its quadratic extension F_p[X]/(X^2-7) for sm_100a.
// Replaces plonky2_field::goldilocks_field::GoldilocksField on the prove() hot path
// (reference call sites /root/reference/src/aggregate_proof.rs:59,105,138,169,212).
// 64x64 products lower to IMAD.WIDE.U32 chains -- the integer (FMA) pipe is this library's compute roofline.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#endif

typedef uint64_t u64;
typedef uint32_t u32;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL

// The three primitives below are on every kernel's inner loop and every kernel here is ALU-pipe bound (ncu: K1 78 %,
// K4 64 %), so the device paths are 32-bit carry chains instead of 64-bit compares and selects:
//   x >= p  <=>  x + eps carries out of 64 bits (eps = 2^32 - 1 = 2^64 - p), and x - p = x + eps (mod 2^64).
GL_HD u64 gl_canon(u64 a) {
#if defined(__CUDA_ARCH__)
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), r0, r1;
  asm("{\n\t"
      ".reg .u32 t0, t1, c;\n\t"
      ".reg .pred q;\n\t"
      "add.cc.u32 t0, %2, 0xFFFFFFFF;\n\t"
      "addc.cc.u32 t1, %3, 0;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "setp.ne.u32 q, c, 0;\n\t"
      "selp.u32 %0, t0, %2, q;\n\t"
      "selp.u32 %1, t1, %3, q;\n\t"
      "}"
      : "=r"(r0), "=r"(r1) : "r"(a0), "r"(a1));
  return ((u64)r1 << 32) | r0;
#else
  return a >= GL_P ? a - GL_P : a;
#endif
}

// canonical + canonical -> canonical
GL_HD u64 gl_add(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 s0, s1, t0, t1, c;\n\t"
      ".reg .pred q;\n\t"
      "add.cc.u32 s0, %2, %4;\n\t"           // s = a + b  (65 bits: c:s)
      "addc.cc.u32 s1, %3, %5;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "add.cc.u32 t0, s0, 0xFFFFFFFF;\n\t"   // t = s - p (mod 2^64); carries iff s >= p
      "addc.cc.u32 t1, s1, 0;\n\t"
      "addc.u32 c, c, 0;\n\t"
      "setp.ne.u32 q, c, 0;\n\t"
      "selp.u32 %0, t0, s0, q;\n\t"
      "selp.u32 %1, t1, s1, q;\n\t"
      "}"
      : "=r"(r0), "=r"(r1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
  return ((u64)r1 << 32) | r0;
#else
  u64 s = a + b;
  return (s < a) ? s + GL_EPS : (s >= GL_P ? s - GL_P : s);
#endif
}
// canonical - canonical -> canonical
GL_HD u64 gl_sub(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 m;\n\t"
      "sub.cc.u32 %0, %2, %4;\n\t"
      "subc.cc.u32 %1, %3, %5;\n\t"
      "subc.u32 m, 0, 0;\n\t"                // 0xFFFFFFFF on borrow: + p = - eps (mod 2^64)
      "sub.cc.u32 %0, %0, m;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
  return ((u64)r1 << 32) | r0;
#else
  u64 d = a - b;
  return (a < b) ? d + GL_P : d;
#endif
}
GL_HD u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }

GL_HD void gl_mul_wide(u64 a, u64 b, u64& lo, u64& hi) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umul64hi(a, b);
#else
  unsigned __int128 m = (unsigned __int128)a * b;
  lo = (u64)m;
  hi = (u64)(m >> 64);
#endif
}
// (hi:lo) mod p, result in [0, 2^64) (not necessarily canonical).  2^64 = 2^32-1, 2^96 = -1 (mod p).
GL_HD u64 gl_reduce128_lazy(u64 lo, u64 hi) {
  u64 hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
  u64 t0 = lo - hi_hi;
  if (lo < hi_hi) t0 -= GL_EPS;
  u64 t1 = (hi_lo << 32) - hi_lo;  // hi_lo * (2^32 - 1)
  u64 r = t0 + t1;
  if (r < t1) r += GL_EPS;
  return r;
}
// any u64 x any u64 -> lazy.
// Device path.  The 128-bit product x3:x2:x1:x0 is left to ptxas (mul.lo.u64 + mul.hi.u64 share four IMAD.WIDE.U32,
// two of them with carry-out / carry-in, seven instructions in four dependent levels -- shorter than anything that
// can be spelled in PTX).  The Goldilocks fold
//     x0 + x1 2^32 + x2 2^64 + x3 2^96  =  (x1:x0) + x2 eps - x3          (2^64 = eps = 2^32-1, 2^96 = -1 mod p)
// is written with 32-bit carry chains instead of 64-bit compares and selects:
//     w  = x2 * eps + x0 = (x2 : x0) - x2                       (fits 64 bits: <= (2^32-1)^2 + 2^32-1)
//     w += x1 * 2^32   -> carry  c1          w -= x3   -> borrow b1
//     r  = w + c1 * eps - b1 * eps            (a carry / borrow of 2^64 is worth eps mod p)
// If only c1 is set the wrapped w is <= 2^64 - 2^33, if only b1 is set it is >= 2^64 - 2^32 + 1, so neither repair
// can wrap; if both are set they cancel in Z/2^64.  12 ALU instructions in 8 dependent levels.
GL_HD u64 gl_mul_lazy(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  const u64 lo = a * b, hi = __umul64hi(a, b);
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 nc, nb;\n\t"
      "sub.cc.u32 %0, %2, %4;\n\t"        // w = (x2:x0) - x2
      "subc.u32 %1, %4, 0;\n\t"
      "add.cc.u32 %1, %1, %3;\n\t"        // w += x1 * 2^32
      "addc.u32 nc, 0, 0;\n\t"            // carry c1 (an add-chain flag must not feed subc: the borrow sense differs)
      "neg.s32 nc, nc;\n\t"               // 0 or 0xFFFFFFFF = c1 * eps
      "sub.cc.u32 %0, %0, %5;\n\t"        // w -= x3
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.u32 nb, 0, 0;\n\t"            // 0 - borrow: 0 or 0xFFFFFFFF = b1 * eps
      "add.cc.u32 %0, %0, nc;\n\t"        // + c1 * eps
      "addc.u32 %1, %1, 0;\n\t"
      "sub.cc.u32 %0, %0, nb;\n\t"        // - b1 * eps
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"((u32)lo), "r"((u32)(lo >> 32)), "r"((u32)hi), "r"((u32)(hi >> 32)));
  return ((u64)r1 << 32) | r0;
#else
  u64 lo, hi;
  gl_mul_wide(a, b, lo, hi);
  return gl_reduce128_lazy(lo, hi);
#endif
}
// a * b + c for any u64 a, b, c -> lazy.  The addend joins the 128-bit product before the fold (no carry out:
// (2^64-1)^2 + 2^64-1 < 2^128), so a multiply-add costs four instructions more than a multiply.
GL_HD u64 gl_mad_lazy(u64 a, u64 b, u64 c) {
#if defined(__CUDA_ARCH__)
  u64 lo = a * b, hi = __umul64hi(a, b);
  lo += c;
  hi += (lo < c);
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 nc, nb;\n\t"
      "sub.cc.u32 %0, %2, %4;\n\t"
      "subc.u32 %1, %4, 0;\n\t"
      "add.cc.u32 %1, %1, %3;\n\t"
      "addc.u32 nc, 0, 0;\n\t"
      "neg.s32 nc, nc;\n\t"
      "sub.cc.u32 %0, %0, %5;\n\t"
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.u32 nb, 0, 0;\n\t"
      "add.cc.u32 %0, %0, nc;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "sub.cc.u32 %0, %0, nb;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"((u32)lo), "r"((u32)(lo >> 32)), "r"((u32)hi), "r"((u32)(hi >> 32)));
  return ((u64)r1 << 32) | r0;
#else
  unsigned __int128 m = (unsigned __int128)a * b + c;
  return gl_reduce128_lazy((u64)m, (u64)(m >> 64));
#endif
}
// any x any -> canonical
GL_HD u64 gl_mul(u64 a, u64 b) { return gl_canon(gl_mul_lazy(a, b)); }
GL_HD u64 gl_sqr(u64 a) { return gl_mul(a, a); }

// a * 2^32 for canonical a -> canonical  (the carry weight in almost every reference constraint)
GL_HD u64 gl_mul_2_32(u64 a) {
  // a = ah*2^32 + al ; a*2^32 = ah*2^64 + al*2^32 = ah*(2^32-1) + al*2^32
  u64 ah = a >> 32, al = a & GL_EPS;
  u64 t = (ah << 32) - ah;       // < 2^64 - 2^33
  u64 u = al << 32;              // < 2^64
  u64 r = t + u;
  if (r < t) r += GL_EPS;
  return gl_canon(r);
}

GL_HD u64 gl_pow(u64 a, u64 e) {
  u64 r = 1;
  while (e) {
    if (e & 1) r = gl_mul(r, a);
    a = gl_mul(a, a);
    e >>= 1;
  }
  return r;
}
GL_HD u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }

// primitive_root_of_unity(log_n) = POWER_OF_TWO_GENERATOR^(2^(32-log_n))
GL_HD u64 gl_root(unsigned log_n) {
  u64 r = 1753635133440165772ULL;
  for (unsigned i = log_n; i < 32; i++) r = gl_mul(r, r);
  return r;
}

struct e2_t {
  u64 a, b;
};
GL_HD e2_t e2_make(u64 a, u64 b) { e2_t r; r.a = a; r.b = b; return r; }
GL_HD e2_t e2_add(e2_t x, e2_t y) { return e2_make(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
GL_HD e2_t e2_sub(e2_t x, e2_t y) { return e2_make(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
GL_HD e2_t e2_mul(e2_t x, e2_t y) {
  u64 bb = gl_mul(x.b, y.b);
  u64 seven_bb = gl_add(gl_add(gl_add(bb, bb), gl_add(bb, bb)), gl_add(gl_add(bb, bb), bb));
  return e2_make(gl_add(gl_mul(x.a, y.a), seven_bb), gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a)));
}
GL_HD e2_t e2_scale(e2_t x, u64 s) { return e2_make(gl_mul(x.a, s), gl_mul(x.b, s)); }
GL_HD e2_t e2_inv(e2_t x) {
  u64 d = gl_inv(gl_sub(gl_mul(x.a, x.a), gl_mul(7, gl_mul(x.b, x.b))));
  return e2_make(gl_mul(x.a, d), gl_mul(gl_neg(x.b), d));
}
GL_HD e2_t e2_pow(e2_t x, u64 e) {
  e2_t r = e2_make(1, 0);
  while (e) {
    if (e & 1) r = e2_mul(r, x);
    x = e2_mul(x, x);
    e >>= 1;
  }
  return r;
}
GL_HD bool e2_eq(e2_t x, e2_t y) { return x.a == y.a && x.b == y.b; }

GL_HD u32 bitrev32(u32 x, u32 bits) {
#if defined(__CUDA_ARCH__)
  return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
  u32 r = 0;
  for (u32 i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
#endif
}
