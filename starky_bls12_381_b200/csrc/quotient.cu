// K4: quotient evaluation -- the reference's constraints at every LDE point, folded with the alpha challenges.
// Replaces starky::prover::compute_quotient_polys + ConstraintConsumer + the per-point callback into
// Stark::eval_packed_generic (fp12_mul.rs:58, calc_pairing_precomp.rs:376, miller_loop.rs:644,
// final_exponentiate.rs:907, ecc_aggregate.rs:92 and the gadgets in fp.rs/fp2.rs/fp6.rs/fp12.rs/g1.rs); SURVEY.md A.8.
//
// The C side cannot call back into Rust per point, so each stark's constraints arrive as a compiled constraint
// program (tools/airgen: symbolic execution of eval_packed_generic -> grouped bytecode, format in
// tools/airgen/compile.py).  One thread owns one LDE position and interprets the (warp-uniform) instruction stream;
// the grid's second dimension splits the program into chunks of whole groups, so that even N = 2048 points fill
// 148 SMs.  acc_j = sum_k alpha_j^(K-1-k) f_k c_k is evaluated as
//     sum_groups  S_g * f_cls(g) * sum_{slots in g} W_slot,j * body_slot
// with W (sums of alpha powers) precomputed per proof, the inner sums accumulated as unreduced 192-bit integers
// (one 64x64->128 multiply-add per challenge per body, no modular reduction) and reduced once per group.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "prover.cuh"

enum { AIR_CODE_PAD = 8 };
enum { OP_NOP = 0, OP_ADD1, OP_ADD2, OP_SHL1, OP_MULS, OP_MUL2, OP_MULC1, OP_MULC2, OP_MUL3C, OP_CONSTI, OP_CONSTC, OP_GROUP };

struct AirProgram {
  uint32_t n_cols = 0, n_pis = 0, degree = 0, K = 0, n_code = 0, n_consts = 0, n_slots = 0, n_groups = 0;
  DevBuf code, consts, slot_off, slot_ks, pw, wt, chunks, part;
  std::vector<uint32_t> group_pc, group_slot;
  uint32_t n_chunks = 0;
};

void air_release_all(sb_ctx* ctx) {
  for (auto& kv : ctx->airs) {
    AirProgram* a = kv.second;
    DevBuf* bufs[] = {&a->code, &a->consts, &a->slot_off, &a->slot_ks, &a->pw, &a->wt, &a->chunks, &a->part};
    for (DevBuf* b : bufs) b->release();
    delete a;
  }
  ctx->airs.clear();
}

// Parses one SBAIRBN1 image (tools/airgen/compile.py: write_airbin) from memory and binds it to `stark_id` on this ctx.
// `what` names the source in error messages (a path, or "embedded:<name>").
static void air_load_image(sb_ctx* ctx, uint32_t stark_id, const unsigned char* img, size_t img_len, const char* path) {
  struct Hdr { char magic[8]; uint32_t v[12]; } h;
  std::vector<u64> code, consts;
  std::vector<uint32_t> slot_off, slot_ks, gpc, gslot;
  size_t off = 0;
  auto rd = [&](void* p, size_t sz, size_t n) {
    if (n == 0) return true;
    if (off + sz * n > img_len) return false;
    memcpy(p, img + off, sz * n);
    off += sz * n;
    return true;
  };
  bool ok = rd(&h, sizeof(h), 1) && memcmp(h.magic, "SBAIRBN1", 8) == 0;
  AirProgram* a = new AirProgram();
  if (ok) {
    a->n_cols = h.v[0]; a->n_pis = h.v[1]; a->degree = h.v[2]; a->K = h.v[3]; a->n_code = h.v[4]; a->n_consts = h.v[5];
    a->n_slots = h.v[6]; a->n_groups = h.v[7];
    ok = (uint64_t)a->n_code * 8 <= img_len && (uint64_t)a->n_consts * 8 <= img_len && (uint64_t)a->n_slots * 4 <= img_len &&
         (uint64_t)a->K * 4 <= img_len && (uint64_t)a->n_groups * 4 <= img_len;
  }
  if (ok) {
    code.resize(a->n_code + 1); consts.resize(a->n_consts + 1); /* consts.back(): spare slot */ slot_off.resize(a->n_slots + 1); slot_ks.resize(a->K);
    gpc.resize(a->n_groups + 1); gslot.resize(a->n_groups + 1);
    ok = rd(code.data(), 8, a->n_code) && rd(consts.data(), 8, a->n_consts) && rd(slot_off.data(), 4, a->n_slots + 1) &&
         rd(slot_ks.data(), 4, a->K) && rd(gpc.data(), 4, a->n_groups + 1) && rd(gslot.data(), 4, a->n_groups + 1);
  }
  if (!ok) { delete a; SB_THROW(SB_EAIR, "malformed constraint program %s", path); }
  // validate: every variable index in range, group table consistent
  const uint32_t n_vars = 2 * a->n_cols + a->n_pis;
  for (uint32_t pc = 0; pc < a->n_code; pc++) {
    u64 w = code[pc];
    unsigned op = w & 15;
    uint32_t v0 = (w >> 8) & 0x3FFFF, v1 = (w >> 26) & 0x3FFFF, v2 = (w >> 44) & 0x3FFFF;
    bool bad = false;
    switch (op) {
      case OP_ADD1: case OP_SHL1: case OP_MULS: bad = v0 >= n_vars; break;
      case OP_ADD2: case OP_MUL2: bad = v0 >= n_vars || v1 >= n_vars; break;
      case OP_MULC1: bad = v0 >= n_vars || v2 >= a->n_consts; break;
      case OP_MULC2: bad = v0 >= n_vars || v1 >= n_vars || v2 >= a->n_consts; break;
      case OP_MUL3C: bad = v0 >= n_vars || v1 >= n_vars || v2 >= n_vars || pc + 1 >= a->n_code; pc++; break;
      case OP_CONSTC: bad = v2 >= a->n_consts; break;
      case OP_GROUP: bad = v0 < 1 || v0 > 4; break;
      case OP_NOP: case OP_CONSTI: break;
      default: bad = true;
    }
    if (bad) { delete a; SB_THROW(SB_EAIR, "constraint program %s: invalid instruction at pc %u", path, pc); }
  }
  // Fast words.  ADD1 / ADD2 / SHL1 / MUL2 whose operands are all LOCAL columns are 85-90 % of every program (limb sums,
  // carries times 2^32, limb products); they are re-encoded for a branch-light path of the interpreter:
  //   bit 7 set, v0 in bits 8..31, v1 in bits 32..55 (one shift / one mask each), op / end / neg0 / neg1 unchanged.
  for (uint32_t pc = 0; pc < a->n_code; pc++) {
    const u64 w = code[pc];
    const unsigned op = w & 15;
    const u64 v0 = (w >> 8) & 0x3FFFF, v1 = (w >> 26) & 0x3FFFF;
    if (op == OP_MUL3C) {       // the inline constant word would look like an instruction to the operand prefetcher
      consts.back() = code[pc + 1];
      code[pc + 1] = (u64)OP_NOP | ((u64)(consts.size() - 1) << 8);
      consts.push_back(0);
      pc++;
      continue;
    }
    const bool two = op == OP_ADD2 || op == OP_MUL2;
    if ((op == OP_ADD1 || op == OP_SHL1 || two) && v0 < a->n_cols && (!two || v1 < a->n_cols))
      code[pc] = (w & 0x7F) | 0x80 | (v0 << 8) | ((two ? v1 : 0) << 32);
  }
  a->n_consts = (uint32_t)consts.size() - 1;
  code.resize(a->n_code + AIR_CODE_PAD, (u64)OP_NOP);   // padding for the look-ahead
  a->code.ensure(8ull * (a->n_code + AIR_CODE_PAD));
  a->consts.ensure(8ull * (a->n_consts + 1));
  a->slot_off.ensure(4ull * (a->n_slots + 1));
  a->slot_ks.ensure(4ull * (a->K + 1));
  CUDA_CHECK(cudaMemcpyAsync(a->code.p, code.data(), 8ull * (a->n_code + AIR_CODE_PAD), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(a->consts.p, consts.data(), 8ull * a->n_consts, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(a->slot_off.p, slot_off.data(), 4ull * (a->n_slots + 1), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(a->slot_ks.p, slot_ks.data(), 4ull * a->K, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  a->group_pc = gpc; a->group_slot = gslot;
  auto it = ctx->airs.find(stark_id);
  if (it != ctx->airs.end()) {
    DevBuf* bufs[] = {&it->second->code, &it->second->consts, &it->second->slot_off, &it->second->slot_ks,
                      &it->second->pw, &it->second->wt, &it->second->chunks, &it->second->part};
    for (DevBuf* b : bufs) b->release();
    delete it->second;
  }
  ctx->airs[stark_id] = a;
}

static void air_load_file(sb_ctx* ctx, uint32_t stark_id, const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) SB_THROW(SB_EAIR, "cannot open constraint program %s", path);
  std::vector<unsigned char> img;
  unsigned char buf[1 << 16];
  size_t got;
  while ((got = fread(buf, 1, sizeof(buf), f)) > 0) img.insert(img.end(), buf, buf + got);
  fclose(f);
  air_load_image(ctx, stark_id, img.data(), img.size(), path);
}

// The five standard programs are linked into the library (air_blobs.S: .incbin of the images the Makefile unpacks from
// starky_bls12_381_b200/air/*.airbin.xz), so sb_prove needs no file at run time.  $SB_AIR_DIR, when set, names a
// directory of <name>.airbin files that take precedence (regenerated programs, experiments).
extern "C" {
extern const unsigned char sb_airbin_fp12_mul[], sb_airbin_fp12_mul_end[];
extern const unsigned char sb_airbin_pairing_precomp[], sb_airbin_pairing_precomp_end[];
extern const unsigned char sb_airbin_miller_loop[], sb_airbin_miller_loop_end[];
extern const unsigned char sb_airbin_final_exp[], sb_airbin_final_exp_end[];
extern const unsigned char sb_airbin_ecc_agg[], sb_airbin_ecc_agg_end[];
}

AirProgram* air_get(sb_ctx* ctx, const sb_params* p) {
  auto it = ctx->airs.find(p->stark_id);
  if (it == ctx->airs.end()) {
    static const char* names[] = {"fp12_mul", "pairing_precomp", "miller_loop", "final_exp", "ecc_agg"};
    if (p->stark_id > SB_STARK_ECC_AGG) SB_THROW(SB_EAIR, "no constraint program loaded for stark id %u (sb_air_load)", p->stark_id);
    const char* env = getenv("SB_AIR_DIR");
    if (env && *env) {
      std::string path = std::string(env) + "/" + names[p->stark_id] + ".airbin";
      air_load_file(ctx, p->stark_id, path.c_str());
    } else {
      const unsigned char* b[5] = {sb_airbin_fp12_mul, sb_airbin_pairing_precomp, sb_airbin_miller_loop, sb_airbin_final_exp, sb_airbin_ecc_agg};
      const unsigned char* e[5] = {sb_airbin_fp12_mul_end, sb_airbin_pairing_precomp_end, sb_airbin_miller_loop_end, sb_airbin_final_exp_end,
                                   sb_airbin_ecc_agg_end};
      const std::string what = std::string("embedded:") + names[p->stark_id];
      air_load_image(ctx, p->stark_id, b[p->stark_id], (size_t)(e[p->stark_id] - b[p->stark_id]), what.c_str());
    }
    it = ctx->airs.find(p->stark_id);
  }
  AirProgram* a = it->second;
  if (a->n_cols != p->n_cols || a->n_pis != p->n_public_inputs)
    SB_THROW(SB_EINVAL, "stark %u: params say %u columns / %u public inputs, constraint program has %u / %u", p->stark_id,
             p->n_cols, p->n_public_inputs, a->n_cols, a->n_pis);
  if (a->degree != p->constraint_degree)
    SB_THROW(SB_EINVAL, "stark %u: constraint_degree %u != program degree %u", p->stark_id, p->constraint_degree, a->degree);
  return a;
}

extern "C" int sb_air_load(sb_ctx* ctx, uint32_t stark_id, const char* path) {
  if (!ctx || !path) return SB_EINVAL;
  try {
    CUDA_CHECK(cudaSetDevice(ctx->device));
    air_load_file(ctx, stark_id, path);
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(ctx, e); }
}

// ---------------------------------------------------------------------------------------------------------
// per-proof weights: pw[j][k] = alpha_j^(K-1-k);  wt[slot][j] = sum_{k in slot} pw[j][k]
// ---------------------------------------------------------------------------------------------------------
__global__ void alpha_pow_kernel(u64* pw, uint32_t K, u64 a0, u64 a1, u64 a2, u64 a3, uint32_t nj) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  u64 al[4] = {a0, a1, a2, a3};
  for (uint32_t j = 0; j < nj; j++) pw[(size_t)j * K + k] = gl_pow(al[j], (u64)(K - 1 - k));
}
__global__ void slot_weight_kernel(u64* wt, const uint32_t* __restrict__ slot_off, const uint32_t* __restrict__ slot_ks,
                                   const u64* __restrict__ pw, uint32_t n_slots, uint32_t K, uint32_t nj) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  for (uint32_t j = 0; j < nj; j++) {
    u64 acc = 0;
    for (uint32_t i = slot_off[s]; i < slot_off[s + 1]; i++) acc = gl_add(acc, pw[(size_t)j * K + slot_ks[i]]);
    wt[(size_t)s * nj + j] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------
// per-shape domain tables, indexed by LDE position: z_last, L_first, L_last, 1/Z_H   
// ---------------------------------------------------------------------------------------------------------
__global__ void domain_kernel(u64* dom, unsigned log_n, unsigned rate_bits, u64 w_N, u64 g, u64 g_inv, u64 seven_n,
                              u64 w_r, u64 n_inv) {
  const uint32_t N = 1u << (log_n + rate_bits), n = 1u << log_n;
  uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= N) return;
  uint32_t J = pos >> log_n, k = pos & (n - 1);
  uint32_t j = bitrev32(J, rate_bits);
  u64 x = gl_mul(7, gl_pow(w_N, (u64)j + ((u64)k << rate_bits)));
  u64 zh = gl_sub(gl_mul(seven_n, gl_pow(w_r, j)), 1);          // x^n - 1 = 7^n * w_{2^r}^j - 1
  u64 t = gl_mul(zh, n_inv);
  dom[pos] = gl_sub(x, g_inv);                                   // z_last
  dom[N + pos] = gl_mul(t, gl_inv(gl_sub(x, 1)));                // L_first = (x^n-1)/(n(x-1))
  dom[2 * (size_t)N + pos] = gl_mul(t, gl_inv(gl_sub(gl_mul(g, x), 1)));  // L_last = (x^n-1)/(n(g x-1))
  dom[3 * (size_t)N + pos] = gl_inv(zh);
}

static const u64* domain_tables(sb_ctx* ctx, unsigned log_n, unsigned rate_bits) {
  uint64_t key = 0x100000000ull | ((uint64_t)log_n << 8) | rate_bits;
  auto it = ctx->coset_scale.find(key);
  if (it != ctx->coset_scale.end()) return it->second.as<u64>();
  DevBuf& b = ctx->coset_scale[key];
  uint32_t N = 1u << (log_n + rate_bits);
  b.ensure(32ull * N);
  u64 g = gl_root(log_n);
  LAUNCH(ctx, domain_kernel, (N + 127) / 128, 128, 0, b.as<u64>(), log_n, rate_bits, gl_root(log_n + rate_bits), g,
         gl_inv(g), gl_pow(7, 1ull << log_n), gl_root(rate_bits), gl_inv(1ull << log_n));
  return b.as<u64>();
}

// ---------------------------------------------------------------------------------------------------------
// the interpreter
// ---------------------------------------------------------------------------------------------------------
struct Acc192 { u64 a, b, c; };
__device__ __forceinline__ void mac192(Acc192& g, u64 x, u64 y) {
  u64 lo = x * y, hi = __umul64hi(x, y);
  asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u64 %2, %2, 0;"
      : "+l"(g.a), "+l"(g.b), "+l"(g.c) : "l"(lo), "l"(hi));
}
// a + b*2^64 + c*2^128 (c < 2^32)  ->  canonical.  2^128 = -2^32 (mod p)
__device__ __forceinline__ u64 reduce192(const Acc192& g) {
  u64 r = gl_canon(gl_reduce128_lazy(g.a, g.b));
  return gl_sub(r, gl_mul_2_32(g.c));
}

// One launch serves both the single-GPU layout (rows = the whole LDE, stride = N, pos0 = 0, halo = NULL) and a row
// block of the sharded layout (SURVEY 8e phase 2: rows = [C][n_local] holding LDE positions pos0 .. pos0 + n_local - 1;
// the "next" row of a position whose successor lives on another rank is read from `halo`, C contiguous values).
template <int D>
__global__ void __launch_bounds__(128) quotient_vm_kernel(
    const u64* __restrict__ lde, size_t stride, uint32_t n_local, uint32_t pos0, const u64* __restrict__ halo,
    uint32_t N, unsigned log_n, uint32_t C, const u64* __restrict__ pis,
    const u64* __restrict__ code, const u64* __restrict__ consts, const ulonglong2* __restrict__ wt,
    const uint4* __restrict__ chunks, const u64* __restrict__ dom, u64* __restrict__ part) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_local) return;
  const uint32_t pos = pos0 + idx;
  const uint32_t n = 1u << log_n;
  const uint32_t pos_next = (pos & ~(n - 1)) | ((pos + 1) & (n - 1));   // next row = same coset, k+1 (wraps)
  const u64* Lb = lde + idx;
  const bool next_local = pos_next - pos0 < n_local;                     // unsigned: also false when pos_next < pos0
  const u64* Nb = next_local ? lde + (pos_next - pos0) : halo;
  const size_t nstride = next_local ? stride : 1;
  const uint4 ch = chunks[blockIdx.y];
  uint32_t pc = ch.x, slot = ch.z;
  const uint32_t pc_end = ch.y;

  auto var = [&](uint32_t v) -> u64 {
    if (v < C) return Lb[(size_t)v * stride];
    if (v < 2 * C) return Nb[(size_t)(v - C) * nstride];
    return __ldg(pis + (v - 2 * C));
  };
  auto class_factor = [&](uint32_t cls) -> u64 { return cls == 1 ? 1 : dom[(size_t)(cls - 2) * N + pos]; };

  u64 acc0 = 0, acc1 = 0, S = 1, T = 0;
  Acc192 g0 = {0, 0, 0}, g1 = {0, 0, 0};
  uint32_t sel_left = 0, cls = 1;
  bool have_group = false;
  const char* Lb8 = (const char*)Lb;
  const uint32_t stride8 = (uint32_t)(stride * 8);   // < 2^32: checked by the host
  auto end_poly = [&]() {
    if (sel_left) {
      S = gl_mul(S, T);
      if (--sel_left == 0) S = gl_mul(S, class_factor(cls));
    } else {
      const ulonglong2 ww = __ldg(wt + slot);
      mac192(g0, T, ww.x);
      mac192(g1, T, ww.y);
      slot++;
    }
    T = 0;
  };
  // Software pipeline.  The interpreter is a serial chain (load operand -> add -> next word) and ncu shows it waiting on
  // memory (long-scoreboard stalls 4.5 per issue, L1 hit rate 36 %): the program is static, so the operands of fast
  // words are requested D words ahead and the code words D + 1 ahead.
  u64 wq[D + 2], pa[D + 1], pb[D + 1];
  auto issue = [&](u64 ww, u64& a, u64& b) {
    const uint32_t l = (uint32_t)ww, h = (uint32_t)(ww >> 32);
    if (l & 0x80u) {
      a = *(const u64*)(Lb8 + (size_t)(l >> 8) * stride8);
      b = *(const u64*)(Lb8 + (size_t)(h & 0xFFFFFFu) * stride8);    // column 0 for one-operand words: harmless
    }
  };
#pragma unroll
  for (int i = 0; i < D + 2; i++) wq[i] = __ldg(code + pc + i);
#pragma unroll
  for (int i = 0; i < D + 1; i++) { pa[i] = 0; pb[i] = 0; issue(wq[i], pa[i], pb[i]); }
  while (pc < pc_end) {
    const u64 w = wq[0], wn = wq[1];
    const u64 wfar = __ldg(code + pc + D + 2);
    const uint32_t wl = (uint32_t)w;
    const u64 a0 = pa[0], b0 = pb[0];
#pragma unroll
    for (int i = 0; i < D + 1; i++) wq[i] = wq[i + 1];
    wq[D + 1] = wfar;
#pragma unroll
    for (int i = 0; i < D; i++) { pa[i] = pa[i + 1]; pb[i] = pb[i + 1]; }
    issue(wq[D], pa[D], pb[D]);
    if (wl & 0x80u) {
      // fast word (see air_load_file): local columns only, operands already in flight
      const unsigned fop = wl & 15u;
      u64 a = a0;
      if (fop == OP_ADD2) {
        T = (wl & 32u) ? gl_sub(T, a) : gl_add(T, a);
        T = (wl & 64u) ? gl_sub(T, b0) : gl_add(T, b0);
      } else {
        if (fop == OP_MUL2) a = gl_mul(a, b0);
        else if (fop == OP_SHL1) a = gl_mul_2_32(a);
        T = (wl & 32u) ? gl_sub(T, a) : gl_add(T, a);
      }
      if (wl & 16u) end_poly();
      pc += 1;
      continue;
    }
    const unsigned op = (unsigned)w & 15u;
    const bool neg = (w >> 5) & 1;
    const uint32_t v0 = (uint32_t)(w >> 8) & 0x3FFFF, v1 = (uint32_t)(w >> 26) & 0x3FFFF, v2 = (uint32_t)(w >> 44) & 0x3FFFF;
    u64 x = 0;
    bool has_x = true;
    switch (op) {
      case OP_ADD1: x = var(v0); break;
      case OP_ADD2: {
        u64 a = var(v0), b = var(v1);
        T = neg ? gl_sub(T, a) : gl_add(T, a);
        T = ((w >> 6) & 1) ? gl_sub(T, b) : gl_add(T, b);
        has_x = false;
        break;
      }
      case OP_SHL1: x = gl_mul_2_32(var(v0)); break;
      case OP_MULS: x = gl_mul(var(v0), (w >> 26) & 0xFFFFFFFFull); break;
      case OP_MUL2: x = gl_mul(var(v0), var(v1)); break;
      case OP_MULC1: x = gl_mul(var(v0), __ldg(consts + v2)); break;
      case OP_MULC2: x = gl_mul(gl_mul(var(v0), var(v1)), __ldg(consts + v2)); break;
      case OP_MUL3C: x = gl_mul(gl_mul(gl_mul(var(v0), var(v1)), var(v2)), __ldg(consts + (wn >> 8))); break;   // next word: NOP | idx
      case OP_CONSTI: x = (w >> 26) & 0xFFFFFFFFull; break;
      case OP_CONSTC: x = __ldg(consts + v2); break;
      case OP_GROUP: {
        if (have_group) {
          acc0 = gl_add(acc0, gl_mul(S, reduce192(g0)));
          acc1 = gl_add(acc1, gl_mul(S, reduce192(g1)));
        }
        have_group = true;
        g0 = {0, 0, 0}; g1 = {0, 0, 0};
        cls = v0; sel_left = v1; T = 0;
        S = sel_left ? 1 : class_factor(cls);
        has_x = false;
        break;
      }
      default: has_x = false; break;
    }
    if (has_x) T = neg ? gl_sub(T, x) : gl_add(T, x);
    if ((w >> 4) & 1) end_poly();     // end of a polynomial
    pc += 1;                           // (the constant slot after a MUL3C is a NOP word)
  }
  if (have_group) {
    acc0 = gl_add(acc0, gl_mul(S, reduce192(g0)));
    acc1 = gl_add(acc1, gl_mul(S, reduce192(g1)));
  }
  part[((size_t)blockIdx.y * 2) * n_local + idx] = acc0;
  part[((size_t)blockIdx.y * 2 + 1) * n_local + idx] = acc1;
}

// out[j][idx] = (sum over chunks) * Z_H(x_pos)^-1,  pos = pos0 + idx
__global__ void quotient_reduce_kernel(const u64* __restrict__ part, uint32_t n_chunks, uint32_t n_local, uint32_t pos0,
                                       const u64* __restrict__ zh_inv, u64* __restrict__ out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * n_local) return;
  uint32_t j = t / n_local, idx = t % n_local;
  u64 acc = 0;
  for (uint32_t c = 0; c < n_chunks; c++) acc = gl_add(acc, part[((size_t)c * 2 + j) * n_local + idx]);
  out[t] = gl_mul(acc, zh_inv[pos0 + idx]);
}

static void build_chunks(sb_ctx* ctx, AirProgram* a, uint32_t want) {
  if (want > a->n_groups) want = a->n_groups;
  if (want < 1) want = 1;
  if (a->n_chunks == want) return;
  std::vector<uint4> tab;
  uint32_t g = 0;
  for (uint32_t c = 0; c < want && g < a->n_groups; c++) {
    // cut the remaining code evenly over the remaining chunks, at group boundaries
    uint32_t pc0 = a->group_pc[g];
    uint64_t target = pc0 + (uint64_t)(a->n_code - pc0) / (want - c);
    uint32_t g1 = g + 1;
    while (g1 < a->n_groups && a->group_pc[g1] < target) g1++;
    if (c + 1 == want) g1 = a->n_groups;
    tab.push_back(make_uint4(pc0, a->group_pc[g1], a->group_slot[g], 0));
    g = g1;
  }
  a->n_chunks = (uint32_t)tab.size();
  a->chunks.ensure(sizeof(uint4) * tab.size());
  CUDA_CHECK(cudaMemcpyAsync(a->chunks.p, tab.data(), sizeof(uint4) * tab.size(), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // tab is a stack vector
}

// q_j(x), j < 2, at the n_local LDE positions pos0 .. pos0 + n_local - 1 held in d_rows ([C][n_local], row stride
// `stride`), into d_out[j][idx].  d_halo: the row that follows the block's last position when that row is not in the
// block (NULL for the whole domain).
void sb_quotient_rows(sb_ctx* ctx, const sb_params* p, const u64* d_rows, size_t stride, uint32_t n_local, uint32_t pos0,
                      const u64* d_halo, const u64* d_pis, const u64* alphas, u64* d_out) {
  if (p->num_challenges != 2) SB_THROW(SB_EINVAL, "num_challenges must be 2 (StarkConfig::standard_fast_config)");
  unsigned qdf = quotient_degree_factor(*p);
  if (ilog2(qdf) != p->rate_bits)
    SB_THROW(SB_EINVAL, "quotient_degree_bits %u != rate_bits %u: unsupported (all five starks have them equal)", ilog2(qdf), p->rate_bits);
  AirProgram* a = air_get(ctx, p);
  const uint32_t N = 1u << (p->log_n + p->rate_bits);
  if ((uint64_t)stride * 8 >= (1ull << 32)) SB_THROW(SB_EINVAL, "row stride %zu too large", stride);
  const unsigned block = n_local < 128 ? n_local : 128;
  const uint32_t xtiles = (n_local + block - 1) / block;
  static const int chunk_mul = [] { const char* e = getenv("SB_QUOTIENT_CHUNK_MUL"); return e ? atoi(e) : 128; }();   // 16: 3.84 / 92.8, 64: 3.39 / 81.4, 128: 3.36 / 79.7, 256: 3.43 / 78.9 ms (PairingPrecomp / FinalExp)
  build_chunks(ctx, a, (uint32_t)((ctx->sm_count * chunk_mul + xtiles - 1) / xtiles));
  a->pw.ensure(16ull * a->K);
  a->wt.ensure(16ull * a->n_slots + 16);
  a->part.ensure(16ull * a->n_chunks * n_local);
  LAUNCH(ctx, alpha_pow_kernel, (a->K + 255) / 256, 256, 0, a->pw.as<u64>(), a->K, alphas[0], alphas[1], 0ull, 0ull, 2u);
  LAUNCH(ctx, slot_weight_kernel, (a->n_slots + 255) / 256, 256, 0, a->wt.as<u64>(), a->slot_off.as<uint32_t>(),
         a->slot_ks.as<uint32_t>(), a->pw.as<u64>(), a->n_slots, a->K, 2u);
  const u64* dom = domain_tables(ctx, p->log_n, p->rate_bits);
  dim3 grid(xtiles, a->n_chunks);
  static const int depth = [] { const char* e = getenv("SB_QUOTIENT_DEPTH"); return e ? atoi(e) : 0; }();
#define QVM_LAUNCH(DD)                                                                                                    \
  LAUNCH(ctx, quotient_vm_kernel<DD>, grid, block, 0, d_rows, stride, n_local, pos0, d_halo, N, p->log_n, p->n_cols, d_pis, \
         a->code.as<u64>(), a->consts.as<u64>(), a->wt.as<ulonglong2>(), a->chunks.as<uint4>(), dom, a->part.as<u64>())
  if (depth <= 0) QVM_LAUNCH(0); else if (depth == 1) QVM_LAUNCH(1); else if (depth == 2) QVM_LAUNCH(2); else if (depth == 3) QVM_LAUNCH(3); else QVM_LAUNCH(5);
#undef QVM_LAUNCH
  LAUNCH(ctx, quotient_reduce_kernel, (2 * n_local + 255) / 256, 256, 0, a->part.as<u64>(), a->n_chunks, n_local, pos0,
         dom + 3ull * N, d_out);
}

// q_j(x) for every LDE position of the committed trace (ctx->lde), j < 2, into d_out[j][pos].
void sb_quotient_device(sb_ctx* ctx, const sb_params* p, const u64* d_pis, const u64* alphas, u64* d_out) {
  const uint32_t N = 1u << (p->log_n + p->rate_bits);
  sb_quotient_rows(ctx, p, ctx->lde.as<u64>(), N, N, 0, nullptr, d_pis, alphas, d_out);
}

// SURVEY 8e phase 2: the quotient values of ONE row block of the sharded LDE (device pointers; the caller exchanged the
// halo row).  Replaces this rank's share of starky::prover::compute_quotient_polys.
extern "C" int sb_quotient_rows_device(sb_ctx* ctx, const sb_params* p, const uint64_t* d_rows, uint32_t rows_per_block,
                                       uint32_t block_index, const uint64_t* d_halo_next_row, const uint64_t* public_inputs,
                                       const uint64_t* alphas, uint64_t* d_out) {
  if (!ctx || !p || !d_rows || !alphas || !d_out || !rows_per_block) return SB_EINVAL;
  try {
    CUDA_CHECK(cudaSetDevice(ctx->device));
    const uint32_t N = 1u << (p->log_n + p->rate_bits), n = 1u << p->log_n;
    if (rows_per_block > N || N % rows_per_block || (uint64_t)block_index * rows_per_block >= N)
      SB_THROW(SB_EINVAL, "row block %u x %u does not tile the %u LDE positions", block_index, rows_per_block, N);
    if (rows_per_block < n && !d_halo_next_row)
      SB_THROW(SB_EINVAL, "a row block shorter than one coset (%u < %u) needs the halo row of its successor", rows_per_block, n);
    ctx->pis.ensure(8ull * (p->n_public_inputs + 1));
    if (p->n_public_inputs) {
      if (!public_inputs) SB_THROW(SB_EINVAL, "public_inputs is NULL");
      CUDA_CHECK(cudaMemcpyAsync(ctx->pis.p, public_inputs, 8ull * p->n_public_inputs, cudaMemcpyHostToDevice, ctx->stream));
    }
    stage_begin(ctx, "quotient");
    sb_quotient_rows(ctx, p, d_rows, rows_per_block, rows_per_block, block_index * rows_per_block, d_halo_next_row,
                     ctx->pis.as<u64>(), alphas, d_out);
    stage_end(ctx, "quotient");
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx);
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(ctx, e); }
}

extern "C" int sb_quotient_values(sb_ctx* ctx, const sb_params* p, const uint64_t* public_inputs, const uint64_t* alphas,
                                  uint64_t* out) {
  if (!ctx || !p || !alphas || !out) return SB_EINVAL;
  try {
    CUDA_CHECK(cudaSetDevice(ctx->device));
    if (!ctx->have_lde || ctx->cur.n_cols != p->n_cols || ctx->cur.log_n != p->log_n || ctx->cur.rate_bits != p->rate_bits)
      SB_THROW(SB_EINVAL, "sb_quotient_values needs a preceding sb_lde_commit with the same shape on this ctx");
    const uint32_t N = 1u << (p->log_n + p->rate_bits), n = 1u << p->log_n;
    ctx->pis.ensure(8ull * (p->n_public_inputs + 1));
    if (p->n_public_inputs) {
      if (!public_inputs) SB_THROW(SB_EINVAL, "public_inputs is NULL");
      CUDA_CHECK(cudaMemcpyAsync(ctx->pis.p, public_inputs, 8ull * p->n_public_inputs, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->qvals.ensure(16ull * N);
    stage_begin(ctx, "quotient");
    sb_quotient_device(ctx, p, ctx->pis.as<u64>(), alphas, ctx->qvals.as<u64>());
    stage_end(ctx, "quotient");
    std::vector<u64> tmp(2ull * N);
    CUDA_CHECK(cudaMemcpyAsync(tmp.data(), ctx->qvals.p, 16ull * N, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx);
    // position J*n + k  ->  natural LDE index bitrev_r(J) + 2^r k
    for (uint32_t j = 0; j < 2; j++)
      for (uint32_t pos = 0; pos < N; pos++) {
        uint32_t J = pos >> p->log_n, k = pos & (n - 1);
        out[(size_t)j * N + bitrev32(J, p->rate_bits) + ((size_t)k << p->rate_bits)] = tmp[(size_t)j * N + pos];
      }
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(ctx, e); }
}
