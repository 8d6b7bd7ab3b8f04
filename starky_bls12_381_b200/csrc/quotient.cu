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

#include <algorithm>

#include "prover.cuh"

enum { AIR_CODE_PAD = 8 };
enum { OP_NOP = 0, OP_ADD1, OP_ADD2, OP_SHL1, OP_MULS, OP_MUL2, OP_MULC1, OP_MULC2, OP_MUL3C, OP_CONSTI, OP_CONSTC, OP_GROUP, OP_RUN };

struct AirProgram {
  uint32_t n_cols = 0, n_pis = 0, degree = 0, K = 0, n_code = 0, n_consts = 0, n_slots = 0, n_groups = 0;
  DevBuf code, consts, slot_off, slot_ks, pw, wt, chunks, part;
  std::vector<uint32_t> group_pc, group_slot;
  uint32_t n_chunks = 0;
  // run form (translate_runs): the same groups, bodies of the common shapes folded into OP_RUN records
  DevBuf code2, slot_off2, slot_ks2, zero;
  std::vector<uint32_t> group_pc2, group_slot2;
  uint32_t n_code2 = 0, n_runs = 0, n_run_bodies = 0, n_bodies = 0;
  bool chunks_are_runs = false;
};

void air_release_all(sb_ctx* ctx) {
  for (auto& kv : ctx->airs) {
    AirProgram* a = kv.second;
    DevBuf* bufs[] = {&a->code, &a->consts, &a->slot_off, &a->slot_ks, &a->pw, &a->wt, &a->chunks, &a->part, &a->code2, &a->slot_off2, &a->slot_ks2, &a->zero};
    for (DevBuf* b : bufs) b->release();
    delete a;
  }
  ctx->airs.clear();
}

// ---------------------------------------------------------------------------------------------------------
// Run form.  ncu on the word-form interpreter (FinalExp, profiles/r2_quotient_fe.txt): ~100 executed instructions per
// bytecode word of which ~9 are arithmetic, and 95 GB of DRAM reads for a 19.3 GB LDE -- airgen emits all constraints of
// one (class, selector) key together, so the columns of one gadget instance are visited in three or four far-apart
// places of the program and every program chunk re-reads them.  The evaluation order is free (the quotient is a sum), so
// the loader rebuilds the program:
//   * groups are ordered by the median column they touch: the groups of one gadget instance become neighbours and a
//     program chunk reads a compact set of columns (ideal traffic at 74 chunks: 68 GB -> 30 GB);
//   * inside a group the bodies are sorted by shape and operands; the reference's gadgets are loops over limbs
//     (fp.rs:443-1553 ...), so bodies of one shape then have column indices in arithmetic progression and fold into
//         OP_RUN | kind | signs | count | imm32     + three words of six (column:20, delta:12) operands
//     records that the kernel decodes once and executes `count` times in one of four small loops (pointer += delta):
//         DIFF  T = v0 - v1                                   (half of all bodies: copy and carry-chain constraints)
//         ONE   T = +-v0 +- imm
//         LIN5  T = +-2^32 v0 +-v1 +-v2 +-v3 +-v4 +- imm      (absent operands read a zero cell)
//         MUL   T = +-v0 v1 +-2^32 v2 +-v3 +-v4 +- imm
//     (>= 99.7 % of the bodies of all five starks); selector factors and the rare other bodies stay in word form;
//   * weight slots are renumbered in the new body order (slot tables permuted on the host).
// ---------------------------------------------------------------------------------------------------------
enum { RUN_DIFF = 0, RUN_ONE = 1, RUN_LIN5 = 2, RUN_MUL = 3, RUN_ZERO_VAR = 0xFFFFF, RUN_WORDS = 4 };
struct RunBody {
  bool ok = false;
  unsigned kind = 0, signs = 0;
  uint32_t imm = 0;
  uint32_t ops[6] = {RUN_ZERO_VAR, RUN_ZERO_VAR, RUN_ZERO_VAR, RUN_ZERO_VAR, RUN_ZERO_VAR, RUN_ZERO_VAR};
  uint32_t slot = 0;                    // weight slot of this body in the word form
  uint32_t pc0 = 0, pc1 = 0;            // its words in the word form
};
// one body = code[pc0, pc1) in word form
static RunBody classify_body(const std::vector<u64>& code, uint32_t pc0, uint32_t pc1) {
  RunBody b;
  b.pc0 = pc0; b.pc1 = pc1;
  unsigned nshl = 0, nmul = 0, nlin = 0, ncst = 0;
  uint32_t shl_v = 0, mul_a = 0, mul_b = 0, lin_v[4] = {0, 0, 0, 0}, imm = 0;
  unsigned shl_s = 0, mul_s = 0, lin_s[4] = {0, 0, 0, 0}, cst_s = 0;
  for (uint32_t pc = pc0; pc < pc1; pc++) {
    const u64 w = code[pc];
    const unsigned op = w & 15, n0 = (w >> 5) & 1, n1 = (w >> 6) & 1;
    const uint32_t v0 = (w >> 8) & 0x3FFFF, v1 = (w >> 26) & 0x3FFFF;
    switch (op) {
      case OP_NOP: break;
      case OP_ADD1: if (nlin >= 4) return b; lin_v[nlin] = v0; lin_s[nlin++] = n0; break;
      case OP_ADD2:
        if (nlin >= 3) return b;
        lin_v[nlin] = v0; lin_s[nlin++] = n0; lin_v[nlin] = v1; lin_s[nlin++] = n1; break;
      case OP_SHL1: if (nshl++) return b; shl_v = v0; shl_s = n0; break;
      case OP_MUL2: if (nmul++) return b; mul_a = v0; mul_b = v1; mul_s = n0; break;
      case OP_CONSTI: if (ncst++) return b; imm = (uint32_t)((w >> 26) & 0xFFFFFFFFu); cst_s = n0; break;
      default: return b;
    }
  }
  if (nmul && nlin > 2) return b;
  if (!nmul && !nshl && !nlin) return b;
  // canonical operand order inside a body: linear terms by (sign, column) so that equal shapes compare equal
  for (unsigned i = 0; i < nlin; i++)
    for (unsigned j = i + 1; j < nlin; j++)
      if (lin_s[j] < lin_s[i] || (lin_s[j] == lin_s[i] && lin_v[j] < lin_v[i])) { std::swap(lin_s[i], lin_s[j]); std::swap(lin_v[i], lin_v[j]); }
  b.ok = true;
  b.imm = ncst ? imm : 0;
  const unsigned imm_sign = ncst ? cst_s << 7 : 0;
  if (!nmul && !nshl && !ncst && nlin == 2 && lin_s[0] != lin_s[1]) {          // sorted: lin_s = (0, 1) -> v0 - v1
    b.kind = RUN_DIFF; b.ops[0] = lin_v[0]; b.ops[1] = lin_v[1];
  } else if (!nmul && !nshl && nlin == 1) {
    b.kind = RUN_ONE; b.ops[0] = lin_v[0]; b.signs = lin_s[0] | imm_sign;
  } else if (!nmul) {
    b.kind = RUN_LIN5;
    if (nshl) { b.ops[0] = shl_v; b.signs |= shl_s; }
    for (unsigned i = 0; i < nlin; i++) { b.ops[1 + i] = lin_v[i]; b.signs |= lin_s[i] << (1 + i); }
    b.signs |= imm_sign;
  } else {
    b.kind = RUN_MUL;
    b.ops[0] = mul_a; b.ops[1] = mul_b; b.signs |= mul_s;
    if (nshl) { b.ops[2] = shl_v; b.signs |= shl_s << 1; }
    for (unsigned i = 0; i < nlin; i++) { b.ops[3 + i] = lin_v[i]; b.signs |= lin_s[i] << (2 + i); }
    b.signs |= imm_sign;
  }
  return b;
}
static void emit_run(std::vector<u64>& out, const RunBody& first, const int* delta, uint32_t count) {
  out.push_back((u64)OP_RUN | (u64)first.kind << 4 | (u64)first.signs << 8 | (u64)count << 16 | (u64)first.imm << 32);
  for (unsigned k = 0; k < 6; k += 2) {
    u64 w = 0;
    for (unsigned h = 0; h < 2; h++) w |= ((u64)first.ops[k + h] | ((u64)((uint32_t)delta[k + h] & 0xFFFu) << 20)) << (32 * h);
    out.push_back(w);
  }
}
struct RunGroup {
  u64 header = 0;
  std::vector<u64> sel_words;
  std::vector<RunBody> bodies;
  double key = 0;
  uint32_t first_slot = 0;
};
// word form (validated) -> run form.  gpc2[g] / gslot2[g] = first word / first weight slot of group g of the result;
// slot_perm[new slot] = slot of the same body in the word form.
static void translate_runs(const std::vector<u64>& code, const std::vector<uint32_t>& gpc, const std::vector<uint32_t>& gslot,
                           uint32_t n_groups, uint32_t n_cols, AirProgram* a, std::vector<u64>& out, std::vector<uint32_t>& gpc2,
                           std::vector<uint32_t>& gslot2, std::vector<uint32_t>& slot_perm) {
  // SB_RUN_KINDS: bit k set = fold bodies of kind k into runs (debugging aid; default all four)
  const char* km = getenv("SB_RUN_KINDS");
  const unsigned run_kind_mask = km ? (unsigned)atoi(km) : 15u;
  std::vector<RunGroup> groups(n_groups);
  for (uint32_t g = 0; g < n_groups; g++) {
    RunGroup& G = groups[g];
    uint32_t pc = gpc[g], slot = gslot[g];
    const uint32_t end = gpc[g + 1];
    if (pc >= end || (code[pc] & 15) != OP_GROUP) SB_THROW(SB_EAIR, "constraint program: group %u does not start with a GROUP word", g);
    uint32_t sel_left = (uint32_t)((code[pc] >> 26) & 0x3FFFF);
    G.header = code[pc++];
    std::vector<uint32_t> cols;
    auto note_cols = [&](u64 w) {
      const unsigned op = w & 15;
      const uint32_t v[3] = {(uint32_t)(w >> 8) & 0x3FFFF, (uint32_t)(w >> 26) & 0x3FFFF, (uint32_t)(w >> 44) & 0x3FFFF};
      const int n = (op == OP_ADD1 || op == OP_SHL1 || op == OP_MULS || op == OP_MULC1) ? 1 : (op == OP_ADD2 || op == OP_MUL2 || op == OP_MULC2) ? 2 : op == OP_MUL3C ? 3 : 0;
      for (int i = 0; i < n; i++) if (v[i] < 2 * n_cols) cols.push_back(v[i] % n_cols);
    };
    while (sel_left && pc < end) {                       // selector factors: word form
      const u64 w = code[pc];
      G.sel_words.push_back(w); note_cols(w); pc++;
      if ((w & 15) == OP_MUL3C) { G.sel_words.push_back(code[pc]); pc++; continue; }
      if ((w >> 4) & 1) sel_left--;
    }
    while (pc < end) {
      uint32_t q = pc;
      for (;;) {                                         // find the end of this body
        if (q >= end) SB_THROW(SB_EAIR, "constraint program: unterminated polynomial in group %u", g);
        const u64 w = code[q];
        note_cols(w);
        if ((w & 15) == OP_MUL3C) { q += 2; continue; }
        q++;
        if ((w >> 4) & 1) break;
      }
      RunBody b = classify_body(code, pc, q);
      if (b.ok && !((run_kind_mask >> b.kind) & 1u)) b.ok = false;
      b.slot = slot++;
      G.bodies.push_back(b);
      a->n_bodies++;
      pc = q;
    }
    if (!cols.empty()) {
      std::nth_element(cols.begin(), cols.begin() + cols.size() / 2, cols.end());
      G.key = cols[cols.size() / 2];
    }
  }
  std::stable_sort(groups.begin(), groups.end(), [](const RunGroup& x, const RunGroup& y) { return x.key < y.key; });
  out.clear(); gpc2.clear(); gslot2.clear(); slot_perm.clear();
  for (RunGroup& G : groups) {
    std::stable_sort(G.bodies.begin(), G.bodies.end(), [](const RunBody& x, const RunBody& y) {
      if (x.ok != y.ok) return x.ok > y.ok;
      if (!x.ok) return false;
      if (x.kind != y.kind) return x.kind < y.kind;
      if (x.signs != y.signs) return x.signs < y.signs;
      if (x.imm != y.imm) return x.imm < y.imm;
      return memcmp(x.ops, y.ops, sizeof(x.ops)) < 0;
    });
    gpc2.push_back((uint32_t)out.size());
    gslot2.push_back((uint32_t)slot_perm.size());
    out.push_back(G.header);
    out.insert(out.end(), G.sel_words.begin(), G.sel_words.end());
    size_t i = 0;
    while (i < G.bodies.size()) {
      const RunBody& b = G.bodies[i];
      if (!b.ok) {
        for (uint32_t k = b.pc0; k < b.pc1; k++) out.push_back(code[k]);
        slot_perm.push_back(b.slot);
        i++;
        continue;
      }
      int delta[6] = {0, 0, 0, 0, 0, 0};
      size_t j = i + 1;
      bool have_delta = false;
      while (j < G.bodies.size() && j - i < 4095) {
        const RunBody& c = G.bodies[j];
        if (!c.ok || c.kind != b.kind || c.signs != b.signs || c.imm != b.imm) break;
        int d[6]; bool fits = true;
        for (int k = 0; k < 6; k++) {
          const uint32_t u = G.bodies[j - 1].ops[k], v = c.ops[k];
          if ((u == RUN_ZERO_VAR) != (v == RUN_ZERO_VAR)) { fits = false; break; }
          d[k] = u == RUN_ZERO_VAR ? 0 : (int)v - (int)u;
          // a progression may not walk from one variable space (local / next / public input) into another
          if (u != RUN_ZERO_VAR && ((u < n_cols) != (v < n_cols) || (u < 2 * n_cols) != (v < 2 * n_cols))) fits = false;
          fits = fits && d[k] >= -2048 && d[k] <= 2047;
        }
        if (!fits || (have_delta && memcmp(d, delta, sizeof(d)) != 0)) break;
        memcpy(delta, d, sizeof(d)); have_delta = true;
        j++;
      }
      emit_run(out, b, delta, (uint32_t)(j - i));
      a->n_runs++; a->n_run_bodies += (uint32_t)(j - i);
      for (size_t k = i; k < j; k++) slot_perm.push_back(G.bodies[k].slot);
      i = j;
    }
  }
  gpc2.push_back((uint32_t)out.size());
  gslot2.push_back((uint32_t)slot_perm.size());
}

// One SBAIRBN1 image (tools/airgen/compile.py: write_airbin) parsed, validated and translated on the host.
struct HostAir {
  uint32_t n_cols = 0, n_pis = 0, degree = 0, K = 0, n_code = 0, n_consts = 0, n_slots = 0, n_groups = 0;
  std::vector<u64> code, consts;                       // word form (code has one spare word, consts one spare slot)
  std::vector<uint32_t> slot_off, slot_ks, gpc, gslot;
  std::vector<u64> code2;                              // run form
  std::vector<uint32_t> gpc2, gslot2, slot_off2, slot_ks2, slot_perm;
  uint32_t n_runs = 0, n_run_bodies = 0, n_bodies = 0;
};
static void air_parse(const unsigned char* img, size_t img_len, const char* path, HostAir& H) {
  struct Hdr { char magic[8]; uint32_t v[12]; } h;
  size_t off = 0;
  auto rd = [&](void* p, size_t sz, size_t n) {
    if (n == 0) return true;
    if (off + sz * n > img_len) return false;
    memcpy(p, img + off, sz * n);
    off += sz * n;
    return true;
  };
  bool ok = rd(&h, sizeof(h), 1) && memcmp(h.magic, "SBAIRBN1", 8) == 0;
  if (ok) {
    H.n_cols = h.v[0]; H.n_pis = h.v[1]; H.degree = h.v[2]; H.K = h.v[3]; H.n_code = h.v[4]; H.n_consts = h.v[5];
    H.n_slots = h.v[6]; H.n_groups = h.v[7];
    ok = (uint64_t)H.n_code * 8 <= img_len && (uint64_t)H.n_consts * 8 <= img_len && (uint64_t)H.n_slots * 4 <= img_len &&
         (uint64_t)H.K * 4 <= img_len && (uint64_t)H.n_groups * 4 <= img_len && H.n_cols < (1u << 18);
  }
  if (ok) {
    H.code.resize(H.n_code + 1); H.consts.resize(H.n_consts + 1); H.slot_off.resize(H.n_slots + 1); H.slot_ks.resize(H.K);
    H.gpc.resize(H.n_groups + 1); H.gslot.resize(H.n_groups + 1);
    ok = rd(H.code.data(), 8, H.n_code) && rd(H.consts.data(), 8, H.n_consts) && rd(H.slot_off.data(), 4, H.n_slots + 1) &&
         rd(H.slot_ks.data(), 4, H.K) && rd(H.gpc.data(), 4, H.n_groups + 1) && rd(H.gslot.data(), 4, H.n_groups + 1);
  }
  if (!ok) SB_THROW(SB_EAIR, "malformed constraint program %s", path);
  // validate: every variable index in range, group / slot tables consistent
  const uint32_t n_vars = 2 * H.n_cols + H.n_pis;
  for (uint32_t pc = 0; pc < H.n_code; pc++) {
    u64 w = H.code[pc];
    unsigned op = w & 15;
    uint32_t v0 = (w >> 8) & 0x3FFFF, v1 = (w >> 26) & 0x3FFFF, v2 = (w >> 44) & 0x3FFFF;
    bool bad = false;
    switch (op) {
      case OP_ADD1: case OP_SHL1: case OP_MULS: bad = v0 >= n_vars; break;
      case OP_ADD2: case OP_MUL2: bad = v0 >= n_vars || v1 >= n_vars; break;
      case OP_MULC1: bad = v0 >= n_vars || v2 >= H.n_consts; break;
      case OP_MULC2: bad = v0 >= n_vars || v1 >= n_vars || v2 >= H.n_consts; break;
      case OP_MUL3C: bad = v0 >= n_vars || v1 >= n_vars || v2 >= n_vars || pc + 1 >= H.n_code; pc++; break;
      case OP_CONSTC: bad = v2 >= H.n_consts; break;
      case OP_GROUP: bad = v0 < 1 || v0 > 4; break;
      case OP_NOP: case OP_CONSTI: break;
      default: bad = true;
    }
    if (bad) SB_THROW(SB_EAIR, "constraint program %s: invalid instruction at pc %u", path, pc);
  }
  for (uint32_t g = 0; g < H.n_groups; g++)
    if (H.gpc[g] >= H.gpc[g + 1] || H.gpc[g + 1] > H.n_code || H.gslot[g] > H.gslot[g + 1] || H.gslot[g + 1] > H.n_slots)
      SB_THROW(SB_EAIR, "constraint program %s: inconsistent group table at group %u", path, g);
  for (uint32_t sl = 0; sl < H.n_slots; sl++)
    if (H.slot_off[sl] > H.slot_off[sl + 1] || H.slot_off[sl + 1] > H.K) SB_THROW(SB_EAIR, "constraint program %s: inconsistent slot table", path);
  // run form (from the validated word form)
  AirProgram counters;
  std::vector<u64> plain(H.code.begin(), H.code.begin() + H.n_code);
  translate_runs(plain, H.gpc, H.gslot, H.n_groups, H.n_cols, &counters, H.code2, H.gpc2, H.gslot2, H.slot_perm);
  H.n_runs = counters.n_runs; H.n_run_bodies = counters.n_run_bodies; H.n_bodies = counters.n_bodies;
  if (H.slot_perm.size() != H.n_slots)
    SB_THROW(SB_EAIR, "constraint program %s: %zu bodies for %u weight slots", path, H.slot_perm.size(), H.n_slots);
  H.slot_off2.push_back(0);
  for (uint32_t ns = 0; ns < H.n_slots; ns++) {
    for (uint32_t i = H.slot_off[H.slot_perm[ns]]; i < H.slot_off[H.slot_perm[ns] + 1]; i++) H.slot_ks2.push_back(H.slot_ks[i]);
    H.slot_off2.push_back((uint32_t)H.slot_ks2.size());
  }
}

// The run form of a constraint program, for inspection and for the CPU test that emulates it (tests/test_air_programs.py):
// sizes first (code2_out == NULL), then the words, the permuted slot tables and the group tables.  Host only.
extern "C" int sb_air_run_form(const void* image, size_t image_len, uint64_t* code2_out, size_t code2_cap, uint32_t* slot_off2_out,
                               uint32_t* slot_ks2_out, uint32_t* group_pc2_out, uint32_t* group_slot2_out, uint32_t* info_out /* [8] */) {
  if (!image || !info_out) return SB_EINVAL;
  try {
    HostAir H;
    air_parse((const unsigned char*)image, image_len, "(memory)", H);
    const uint32_t info[8] = {(uint32_t)H.code2.size(), H.n_slots, H.K, H.n_groups, H.n_runs, H.n_run_bodies, H.n_bodies, H.n_cols};
    memcpy(info_out, info, sizeof(info));
    if (!code2_out) return SB_OK;
    if (code2_cap < H.code2.size()) SB_THROW(SB_EINVAL, "buffer of %zu words for %zu", code2_cap, H.code2.size());
    memcpy(code2_out, H.code2.data(), 8 * H.code2.size());
    if (slot_off2_out) memcpy(slot_off2_out, H.slot_off2.data(), 4 * H.slot_off2.size());
    if (slot_ks2_out) memcpy(slot_ks2_out, H.slot_ks2.data(), 4 * H.slot_ks2.size());
    if (group_pc2_out) memcpy(group_pc2_out, H.gpc2.data(), 4 * H.gpc2.size());
    if (group_slot2_out) memcpy(group_slot2_out, H.gslot2.data(), 4 * H.gslot2.size());
    return SB_OK;
  } catch (const SbError& e) { return sb_fail(nullptr, e); }
}

// Binds one image to `stark_id` on this ctx.  `path` names the source in error messages (a path, or "embedded:<name>").
static void air_load_image(sb_ctx* ctx, uint32_t stark_id, const unsigned char* img, size_t img_len, const char* path) {
  HostAir H;
  air_parse(img, img_len, path, H);
  AirProgram* a = new AirProgram();
  a->n_cols = H.n_cols; a->n_pis = H.n_pis; a->degree = H.degree; a->K = H.K; a->n_code = H.n_code; a->n_consts = H.n_consts;
  a->n_slots = H.n_slots; a->n_groups = H.n_groups;
  a->n_runs = H.n_runs; a->n_run_bodies = H.n_run_bodies; a->n_bodies = H.n_bodies;
  a->group_pc2 = H.gpc2; a->group_slot2 = H.gslot2;
  std::vector<u64>& code = H.code; std::vector<u64>& consts = H.consts;
  std::vector<uint32_t>&slot_off = H.slot_off, &slot_ks = H.slot_ks, &gpc = H.gpc, &gslot = H.gslot, &slot_off2 = H.slot_off2, &slot_ks2 = H.slot_ks2;
  std::vector<u64>& code2 = H.code2;
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
  a->n_code2 = (uint32_t)code2.size();
  code2.resize(code2.size() + AIR_CODE_PAD, (u64)OP_NOP);
  a->code2.ensure(8ull * code2.size());
  a->slot_off2.ensure(4ull * (a->n_slots + 1));
  a->slot_ks2.ensure(4ull * (a->K + 1));
  a->zero.ensure(64);
  CUDA_CHECK(cudaMemsetAsync(a->zero.p, 0, 64, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(a->code2.p, code2.data(), 8ull * code2.size(), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(a->slot_off2.p, slot_off2.data(), 4ull * (a->n_slots + 1), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(a->slot_ks2.p, slot_ks2.data(), 4ull * a->K, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  auto it = ctx->airs.find(stark_id);
  if (it != ctx->airs.end()) {
    DevBuf* bufs[] = {&it->second->code, &it->second->consts, &it->second->slot_off, &it->second->slot_ks,
                      &it->second->pw, &it->second->wt, &it->second->chunks, &it->second->part, &it->second->code2, &it->second->slot_off2,
                      &it->second->slot_ks2, &it->second->zero};
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

// ---------------------------------------------------------------------------------------------------------
// the run-form evaluator (translate_runs above): one thread owns one LDE position, the record stream is warp-uniform.
// Code size matters: the first version specialised twelve shapes, 108 KB of SASS, and a third of the warp samples were
// instruction-fetch stalls (stall_no_inst 33 %); the four loops below are ~0.6 K instructions.
// ---------------------------------------------------------------------------------------------------------
#ifndef QRUN_MIN_BLOCKS
#define QRUN_MIN_BLOCKS 8
#endif
// 128-bit accumulator of a handful of signed terms, kept non-negative by starting from a multiple of p:
//   BIAS = 2^36 p = 2^100 - 2^68 + 2^36  >  the magnitude of any body (one 2^32-weighted term, one reduced product and a
//   few plain terms: < 2^98), so BIAS + T is a 128-bit non-negative integer congruent to T and folds like a product.
struct Acc128 { u64 lo, hi; };
// acc += neg ? -x : x for a 64-bit x (M = neg ? ~0 : 0; the +1 of each two's-complement negation is in the start value)
__device__ __forceinline__ void acc128_add(Acc128& a, u64 x, u64 M) {
  const u64 t = x ^ M;
  a.lo += t;
  a.hi += M + (a.lo < t);
}
// acc += neg ? -(x 2^32) : x 2^32
__device__ __forceinline__ void acc128_add_shl32(Acc128& a, u64 x, u64 M) {
  const u64 tl = (x << 32) ^ M, th = (x >> 32) ^ M;
  a.lo += tl;
  a.hi += th + (a.lo < tl);
}
// (hi:lo) mod p -> canonical (2^64 = eps, 2^96 = -1)
__device__ __forceinline__ u64 acc128_reduce(const Acc128& a) { return gl_canon(gl_reduce128_lazy(a.lo, a.hi)); }

struct RunEnv {
  const char* Lb8; const char* Nb8; const u64* pis; const char* zero;
  uint32_t stride8, nstride8, C;
};
// operand k of a record: pointer to this position's value and byte step per body
__device__ __forceinline__ void run_operand(const RunEnv& e, const u64 (&opw)[3], int k, const char*& ptr, long long& step) {
  const u64 w = opw[k >> 1];
  const u32 half = (k & 1) ? (u32)(w >> 32) : (u32)w;
  const uint32_t var = half & 0xFFFFFu;
  const int delta = (int)half >> 20;
  if (var < e.C) { ptr = e.Lb8 + (size_t)var * e.stride8; step = (long long)delta * e.stride8; }
  else if (var < 2 * e.C) { ptr = e.Nb8 + (size_t)(var - e.C) * e.nstride8; step = (long long)delta * e.nstride8; }
  else if (var == RUN_ZERO_VAR) { ptr = e.zero; step = 0; }
  else { ptr = (const char*)(e.pis + (var - 2 * e.C)); step = 8ll * delta; }
}

__global__ void __launch_bounds__(128, QRUN_MIN_BLOCKS) quotient_run_kernel(
    const u64* __restrict__ lde, size_t stride, uint32_t n_local, uint32_t pos0, const u64* __restrict__ halo,
    uint32_t N, unsigned log_n, uint32_t C, const u64* __restrict__ pis,
    const u64* __restrict__ code, const u64* __restrict__ consts, const ulonglong2* __restrict__ wt,
    const uint4* __restrict__ chunks, const u64* __restrict__ dom, const u64* __restrict__ zero_cell, u64* __restrict__ part) {
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
  RunEnv env;
  env.Lb8 = (const char*)Lb; env.Nb8 = (const char*)Nb; env.pis = pis; env.zero = (const char*)zero_cell;
  env.stride8 = (uint32_t)(stride * 8); env.nstride8 = (uint32_t)(nstride * 8); env.C = C;   // < 2^32: checked by the host

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
  // The record stream is static: the words of the NEXT record are requested while the current one executes (a run record is
  // RUN_WORDS = 4 words, everything else one word, MUL3C two; the code is padded by AIR_CODE_PAD words).
  auto record_len = [](u64 hw) -> uint32_t {
    const unsigned o = (unsigned)hw & 15u;
    return o == OP_RUN ? (uint32_t)RUN_WORDS : (o == OP_MUL3C ? 2u : 1u);
  };
  u64 w = __ldg(code + pc);
  u64 opw[3] = {__ldg(code + pc + 1), __ldg(code + pc + 2), __ldg(code + pc + 3)};
  while (pc < pc_end) {
    const uint32_t pc_next = pc + record_len(w);
    const u64 w_next = __ldg(code + pc_next);
    const u64 o_next[3] = {__ldg(code + pc_next + 1), __ldg(code + pc_next + 2), __ldg(code + pc_next + 3)};
    const unsigned op = (unsigned)w & 15u;
    if (op == OP_RUN) {
      const unsigned kind = (unsigned)(w >> 4) & 7u;
      const uint32_t signs = (uint32_t)(w >> 8) & 0xFFu, count = (uint32_t)(w >> 16) & 0xFFFu, imm = (uint32_t)(w >> 32);
      const ulonglong2* wts = wt + slot;
      if (kind == RUN_DIFF) {
        const char *pa, *pb; long long sa, sb;
        run_operand(env, opw, 0, pa, sa);
        run_operand(env, opw, 1, pb, sb);
        u64 a = *(const u64*)pa, b = *(const u64*)pb;
#pragma unroll 1
        for (uint32_t i = 0; i < count; i++) {
          const ulonglong2 ww = __ldg(wts + i);
          const u64 Tb = gl_sub(a, b);
          pa += sa; pb += sb;
          if (i + 1 < count) { a = *(const u64*)pa; b = *(const u64*)pb; }     // next body's operands under this body's multiplies
          mac192(g0, Tb, ww.x);
          mac192(g1, Tb, ww.y);
        }
      } else if (kind == RUN_ONE) {
        const char* pa; long long sa;
        run_operand(env, opw, 0, pa, sa);
        const u64 c = (signs & 0x80u) ? gl_neg((u64)imm) : (u64)imm;
        u64 a = *(const u64*)pa;
#pragma unroll 1
        for (uint32_t i = 0; i < count; i++) {
          const ulonglong2 ww = __ldg(wts + i);
          const u64 Tb = gl_add((signs & 1u) ? gl_neg(a) : a, c);
          pa += sa;
          if (i + 1 < count) a = *(const u64*)pa;
          mac192(g0, Tb, ww.x);
          mac192(g1, Tb, ww.y);
        }
      } else {
        // LIN5: +-2^32 v0 +-v1..v4 +- imm        MUL: +-v0 v1 +-2^32 v2 +-v3 +-v4 +- imm
        const bool is_mul = kind == RUN_MUL;
        const char* ptr[5]; long long step[5];
#pragma unroll
        for (int k = 0; k < 5; k++) run_operand(env, opw, k, ptr[k], step[k]);
        u64 m[5];
#pragma unroll
        for (int t = 0; t < 5; t++) m[t] = 0ull - (u64)((signs >> t) & 1u);
        // start value: BIAS + (number of negated terms) +- imm
        Acc128 start = {0x0000001000000000ull, 0x0000000FFFFFFFF0ull};
        const uint32_t n_terms = is_mul ? 4u : 5u;
        acc128_add(start, (u64)__popc(signs & ((1u << n_terms) - 1u)), 0ull);
        if (signs & 0x80u) { acc128_add(start, (u64)imm, ~0ull); acc128_add(start, 1ull, 0ull); }
        else acc128_add(start, (u64)imm, 0ull);
        u64 v[5];
#pragma unroll
        for (int k = 0; k < 5; k++) v[k] = *(const u64*)ptr[k];
#pragma unroll 1
        for (uint32_t i = 0; i < count; i++) {
          const ulonglong2 ww = __ldg(wts + i);
          Acc128 acc = start;
          if (is_mul) {
            acc128_add(acc, gl_mul_lazy(v[0], v[1]), m[0]);
            acc128_add_shl32(acc, v[2], m[1]);
            acc128_add(acc, v[3], m[2]);
            acc128_add(acc, v[4], m[3]);
          } else {
            acc128_add_shl32(acc, v[0], m[0]);
            acc128_add(acc, v[1], m[1]);
            acc128_add(acc, v[2], m[2]);
            acc128_add(acc, v[3], m[3]);
            acc128_add(acc, v[4], m[4]);
          }
#pragma unroll
          for (int q = 0; q < 5; q++) ptr[q] += step[q];
          if (i + 1 < count) {
#pragma unroll
            for (int q = 0; q < 5; q++) v[q] = *(const u64*)ptr[q];
          }
          const u64 Tb = acc128_reduce(acc);
          mac192(g0, Tb, ww.x);
          mac192(g1, Tb, ww.y);
        }
      }
      slot += count;
      pc = pc_next; w = w_next; opw[0] = o_next[0]; opw[1] = o_next[1]; opw[2] = o_next[2];
      continue;
    }
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
      case OP_MUL3C: x = gl_mul(gl_mul(gl_mul(var(v0), var(v1)), var(v2)), gl_canon(opw[0])); break;   // inline constant word
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
    if ((w >> 4) & 1) {                // end of a polynomial
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
    }
    pc = pc_next; w = w_next; opw[0] = o_next[0]; opw[1] = o_next[1]; opw[2] = o_next[2];
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

static void build_chunks(sb_ctx* ctx, AirProgram* a, uint32_t want, bool runs) {
  if (want > a->n_groups) want = a->n_groups;
  if (want < 1) want = 1;
  if (a->n_chunks == want && a->chunks_are_runs == runs) return;
  a->chunks_are_runs = runs;
  const std::vector<uint32_t>& gpc = runs ? a->group_pc2 : a->group_pc;
  const uint32_t n_code = runs ? a->n_code2 : a->n_code;
  std::vector<uint4> tab;
  uint32_t g = 0;
  for (uint32_t c = 0; c < want && g < a->n_groups; c++) {
    // cut the remaining code evenly over the remaining chunks, at group boundaries
    uint32_t pc0 = gpc[g];
    uint64_t target = pc0 + (uint64_t)(n_code - pc0) / (want - c);
    uint32_t g1 = g + 1;
    while (g1 < a->n_groups && gpc[g1] < target) g1++;
    if (c + 1 == want) g1 = a->n_groups;
    tab.push_back(make_uint4(pc0, gpc[g1], (runs ? a->group_slot2 : a->group_slot)[g], 0));
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
  // SB_QUOTIENT_VM=1: the word-form interpreter of round 1 (A/B and the tests); default: the run-form evaluator
  const char* vm_env = getenv("SB_QUOTIENT_VM");
  const bool use_vm = vm_env && atoi(vm_env) != 0;
  build_chunks(ctx, a, (uint32_t)((ctx->sm_count * chunk_mul + xtiles - 1) / xtiles), !use_vm);
  a->pw.ensure(16ull * a->K);
  a->wt.ensure(16ull * a->n_slots + 16);
  a->part.ensure(16ull * a->n_chunks * n_local);
  LAUNCH(ctx, alpha_pow_kernel, (a->K + 255) / 256, 256, 0, a->pw.as<u64>(), a->K, alphas[0], alphas[1], 0ull, 0ull, 2u);
  LAUNCH(ctx, slot_weight_kernel, (a->n_slots + 255) / 256, 256, 0, a->wt.as<u64>(), (use_vm ? a->slot_off : a->slot_off2).as<uint32_t>(),
         (use_vm ? a->slot_ks : a->slot_ks2).as<uint32_t>(), a->pw.as<u64>(), a->n_slots, a->K, 2u);
  const u64* dom = domain_tables(ctx, p->log_n, p->rate_bits);
  dim3 grid(xtiles, a->n_chunks);
  static const int depth = [] { const char* e = getenv("SB_QUOTIENT_DEPTH"); return e ? atoi(e) : 0; }();
#define QVM_LAUNCH(DD)                                                                                                    \
  LAUNCH(ctx, quotient_vm_kernel<DD>, grid, block, 0, d_rows, stride, n_local, pos0, d_halo, N, p->log_n, p->n_cols, d_pis, \
         a->code.as<u64>(), a->consts.as<u64>(), a->wt.as<ulonglong2>(), a->chunks.as<uint4>(), dom, a->part.as<u64>())
  if (!use_vm) {
    LAUNCH(ctx, quotient_run_kernel, grid, block, 0, d_rows, stride, n_local, pos0, d_halo, N, p->log_n, p->n_cols, d_pis,
           a->code2.as<u64>(), a->consts.as<u64>(), a->wt.as<ulonglong2>(), a->chunks.as<uint4>(), dom, a->zero.as<u64>(), a->part.as<u64>());
  } else if (depth <= 0) QVM_LAUNCH(0); else if (depth == 1) QVM_LAUNCH(1); else if (depth == 2) QVM_LAUNCH(2); else if (depth == 3) QVM_LAUNCH(3); else QVM_LAUNCH(5);
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
