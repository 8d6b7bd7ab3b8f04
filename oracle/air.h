// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// Straight evaluation of a stark's constraints in the reference's emission order, the way
// starky::constraint_consumer::ConstraintConsumer does:
//     constraint(c):            acc_j <- acc_j * alpha_j + c
//     constraint_transition(c): constraint(c * z_last)
//     constraint_first_row(c):  constraint(c * lagrange_basis_first)
//     constraint_last_row(c):   constraint(c * lagrange_basis_last)
// The constraints themselves are the expression DAG ("flat AIR", file format SBAIR001) that
// tools/airgen extracts from the reference's eval_packed_generic bodies
// (fp12_mul.rs:58, calc_pairing_precomp.rs:376, miller_loop.rs:644, final_exponentiate.rs:907,
// ecc_aggregate.rs:92 and the gadget functions they call).  This evaluator does no regrouping,
// no selector hoisting and no explicit alpha-power weights -- it is the plain Horner fold, so it
// checks the GPU's restructured evaluation independently.
#pragma once
#include "gl.h"
#include <stdio.h>
#include <string.h>
#include <string>

namespace orc {

enum { AIR_CONST = 0, AIR_LOCAL = 1, AIR_NEXT = 2, AIR_PI = 3, AIR_ADD = 4, AIR_SUB = 5, AIR_MUL = 6 };
enum { CLS_PLAIN = 1, CLS_TRANSITION = 2, CLS_FIRST = 3, CLS_LAST = 4 };

struct AirNode { uint32_t op, a, b; };
struct AirCons { uint32_t cls, node; };
struct AirHeader { char magic[8]; uint32_t n_cols, n_pis, degree, n_consts, n_nodes, n_constraints, reserved[2]; };

struct Air {
  AirHeader h;
  std::vector<u64> consts;
  std::vector<AirNode> nodes;
  std::vector<AirCons> cons;
  bool load(const char* path, std::string* err) {
    FILE* f = fopen(path, "rb");
    if (!f) { if (err) *err = std::string("cannot open ") + path; return false; }
    bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "SBAIR001", 8) == 0;
    if (ok) {
      consts.resize(h.n_consts); nodes.resize(h.n_nodes); cons.resize(h.n_constraints);
      ok = (h.n_consts == 0 || fread(consts.data(), 8, h.n_consts, f) == h.n_consts) &&
           (h.n_nodes == 0 || fread(nodes.data(), sizeof(AirNode), h.n_nodes, f) == h.n_nodes) &&
           (h.n_constraints == 0 || fread(cons.data(), sizeof(AirCons), h.n_constraints, f) == h.n_constraints);
    }
    fclose(f);
    if (!ok && err) *err = std::string("malformed AIR file ") + path;
    return ok;
  }

  // Base-field evaluation at one LDE point.  scratch must hold n_nodes values.
  void eval_base(const u64* local, const u64* next, const u64* pis, u64 z_last, u64 l_first, u64 l_last,
                 const u64* alphas, unsigned n_alphas, u64* acc, u64* scratch) const {
    for (size_t i = 0; i < nodes.size(); i++) {
      const AirNode& nd = nodes[i];
      switch (nd.op) {
        case AIR_CONST: scratch[i] = consts[nd.a]; break;
        case AIR_LOCAL: scratch[i] = local[nd.a]; break;
        case AIR_NEXT: scratch[i] = next[nd.a]; break;
        case AIR_PI: scratch[i] = pis[nd.a]; break;
        case AIR_ADD: scratch[i] = gl_add(scratch[nd.a], scratch[nd.b]); break;
        case AIR_SUB: scratch[i] = gl_sub(scratch[nd.a], scratch[nd.b]); break;
        default: scratch[i] = gl_mul(scratch[nd.a], scratch[nd.b]); break;
      }
    }
    for (unsigned j = 0; j < n_alphas; j++) acc[j] = 0;
    for (size_t k = 0; k < cons.size(); k++) {
      u64 c = scratch[cons[k].node];
      switch (cons[k].cls) {
        case CLS_TRANSITION: c = gl_mul(c, z_last); break;
        case CLS_FIRST: c = gl_mul(c, l_first); break;
        case CLS_LAST: c = gl_mul(c, l_last); break;
        default: break;
      }
      for (unsigned j = 0; j < n_alphas; j++) acc[j] = gl_add(gl_mul(acc[j], alphas[j]), c);
    }
  }

  // Extension-field evaluation at zeta (verifier).
  void eval_ext(const E2* local, const E2* next, const u64* pis, E2 z_last, E2 l_first, E2 l_last,
                const u64* alphas, unsigned n_alphas, E2* acc, E2* scratch) const {
    for (size_t i = 0; i < nodes.size(); i++) {
      const AirNode& nd = nodes[i];
      switch (nd.op) {
        case AIR_CONST: scratch[i] = e2(consts[nd.a]); break;
        case AIR_LOCAL: scratch[i] = local[nd.a]; break;
        case AIR_NEXT: scratch[i] = next[nd.a]; break;
        case AIR_PI: scratch[i] = e2(pis[nd.a]); break;
        case AIR_ADD: scratch[i] = e2_add(scratch[nd.a], scratch[nd.b]); break;
        case AIR_SUB: scratch[i] = e2_sub(scratch[nd.a], scratch[nd.b]); break;
        default: scratch[i] = e2_mul(scratch[nd.a], scratch[nd.b]); break;
      }
    }
    for (unsigned j = 0; j < n_alphas; j++) acc[j] = e2(0);
    for (size_t k = 0; k < cons.size(); k++) {
      E2 c = scratch[cons[k].node];
      switch (cons[k].cls) {
        case CLS_TRANSITION: c = e2_mul(c, z_last); break;
        case CLS_FIRST: c = e2_mul(c, l_first); break;
        case CLS_LAST: c = e2_mul(c, l_last); break;
        default: break;
      }
      for (unsigned j = 0; j < n_alphas; j++) acc[j] = e2_add(e2_scale(acc[j], alphas[j]), c);
    }
  }
};

}  // namespace orc
