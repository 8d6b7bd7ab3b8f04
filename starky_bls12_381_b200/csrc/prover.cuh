// Internal declarations shared by capi.cu / prover.cu / quotient.cu / fri.cu.
#pragma once
#include "common.cuh"

// proof layout (include/starky_b200.h sb_proof_layout)
sb_proof_layout proof_layout(const sb_params& p);
uint32_t fri_step_path_len(const sb_proof_layout& l, uint32_t round);
uint64_t fri_step_offset(const sb_proof_layout& l, uint32_t round);
std::vector<unsigned> fri_arities(const sb_params& p);
static inline unsigned quotient_degree_factor(const sb_params& p) { return p.constraint_degree > 1 ? p.constraint_degree - 1 : 1; }

// capi.cu
void check_params(const sb_params* p);
int sb_fail(sb_ctx* ctx, const SbError& e);
const u64* ingest_trace(sb_ctx* ctx, const sb_params* p, const void* trace, int layout);
void commit_trace(sb_ctx* ctx, const sb_params* p, const u64* d_values);
void ingest_and_commit_trace(sb_ctx* ctx, const sb_params* p, const void* trace, int layout, cudaEvent_t h2d_done);
const u64* tree_cap_ptr(const u64* d_tree, size_t n_leaves, unsigned cap_height);

// quotient.cu
void air_release_all(sb_ctx* ctx);

// fri.cu
void sb_bitrev_permute_device(sb_ctx* ctx, const u64* d_in, u64* d_out, unsigned log_size, uint32_t count);

// prover.cu
sb_proof* proof_alloc(const sb_params& p);

// group.cu
void multi_init(sb_ctx* ctx, const int* devices, int n);
void multi_destroy(sb_ctx* ctx);
int multi_prove(sb_ctx* ctx, const sb_params* p, const void* trace, int layout, const uint64_t* public_inputs, sb_proof** out);
