/*
 * S1 / S2: one-hop neighbour sampling WITHOUT replacement over a CSR graph held in WholeMemory.
 *
 * Drop-in for the entry points of /root/reference/cpp/include/wholememory/wholegraph_op.h:31-73 (same names, same
 * argument order and types -- that is the boundary a binding links against); kernels replaced:
 *   cpp/src/wholegraph_ops/unweighted_sample_without_replacement_func.cuh:29-271,
 *   cpp/src/wholegraph_ops/weighted_sample_without_replacement_func.cuh:208-291,
 *   cpp/src/wholegraph_ops/sample_comm.cuh:14-48.
 *
 * Shared contract of the two samplers
 *   csr_row_ptr      int64 [V+1]           WholeMemory (any type) or local device tensor
 *   csr_col          int32 | int64 [E]
 *   csr_weight       fp32 | fp64 [E]       (weighted form only; A-Res keys log2(u)/w, the M largest win)
 *   center_nodes     int32 | int64 [n]     local device tensor
 *   max_sample_count M: <= 0 copies the whole adjacency in CSR order; > 1024 returns WHOLEMEMORY_NOT_IMPLEMENTED
 *   sample_offset    int32 [n+1]           caller-allocated; sample_offset[i+1] - sample_offset[i] = min(deg_i, M)
 *   dest / center_localid / edge_gid       col dtype / int32 / int64 [sample_offset[n]], allocated through
 *                                          p_env_fns->output_fns into the given contexts; the last two may be NULL
 *   random_seed      stream of row b, slot j is PCG(seed, subsequence = b * block + j) (csrc/pcg.cuh)
 * One host synchronisation per call (to size the outputs).
 */
#pragma once

#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_tensor.h>

#ifdef __cplusplus
extern "C" {
#endif

/* uniform: Fisher-Yates over the row's positions, order of the M outputs is part of the contract */
wholememory_error_code_t wholegraph_csr_unweighted_sample_without_replacement(
  wholememory_tensor_t wm_csr_row_ptr_tensor, wholememory_tensor_t wm_csr_col_ptr_tensor,
  wholememory_tensor_t center_nodes_tensor, int max_sample_count, wholememory_tensor_t output_sample_offset_tensor,
  void* output_dest_memory_context, void* output_center_localid_memory_context, void* output_edge_gid_memory_context,
  unsigned long long random_seed, wholememory_env_func_t* p_env_fns, void* stream);

/* weight-biased: per-row outputs are a set (the reference's tests compare them sorted) */
wholememory_error_code_t wholegraph_csr_weighted_sample_without_replacement(
  wholememory_tensor_t wm_csr_row_ptr_tensor, wholememory_tensor_t wm_csr_col_ptr_tensor,
  wholememory_tensor_t wm_csr_weight_ptr_tensor, wholememory_tensor_t center_nodes_tensor, int max_sample_count,
  wholememory_tensor_t output_sample_offset_tensor, void* output_dest_memory_context,
  void* output_center_localid_memory_context, void* output_edge_gid_memory_context, unsigned long long random_seed,
  wholememory_env_func_t* p_env_fns, void* stream);

/* host twins of the device random stream (reference: cpp/src/wholegraph_ops/raft_random_gen.cu:15-96): `output` is a
 * host tensor (int32 / int64, resp. fp32) filled with the first draws of stream (random_seed, subsequence) */
wholememory_error_code_t generate_random_positive_int_cpu(int64_t random_seed, int64_t subsequence, wholememory_tensor_t output);
wholememory_error_code_t generate_exponential_distribution_negative_float_cpu(int64_t random_seed, int64_t subsequence,
                                                                              wholememory_tensor_t output);

#ifdef __cplusplus
}
#endif
