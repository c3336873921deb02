/*
 * S1 / S2: one-hop neighbour sampling WITHOUT replacement over a CSR graph held in WholeMemory.
 *
 * Replaces /root/reference/cpp/include/wholememory/wholegraph_op.h:31-73 (same signatures);
 * kernels replaced: cpp/src/wholegraph_ops/unweighted_sample_without_replacement_func.cuh:29-271,
 *                   cpp/src/wholegraph_ops/weighted_sample_without_replacement_func.cuh:208-291,
 *                   cpp/src/wholegraph_ops/sample_comm.cuh:14-48.
 *
 * row_ptr must be int64, col int32|int64, weights fp32|fp64, center nodes int32|int64,
 * output_sample_offset a caller-allocated int32[n+1].  dest (col dtype), center_localid (int32)
 * and edge_gid (int64) are allocated through p_env_fns->output_fns; the last two contexts may
 * be NULL.  max_sample_count <= 0 copies the full adjacency in CSR order.
 */
#pragma once

#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_tensor.h>

#ifdef __cplusplus
extern "C" {
#endif

wholememory_error_code_t wholegraph_csr_unweighted_sample_without_replacement(
  wholememory_tensor_t wm_csr_row_ptr_tensor,
  wholememory_tensor_t wm_csr_col_ptr_tensor,
  wholememory_tensor_t center_nodes_tensor,
  int max_sample_count,
  wholememory_tensor_t output_sample_offset_tensor,
  void* output_dest_memory_context,
  void* output_center_localid_memory_context,
  void* output_edge_gid_memory_context,
  unsigned long long random_seed,
  wholememory_env_func_t* p_env_fns,
  void* stream);

wholememory_error_code_t wholegraph_csr_weighted_sample_without_replacement(
  wholememory_tensor_t wm_csr_row_ptr_tensor,
  wholememory_tensor_t wm_csr_col_ptr_tensor,
  wholememory_tensor_t wm_csr_weight_ptr_tensor,
  wholememory_tensor_t center_nodes_tensor,
  int max_sample_count,
  wholememory_tensor_t output_sample_offset_tensor,
  void* output_dest_memory_context,
  void* output_center_localid_memory_context,
  void* output_edge_gid_memory_context,
  unsigned long long random_seed,
  wholememory_env_func_t* p_env_fns,
  void* stream);

/* host twins of the device random stream (reference: cpp/src/wholegraph_ops/raft_random_gen.cu:15-96) */
wholememory_error_code_t generate_random_positive_int_cpu(int64_t random_seed,
                                                          int64_t subsequence,
                                                          wholememory_tensor_t output);
wholememory_error_code_t generate_exponential_distribution_negative_float_cpu(
  int64_t random_seed, int64_t subsequence, wholememory_tensor_t output);

#ifdef __cplusplus
}
#endif
