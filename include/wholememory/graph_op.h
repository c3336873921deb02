/*
 * S3: append-unique (hop-to-hop renumbering).
 *
 * Replaces /root/reference/cpp/include/wholememory/graph_op.h:27-33 (same signature);
 * kernels replaced: cpp/src/graph_ops/append_unique_func.cuh:201-341.
 *   unique       = targets ++ (neighbours not in targets, de-duplicated)
 *   raw_to_unique[k] = position of neighbours[k] in unique   (int32, optional)
 * New ids are assigned in FIRST-OCCURRENCE order (deterministic; the reference's GPU order is
 * hash-slot order and its tests sort the tail before comparing).
 * csr_add_self_loop (graph_op.h:44-48) is a GAT helper outside the hot path and is not exported.
 */
#pragma once

#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_tensor.h>

#ifdef __cplusplus
extern "C" {
#endif

/* target_nodes / neighbor_nodes: int32 | int64 local device tensors of one dtype; the unique list (same dtype) is allocated
 * through p_env_fns->output_fns into output_unique_node_memory_context; the mapping tensor (int32 [len(neighbors)]) is
 * caller-allocated and may be NULL.  One host synchronisation (to size the unique list). */
wholememory_error_code_t graph_append_unique(wholememory_tensor_t target_nodes_tensor, wholememory_tensor_t neighbor_nodes_tensor,
                                             void* output_unique_node_memory_context,
                                             wholememory_tensor_t output_neighbor_raw_to_unique_mapping_tensor,
                                             wholememory_env_func_t* p_env_fns, void* stream);

#ifdef __cplusplus
}
#endif
