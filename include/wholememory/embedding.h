/*
 * WholeMemory embedding: the non-cached feature table that feeds each mini-batch.
 *
 * Replaces, for the feature-fetch path only, /root/reference/cpp/include/wholememory/embedding.h:
 *   wholememory_create_embedding            :119-128   (cache_policy must be NULL)
 *   wholememory_destroy_embedding           :135-136
 *   wholememory_embedding_get_embedding_tensor :143-144
 *   wholememory_embedding_gather            :173-178   (== noncached_embedding::gather,
 *                                            cpp/src/wholememory/embedding.cpp:545-554,1045-1073)
 * Cache policies and sparse optimizers (trainable embeddings) are outside the hot path
 * (SURVEY.md §8f row 4): passing a non-NULL cache policy returns WHOLEMEMORY_NOT_SUPPORTED.
 */
#pragma once

#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory_tensor.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wholememory_embedding_cache_policy_* wholememory_embedding_cache_policy_t;
typedef struct wholememory_embedding_* wholememory_embedding_t;

wholememory_error_code_t wholememory_create_embedding(
  wholememory_embedding_t* wholememory_embedding,
  wholememory_tensor_description_t* embedding_tensor_description,
  wholememory_comm_t comm,
  wholememory_memory_type_t memory_type,
  wholememory_memory_location_t memory_location,
  wholememory_embedding_cache_policy_t cache_policy,
  size_t* embedding_entry_partition = nullptr,
  int user_defined_sms              = -1,
  int round_robin_size              = 0);

wholememory_error_code_t wholememory_destroy_embedding(
  wholememory_embedding_t wholememory_embedding);

wholememory_tensor_t wholememory_embedding_get_embedding_tensor(
  wholememory_embedding_t wholememory_embedding);

wholememory_error_code_t wholememory_embedding_gather(wholememory_embedding_t wholememory_embedding,
                                                      wholememory_tensor_t indices,
                                                      wholememory_tensor_t output,
                                                      bool adjust_cache,
                                                      wholememory_env_func_t* p_env_fns,
                                                      int64_t stream_int);

#ifdef __cplusplus
}
#endif
