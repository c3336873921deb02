/*
 * WholeMemory embedding: the non-cached feature table that feeds each mini-batch.
 *
 * Replaces, for the feature-fetch path only, /root/reference/cpp/include/wholememory/embedding.h:
 *   wholememory_create_embedding            :119-128   (cache_policy must be NULL)
 *   wholememory_destroy_embedding           :135-136
 *   wholememory_embedding_get_embedding_tensor :143-144
 *   wholememory_embedding_gather            :173-178   (== noncached_embedding::gather,
 *                                            cpp/src/wholememory/embedding.cpp:545-554,1045-1073)
 * Trainable embeddings, same declarations as the reference's embedding.h:
 *   wholememory_optimizer_type_t, wholememory_create_embedding_optimizer, wholememory_optimizer_set_parameter,
 *   wholememory_destroy_embedding_optimizer                                   :45-80
 *   wholememory_embedding_set_optimizer                                        :151-152
 *   wholememory_embedding_gather_gradient_apply                                :191-198
 *   wholememory_embedding_get_optimizer_state_names / _get_optimizer_state     :205-215
 * (implementation: cugraph-gnn_b200/csrc/embedding_optimizer.cu -- publish/pull over peer-mapped mailboxes instead of
 * the reference's two NCCL all-to-alls).  Cache policies are not provided: tables live in HBM; passing a non-NULL
 * cache policy returns WHOLEMEMORY_NOT_SUPPORTED.
 */
#pragma once

#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory_tensor.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wholememory_embedding_cache_policy_* wholememory_embedding_cache_policy_t;
typedef struct wholememory_embedding_optimizer_* wholememory_embedding_optimizer_t;
typedef struct wholememory_embedding_* wholememory_embedding_t;

enum wholememory_optimizer_type_t {
  WHOLEMEMORY_OPT_NONE = 0,  /* no optimizer */
  WHOLEMEMORY_OPT_SGD,       /* parameters: weight_decay */
  WHOLEMEMORY_OPT_LAZY_ADAM, /* weight_decay, epsilon, beta1, beta2, adam_w; states m, v, beta12t */
  WHOLEMEMORY_OPT_RMSPROP,   /* weight_decay, epsilon, alpha; state v */
  WHOLEMEMORY_OPT_ADAGRAD,   /* weight_decay, epsilon; state state_sum */
};

wholememory_error_code_t wholememory_create_embedding_optimizer(wholememory_embedding_optimizer_t* optimizer,
                                                                wholememory_optimizer_type_t optimizer_type);

/* value points to a float */
wholememory_error_code_t wholememory_optimizer_set_parameter(wholememory_embedding_optimizer_t optimizer,
                                                             const char* parameter_name,
                                                             void* value);

void wholememory_destroy_embedding_optimizer(wholememory_embedding_optimizer_t optimizer);

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

/* collective over the embedding's communicator: allocates the per-row states (partitioned like the table) */
wholememory_error_code_t wholememory_embedding_set_optimizer(wholememory_embedding_t wholememory_embedding,
                                                             wholememory_embedding_optimizer_t optimizer);

/* collective: every rank passes ITS (indices, fp32 gradient rows); rows may repeat within and across ranks -- their
 * gradients are summed before the optimizer step, which runs once per distinct row on the rank that owns it */
wholememory_error_code_t wholememory_embedding_gather_gradient_apply(wholememory_embedding_t wholememory_embedding,
                                                                     wholememory_tensor_t indices,
                                                                     wholememory_tensor_t grads,
                                                                     bool adjust_cache,
                                                                     float lr,
                                                                     wholememory_env_func_t* p_env_fns,
                                                                     int64_t stream_int);

/* nullptr-terminated */
const char* const* wholememory_embedding_get_optimizer_state_names(wholememory_embedding_t wholememory_embedding);

wholememory_tensor_t wholememory_embedding_get_optimizer_state(wholememory_embedding_t wholememory_embedding,
                                                               const char* name);

#ifdef __cplusplus
}
#endif
