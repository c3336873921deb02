/*
 * WholeMemory tensors: a descriptor on top of a WholeMemory handle, or a non-owning view of a
 * plain device/host pointer (how torch tensors are passed to the ops).
 *
 * Replaces /root/reference/cpp/include/wholememory/wholememory_tensor.h:20-186 (same signatures).
 */
#pragma once

#include <wholememory/tensor_description.h>
#include <wholememory/wholememory.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wholememory_tensor_* wholememory_tensor_t;

wholememory_error_code_t wholememory_create_tensor(
  wholememory_tensor_t* wholememory_tensor,
  wholememory_tensor_description_t* tensor_description,
  wholememory_comm_t comm,
  wholememory_memory_type_t memory_type,
  wholememory_memory_location_t memory_location,
  size_t* tensor_entry_partition = nullptr);
wholememory_error_code_t wholememory_destroy_tensor(wholememory_tensor_t wholememory_tensor);
wholememory_error_code_t wholememory_make_tensor_from_pointer(
  wholememory_tensor_t* wholememory_tensor,
  void* storage_ptr,
  wholememory_tensor_description_t* tensor_description);
wholememory_error_code_t wholememory_make_tensor_from_handle(
  wholememory_tensor_t* wholememory_tensor,
  wholememory_handle_t wholememory_handle,
  wholememory_tensor_description_t* tensor_description);
bool wholememory_tensor_has_handle(wholememory_tensor_t wholememory_tensor);
wholememory_handle_t wholememory_tensor_get_memory_handle(wholememory_tensor_t wholememory_tensor);
wholememory_tensor_description_t* wholememory_tensor_get_tensor_description(
  wholememory_tensor_t wholememory_tensor);
wholememory_error_code_t wholememory_tensor_get_global_reference(
  wholememory_tensor_t wholememory_tensor, wholememory_gref_t* wholememory_gref);
wholememory_error_code_t wholememory_tensor_map_local_tensor(
  wholememory_tensor_t wholememory_tensor, wholememory_tensor_t* local_tensor);
void* wholememory_tensor_get_data_pointer(wholememory_tensor_t wholememory_tensor);
wholememory_error_code_t wholememory_tensor_get_entry_offsets(
  size_t* entry_offsets, wholememory_tensor_t wholememory_tensor);
wholememory_error_code_t wholememory_tensor_get_entry_partition_sizes(
  size_t* entry_partition, wholememory_tensor_t wholememory_tensor);
wholememory_error_code_t wholememory_tensor_get_local_entry_count(
  size_t* local_entry_count, wholememory_tensor_t wholememory_tensor);
wholememory_error_code_t wholememory_tensor_get_local_entry_start(
  size_t* local_entry_start, wholememory_tensor_t wholememory_tensor);
wholememory_error_code_t wholememory_tensor_get_subtensor(
  wholememory_tensor_t wholememory_tensor,
  int64_t* starts,
  int64_t* ends,
  wholememory_tensor_t* sub_wholememory_tensor);
wholememory_tensor_t wholememory_tensor_get_root(wholememory_tensor_t wholememory_tensor);

#define WM_TENSOR_COUNT_DEBUG
int64_t get_wholememory_tensor_count();

#ifdef __cplusplus
}
#endif
