/*
 * dtype enum and array / matrix / tensor descriptors.
 *
 * Replaces (same names, same field order, same enum values):
 *   /root/reference/cpp/include/wholememory/tensor_description.h:19-30   (wholememory_dtype_t)
 *   /root/reference/cpp/include/wholememory/tensor_description.h:56-91   (array / matrix / tensor descriptors)
 *   /root/reference/cpp/include/wholememory/tensor_description.h:100-236 (helpers)
 * All sizes, strides and storage offsets count ELEMENTS, never bytes.
 */
#pragma once

#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

enum wholememory_dtype_t {
  WHOLEMEMORY_DT_UNKNOWN = 0,
  WHOLEMEMORY_DT_FLOAT,
  WHOLEMEMORY_DT_HALF,
  WHOLEMEMORY_DT_DOUBLE,
  WHOLEMEMORY_DT_BF16,
  WHOLEMEMORY_DT_INT,
  WHOLEMEMORY_DT_INT64,
  WHOLEMEMORY_DT_INT16,
  WHOLEMEMORY_DT_INT8,
  WHOLEMEMORY_DT_COUNT,
};

size_t wholememory_dtype_get_element_size(wholememory_dtype_t dtype); /* (size_t)-1 on invalid dtype */
bool wholememory_dtype_is_floating_number(wholememory_dtype_t dtype);
bool wholememory_dtype_is_integer_number(wholememory_dtype_t dtype);

struct wholememory_array_description_t {
  int64_t size;
  int64_t storage_offset;
  wholememory_dtype_t dtype;
};

struct wholememory_matrix_description_t {
  int64_t sizes[2]; /* rows, columns */
  int64_t stride;   /* elements between consecutive rows */
  int64_t storage_offset;
  wholememory_dtype_t dtype;
};

#define WHOLEMEMORY_MAX_TENSOR_DIM (8)

struct wholememory_tensor_description_t {
  int64_t sizes[WHOLEMEMORY_MAX_TENSOR_DIM];
  int64_t strides[WHOLEMEMORY_MAX_TENSOR_DIM];
  int64_t storage_offset;
  int dim;
  wholememory_dtype_t dtype;
};

wholememory_array_description_t wholememory_create_array_desc(int64_t size,
                                                              int64_t storage_offset,
                                                              wholememory_dtype_t dtype);
wholememory_matrix_description_t wholememory_create_matrix_desc(int64_t sizes[2],
                                                                int64_t stride,
                                                                int64_t storage_offset,
                                                                wholememory_dtype_t dtype);
void wholememory_initialize_tensor_desc(wholememory_tensor_description_t* p_tensor_description);
void wholememory_copy_array_desc_to_matrix(wholememory_matrix_description_t* p_matrix_description,
                                           wholememory_array_description_t* p_array_description);
void wholememory_copy_array_desc_to_tensor(wholememory_tensor_description_t* p_tensor_description,
                                           wholememory_array_description_t* p_array_description);
void wholememory_copy_matrix_desc_to_tensor(wholememory_tensor_description_t* p_tensor_description,
                                            wholememory_matrix_description_t* p_matrix_description);
bool wholememory_convert_tensor_desc_to_array(wholememory_array_description_t* p_array_description,
                                              wholememory_tensor_description_t* p_tensor_description);
bool wholememory_convert_tensor_desc_to_matrix(
  wholememory_matrix_description_t* p_matrix_description,
  wholememory_tensor_description_t* p_tensor_description);
int64_t wholememory_get_memory_element_count_from_array(
  wholememory_array_description_t* p_array_description);
int64_t wholememory_get_memory_size_from_array(wholememory_array_description_t* p_array_description);
int64_t wholememory_get_memory_element_count_from_matrix(
  wholememory_matrix_description_t* p_matrix_description);
int64_t wholememory_get_memory_size_from_matrix(
  wholememory_matrix_description_t* p_matrix_description);
int64_t wholememory_get_memory_element_count_from_tensor(
  wholememory_tensor_description_t* p_tensor_description);
int64_t wholememory_get_memory_size_from_tensor(
  wholememory_tensor_description_t* p_tensor_description);
bool wholememory_squeeze_tensor(wholememory_tensor_description_t* p_tensor_description, int dim);
bool wholememory_unsqueeze_tensor(wholememory_tensor_description_t* p_tensor_description, int dim);

#ifdef __cplusplus
}
#endif
