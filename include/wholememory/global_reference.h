/*
 * Global reference of a WholeMemory allocation: what a kernel needs to address memory that is
 * striped over the GPUs of one NVSwitch box.
 *
 * Replaces (same type name, same fields, same meaning):
 *   /root/reference/cpp/include/wholememory/global_reference.h:19-28  (wholememory_gref_t)
 *   /root/reference/cpp/include/wholememory/global_reference.h:36     (create_continuous_global_reference)
 */
#pragma once

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

struct wholememory_gref_t {
  void* pointer;               /* CONTINUOUS: flat base pointer. CHUNKED: DEVICE array of world_size base pointers */
  size_t* rank_memory_offsets; /* DEVICE array, world_size + 1 byte offsets (start of every rank's chunk) */
  int world_size;
  size_t stride;   /* 0 for CONTINUOUS; bytes per chunk for CHUNKED */
  bool same_chunk; /* true: rank == byte_offset / stride */
};

wholememory_gref_t wholememory_create_continuous_global_reference(void* ptr);

#ifdef __cplusplus
}
#endif
