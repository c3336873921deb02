/*
 * Caller-supplied allocation callbacks.  Ops whose output size is only known on the device
 * (samplers, append-unique) allocate their outputs through `output_fns`, scratch through
 * `temporary_fns`; the Python layer implements both with torch.empty so results come back as
 * torch.Tensor without a copy.
 *
 * Replaces (same names and layouts):
 *   /root/reference/cpp/include/wholememory/env_func_ptrs.h:22-62
 */
#pragma once

#include <wholememory/tensor_description.h>

#ifdef __cplusplus
extern "C" {
#endif

enum wholememory_memory_allocation_type_t {
  WHOLEMEMORY_MA_NONE = 0,
  WHOLEMEMORY_MA_DEVICE,
  WHOLEMEMORY_MA_HOST,
  WHOLEMEMORY_MA_PINNED,
};

/* A "memory context" is the caller's handle for ONE allocation (in Python: an object that ends up holding the torch
 * tensor); `global_context` is passed back verbatim to every callback. */
typedef void (*wholememory_create_memory_context_func_t)(void** memory_context, void* global_context);
typedef void (*wholememory_destroy_memory_context_func_t)(void* memory_context, void* global_context);
/* returns the data pointer of a fresh allocation described by `desc` (sizes, dtype) in the requested kind of memory */
typedef void* (*wholememory_malloc_func_t)(wholememory_tensor_description_t* desc,
                                           wholememory_memory_allocation_type_t memory_allocation_type, void* memory_context,
                                           void* global_context);
typedef void (*wholememory_free_func_t)(void* memory_context, void* global_context);

/* scratch that lives for one op: context created and destroyed by the op itself */
struct wholememory_temp_memory_func_t {
  wholememory_create_memory_context_func_t create_memory_context_fn;
  wholememory_destroy_memory_context_func_t destroy_memory_context_fn;
  wholememory_malloc_func_t malloc_fn;
  wholememory_free_func_t free_fn;
  void* global_context;
};
/* results handed back to the caller: the caller creates the context, the op fills it through malloc_fn */
struct wholememory_output_memory_func_t {
  wholememory_malloc_func_t malloc_fn;
  wholememory_free_func_t free_fn;
  void* global_context;
};
struct wholememory_env_func_t {
  wholememory_temp_memory_func_t temporary_fns;
  wholememory_output_memory_func_t output_fns;
};

/*
 * Built-in environment backed by cudaMallocAsync/cudaFreeAsync on the op's stream and a small
 * registry of output contexts; lets C/C++ callers use the ops without writing callbacks
 * (the reference ships the analogous default in cpp/src/wholememory/env_func_ptrs.cpp).
 * An output context created here is a `wholememory_default_output_t`.
 */
struct wholememory_default_output_t {
  void* ptr;
  wholememory_tensor_description_t desc;
};
wholememory_env_func_t* wholememory_get_default_env_func();
void wholememory_default_output_release(wholememory_default_output_t* out);

#ifdef __cplusplus
}
#endif
