// TEST INFRASTRUCTURE ONLY: drives the product's one-hop samplers S1 / S2 (csrc/sample.cu: wholegraph_csr_unweighted_ /
// weighted_sample_without_replacement, host code and kernels) on the CPU through tests/emu; operands live in exactly-sized,
// guard-fenced emulated device allocations.
#include "cuda_emu.h"

#include "wm_common.cuh"

#include <wholememory/wholegraph_op.h>

#include <cstdlib>

extern "C" long long emu_guard_violations();

namespace {

struct Out {
  void* p         = nullptr;
  long long count = 0;
  int elt         = 0;
};
void* out_malloc(wholememory_tensor_description_t* d, wholememory_memory_allocation_type_t, void* ctx, void*)
{
  Out* o   = static_cast<Out*>(ctx);
  o->count = d->sizes[0];
  o->elt   = (int)wholememory_dtype_get_element_size(d->dtype);
  cudaMalloc(&o->p, (size_t)std::max<long long>(o->count, 1) * o->elt);
  return o->p;
}
void out_free(void* ctx, void*) { cudaFree(static_cast<Out*>(ctx)->p); }
struct TempCtx {
  void* p = nullptr;
};
void temp_create(void** ctx, void*) { *ctx = new TempCtx(); }
void temp_destroy(void* ctx, void*) { delete static_cast<TempCtx*>(ctx); }
void* temp_malloc(wholememory_tensor_description_t* d, wholememory_memory_allocation_type_t, void* ctx, void*)
{
  cudaMalloc(&static_cast<TempCtx*>(ctx)->p, (size_t)std::max<long long>(d->sizes[0], 1) * wholememory_dtype_get_element_size(d->dtype));
  return static_cast<TempCtx*>(ctx)->p;
}
void temp_free(void* ctx, void*)
{
  cudaFree(static_cast<TempCtx*>(ctx)->p);
  static_cast<TempCtx*>(ctx)->p = nullptr;
}
wholememory_tensor_ tensor1d(void* p, long long n, wholememory_dtype_t dt)
{
  wholememory_tensor_ t;
  wholememory_initialize_tensor_desc(&t.desc);
  t.desc.dim = 1; t.desc.sizes[0] = n; t.desc.strides[0] = 1; t.desc.dtype = dt;
  t.storage_ptr = p;
  return t;
}
void* dev_copy(const void* host, size_t bytes)
{
  void* p = nullptr;
  cudaMalloc(&p, bytes);
  if (host && bytes) std::memcpy(p, host, bytes);
  return p;
}

}  // namespace

extern "C" {

// weight == NULL: S1 (uniform), else S2 (A-Res).  offsets int32 [n+1]; dest (col dtype) / lid int32 / gid int64 hold up to cap
// entries.  Returns the number of sampled edges, -(1000 + code) for a product error, -102 for a write outside an allocation.
long long emu_one_hop(const long long* row_ptr, long long V, const void* col, long long E, int col_is_int64, const void* weight,
                      int weight_is_double, const void* centers, long long n, int centers_is_int64, int M, unsigned long long seed, int* offsets,
                      void* dest, int* lid, long long* gid, long long cap)
{
  const size_t ce = col_is_int64 ? 8 : 4;
  void* d_rp  = dev_copy(row_ptr, (size_t)(V + 1) * 8);
  void* d_col = dev_copy(col, (size_t)E * ce);
  void* d_w   = weight ? dev_copy(weight, (size_t)E * (weight_is_double ? 8 : 4)) : nullptr;
  void* d_c   = dev_copy(centers, (size_t)n * (centers_is_int64 ? 8 : 4));
  void* d_off = dev_copy(nullptr, (size_t)(n + 1) * 4);
  wholememory_tensor_ trp = tensor1d(d_rp, V + 1, WHOLEMEMORY_DT_INT64), tcol = tensor1d(d_col, E, col_is_int64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT);
  wholememory_tensor_ tw  = tensor1d(d_w, E, weight_is_double ? WHOLEMEMORY_DT_DOUBLE : WHOLEMEMORY_DT_FLOAT);
  wholememory_tensor_ tc  = tensor1d(d_c, n, centers_is_int64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT);
  wholememory_tensor_ to  = tensor1d(d_off, n + 1, WHOLEMEMORY_DT_INT);
  wholememory_env_func_t env;
  std::memset(&env, 0, sizeof(env));
  env.temporary_fns.create_memory_context_fn  = temp_create;
  env.temporary_fns.destroy_memory_context_fn = temp_destroy;
  env.temporary_fns.malloc_fn                 = temp_malloc;
  env.temporary_fns.free_fn                   = temp_free;
  env.output_fns.malloc_fn                    = out_malloc;
  env.output_fns.free_fn                      = out_free;
  Out o_dest, o_lid, o_gid;
  int rc = weight ? wholegraph_csr_weighted_sample_without_replacement(&trp, &tcol, &tw, &tc, M, &to, &o_dest, &o_lid, &o_gid, seed, &env, nullptr)
                  : wholegraph_csr_unweighted_sample_without_replacement(&trp, &tcol, &tc, M, &to, &o_dest, &o_lid, &o_gid, seed, &env, nullptr);
  long long result;
  if (rc != WHOLEMEMORY_SUCCESS) {
    result = -(1000 + rc);
  } else if (emu_guard_violations() != 0 || o_dest.count > cap || o_lid.count != o_dest.count || o_gid.count != o_dest.count) {
    result = -102;
  } else {
    result = o_dest.count;
    std::memcpy(offsets, d_off, (size_t)(n + 1) * 4);
    std::memcpy(dest, o_dest.p, (size_t)result * ce);
    std::memcpy(lid, o_lid.p, (size_t)result * 4);
    std::memcpy(gid, o_gid.p, (size_t)result * 8);
  }
  for (void* p : {d_rp, d_col, d_w, d_c, d_off, o_dest.p, o_lid.p, o_gid.p})
    cudaFree(p);
  if (result >= 0 && emu_guard_violations() != 0) result = -102;
  return result;
}

}  // extern "C"
