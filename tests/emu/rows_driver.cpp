// TEST INFRASTRUCTURE ONLY: drives the product's gather / scatter (csrc/gather_scatter.cu: rows_op and its kernels) and CSR
// aggregation (csrc/aggregate.cu) on the CPU through tests/emu.  Every operand is copied into an emulated "device" allocation
// of EXACTLY its size, so that an access past the end of a table, an index array or an output is a guard violation (or an
// AddressSanitizer report when built with WGB_EMU_ASAN=1) instead of a silent read of neighbouring memory.
#include "cuda_emu.h"

#include "wm_common.cuh"

#include <wholememory/b200_ops.h>
#include <wholememory/graph_op.h>

#include <vector>

extern "C" long long emu_guard_violations();
extern "C" long long emu_live_allocations();

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t n = 0;
  DevBuf(const void* host, size_t bytes) : n(bytes)
  {
    cudaMalloc(&p, bytes);
    if (host && bytes) std::memcpy(p, host, bytes);
  }
  ~DevBuf() { cudaFree(p); }
};

wholememory_tensor_ tensor1d(void* p, long long n, wholememory_dtype_t dt)
{
  wholememory_tensor_ t;
  wholememory_initialize_tensor_desc(&t.desc);
  t.desc.dim = 1; t.desc.sizes[0] = n; t.desc.strides[0] = 1; t.desc.dtype = dt;
  t.storage_ptr = p;
  return t;
}
wholememory_tensor_ tensor2d(void* p, long long rows, long long dim, long long stride, long long storage_offset, wholememory_dtype_t dt)
{
  wholememory_tensor_ t;
  wholememory_initialize_tensor_desc(&t.desc);
  t.desc.dim = 2; t.desc.sizes[0] = rows; t.desc.sizes[1] = dim; t.desc.strides[0] = stride; t.desc.strides[1] = 1;
  t.desc.storage_offset = storage_offset; t.desc.dtype = dt;
  t.storage_ptr = p;
  return t;
}
size_t elt(int dt) { return wholememory_dtype_get_element_size((wholememory_dtype_t)dt); }

}  // namespace

extern "C" {

// table: (storage_offset + rows * stride) elements of table_dtype; dense: n * dense_stride elements of dense_dtype (row-major, no
// offset).  scatter = 0: dense <- table[idx]; 1: table[idx] <- dense.  hot_slot (int32 [rows]) / hot_rows ([H, dim] of
// table_dtype, stride dim): the replicated hot rows consulted by same-dtype gathers, or NULL.  Returns the product's error code,
// or -102 for a write outside an allocation.
int emu_rows_op(void* table, long long rows, long long dim, long long stride, long long storage_offset, int table_dtype, const void* idx,
                long long n, int idx_is_int64, void* dense, long long dense_stride, int dense_dtype, int scatter, const int* hot_slot,
                const void* hot_rows, long long hot_count, int sms)
{
  const size_t table_bytes = (size_t)(storage_offset + rows * stride) * elt(table_dtype);
  const size_t dense_bytes = (size_t)(n * dense_stride) * elt(dense_dtype);
  int rc;
  {
    DevBuf d_table(table, table_bytes), d_idx(idx, (size_t)n * (idx_is_int64 ? 8 : 4)), d_dense(dense, dense_bytes);
    DevBuf d_slot(hot_slot, hot_slot ? (size_t)rows * 4 : 0), d_hot(hot_rows, hot_rows ? (size_t)hot_count * dim * elt(table_dtype) : 0);
    wholememory_tensor_ t  = tensor2d(d_table.p, rows, dim, stride, storage_offset, (wholememory_dtype_t)table_dtype);
    wholememory_tensor_ ix = tensor1d(d_idx.p, n, idx_is_int64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT);
    wholememory_tensor_ de = tensor2d(d_dense.p, n, dim, dense_stride, 0, (wholememory_dtype_t)dense_dtype);
    wgb_hot_rows hot;
    if (hot_slot) {
      hot.slot         = static_cast<const int*>(d_slot.p);
      hot.rows         = static_cast<const char*>(d_hot.p);
      hot.stride_bytes = (unsigned long long)dim * elt(table_dtype);
    }
    rc = wgb::rows_op(&t, &ix, &de, nullptr, sms, scatter != 0, hot_slot ? &hot : nullptr);
    if (rc == WHOLEMEMORY_SUCCESS && emu_guard_violations() != 0) rc = -102;
    if (scatter) std::memcpy(table, d_table.p, table_bytes);
    else std::memcpy(dense, d_dense.p, dense_bytes);
  }
  if (rc == WHOLEMEMORY_SUCCESS && (emu_guard_violations() != 0 || emu_live_allocations() != 0)) rc = -102;
  return rc;
}

// out[i, :] = sum | mean over e in [indptr[i], indptr[i+1]) of x[map[indices[e]], :]   (map may be NULL); out fp32 [n_dst, F]
int emu_csr_aggregate(const void* indptr, int indptr_is_int64, long long n_dst, const void* indices, int indices_is_int64, long long nnz,
                      const long long* map, long long map_len, const void* x, long long x_rows, long long F, int x_dtype, int reduce, float* out)
{
  int rc;
  {
    DevBuf d_ptr(indptr, (size_t)(n_dst + 1) * (indptr_is_int64 ? 8 : 4)), d_ind(indices, (size_t)nnz * (indices_is_int64 ? 8 : 4));
    DevBuf d_map(map, map ? (size_t)map_len * 8 : 0), d_x(x, (size_t)x_rows * F * elt(x_dtype)), d_out(nullptr, (size_t)n_dst * F * 4);
    wholememory_tensor_ tp = tensor1d(d_ptr.p, n_dst + 1, indptr_is_int64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT);
    wholememory_tensor_ ti = tensor1d(d_ind.p, nnz, indices_is_int64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT);
    wholememory_tensor_ tm = tensor1d(d_map.p, map_len, WHOLEMEMORY_DT_INT64);
    wholememory_tensor_ tx = tensor2d(d_x.p, x_rows, F, F, 0, (wholememory_dtype_t)x_dtype);
    wholememory_tensor_ to = tensor2d(d_out.p, n_dst, F, F, 0, WHOLEMEMORY_DT_FLOAT);
    rc = wholegraph_csr_aggregate(&tp, &ti, map ? &tm : nullptr, &tx, reduce, &to, nullptr);
    if (rc == WHOLEMEMORY_SUCCESS && emu_guard_violations() != 0) rc = -102;
    std::memcpy(out, d_out.p, (size_t)n_dst * F * 4);
  }
  if (rc == WHOLEMEMORY_SUCCESS && (emu_guard_violations() != 0 || emu_live_allocations() != 0)) rc = -102;
  return rc;
}

// S3: graph_append_unique (csrc/append_unique.cu).  unique_out holds T + N entries of the id type; returns the unique count, or a
// negative error (-(1000 + code) for a product error code, -102 for a write outside an allocation).
namespace {
struct TempCtx {
  void* p = nullptr;
};
void temp_create(void** ctx, void*) { *ctx = new TempCtx(); }
void temp_destroy(void* ctx, void*) { delete static_cast<TempCtx*>(ctx); }
void* temp_malloc(wholememory_tensor_description_t* d, wholememory_memory_allocation_type_t, void* ctx, void*)
{
  size_t bytes = (size_t)std::max<long long>(d->sizes[0], 1) * wholememory_dtype_get_element_size(d->dtype);
  cudaMalloc(&static_cast<TempCtx*>(ctx)->p, bytes);
  return static_cast<TempCtx*>(ctx)->p;
}
void temp_free(void* ctx, void*)
{
  cudaFree(static_cast<TempCtx*>(ctx)->p);
  static_cast<TempCtx*>(ctx)->p = nullptr;
}
struct OutCtx {
  void* p         = nullptr;
  long long count = 0;
};
void* out_malloc2(wholememory_tensor_description_t* d, wholememory_memory_allocation_type_t, void* ctx, void*)
{
  OutCtx* o = static_cast<OutCtx*>(ctx);
  o->count  = d->sizes[0];
  cudaMalloc(&o->p, (size_t)std::max<long long>(o->count, 1) * wholememory_dtype_get_element_size(d->dtype));
  return o->p;
}
void out_free2(void* ctx, void*) { cudaFree(static_cast<OutCtx*>(ctx)->p); }
}  // namespace

long long emu_append_unique(const void* targets, long long T, const void* neighbors, long long N, int is_int64, void* unique_out, int* raw_to_unique)
{
  const size_t e = is_int64 ? 8 : 4;
  long long result;
  {
    DevBuf d_t(targets, (size_t)T * e), d_n(neighbors, (size_t)N * e), d_r(nullptr, (size_t)N * 4);
    const wholememory_dtype_t dt = is_int64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT;
    wholememory_tensor_ tt = tensor1d(d_t.p, T, dt), tn = tensor1d(d_n.p, N, dt), tr = tensor1d(d_r.p, N, WHOLEMEMORY_DT_INT);
    wholememory_env_func_t env;
    std::memset(&env, 0, sizeof(env));
    env.temporary_fns.create_memory_context_fn  = temp_create;
    env.temporary_fns.destroy_memory_context_fn = temp_destroy;
    env.temporary_fns.malloc_fn                 = temp_malloc;
    env.temporary_fns.free_fn                   = temp_free;
    env.output_fns.malloc_fn                    = out_malloc2;
    env.output_fns.free_fn                      = out_free2;
    OutCtx out;
    int rc = graph_append_unique(&tt, &tn, &out, raw_to_unique ? &tr : nullptr, &env, nullptr);
    if (rc != WHOLEMEMORY_SUCCESS) {
      result = -(1000 + rc);
    } else {
      result = emu_guard_violations() != 0 ? -102 : out.count;
      if (result >= 0) {
        std::memcpy(unique_out, out.p, (size_t)out.count * e);
        if (raw_to_unique) std::memcpy(raw_to_unique, d_r.p, (size_t)N * 4);
      }
    }
    cudaFree(out.p);
  }
  if (result >= 0 && (emu_guard_violations() != 0 || emu_live_allocations() != 0)) result = -102;
  return result;
}

}  // extern "C"
