// TEST INFRASTRUCTURE ONLY: calls the product's multi-hop sampler entry points (csrc/multihop.cu, compiled for the CPU by
// tests/emu: launches rewritten by emu_preprocess.py, device code run by cuda_emu.h, runtime from emu_runtime.cpp) on
// plain host arrays and hands the outputs back to the Python test, which compares them with the oracle.
#include "cuda_emu.h"

#include "wm_common.cuh"

#include <wholememory/b200_ops.h>

#include <cstdlib>
#include <vector>

extern "C" long long emu_guard_violations();
extern "C" long long emu_live_allocations();

namespace {

struct Out {
  void* ptr       = nullptr;
  long long count = 0;
  int elt         = 0;
};

void* out_malloc(wholememory_tensor_description_t* d, wholememory_memory_allocation_type_t, void* ctx, void*)
{
  Out* o   = static_cast<Out*>(ctx);
  o->count = d->sizes[0];
  o->elt   = (int)wholememory_dtype_get_element_size(d->dtype);
  o->ptr   = std::malloc((size_t)std::max<long long>(o->count, 1) * o->elt);
  return o->ptr;
}
void out_free(void* ctx, void*)
{
  Out* o = static_cast<Out*>(ctx);
  std::free(o->ptr);
  o->ptr = nullptr;
}

wholememory_tensor_ make_tensor(const void* p, long long n, wholememory_dtype_t dt)
{
  wholememory_tensor_ t;
  wholememory_initialize_tensor_desc(&t.desc);
  t.desc.dim        = 1;
  t.desc.sizes[0]   = n;
  t.desc.strides[0] = 1;
  t.desc.dtype      = dt;
  t.storage_ptr     = const_cast<void*>(p);
  return t;
}

}  // namespace

extern "C" {

// outs[k] receive malloc'ed arrays (freed by emu_free): heterogeneous order = majors, minors, edge_id, edge_type,
// label_type_hop_offsets, renumber_map, renumber_map_offsets, edge_renumber_map, edge_renumber_map_offsets, step base;
// homogeneous order = majors, minors, edge_id, label_hop_offsets, renumber_map, renumber_map_offsets, major_offsets, step base;
// outs[10] = seed local ids (want_seed_ids).  out_ptr / out_count / out_elt hold 11 entries.
// col_is_int64: dtype of the cols; edge_time == NULL: the plain (non-temporal) call; weight == NULL: uniform.
int emu_multihop(int T, const long long* const* row_ptr, long long V, const void* const* col, const long long* num_edges, int col_is_int64,
                 const long long* const* edge_time, const void* const* weight, int weight_is_double, const long long* const* edge_id, const long long* vto, int Vt, int hetero,
                 const long long* seeds, const long long* seed_times, long long S, const long long* label_offsets, long long B, const int* fanout,
                 int hops, unsigned long long random_state, int cmp, int flags, int reps, int want_seed_ids, const long long* num_times,
                 void** out_ptr, long long* out_count, int* out_elt)
{
  std::vector<wholememory_tensor_> rp(T), cl(T), tm(T), ei(T), wt(T);
  std::vector<wholememory_tensor_t> rp_h(T), cl_h(T), tm_h(T), ei_h(T), wt_h(T);
  bool any_eid = false;
  for (int t = 0; t < T; t++) {
    rp[t]   = make_tensor(row_ptr[t], V + 1, WHOLEMEMORY_DT_INT64);
    cl[t]   = make_tensor(col[t], num_edges[t], col_is_int64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT);
    rp_h[t] = &rp[t];
    cl_h[t] = &cl[t];
    if (edge_time) {
      tm[t]   = make_tensor(edge_time[t], num_times ? num_times[t] : num_edges[t], WHOLEMEMORY_DT_INT64);  // real length: the entry point checks it
      tm_h[t] = &tm[t];
    }
    if (weight) {
      wt[t]   = make_tensor(weight[t], num_edges[t], weight_is_double ? WHOLEMEMORY_DT_DOUBLE : WHOLEMEMORY_DT_FLOAT);
      wt_h[t] = &wt[t];
    }
    ei_h[t] = nullptr;
    if (edge_id && edge_id[t]) {
      ei[t]   = make_tensor(edge_id[t], num_edges[t], WHOLEMEMORY_DT_INT64);
      ei_h[t] = &ei[t];
      any_eid = true;
    }
  }
  wholememory_tensor_ sd = make_tensor(seeds, S, WHOLEMEMORY_DT_INT64), st = make_tensor(seed_times, S, WHOLEMEMORY_DT_INT64),
                      lo = make_tensor(label_offsets, B + 1, WHOLEMEMORY_DT_INT64);
  wholememory_env_func_t env;
  std::memset(&env, 0, sizeof(env));
  env.output_fns.malloc_fn = out_malloc;
  env.output_fns.free_fn   = out_free;
  wholegraph_multihop_sampler_t sp = nullptr;
  if (wholegraph_create_multihop_sampler(&sp) != WHOLEMEMORY_SUCCESS) return -100;
  Out outs[11];  // [10]: local id of every input seed (wholegraph_multihop_seed_local_ids), when asked for
  int rc = -101;
  for (int rep = 0; rep < reps; rep++) {  // reps > 1: the same call again on one sampler object (scratch reuse, epochs, tickets)
    for (auto& o : outs) {
      std::free(o.ptr);
      o = Out();
    }
    if (edge_time) {
      rc = wholegraph_temporal_multihop_neighbor_sample_begin(sp, T, rp_h.data(), cl_h.data(), weight ? wt_h.data() : nullptr, tm_h.data(),
                                                              any_eid ? ei_h.data() : nullptr, vto, Vt, hetero, &sd, &st, &lo, fanout, hops,
                                                              random_state, cmp, flags, nullptr);
    } else if (hetero) {
      rc = wholegraph_hetero_multihop_neighbor_sample_begin(sp, T, rp_h.data(), cl_h.data(), weight ? wt_h.data() : nullptr, any_eid ? ei_h.data() : nullptr, vto, Vt, &sd,
                                                            &lo, fanout, hops, random_state, flags, nullptr);
    } else {
      rc = wholegraph_multihop_neighbor_sample_begin(sp, rp_h[0], cl_h[0], weight ? wt_h[0] : nullptr, ei_h[0], &sd, &lo, fanout, hops, random_state, flags, nullptr);
    }
    if (rc != WHOLEMEMORY_SUCCESS) break;
    if (hetero) {
      rc = wholegraph_hetero_multihop_neighbor_sample_finish(sp, &outs[0], &outs[1], &outs[2], &outs[3], &outs[4], &outs[5], &outs[6], &outs[7],
                                                             &outs[8], &outs[9], &env, nullptr);
    } else {
      const bool csr = (flags & WHOLEGRAPH_MULTIHOP_CSR) != 0;
      rc = wholegraph_multihop_neighbor_sample_finish(sp, csr ? nullptr : &outs[0], &outs[1], &outs[2], &outs[3], &outs[4], &outs[5],
                                                      csr ? &outs[6] : nullptr, &outs[7], &env, nullptr);
    }
    if (rc != WHOLEMEMORY_SUCCESS) break;
    if (want_seed_ids) {
      rc = wholegraph_multihop_seed_local_ids(sp, &outs[10], &env, nullptr);
      if (rc != WHOLEMEMORY_SUCCESS) break;
    }
  }
  const long long dirty = emu_guard_violations();  // before the scratch is released: every buffer of the call is still live
  wholegraph_destroy_multihop_sampler(sp);
  if (rc == WHOLEMEMORY_SUCCESS && (dirty != 0 || emu_guard_violations() != 0)) rc = -102;  // a write outside a device allocation
  if (rc == WHOLEMEMORY_SUCCESS && emu_live_allocations() != 0) rc = -103;                  // the sampler object leaked device memory
  for (int k = 0; k < 11; k++) {
    out_ptr[k]   = outs[k].ptr;
    out_count[k] = outs[k].count;
    out_elt[k]   = outs[k].elt;
  }
  return rc;
}

void emu_free(void* p) { std::free(p); }

}  // extern "C"
