// S1 / S2: one-hop neighbour sampling without replacement over CSR in WholeMemory (sm_100a).
//
// Replaces cpp/src/wholegraph_ops/unweighted_sample_without_replacement_func.cuh:29-464,
// weighted_sample_without_replacement_func.cuh:208-653 and sample_comm.cuh:14-48.
//
// What changed versus the reference:
//  * count + exclusive scan are ONE single-pass kernel (decoupled look-back) instead of a count
//    kernel + thrust::exclusive_scan; one host sync per call (to size the outputs) instead of two.
//  * fan-outs <= 32 (every BASELINE config) run a sub-warp-per-seed kernel: 8/16/32 lanes own one
//    seed row, a warp first loads row_ptr for 32 seeds at once (32 independent random reads in
//    flight), the Fisher-Yates chain a[i]=Q[r[i]], Q[r[i]]=Q[N-i-1] is resolved with
//    __match_any_sync + one shared int per lane + log2(M) shuffle pointer-jumps -- no CUB block
//    radix sort, no __syncthreads, 128 B of shared memory per warp.
//  * the RAFT-compatible PCG stream of virtual thread (b*T + j) is reached through a table-driven
//    affine skip (pcg.cuh) so the random numbers are bit-identical to the reference geometry
//    (block = one seed, T = 32*warp_count[(M-1)/32] threads, ITEMS draws per thread) no matter how
//    the work is actually mapped to lanes.
//  * all graph arrays are addressed through wgb::ChunkRef, i.e. a CSR striped over the GPUs of the
//    box is read by P2P loads from inside the kernel.

#include "wm_common.cuh"
#include "pcg.cuh"
#include "sample_device.cuh"

#include <mutex>

namespace wgb {

// ---- device skip table ----------------------------------------------------------------------------
const Affine* skip_table_device()
{
  static std::mutex mu;
  static Affine* tabs[64] = {};
  int dev                 = 0;
  WGB_CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (tabs[dev] == nullptr) {
    std::vector<Affine> host(kSkipTabSize);
    for (int p = 0; p < kSkipTabBytes; p++) {
      Affine unit = affine_skip_loop(1ULL << (8 * p));  // skip by 256^p
      Affine acc{1ULL, 0ULL};
      for (int v = 0; v < 256; v++) {
        host[p * 256 + v] = acc;
        acc               = affine_then(acc, unit);
      }
    }
    Affine* d = nullptr;
    WGB_CUDA_TRY(cudaMalloc(&d, sizeof(Affine) * kSkipTabSize));
    WGB_CUDA_TRY(cudaMemcpy(d, host.data(), sizeof(Affine) * kSkipTabSize, cudaMemcpyHostToDevice));
    tabs[dev] = d;
  }
  return tabs[dev];
}

// ---- host launchers ---------------------------------------------------------------------------------
struct SampleArgs {
  ChunkRef row_ptr;
  unsigned long long row_ptr_off;  // elements
  ChunkRef col;
  unsigned long long col_off;
  wholememory_dtype_t col_dtype;
  ChunkRef wgt;
  unsigned long long wgt_off;
  wholememory_dtype_t wgt_dtype;
  bool weighted;
  const void* centers;
  wholememory_dtype_t center_dtype;
  int n;
  int M;
  unsigned long long seed;
  int* offsets;  // n + 1, device
  void* out_dest;
  int* out_lid;
  long long* out_gid;
  cudaStream_t stream;
};

template <typename IdT>
static void launch_count_scan(const SampleArgs& a, void* tile_state_mem)
{
  int tiles    = (a.n + kScanTile) / kScanTile;  // covers n + 1 outputs
  bool chunked = a.row_ptr.world > 1;
  WGB_CUDA_TRY(cudaMemsetAsync(tile_state_mem, 0, scan_state_bytes(tiles), a.stream));
  auto* state  = static_cast<unsigned long long*>(tile_state_mem);
  auto* ticket = reinterpret_cast<unsigned int*>(state + tiles);
  if (chunked)
    count_scan_kernel<IdT, true><<<tiles, kScanBlock, 0, a.stream>>>(a.row_ptr, a.row_ptr_off, static_cast<const IdT*>(a.centers), a.n, a.M, a.offsets, state, ticket);
  else
    count_scan_kernel<IdT, false><<<tiles, kScanBlock, 0, a.stream>>>(a.row_ptr, a.row_ptr_off, static_cast<const IdT*>(a.centers), a.n, a.M, a.offsets, state, ticket);
  WGB_CHECK_LAUNCH();
}

template <typename IdT, typename ColT, bool CHUNKED>
static void launch_uniform(const SampleArgs& a)
{
  const IdT* centers = static_cast<const IdT*>(a.centers);
  ColT* out          = static_cast<ColT*>(a.out_dest);
  int sms            = num_sms();
  if (a.M <= 0) {
    int warps_needed = a.n;
    int grid         = std::max(1, std::min((warps_needed + 7) / 8, sms * 8));
    sample_all_kernel<IdT, ColT, CHUNKED><<<grid, 256, 0, a.stream>>>(a.row_ptr, a.row_ptr_off, a.col, a.col_off, centers, a.n, a.offsets, out, a.out_lid, a.out_gid);
  } else if (a.M <= 32) {
    const Affine* tab = skip_table_device();
    int batches       = (a.n + 31) / 32;
    int grid          = std::max(1, std::min((batches + 7) / 8, sms * 8));
    if (a.M <= 8)
      uniform_small_kernel<IdT, ColT, 8, CHUNKED><<<grid, 256, 0, a.stream>>>(a.row_ptr, a.row_ptr_off, a.col, a.col_off, centers, a.n, a.M, a.seed, a.offsets, out, a.out_lid, a.out_gid, tab);
    else if (a.M <= 16)
      uniform_small_kernel<IdT, ColT, 16, CHUNKED><<<grid, 256, 0, a.stream>>>(a.row_ptr, a.row_ptr_off, a.col, a.col_off, centers, a.n, a.M, a.seed, a.offsets, out, a.out_lid, a.out_gid, tab);
    else
      uniform_small_kernel<IdT, ColT, 32, CHUNKED><<<grid, 256, 0, a.stream>>>(a.row_ptr, a.row_ptr_off, a.col, a.col_off, centers, a.n, a.M, a.seed, a.offsets, out, a.out_lid, a.out_gid, tab);
  } else {
    const Affine* tab = skip_table_device();
    int grid          = std::max(1, std::min(a.n, sms * 16));
    uniform_general_kernel<IdT, ColT, CHUNKED><<<grid, kGeneralBlock, 0, a.stream>>>(a.row_ptr, a.row_ptr_off, a.col, a.col_off, centers, a.n, a.M, a.seed, a.offsets, out, a.out_lid, a.out_gid, tab);
  }
  WGB_CHECK_LAUNCH();
}

template <typename IdT, typename ColT, typename WT, bool CHUNKED>
static void launch_weighted(const SampleArgs& a)
{
  const IdT* centers = static_cast<const IdT*>(a.centers);
  ColT* out          = static_cast<ColT*>(a.out_dest);
  int sms            = num_sms();
  if (a.M <= 0) {
    int grid = std::max(1, std::min((a.n + 7) / 8, sms * 8));
    sample_all_kernel<IdT, ColT, CHUNKED><<<grid, 256, 0, a.stream>>>(a.row_ptr, a.row_ptr_off, a.col, a.col_off, centers, a.n, a.offsets, out, a.out_lid, a.out_gid);
  } else {
    const Affine* tab = skip_table_device();
    int grid          = std::max(1, std::min(a.n, sms * 8));
    if (a.M <= 256)
      weighted_kernel<IdT, ColT, WT, 128, CHUNKED><<<grid, 128, 0, a.stream>>>(a.row_ptr, a.row_ptr_off, a.col, a.col_off, a.wgt, a.wgt_off, centers, a.n, a.M, a.seed, a.offsets, out, a.out_lid, a.out_gid, tab);
    else
      weighted_kernel<IdT, ColT, WT, 256, CHUNKED><<<grid, 256, 0, a.stream>>>(a.row_ptr, a.row_ptr_off, a.col, a.col_off, a.wgt, a.wgt_off, centers, a.n, a.M, a.seed, a.offsets, out, a.out_lid, a.out_gid, tab);
  }
  WGB_CHECK_LAUNCH();
}

template <typename IdT, typename ColT>
static void launch_sample(const SampleArgs& a)
{
  bool chunked = a.row_ptr.world > 1 || a.col.world > 1 || (a.weighted && a.wgt.world > 1);
  if (!a.weighted) {
    if (chunked) launch_uniform<IdT, ColT, true>(a);
    else launch_uniform<IdT, ColT, false>(a);
  } else if (a.wgt_dtype == WHOLEMEMORY_DT_FLOAT) {
    if (chunked) launch_weighted<IdT, ColT, float, true>(a);
    else launch_weighted<IdT, ColT, float, false>(a);
  } else {
    if (chunked) launch_weighted<IdT, ColT, double, true>(a);
    else launch_weighted<IdT, ColT, double, false>(a);
  }
}

static wholememory_error_code_t sample_op(wholememory_tensor_t row_ptr_t, wholememory_tensor_t col_t,
                                          wholememory_tensor_t wgt_t, bool weighted,
                                          wholememory_tensor_t centers_t, int max_sample_count,
                                          wholememory_tensor_t offsets_t, void* dest_ctx, void* lid_ctx,
                                          void* gid_ctx, unsigned long long seed, wholememory_env_func_t* env,
                                          void* stream)
{
  const char* what = weighted ? "wholegraph_csr_weighted_sample_without_replacement"
                              : "wholegraph_csr_unweighted_sample_without_replacement";
  if (!row_ptr_t || !col_t || !centers_t || !offsets_t || !env || (weighted && !wgt_t)) return WHOLEMEMORY_INVALID_INPUT;
  // argument checks mirror cpp/src/wholegraph_ops/unweighted_sample_without_replacement.cpp:24-124
  auto* rd = wholememory_tensor_get_tensor_description(row_ptr_t);
  auto* cd = wholememory_tensor_get_tensor_description(col_t);
  auto* nd = wholememory_tensor_get_tensor_description(centers_t);
  auto* od = wholememory_tensor_get_tensor_description(offsets_t);
  if (rd->dim != 1 || cd->dim != 1 || nd->dim != 1 || od->dim != 1) {
    log_msg(LEVEL_ERROR, "%s: all tensors must be 1-D", what);
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (rd->dtype != WHOLEMEMORY_DT_INT64) {
    log_msg(LEVEL_ERROR, "%s: csr_row_ptr must be int64", what);
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (cd->dtype != WHOLEMEMORY_DT_INT && cd->dtype != WHOLEMEMORY_DT_INT64) return WHOLEMEMORY_INVALID_INPUT;
  if (nd->dtype != WHOLEMEMORY_DT_INT && nd->dtype != WHOLEMEMORY_DT_INT64) return WHOLEMEMORY_INVALID_INPUT;
  if (od->dtype != WHOLEMEMORY_DT_INT) {
    log_msg(LEVEL_ERROR, "%s: output_sample_offset must be int32", what);
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (od->sizes[0] != nd->sizes[0] + 1) {
    log_msg(LEVEL_ERROR, "%s: output_sample_offset size must be center node count + 1", what);
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (weighted) {
    auto* wd = wholememory_tensor_get_tensor_description(wgt_t);
    if (wd->dim != 1 || (wd->dtype != WHOLEMEMORY_DT_FLOAT && wd->dtype != WHOLEMEMORY_DT_DOUBLE)) return WHOLEMEMORY_INVALID_INPUT;
    if (wd->sizes[0] != cd->sizes[0]) return WHOLEMEMORY_INVALID_INPUT;
  }
  if (dest_ctx == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  if (max_sample_count > 1024) {
    // reference: large_sample_kernel / segmented-sort paths (func.cuh:51-113, weighted :530-590); no CPU
    // reference exists for them and no configuration uses fan-outs > 1024.
    log_msg(LEVEL_ERROR, "%s: max_sample_count > 1024 is not supported", what);
    return WHOLEMEMORY_NOT_IMPLEMENTED;
  }
  return guarded(what, [&] {
    wholememory_tensor_t centers_root = wholememory_tensor_get_root(centers_t);
    wholememory_tensor_t offsets_root = wholememory_tensor_get_root(offsets_t);
    WGB_EXPECTS(!centers_root->is_wholememory && !offsets_root->is_wholememory, "center nodes / offsets must be local tensors");
    SampleArgs a;
    a.row_ptr      = make_chunk_ref(row_ptr_t);
    a.row_ptr_off  = (unsigned long long)rd->storage_offset;
    a.col          = make_chunk_ref(col_t);
    a.col_off      = (unsigned long long)cd->storage_offset;
    a.col_dtype    = cd->dtype;
    a.weighted     = weighted;
    a.wgt_off      = 0;
    a.wgt_dtype    = WHOLEMEMORY_DT_FLOAT;
    memset(&a.wgt, 0, sizeof(a.wgt));
    if (weighted) {
      auto* wd    = wholememory_tensor_get_tensor_description(wgt_t);
      a.wgt       = make_chunk_ref(wgt_t);
      a.wgt_off   = (unsigned long long)wd->storage_offset;
      a.wgt_dtype = wd->dtype;
    }
    a.centers      = static_cast<const char*>(centers_root->storage_ptr) + nd->storage_offset * dtype_size(nd->dtype);
    a.center_dtype = nd->dtype;
    WGB_EXPECTS(nd->sizes[0] < (1LL << 31) - kScanTile, "too many center nodes for one call");
    a.n       = (int)nd->sizes[0];
    a.M       = max_sample_count;
    a.seed    = seed;
    a.offsets = reinterpret_cast<int*>(static_cast<char*>(offsets_root->storage_ptr) + od->storage_offset * 4);
    a.stream  = as_stream(stream);

    // 1. counts + exclusive scan, one kernel
    {
      int tiles = (a.n + kScanTile) / kScanTile;
      temp_memory state(env);
      void* state_mem = state.bytes((int64_t)scan_state_bytes(tiles));
      if (nd->dtype == WHOLEMEMORY_DT_INT) launch_count_scan<int32_t>(a, state_mem);
      else launch_count_scan<int64_t>(a, state_mem);
      // 2. the only host sync: the output size
      int total = 0;
      WGB_CUDA_TRY(cudaMemcpyAsync(&total, a.offsets + a.n, sizeof(int), cudaMemcpyDeviceToHost, a.stream));
      WGB_CUDA_TRY(cudaStreamSynchronize(a.stream));
      WGB_EXPECTS(total >= 0, "sample count overflowed int32");
      a.out_dest = output_alloc(env, dest_ctx, total, cd->dtype);
      a.out_lid  = lid_ctx ? static_cast<int*>(output_alloc(env, lid_ctx, total, WHOLEMEMORY_DT_INT)) : nullptr;
      a.out_gid  = gid_ctx ? static_cast<long long*>(output_alloc(env, gid_ctx, total, WHOLEMEMORY_DT_INT64)) : nullptr;
      if (total == 0 || a.n == 0) return;
    }
    // 3. sample
    if (nd->dtype == WHOLEMEMORY_DT_INT) {
      if (cd->dtype == WHOLEMEMORY_DT_INT) launch_sample<int32_t, int32_t>(a);
      else launch_sample<int32_t, int64_t>(a);
    } else {
      if (cd->dtype == WHOLEMEMORY_DT_INT) launch_sample<int64_t, int32_t>(a);
      else launch_sample<int64_t, int64_t>(a);
    }
  });
}

}  // namespace wgb

extern "C" {

wholememory_error_code_t wholegraph_csr_unweighted_sample_without_replacement(
  wholememory_tensor_t wm_csr_row_ptr_tensor, wholememory_tensor_t wm_csr_col_ptr_tensor,
  wholememory_tensor_t center_nodes_tensor, int max_sample_count, wholememory_tensor_t output_sample_offset_tensor,
  void* output_dest_memory_context, void* output_center_localid_memory_context,
  void* output_edge_gid_memory_context, unsigned long long random_seed, wholememory_env_func_t* p_env_fns,
  void* stream)
{
  return wgb::sample_op(wm_csr_row_ptr_tensor, wm_csr_col_ptr_tensor, nullptr, false, center_nodes_tensor,
                        max_sample_count, output_sample_offset_tensor, output_dest_memory_context,
                        output_center_localid_memory_context, output_edge_gid_memory_context, random_seed,
                        p_env_fns, stream);
}

wholememory_error_code_t wholegraph_csr_weighted_sample_without_replacement(
  wholememory_tensor_t wm_csr_row_ptr_tensor, wholememory_tensor_t wm_csr_col_ptr_tensor,
  wholememory_tensor_t wm_csr_weight_ptr_tensor, wholememory_tensor_t center_nodes_tensor, int max_sample_count,
  wholememory_tensor_t output_sample_offset_tensor, void* output_dest_memory_context,
  void* output_center_localid_memory_context, void* output_edge_gid_memory_context,
  unsigned long long random_seed, wholememory_env_func_t* p_env_fns, void* stream)
{
  return wgb::sample_op(wm_csr_row_ptr_tensor, wm_csr_col_ptr_tensor, wm_csr_weight_ptr_tensor, true,
                        center_nodes_tensor, max_sample_count, output_sample_offset_tensor,
                        output_dest_memory_context, output_center_localid_memory_context,
                        output_edge_gid_memory_context, random_seed, p_env_fns, stream);
}

// host twins of the device stream (reference: cpp/src/wholegraph_ops/raft_random_gen.cu:15-96)
wholememory_error_code_t generate_random_positive_int_cpu(int64_t random_seed, int64_t subsequence,
                                                          wholememory_tensor_t output)
{
  if (!output) return WHOLEMEMORY_INVALID_INPUT;
  auto d = *wholememory_tensor_get_tensor_description(output);
  if (d.dim != 1) return WHOLEMEMORY_INVALID_INPUT;
  if (d.dtype != WHOLEMEMORY_DT_INT64 && d.dtype != WHOLEMEMORY_DT_INT) return WHOLEMEMORY_INVALID_INPUT;
  void* p = wholememory_tensor_get_data_pointer(output);
  if (!p && d.sizes[0] > 0) return WHOLEMEMORY_INVALID_INPUT;
  wgb::Pcg rng;
  rng.init_loop((unsigned long long)random_seed, (unsigned long long)subsequence);
  for (int64_t i = 0; i < d.sizes[0]; i++) {
    if (d.dtype == WHOLEMEMORY_DT_INT) static_cast<int*>(p)[i] = rng.next_i32();
    else static_cast<int64_t*>(p)[i] = (int64_t)(rng.next_u64() & 0x7fffffffffffffffULL);
  }
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t generate_exponential_distribution_negative_float_cpu(int64_t random_seed,
                                                                              int64_t subsequence,
                                                                              wholememory_tensor_t output)
{
  if (!output) return WHOLEMEMORY_INVALID_INPUT;
  auto d = *wholememory_tensor_get_tensor_description(output);
  if (d.dim != 1 || d.dtype != WHOLEMEMORY_DT_FLOAT) return WHOLEMEMORY_INVALID_INPUT;
  float* p = static_cast<float*>(wholememory_tensor_get_data_pointer(output));
  if (!p && d.sizes[0] > 0) return WHOLEMEMORY_INVALID_INPUT;
  wgb::Pcg rng;
  rng.init_loop((unsigned long long)random_seed, (unsigned long long)subsequence);
  for (int64_t i = 0; i < d.sizes[0]; i++) {
    float u               = rng.next_float();
    u                     = (float)(-(0.5 + 0.5 * (double)u));
    unsigned long long r2 = 0;
    int extra             = -1;
    do {
      r2 = rng.next_u64();
      extra++;
    } while (!r2);
    int lz = 0;
    for (unsigned long long t = r2; !(t >> 63); t <<= 1)
      lz++;
    int one_bit = lz + extra * 64;
    u           = (float)((double)u * pow(2.0, -one_bit));
    p[i]        = (float)(log1p((double)u) / log(2.0));
  }
  return WHOLEMEMORY_SUCCESS;
}

}  // extern "C"
