// S3: graph_append_unique -- hop-to-hop renumbering (sm_100a).
//
// Replaces cpp/src/graph_ops/append_unique_func.cuh:201-341 (InsertKeys / CountBucket / AssignValue /
// ComputeOutputUniqueNeighborAndCount / RetrieveKeys kernels + thrust scan + 3 stream syncs, ids of new
// neighbours in racy hash-slot order).
//
// Here: open-addressing table of 16-byte slots {key, first_position}; every key
// records with one atomicMin the FIRST position at which it occurs (targets occupy positions 0..T-1,
// neighbour k position T+k).  A neighbour is "new" iff its slot's first position is its own, so a
// single-pass flag+scan over the neighbours yields ids in first-occurrence order -- deterministic and
// equal to the reference's host algorithm (cpp/tests/graph_ops/append_unique_test_utils.cu:53-119).
// 3 kernels + 1 memset, one host sync (to size the output).

#include "wm_common.cuh"
#include "sample_device.cuh"

namespace wgb {

struct UniqSlot {
  long long key;
  unsigned long long first_pos;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

constexpr long long kEmptyKey = -1LL;

__device__ __forceinline__ unsigned int uniq_insert(UniqSlot* table, unsigned int mask, long long key, unsigned long long pos)
{
  unsigned int slot = (unsigned int)mix64((unsigned long long)key) & mask;
  while (true) {
    long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(&table[slot].key),
                                          (unsigned long long)kEmptyKey, (unsigned long long)key);
    if (prev == kEmptyKey || prev == key) break;
    slot = (slot + 1) & mask;
  }
  atomicMin(&table[slot].first_pos, pos);
  return slot;
}

template <typename KeyT>
__global__ void __launch_bounds__(256) uniq_insert_kernel(UniqSlot* table, unsigned int mask, const KeyT* __restrict__ targets,
                                                          int T, const KeyT* __restrict__ neighbors, int Nn,
                                                          unsigned int* __restrict__ slot_of)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)T + Nn; i += (long long)gridDim.x * blockDim.x) {
    if (i < T) {
      uniq_insert(table, mask, (long long)targets[i], (unsigned long long)i);
    } else {
      int k      = (int)(i - T);
      slot_of[k] = uniq_insert(table, mask, (long long)neighbors[k], (unsigned long long)i);
    }
  }
}

// rank[k] = number of first-occurrence new neighbours before k; rank[Nn] = total
__global__ void __launch_bounds__(kScanBlock) uniq_rank_kernel(const UniqSlot* __restrict__ table, const unsigned int* __restrict__ slot_of,
                                                               int T, int Nn, int* __restrict__ rank,
                                                               unsigned long long* state, unsigned int* ticket)
{
  const int tile       = blockIdx.x;
  const long long base = (long long)tile * kScanTile + (long long)threadIdx.x * kScanItems;
  unsigned int v[kScanItems];
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    long long i = base + k;
    v[k]        = (i < Nn && table[slot_of[i]].first_pos == (unsigned long long)(T + i)) ? 1u : 0u;
  }
  unsigned int agg          = block_scan_items(v);
  unsigned long long prefix = scan_tile_prefix(state, tile, agg);
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    long long i = base + k;
    if (i <= Nn) rank[i] = (int)(prefix + v[k]);
  }
}

template <typename KeyT>
__global__ void __launch_bounds__(256) uniq_finalize_kernel(const UniqSlot* __restrict__ table, const unsigned int* __restrict__ slot_of,
                                                            const KeyT* __restrict__ targets, int T,
                                                            const KeyT* __restrict__ neighbors, int Nn,
                                                            const int* __restrict__ rank, KeyT* __restrict__ unique_out,
                                                            int* __restrict__ raw_to_unique)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)T + Nn; i += (long long)gridDim.x * blockDim.x) {
    if (i < T) {
      unique_out[i] = targets[i];
    } else {
      int k                  = (int)(i - T);
      unsigned long long pos = table[slot_of[k]].first_pos;
      int id                 = pos < (unsigned long long)T ? (int)pos : T + rank[pos - T];
      if (raw_to_unique) raw_to_unique[k] = id;
      if (pos == (unsigned long long)i) unique_out[id] = neighbors[k];
    }
  }
}

template <typename KeyT>
static void append_unique_typed(const KeyT* targets, int T, const KeyT* neighbors, int Nn, void* unique_ctx,
                                int* raw_to_unique, wholememory_dtype_t dt, wholememory_env_func_t* env,
                                cudaStream_t stream)
{
  long long total = (long long)T + Nn;
  unsigned int slots = 1024;
  while ((long long)slots < 2 * total)
    slots <<= 1;
  temp_memory table_mem(env), slot_mem(env), rank_mem(env), state_mem(env);
  auto* table   = static_cast<UniqSlot*>(table_mem.bytes((int64_t)slots * sizeof(UniqSlot)));
  auto* slot_of = static_cast<unsigned int*>(slot_mem.device(std::max(Nn, 1), WHOLEMEMORY_DT_INT));
  int* rank     = static_cast<int*>(rank_mem.device(Nn + 1, WHOLEMEMORY_DT_INT));
  int tiles     = (Nn + kScanTile) / kScanTile;
  void* st      = state_mem.bytes((int64_t)scan_state_bytes(tiles));
  WGB_CUDA_TRY(cudaMemsetAsync(table, 0xFF, (size_t)slots * sizeof(UniqSlot), stream));
  WGB_CUDA_TRY(cudaMemsetAsync(st, 0, scan_state_bytes(tiles), stream));
  int sms  = num_sms();
  int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sms * 8));
  if (total > 0) {
    uniq_insert_kernel<KeyT><<<grid, 256, 0, stream>>>(table, slots - 1, targets, T, neighbors, Nn, slot_of);
    WGB_CHECK_LAUNCH();
  }
  auto* state  = static_cast<unsigned long long*>(st);
  auto* ticket = reinterpret_cast<unsigned int*>(state + tiles);
  uniq_rank_kernel<<<tiles, kScanBlock, 0, stream>>>(table, slot_of, T, Nn, rank, state, ticket);
  WGB_CHECK_LAUNCH();
  int new_count = 0;
  WGB_CUDA_TRY(cudaMemcpyAsync(&new_count, rank + Nn, sizeof(int), cudaMemcpyDeviceToHost, stream));
  WGB_CUDA_TRY(cudaStreamSynchronize(stream));
  auto* unique_out = static_cast<KeyT*>(output_alloc(env, unique_ctx, (int64_t)T + new_count, dt));
  if (total > 0) {
    uniq_finalize_kernel<KeyT><<<grid, 256, 0, stream>>>(table, slot_of, targets, T, neighbors, Nn, rank, unique_out, raw_to_unique);
    WGB_CHECK_LAUNCH();
  }
  // temp buffers are released by the callbacks when this scope ends; with stream-ordered or
  // caching allocators (torch) that is safe without another sync.
}

}  // namespace wgb

extern "C" wholememory_error_code_t graph_append_unique(wholememory_tensor_t target_nodes_tensor,
                                                        wholememory_tensor_t neighbor_nodes_tensor,
                                                        void* output_unique_node_memory_context,
                                                        wholememory_tensor_t output_neighbor_raw_to_unique_mapping_tensor,
                                                        wholememory_env_func_t* p_env_fns, void* stream)
{
  using namespace wgb;
  // argument checks mirror cpp/src/graph_ops/append_unique.cpp
  if (!target_nodes_tensor || !neighbor_nodes_tensor || !output_unique_node_memory_context || !p_env_fns) return WHOLEMEMORY_INVALID_INPUT;
  auto* td = wholememory_tensor_get_tensor_description(target_nodes_tensor);
  auto* nd = wholememory_tensor_get_tensor_description(neighbor_nodes_tensor);
  if (td->dim != 1 || nd->dim != 1) {
    log_msg(LEVEL_ERROR, "graph_append_unique: target and neighbor tensors must be 1-D");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (td->dtype != nd->dtype || (td->dtype != WHOLEMEMORY_DT_INT && td->dtype != WHOLEMEMORY_DT_INT64)) {
    log_msg(LEVEL_ERROR, "graph_append_unique: target and neighbor must share dtype int32 or int64");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  int* r2u = nullptr;
  if (output_neighbor_raw_to_unique_mapping_tensor) {
    auto* md = wholememory_tensor_get_tensor_description(output_neighbor_raw_to_unique_mapping_tensor);
    if (md->dim != 1 || md->dtype != WHOLEMEMORY_DT_INT || md->sizes[0] != nd->sizes[0]) {
      log_msg(LEVEL_ERROR, "graph_append_unique: raw_to_unique mapping must be int32[neighbor count]");
      return WHOLEMEMORY_INVALID_INPUT;
    }
    r2u = static_cast<int*>(wholememory_tensor_get_data_pointer(output_neighbor_raw_to_unique_mapping_tensor));
  }
  return guarded("graph_append_unique", [&] {
    WGB_EXPECTS(td->sizes[0] + nd->sizes[0] < (1LL << 30), "too many nodes for one append_unique call");
    void* tp = wholememory_tensor_get_data_pointer(target_nodes_tensor);
    void* np = wholememory_tensor_get_data_pointer(neighbor_nodes_tensor);
    if (td->dtype == WHOLEMEMORY_DT_INT)
      append_unique_typed<int32_t>(static_cast<int32_t*>(tp), (int)td->sizes[0], static_cast<int32_t*>(np), (int)nd->sizes[0],
                                   output_unique_node_memory_context, r2u, td->dtype, p_env_fns, as_stream(stream));
    else
      append_unique_typed<int64_t>(static_cast<int64_t*>(tp), (int)td->sizes[0], static_cast<int64_t*>(np), (int)nd->sizes[0],
                                   output_unique_node_memory_context, r2u, td->dtype, p_env_fns, as_stream(stream));
  });
}
