// Trainable WholeMemory embeddings: sparse gradient apply with SGD / LazyAdam / AdaGrad / RMSProp (sm_100a).
//
// Replaces, behind the same C entry points,
//   wholememory_embedding_gather_gradient_apply      cpp/src/wholememory/embedding.cpp:136-315
//   optimizer steps                                   cpp/src/wholememory_ops/functions/embedding_optimizer_func.cu:169-1024
//   optimizer objects / states                        cpp/src/wholememory/embedding_optimizer.cpp:54-527
// of the reference, whose multi-GPU path is: bucket ids by owner -> NCCL all-to-all of ids -> NCCL all-to-all of
// gradient rows -> dedup (sort + segmented sum) -> optimizer step on the local rows.
//
// B200-first: no collective on the data path.  Every rank PUBLISHES its (index, gradient row) pairs into its own
// chunk of a peer-mapped mailbox (a local device-to-device copy); after one barrier every owner PULLS what it owns:
//   select   ordered compaction over all ranks' indices (P2P reads of 8 B per entry) of those in [lo, hi)
//   sort     stable radix sort of (index, mailbox position)   -> duplicates of a row are adjacent, in publish order
//   apply    one warp per distinct row: sums the duplicate gradient rows straight out of the publishers' mailboxes
//            (NVLink loads, like the feature gather) in a fixed order and applies the optimizer to the local row
// Gradient bytes cross NVLink exactly once, are never staged twice, and the result is deterministic.
//
// Update rules are the reference's, element for element (embedding_optimizer_func.cu:205-213, 389-420, 655-668, 865-880).

#include "wm_common.cuh"
#include "sample_device.cuh"

#include <wholememory/embedding.h>

#include <cub/device/device_radix_sort.cuh>

#include <cstring>

struct wholememory_embedding_optimizer_ {
  wholememory_optimizer_type_t type = WHOLEMEMORY_OPT_NONE;
  float weight_decay = 0.0f;
  float epsilon      = 1e-8f;
  float beta1        = 0.9f;
  float beta2        = 0.999f;
  float adam_w       = 0.0f;
  float alpha        = 0.99f;
};

namespace wgb {

struct OptParams {
  int type;
  float weight_decay, epsilon, beta1, beta2, alpha, lr;
  bool adam_w;
};

template <typename T>
__device__ __forceinline__ float emb_load(const T* p)
{
  return static_cast<float>(*p);
}
template <>
__device__ __forceinline__ float emb_load<__half>(const __half* p)
{
  return __half2float(*p);
}
template <>
__device__ __forceinline__ float emb_load<__nv_bfloat16>(const __nv_bfloat16* p)
{
  return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void emb_store(T* p, float v)
{
  *p = static_cast<T>(v);
}
template <>
__device__ __forceinline__ void emb_store<__half>(__half* p, float v)
{
  *p = __float2half_rn(v);
}
template <>
__device__ __forceinline__ void emb_store<__nv_bfloat16>(__nv_bfloat16* p, float v)
{
  *p = __float2bfloat16_rn(v);
}

// own (index, gradient row) pairs -> own chunk of the mailboxes
template <typename IdxT>
__global__ void __launch_bounds__(256) publish_kernel(const IdxT* __restrict__ indices, const float* __restrict__ grads,
                                                      long long grad_stride, long long n, int dim,
                                                      long long* __restrict__ box_idx, float* __restrict__ box_grad)
{
  const long long total = n * (long long)dim;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / dim;
    const int d       = (int)(i - r * dim);
    box_grad[i]       = grads[r * grad_stride + d];
    if (d == 0) box_idx[r] = (long long)indices[r];
  }
}

struct RankCounts {
  long long n[kMaxWorld];
};

// ordered compaction of the mailbox entries this rank owns: (index, mailbox position)
template <bool CHUNKED>
__global__ void __launch_bounds__(kScanBlock) select_owned_kernel(ChunkRef box_idx, RankCounts counts, long long cap, int world,
                                                                  long long lo, long long hi, long long* __restrict__ out_idx,
                                                                  long long* __restrict__ out_pos, long long* __restrict__ out_count,
                                                                  unsigned long long* state, unsigned int* ticket)
{
  const long long total = cap * world;
  while (true) {
    const int tile = take_ticket(ticket);
    if ((long long)tile * kScanTile > total) return;
    const long long base = (long long)tile * kScanTile + (long long)threadIdx.x * kScanItems;
    unsigned int v[kScanItems];
    long long idx[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      const long long p = base + k;
      bool own          = false;
      idx[k]            = -1;
      if (p < total) {
        const int r       = (int)(p / cap);
        const long long j = p - (long long)r * cap;
        if (j < counts.n[r]) {
          idx[k] = load_i64<CHUNKED>(box_idx, (unsigned long long)p);
          own    = idx[k] >= lo && idx[k] < hi;
        }
      }
      v[k] = own ? 1u : 0u;
    }
    unsigned int flags = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
      flags |= v[k] << k;
    unsigned int agg          = block_scan_items(v);
    unsigned long long prefix = scan_tile_prefix(state, tile, agg);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      const long long p = base + k;
      if ((flags >> k) & 1u) {
        out_idx[prefix + v[k]] = idx[k];
        out_pos[prefix + v[k]] = p;
      }
      if (p == total) *out_count = (long long)(prefix + v[k]);
    }
  }
}

struct StatePtrs {
  float* a;        // m | state_sum | v
  float* b;        // v (adam)
  float* per_row;  // beta12t [rows, 2] (adam)
};

// one warp per sorted entry; only the first entry of a run of equal indices works
template <typename EmbT, bool CHUNKED>
__global__ void __launch_bounds__(256) apply_kernel(const long long* __restrict__ sorted_idx, const long long* __restrict__ sorted_pos,
                                                    const long long* __restrict__ m_dev, ChunkRef box_grad, int dim, long long lo,
                                                    EmbT* __restrict__ emb, long long emb_stride, StatePtrs st, OptParams op)
{
  const long long m     = *m_dev;
  const int lane        = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < m; i += warps) {
    const long long idx = sorted_idx[i];
    if (i > 0 && sorted_idx[i - 1] == idx) continue;
    long long run = 1;
    while (i + run < m && sorted_idx[i + run] == idx)
      run++;
    const long long row = idx - lo;
    EmbT* w             = emb + row * emb_stride;
    float b1t = 1.f, b2t = 1.f;
    if (op.type == WHOLEMEMORY_OPT_LAZY_ADAM) {
      b1t = st.per_row[row * 2] * op.beta1;
      b2t = st.per_row[row * 2 + 1] * op.beta2;
    }
    for (int d = lane; d < dim; d += 32) {
      float g = 0.f;
      for (long long k = 0; k < run; k++) {  // publish order: deterministic sum
        const unsigned long long off = ((unsigned long long)sorted_pos[i + k] * (unsigned long long)dim + (unsigned long long)d) * 4ULL;
        g += __ldg(reinterpret_cast<const float*>(box_grad.at<CHUNKED>(off)));
      }
      float x = emb_load<EmbT>(w + d);
      const long long e = row * (long long)dim + d;
      if (op.type == WHOLEMEMORY_OPT_SGD) {
        g += op.weight_decay * x;
        x -= op.lr * g;
      } else if (op.type == WHOLEMEMORY_OPT_LAZY_ADAM) {
        if (op.adam_w) x -= op.lr * op.weight_decay * x;
        else g = g + op.weight_decay * x;
        float mm = st.a[e], vv = st.b[e];
        mm       = op.beta1 * mm + (1 - op.beta1) * g;
        vv       = op.beta2 * vv + (1 - op.beta2) * g * g;
        float mhat = mm / (1 - b1t);
        float vhat = vv / (1 - b2t);
        x          = x - op.lr * mhat / (sqrtf(vhat) + op.epsilon);
        st.a[e]    = mm;
        st.b[e]    = vv;
      } else if (op.type == WHOLEMEMORY_OPT_ADAGRAD) {
        g        = g + op.weight_decay * x;
        float s  = st.a[e];
        s        = s + g * g;
        x        = x - op.lr * g / (sqrtf(s) + op.epsilon);
        st.a[e]  = s;
      } else {  // RMSProp
        g        = g + op.weight_decay * x;
        float vv = st.a[e];
        vv       = op.alpha * vv + (1 - op.alpha) * g * g;
        x        = x - op.lr * g / (sqrtf(vv) + op.epsilon);
        st.a[e]  = vv;
      }
      emb_store<EmbT>(w + d, x);
    }
    if (op.type == WHOLEMEMORY_OPT_LAZY_ADAM) {
      __syncwarp();
      if (lane == 0) {
        st.per_row[row * 2]     = b1t;
        st.per_row[row * 2 + 1] = b2t;
      }
    }
  }
}

__global__ void fill_kernel(float* p, float v, long long n)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

static void* scratch(wholememory_embedding_t e, int slot, size_t bytes)
{
  if (bytes < 256) bytes = 256;
  if (e->scratch_bytes[slot] < bytes) {
    if (e->scratch[slot]) WGB_CUDA_TRY(cudaFree(e->scratch[slot]));
    e->scratch[slot]       = nullptr;
    e->scratch_bytes[slot] = 0;
    WGB_CUDA_TRY(cudaMalloc(&e->scratch[slot], bytes + bytes / 4));
    e->scratch_bytes[slot] = bytes + bytes / 4;
  }
  return e->scratch[slot];
}

static wholememory_comm_t comm_of(wholememory_tensor_t t)
{
  wholememory_comm_t c = nullptr;
  wholememory_handle_t h = wholememory_tensor_get_memory_handle(t);
  WGB_EXPECTS(h != nullptr, "embedding tensor has no WholeMemory handle");
  WGB_EXPECTS(wholememory_get_communicator(&c, h) == WHOLEMEMORY_SUCCESS, "no communicator");
  return c;
}

static wholememory_tensor_t make_like(wholememory_tensor_t emb_tensor, int64_t cols, wholememory_dtype_t dt, const std::vector<size_t>* part)
{
  auto* ed = wholememory_tensor_get_tensor_description(emb_tensor);
  wholememory_handle_t h = wholememory_tensor_get_memory_handle(emb_tensor);
  wholememory_tensor_description_t d;
  wholememory_initialize_tensor_desc(&d);
  d.dim            = 2;
  d.dtype          = dt;
  d.sizes[0]       = ed->sizes[0];
  d.sizes[1]       = cols;
  d.strides[0]     = cols;
  d.strides[1]     = 1;
  d.storage_offset = 0;
  wholememory_tensor_t t = nullptr;
  std::vector<size_t> p;
  if (part) p = *part;
  auto err = wholememory_create_tensor(&t, &d, comm_of(emb_tensor), wholememory_get_memory_type(h), wholememory_get_memory_location(h),
                                       part ? p.data() : nullptr);
  WGB_EXPECTS(err == WHOLEMEMORY_SUCCESS, "could not allocate an optimizer state tensor");
  return t;
}

static void fill_local(wholememory_tensor_t t, float v)
{
  wholememory_tensor_t local = nullptr;
  WGB_EXPECTS(wholememory_tensor_map_local_tensor(t, &local) == WHOLEMEMORY_SUCCESS, "map_local_tensor failed");
  auto* d = wholememory_tensor_get_tensor_description(local);
  long long n = d->sizes[0] * d->sizes[1];
  if (n > 0) {
    fill_kernel<<<(int)std::min<long long>((n + 255) / 256, 1184), 256>>>(static_cast<float*>(wholememory_tensor_get_data_pointer(local)), v, n);
    WGB_CHECK_LAUNCH();
  }
  wholememory_destroy_tensor(local);
  WGB_CUDA_TRY(cudaDeviceSynchronize());
}

void embedding_release_training_state(wholememory_embedding_t e)
{
  for (auto& s : e->states)
    if (s.second) wholememory_destroy_tensor(s.second);
  e->states.clear();
  e->state_names.clear();
  if (e->inbox_idx) wholememory_destroy_tensor(e->inbox_idx);
  if (e->inbox_grad) wholememory_destroy_tensor(e->inbox_grad);
  e->inbox_idx = e->inbox_grad = nullptr;
  e->inbox_cap                 = 0;
  for (int i = 0; i < 6; i++) {
    if (e->scratch[i]) cudaFree(e->scratch[i]);
    e->scratch[i]       = nullptr;
    e->scratch_bytes[i] = 0;
  }
  cudaGetLastError();
}

static std::vector<size_t> entry_partition(wholememory_tensor_t t, int world)
{
  std::vector<size_t> offs(world + 1);
  WGB_EXPECTS(wholememory_tensor_get_entry_offsets(offs.data(), t) == WHOLEMEMORY_SUCCESS, "entry offsets");
  std::vector<size_t> part(world);
  for (int r = 0; r < world; r++)
    part[r] = offs[r + 1] - offs[r];
  return part;
}

template <typename EmbT>
static void launch_apply(bool chunked, int grid, cudaStream_t st, const long long* si, const long long* sp, const long long* m_dev,
                         const ChunkRef& grad_ref, int dim, long long lo, void* emb, long long stride, StatePtrs sps, OptParams op)
{
  if (chunked) apply_kernel<EmbT, true><<<grid, 256, 0, st>>>(si, sp, m_dev, grad_ref, dim, lo, static_cast<EmbT*>(emb), stride, sps, op);
  else apply_kernel<EmbT, false><<<grid, 256, 0, st>>>(si, sp, m_dev, grad_ref, dim, lo, static_cast<EmbT*>(emb), stride, sps, op);
  WGB_CHECK_LAUNCH();
}

static void gradient_apply(wholememory_embedding_t e, wholememory_tensor_t indices, wholememory_tensor_t grads, float lr, cudaStream_t st)
{
  WGB_EXPECTS(e->optimizer != nullptr && e->optimizer->type != WHOLEMEMORY_OPT_NONE, "the embedding has no optimizer");
  embedding_drop_hot_rows(e);  // a replica of rows that are about to change would go stale
  auto* ed = wholememory_tensor_get_tensor_description(e->tensor);
  auto* id = wholememory_tensor_get_tensor_description(indices);
  auto* gd = wholememory_tensor_get_tensor_description(grads);
  WGB_CHECK_INPUT(id->dim == 1 && (id->dtype == WHOLEMEMORY_DT_INT || id->dtype == WHOLEMEMORY_DT_INT64), "indices must be 1-D int32/int64");
  WGB_CHECK_INPUT(gd->dim == 2 && gd->dtype == WHOLEMEMORY_DT_FLOAT && gd->sizes[0] == id->sizes[0] && gd->sizes[1] == ed->sizes[1] && gd->strides[1] == 1,
                  "grads must be fp32 [len(indices), embedding dim]");
  wholememory_comm_t comm = comm_of(e->tensor);
  const int world = comm->size, rank = comm->rank;
  const int dim   = (int)ed->sizes[1];
  const long long n = id->sizes[0];
  const int sms     = num_sms();

  // ---- counts of every rank, mailbox capacity (collective) -------------------------------------------------
  RankCounts counts;
  memset(&counts, 0, sizeof(counts));
  long long all_n[kMaxWorld];
  comm_allgather(comm, &n, all_n, sizeof(long long));
  long long most = 1;
  for (int r = 0; r < world; r++) {
    counts.n[r] = all_n[r];
    most        = std::max(most, all_n[r]);
  }
  if ((size_t)most > e->inbox_cap) {
    if (e->inbox_idx) wholememory_destroy_tensor(e->inbox_idx);
    if (e->inbox_grad) wholememory_destroy_tensor(e->inbox_grad);
    size_t cap = 1024;
    while (cap < (size_t)most)
      cap *= 2;
    wholememory_handle_t h = wholememory_tensor_get_memory_handle(e->tensor);
    std::vector<size_t> part(world, cap);
    wholememory_tensor_description_t d;
    wholememory_initialize_tensor_desc(&d);
    d.dim = 1; d.dtype = WHOLEMEMORY_DT_INT64; d.sizes[0] = (int64_t)(cap * world); d.strides[0] = 1; d.storage_offset = 0;
    WGB_EXPECTS(wholememory_create_tensor(&e->inbox_idx, &d, comm, wholememory_get_memory_type(h), WHOLEMEMORY_ML_DEVICE, part.data()) == WHOLEMEMORY_SUCCESS, "mailbox");
    d.dim = 2; d.dtype = WHOLEMEMORY_DT_FLOAT; d.sizes[0] = (int64_t)(cap * world); d.sizes[1] = dim; d.strides[0] = dim; d.strides[1] = 1;
    WGB_EXPECTS(wholememory_create_tensor(&e->inbox_grad, &d, comm, wholememory_get_memory_type(h), WHOLEMEMORY_ML_DEVICE, part.data()) == WHOLEMEMORY_SUCCESS, "mailbox");
    e->inbox_cap = cap;
  }
  const long long cap = (long long)e->inbox_cap;

  // ---- publish -------------------------------------------------------------------------------------------------
  wholememory_tensor_t li = nullptr, lg = nullptr;
  WGB_EXPECTS(wholememory_tensor_map_local_tensor(e->inbox_idx, &li) == WHOLEMEMORY_SUCCESS, "map mailbox");
  WGB_EXPECTS(wholememory_tensor_map_local_tensor(e->inbox_grad, &lg) == WHOLEMEMORY_SUCCESS, "map mailbox");
  long long* box_idx = static_cast<long long*>(wholememory_tensor_get_data_pointer(li));
  float* box_grad    = static_cast<float*>(wholememory_tensor_get_data_pointer(lg));
  wholememory_destroy_tensor(li);
  wholememory_destroy_tensor(lg);
  if (n > 0) {
    int grid = (int)std::min<long long>((n * dim + 255) / 256, (long long)sms * 8);
    const float* g = static_cast<const float*>(wholememory_tensor_get_data_pointer(grads)) + gd->storage_offset;
    if (id->dtype == WHOLEMEMORY_DT_INT)
      publish_kernel<int><<<grid, 256, 0, st>>>(static_cast<const int*>(wholememory_tensor_get_data_pointer(indices)) + id->storage_offset, g, gd->strides[0], n, dim, box_idx, box_grad);
    else
      publish_kernel<long long><<<grid, 256, 0, st>>>(static_cast<const long long*>(wholememory_tensor_get_data_pointer(indices)) + id->storage_offset, g, gd->strides[0], n, dim, box_idx, box_grad);
    WGB_CHECK_LAUNCH();
  }
  if (world > 1) {
    WGB_CUDA_TRY(cudaStreamSynchronize(st));
    comm_barrier(comm);  // every mailbox is complete
  }

  // ---- pull: select what this rank owns, sort, apply ---------------------------------------------------------------
  std::vector<size_t> offs(world + 1);
  WGB_EXPECTS(wholememory_tensor_get_entry_offsets(offs.data(), e->tensor) == WHOLEMEMORY_SUCCESS, "entry offsets");
  const long long lo = (long long)offs[rank], hi = (long long)offs[rank + 1];
  const long long total = cap * world;
  long long total_published = 0;
  for (int r = 0; r < world; r++)
    total_published += counts.n[r];
  const size_t list_bytes = sizeof(long long) * (size_t)std::max<long long>(total_published, 1);
  long long* own_idx  = static_cast<long long*>(scratch(e, 0, list_bytes));
  long long* own_pos  = static_cast<long long*>(scratch(e, 1, list_bytes));
  long long* sort_idx = static_cast<long long*>(scratch(e, 2, list_bytes));
  long long* sort_pos = static_cast<long long*>(scratch(e, 3, list_bytes));
  const int tiles     = (int)((total + kScanTile) / kScanTile);
  const size_t sbytes = scan_state_bytes(tiles) + 16;
  unsigned long long* state = static_cast<unsigned long long*>(scratch(e, 4, sbytes));
  WGB_CUDA_TRY(cudaMemsetAsync(state, 0, sbytes, st));
  long long* m_dev     = reinterpret_cast<long long*>(state + tiles + 1);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(state + tiles);
  ChunkRef idx_ref  = make_chunk_ref(e->inbox_idx);
  ChunkRef grad_ref = make_chunk_ref(e->inbox_grad);
  const bool chunked = idx_ref.world > 1;
  const int sgrid    = std::min(tiles, sms * 8);
  if (chunked) select_owned_kernel<true><<<sgrid, kScanBlock, 0, st>>>(idx_ref, counts, cap, world, lo, hi, own_idx, own_pos, m_dev, state, ticket);
  else select_owned_kernel<false><<<sgrid, kScanBlock, 0, st>>>(idx_ref, counts, cap, world, lo, hi, own_idx, own_pos, m_dev, state, ticket);
  WGB_CHECK_LAUNCH();
  long long m_host = 0;
  WGB_CUDA_TRY(cudaMemcpyAsync(&m_host, m_dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
  WGB_CUDA_TRY(cudaStreamSynchronize(st));
  if (m_host > 0) {
    int end_bit = 1;
    while (end_bit < 64 && (1ULL << end_bit) <= (unsigned long long)ed->sizes[0])
      end_bit++;
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, own_idx, sort_idx, own_pos, sort_pos, (int)m_host, 0, end_bit, st);
    void* temp = scratch(e, 5, temp_bytes);
    WGB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, own_idx, sort_idx, own_pos, sort_pos, (int)m_host, 0, end_bit, st));
    ++g_kernel_launches;

    wholememory_tensor_t local_emb = nullptr;
    WGB_EXPECTS(wholememory_tensor_map_local_tensor(e->tensor, &local_emb) == WHOLEMEMORY_SUCCESS, "map embedding");
    void* emb_ptr = wholememory_tensor_get_data_pointer(local_emb);
    wholememory_destroy_tensor(local_emb);
    StatePtrs sps{nullptr, nullptr, nullptr};
    auto local_state = [&](const char* name) -> float* {
      for (auto& s : e->states)
        if (s.first == name) {
          wholememory_tensor_t l = nullptr;
          WGB_EXPECTS(wholememory_tensor_map_local_tensor(s.second, &l) == WHOLEMEMORY_SUCCESS, "map state");
          float* p = static_cast<float*>(wholememory_tensor_get_data_pointer(l));
          wholememory_destroy_tensor(l);
          return p;
        }
      throw logic_error(std::string("missing optimizer state ") + name);
    };
    const auto* o = e->optimizer;
    OptParams op{(int)o->type, o->weight_decay, o->epsilon, o->beta1, o->beta2, o->alpha, lr, o->adam_w > 0.5f};
    if (o->type == WHOLEMEMORY_OPT_LAZY_ADAM) {
      sps.a = local_state("m");
      sps.b = local_state("v");
      sps.per_row = local_state("beta12t");
    } else if (o->type == WHOLEMEMORY_OPT_ADAGRAD) {
      sps.a = local_state("state_sum");
    } else if (o->type == WHOLEMEMORY_OPT_RMSPROP) {
      sps.a = local_state("v");
    }
    const int grid = (int)std::min<long long>((m_host + 7) / 8, (long long)sms * 8);
    const bool gchunked = grad_ref.world > 1;
    switch (ed->dtype) {
      case WHOLEMEMORY_DT_FLOAT: launch_apply<float>(gchunked, grid, st, sort_idx, sort_pos, m_dev, grad_ref, dim, lo, emb_ptr, ed->strides[0], sps, op); break;
      case WHOLEMEMORY_DT_HALF: launch_apply<__half>(gchunked, grid, st, sort_idx, sort_pos, m_dev, grad_ref, dim, lo, emb_ptr, ed->strides[0], sps, op); break;
      case WHOLEMEMORY_DT_BF16: launch_apply<__nv_bfloat16>(gchunked, grid, st, sort_idx, sort_pos, m_dev, grad_ref, dim, lo, emb_ptr, ed->strides[0], sps, op); break;
      default: throw invalid_input("trainable embeddings must be fp32, fp16 or bf16");
    }
  }
  if (world > 1) {
    WGB_CUDA_TRY(cudaStreamSynchronize(st));
    comm_barrier(comm);  // nobody overwrites a mailbox that is still being read
  }
}

}  // namespace wgb

extern "C" {

wholememory_error_code_t wholememory_create_embedding_optimizer(wholememory_embedding_optimizer_t* optimizer,
                                                                wholememory_optimizer_type_t optimizer_type)
{
  if (!optimizer) return WHOLEMEMORY_INVALID_INPUT;
  if (optimizer_type < WHOLEMEMORY_OPT_NONE || optimizer_type > WHOLEMEMORY_OPT_ADAGRAD) return WHOLEMEMORY_INVALID_INPUT;
  auto* o     = new wholememory_embedding_optimizer_();
  o->type     = optimizer_type;
  *optimizer  = o;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_optimizer_set_parameter(wholememory_embedding_optimizer_t o, const char* name, void* value)
{
  if (!o || !name || !value) return WHOLEMEMORY_INVALID_INPUT;
  const float v = *static_cast<float*>(value);
  const std::string k(name);
  // parameter sets per optimizer: cpp/src/wholememory/embedding_optimizer.cpp:100-108, 159-179, 286-296, 385-399
  auto has = [&](std::initializer_list<const char*> names) {
    for (auto* n : names)
      if (k == n) return true;
    return false;
  };
  bool ok = false;
  switch (o->type) {
    case WHOLEMEMORY_OPT_SGD: ok = has({"weight_decay"}); break;
    case WHOLEMEMORY_OPT_LAZY_ADAM: ok = has({"weight_decay", "epsilon", "beta1", "beta2", "adam_w"}); break;
    case WHOLEMEMORY_OPT_ADAGRAD: ok = has({"weight_decay", "epsilon"}); break;
    case WHOLEMEMORY_OPT_RMSPROP: ok = has({"weight_decay", "epsilon", "alpha"}); break;
    default: ok = false;
  }
  if (!ok) return WHOLEMEMORY_INVALID_INPUT;
  if (k == "weight_decay") o->weight_decay = v;
  else if (k == "epsilon") o->epsilon = v;
  else if (k == "beta1") o->beta1 = v;
  else if (k == "beta2") o->beta2 = v;
  else if (k == "adam_w") o->adam_w = v;
  else if (k == "alpha") o->alpha = v;
  return WHOLEMEMORY_SUCCESS;
}

void wholememory_destroy_embedding_optimizer(wholememory_embedding_optimizer_t optimizer) { delete optimizer; }

wholememory_error_code_t wholememory_embedding_set_optimizer(wholememory_embedding_t e, wholememory_embedding_optimizer_t o)
{
  using namespace wgb;
  if (!e || !o) return WHOLEMEMORY_INVALID_INPUT;
  if (e->optimizer != nullptr) return WHOLEMEMORY_INVALID_INPUT;  // "optimizer can only be set once"
  return guarded("wholememory_embedding_set_optimizer", [&] {
    auto* ed = wholememory_tensor_get_tensor_description(e->tensor);
    WGB_CHECK_INPUT(ed->dtype == WHOLEMEMORY_DT_FLOAT || ed->dtype == WHOLEMEMORY_DT_HALF || ed->dtype == WHOLEMEMORY_DT_BF16,
                    "trainable embeddings must be fp32, fp16 or bf16");
    wholememory_comm_t comm = comm_of(e->tensor);
    std::vector<size_t> part = entry_partition(e->tensor, comm->size);
    auto add = [&](const char* name, int64_t cols, float init) {
      wholememory_tensor_t t = make_like(e->tensor, cols, WHOLEMEMORY_DT_FLOAT, &part);
      fill_local(t, init);
      e->states.emplace_back(name, t);
    };
    const int64_t dim = ed->sizes[1];
    switch (o->type) {  // state names: embedding_optimizer.cpp:108, 179, 296, 399
      case WHOLEMEMORY_OPT_LAZY_ADAM:
        add("m", dim, 0.f);
        add("v", dim, 0.f);
        add("beta12t", 2, 1.f);
        break;
      case WHOLEMEMORY_OPT_ADAGRAD: add("state_sum", dim, 0.f); break;
      case WHOLEMEMORY_OPT_RMSPROP: add("v", dim, 0.f); break;
      default: break;
    }
    e->state_names.clear();
    for (auto& s : e->states)
      e->state_names.push_back(s.first.c_str());
    e->state_names.push_back(nullptr);
    e->optimizer = o;
    comm_barrier(comm);
  });
}

wholememory_error_code_t wholememory_embedding_gather_gradient_apply(wholememory_embedding_t e, wholememory_tensor_t indices,
                                                                     wholememory_tensor_t grads, bool /*adjust_cache*/, float lr,
                                                                     wholememory_env_func_t* /*p_env_fns*/, int64_t stream_int)
{
  if (!e || !indices || !grads) return WHOLEMEMORY_INVALID_INPUT;
  return wgb::guarded("wholememory_embedding_gather_gradient_apply",
                      [&] { wgb::gradient_apply(e, indices, grads, lr, reinterpret_cast<cudaStream_t>(stream_int)); });
}

const char* const* wholememory_embedding_get_optimizer_state_names(wholememory_embedding_t e)
{
  static const char* const none[] = {nullptr};
  if (!e || e->state_names.empty()) return none;
  return e->state_names.data();
}

wholememory_tensor_t wholememory_embedding_get_optimizer_state(wholememory_embedding_t e, const char* name)
{
  if (!e || !name) return nullptr;
  for (auto& s : e->states)
    if (s.first == name) return s.second;
  return nullptr;
}

}  // extern "C"
