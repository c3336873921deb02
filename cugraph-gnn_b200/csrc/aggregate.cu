// A1: CSR neighbourhood aggregation (mean / sum SpMM of the sampled block) on sm_100a.
//
// The reference has no aggregation kernel: its models call torch_geometric SAGEConv / GCNConv, whose
// message passing scatter-adds x[row[e]] into out[col[e]] over the COO block the loader expanded from the
// sampler's CSR (python/cugraph-pyg/cugraph_pyg/sampler/sampler.py:55-65,
// python/cugraph-pyg/cugraph_pyg/examples/gcn_dist_mnmg.py:239,
// python/pylibwholegraph/pylibwholegraph/torch/gnn_model.py:119-125).
//
// Here the sampler's CSR (major_offsets / minors) is consumed directly:
//     out[i, :] = reduce_{e in [indptr[i], indptr[i+1])}  x[ map[indices[e]], : ]      reduce = sum | mean
// HBM-bound (nnz * F * elt bytes read, n_dst * F * 4 written; 2 flop per loaded element): a sub-warp owns one
// destination row, every lane keeps one 16-byte slice of the feature row in fp32 registers, the edge loop is
// unrolled 4x so that four 512-byte source rows are in flight per warp.  No atomics, deterministic sum order
// (CSR order).  `x` may be a WholeMemory tensor and `map` the renumber map, which fuses the feature gather
// of the first layer into the aggregation: rows are then pulled from peer GPUs by P2P loads, and the
// gathered matrix never exists.  Tensor cores are not used: there is no reuse to feed them.

#include "wm_common.cuh"

#include <wholememory/b200_ops.h>

namespace wgb {

template <typename T>
struct Vec16 {
  static constexpr int kElts = 16 / sizeof(T);
};

template <typename T>
__device__ __forceinline__ void accumulate16(float* acc, const uint4& raw);

template <>
__device__ __forceinline__ void accumulate16<float>(float* acc, const uint4& raw)
{
  acc[0] += __uint_as_float(raw.x);
  acc[1] += __uint_as_float(raw.y);
  acc[2] += __uint_as_float(raw.z);
  acc[3] += __uint_as_float(raw.w);
}
template <>
__device__ __forceinline__ void accumulate16<__half>(float* acc, const uint4& raw)
{
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float2 f = __half22float2(h[k]);
    acc[2 * k] += f.x;
    acc[2 * k + 1] += f.y;
  }
}
template <>
__device__ __forceinline__ void accumulate16<__nv_bfloat16>(float* acc, const uint4& raw)
{
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float2 f = __bfloat1622float2(h[k]);
    acc[2 * k] += f.x;
    acc[2 * k + 1] += f.y;
  }
}

__device__ __forceinline__ uint4 ld_row16(const char* p)
{
#ifdef WGB_HOST_EMULATION  // tests/emu compiles this file with g++ (logic check on the CPU)
  return *reinterpret_cast<const uint4*>(p);
#else
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#endif
}

constexpr int kAggUnroll = 4;

// G lanes per destination row (G * 16 bytes >= one pass over the row; rows wider than 512 B loop).
template <typename T, typename IdxT, typename PtrT, int G, bool CHUNKED, bool HAS_MAP>
__global__ void __launch_bounds__(256) csr_aggregate_kernel(const PtrT* __restrict__ indptr, const IdxT* __restrict__ indices,
                                                            const long long* __restrict__ map, long long n_dst,
                                                            ChunkRef x, unsigned long long x_off_bytes,
                                                            unsigned long long x_stride_bytes, int row_vecs, bool mean,
                                                            float* __restrict__ out, long long out_stride)
{
  constexpr int E    = Vec16<T>::kElts;
  constexpr int RPW  = 32 / G;  // rows per warp
  const int lane     = threadIdx.x & 31;
  const int g        = lane & (G - 1);
  const int sub      = lane / G;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long wrow = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW; wrow < n_dst; wrow += warps * RPW) {
    const long long i = wrow + sub;
    if (i >= n_dst) continue;
    const long long s = (long long)indptr[i], e = (long long)indptr[i + 1];
    const float scale = (mean && e > s) ? 1.0f / (float)(e - s) : 1.0f;
    for (int v0 = 0; v0 < row_vecs; v0 += G) {
      const int v       = v0 + g;
      const bool active = v < row_vecs;
      float acc[E];
#pragma unroll
      for (int k = 0; k < E; k++)
        acc[k] = 0.f;
      long long p = s;
      for (; p + kAggUnroll <= e; p += kAggUnroll) {
        uint4 raw[kAggUnroll];
#pragma unroll
        for (int u = 0; u < kAggUnroll; u++) {
          long long src = (long long)indices[p + u];
          if (HAS_MAP) src = map[src];
          const char* rp = x.at<CHUNKED>(x_off_bytes + (unsigned long long)src * x_stride_bytes) + (unsigned long long)v * 16;
          if (active) raw[u] = ld_row16(rp);
        }
#pragma unroll
        for (int u = 0; u < kAggUnroll; u++)
          if (active) accumulate16<T>(acc, raw[u]);
      }
      for (; p < e; p++) {
        long long src = (long long)indices[p];
        if (HAS_MAP) src = map[src];
        const char* rp = x.at<CHUNKED>(x_off_bytes + (unsigned long long)src * x_stride_bytes) + (unsigned long long)v * 16;
        if (active) {
          uint4 raw = ld_row16(rp);
          accumulate16<T>(acc, raw);
        }
      }
      if (active) {
        float* o = out + i * out_stride + (long long)v * E;
#pragma unroll
        for (int k = 0; k < E; k += 4)
          *reinterpret_cast<float4*>(o + k) = make_float4(acc[k] * scale, acc[k + 1] * scale, acc[k + 2] * scale, acc[k + 3] * scale);
      }
    }
  }
}

// backward of the aggregation w.r.t. x (local x only): grad_x[indices[e], :] += grad_out[i, :] * scale_i
template <typename IdxT, typename PtrT>
__global__ void __launch_bounds__(256) csr_aggregate_backward_kernel(const PtrT* __restrict__ indptr,
                                                                     const IdxT* __restrict__ indices, long long n_dst,
                                                                     const float* __restrict__ grad_out,
                                                                     long long go_stride, int dim, bool mean,
                                                                     float* __restrict__ grad_x, long long gx_stride)
{
  const int lane        = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_dst; i += warps) {
    const long long s = (long long)indptr[i], e = (long long)indptr[i + 1];
    const float scale = (mean && e > s) ? 1.0f / (float)(e - s) : 1.0f;
    for (int c = lane; c < dim; c += 32) {
      const float gval = grad_out[i * go_stride + c] * scale;
      for (long long p = s; p < e; p++)
        atomicAdd(grad_x + (long long)indices[p] * gx_stride + c, gval);
    }
  }
}

template <typename T, typename IdxT, typename PtrT, int G>
static void launch_agg(const void* indptr, const void* indices, const long long* map, long long n_dst, const ChunkRef& x,
                       unsigned long long x_off_bytes, unsigned long long x_stride_bytes, int row_vecs, bool mean, float* out,
                       long long out_stride, cudaStream_t st)
{
  constexpr int RPW = 32 / G;
  long long warps_needed = (n_dst + RPW - 1) / RPW;
  int grid = (int)std::max<long long>(1, std::min<long long>((warps_needed + 7) / 8, (long long)num_sms() * 8));
  const bool chunked = x.world > 1;
#define WGB_AGG_LAUNCH(CH, MAP)                                                                                          \
  csr_aggregate_kernel<T, IdxT, PtrT, G, CH, MAP><<<grid, 256, 0, st>>>(static_cast<const PtrT*>(indptr),              \
                                                                        static_cast<const IdxT*>(indices), map, n_dst, x, \
                                                                        x_off_bytes, x_stride_bytes, row_vecs, mean, out, \
                                                                        out_stride)
  if (chunked) {
    if (map) WGB_AGG_LAUNCH(true, true);
    else WGB_AGG_LAUNCH(true, false);
  } else {
    if (map) WGB_AGG_LAUNCH(false, true);
    else WGB_AGG_LAUNCH(false, false);
  }
#undef WGB_AGG_LAUNCH
  WGB_CHECK_LAUNCH();
}

template <typename T, typename IdxT, typename PtrT>
static void dispatch_group(int row_vecs, const void* indptr, const void* indices, const long long* map, long long n_dst,
                           const ChunkRef& x, unsigned long long xo, unsigned long long xs, bool mean, float* out,
                           long long out_stride, cudaStream_t st)
{
  if (row_vecs <= 4) launch_agg<T, IdxT, PtrT, 4>(indptr, indices, map, n_dst, x, xo, xs, row_vecs, mean, out, out_stride, st);
  else if (row_vecs <= 8) launch_agg<T, IdxT, PtrT, 8>(indptr, indices, map, n_dst, x, xo, xs, row_vecs, mean, out, out_stride, st);
  else if (row_vecs <= 16) launch_agg<T, IdxT, PtrT, 16>(indptr, indices, map, n_dst, x, xo, xs, row_vecs, mean, out, out_stride, st);
  else launch_agg<T, IdxT, PtrT, 32>(indptr, indices, map, n_dst, x, xo, xs, row_vecs, mean, out, out_stride, st);
}

}  // namespace wgb

extern "C" {

wholememory_error_code_t wholegraph_csr_aggregate(wholememory_tensor_t indptr, wholememory_tensor_t indices,
                                                  wholememory_tensor_t gather_map, wholememory_tensor_t x, int reduce,
                                                  wholememory_tensor_t out, void* stream)
{
  using namespace wgb;
  if (!indptr || !indices || !x || !out) return WHOLEMEMORY_INVALID_INPUT;
  auto* pd = wholememory_tensor_get_tensor_description(indptr);
  auto* id = wholememory_tensor_get_tensor_description(indices);
  auto* xd = wholememory_tensor_get_tensor_description(x);
  auto* od = wholememory_tensor_get_tensor_description(out);
  if (pd->dim != 1 || id->dim != 1 || xd->dim != 2 || od->dim != 2) return WHOLEMEMORY_INVALID_INPUT;
  if ((pd->dtype != WHOLEMEMORY_DT_INT && pd->dtype != WHOLEMEMORY_DT_INT64) ||
      (id->dtype != WHOLEMEMORY_DT_INT && id->dtype != WHOLEMEMORY_DT_INT64))
    return WHOLEMEMORY_INVALID_INPUT;
  if (xd->dtype != WHOLEMEMORY_DT_FLOAT && xd->dtype != WHOLEMEMORY_DT_HALF && xd->dtype != WHOLEMEMORY_DT_BF16) return WHOLEMEMORY_INVALID_INPUT;
  if (od->dtype != WHOLEMEMORY_DT_FLOAT) return WHOLEMEMORY_INVALID_INPUT;
  if (od->sizes[1] != xd->sizes[1] || od->sizes[0] != pd->sizes[0] - 1) return WHOLEMEMORY_INVALID_INPUT;
  if (reduce != WHOLEGRAPH_AGG_SUM && reduce != WHOLEGRAPH_AGG_MEAN) return WHOLEMEMORY_INVALID_INPUT;
  if (gather_map) {
    auto* md = wholememory_tensor_get_tensor_description(gather_map);
    if (md->dim != 1 || md->dtype != WHOLEMEMORY_DT_INT64) return WHOLEMEMORY_INVALID_INPUT;
  }
  return guarded("wholegraph_csr_aggregate", [&] {
    const long long n_dst = pd->sizes[0] - 1;
    if (n_dst <= 0 || xd->sizes[1] == 0) return;
    const size_t elt = dtype_size(xd->dtype);
    const unsigned long long row_bytes = (unsigned long long)xd->sizes[1] * elt;
    ChunkRef xr = make_chunk_ref(x);
    unsigned long long align_or = row_bytes | ((unsigned long long)xd->strides[0] * elt) | ((unsigned long long)xd->storage_offset * elt);
    for (int r = 0; r < xr.world; r++)
      align_or |= reinterpret_cast<unsigned long long>(xr.base[r]) | (r > 0 ? xr.start[r] : 0ULL);
    float* outp = static_cast<float*>(wholememory_tensor_get_data_pointer(out));
    align_or |= reinterpret_cast<unsigned long long>(outp) | ((unsigned long long)od->strides[0] * 4ULL);
    WGB_EXPECTS(align_or % 16 == 0, "csr_aggregate needs 16-byte aligned feature rows (dim * elt, strides and pointers multiples of 16 B)");
    const int row_vecs = (int)(row_bytes / 16);
    const void* ip = wholememory_tensor_get_data_pointer(indptr);
    const void* ix = wholememory_tensor_get_data_pointer(indices);
    const long long* mp = gather_map ? static_cast<const long long*>(wholememory_tensor_get_data_pointer(gather_map)) : nullptr;
    const unsigned long long xo = (unsigned long long)xd->storage_offset * elt, xs = (unsigned long long)xd->strides[0] * elt;
    const bool mean = reduce == WHOLEGRAPH_AGG_MEAN;
    cudaStream_t st = as_stream(stream);
    auto by_index = [&](auto t) {
      using T = decltype(t);
      if (id->dtype == WHOLEMEMORY_DT_INT) {
        if (pd->dtype == WHOLEMEMORY_DT_INT) dispatch_group<T, int, int>(row_vecs, ip, ix, mp, n_dst, xr, xo, xs, mean, outp, od->strides[0], st);
        else dispatch_group<T, int, long long>(row_vecs, ip, ix, mp, n_dst, xr, xo, xs, mean, outp, od->strides[0], st);
      } else {
        if (pd->dtype == WHOLEMEMORY_DT_INT) dispatch_group<T, long long, int>(row_vecs, ip, ix, mp, n_dst, xr, xo, xs, mean, outp, od->strides[0], st);
        else dispatch_group<T, long long, long long>(row_vecs, ip, ix, mp, n_dst, xr, xo, xs, mean, outp, od->strides[0], st);
      }
    };
    if (xd->dtype == WHOLEMEMORY_DT_FLOAT) by_index(float{});
    else if (xd->dtype == WHOLEMEMORY_DT_HALF) by_index(__half{});
    else by_index(__nv_bfloat16{});
  });
}

wholememory_error_code_t wholegraph_csr_aggregate_backward(wholememory_tensor_t indptr, wholememory_tensor_t indices,
                                                           wholememory_tensor_t grad_out, int reduce,
                                                           wholememory_tensor_t grad_x, void* stream)
{
  using namespace wgb;
  if (!indptr || !indices || !grad_out || !grad_x) return WHOLEMEMORY_INVALID_INPUT;
  auto* pd = wholememory_tensor_get_tensor_description(indptr);
  auto* id = wholememory_tensor_get_tensor_description(indices);
  auto* gd = wholememory_tensor_get_tensor_description(grad_out);
  auto* xd = wholememory_tensor_get_tensor_description(grad_x);
  if (pd->dim != 1 || id->dim != 1 || gd->dim != 2 || xd->dim != 2) return WHOLEMEMORY_INVALID_INPUT;
  if (gd->dtype != WHOLEMEMORY_DT_FLOAT || xd->dtype != WHOLEMEMORY_DT_FLOAT || gd->sizes[1] != xd->sizes[1]) return WHOLEMEMORY_INVALID_INPUT;
  if ((pd->dtype != WHOLEMEMORY_DT_INT && pd->dtype != WHOLEMEMORY_DT_INT64) ||
      (id->dtype != WHOLEMEMORY_DT_INT && id->dtype != WHOLEMEMORY_DT_INT64))
    return WHOLEMEMORY_INVALID_INPUT;
  return guarded("wholegraph_csr_aggregate_backward", [&] {
    const long long n_dst = pd->sizes[0] - 1;
    if (n_dst <= 0) return;
    int grid        = (int)std::max<long long>(1, std::min<long long>((n_dst + 7) / 8, (long long)num_sms() * 8));
    cudaStream_t st = as_stream(stream);
    const void* ip  = wholememory_tensor_get_data_pointer(indptr);
    const void* ix  = wholememory_tensor_get_data_pointer(indices);
    const float* go = static_cast<const float*>(wholememory_tensor_get_data_pointer(grad_out));
    float* gx       = static_cast<float*>(wholememory_tensor_get_data_pointer(grad_x));
    const bool mean = reduce == WHOLEGRAPH_AGG_MEAN;
    const int dim   = (int)gd->sizes[1];
#define WGB_BWD(IT, PT)                                                                                                  \
  csr_aggregate_backward_kernel<IT, PT><<<grid, 256, 0, st>>>(static_cast<const PT*>(ip), static_cast<const IT*>(ix), n_dst, go, \
                                                              gd->strides[0], dim, mean, gx, xd->strides[0])
    if (id->dtype == WHOLEMEMORY_DT_INT) {
      if (pd->dtype == WHOLEMEMORY_DT_INT) WGB_BWD(int, int);
      else WGB_BWD(int, long long);
    } else {
      if (pd->dtype == WHOLEMEMORY_DT_INT) WGB_BWD(long long, int);
      else WGB_BWD(long long, long long);
    }
#undef WGB_BWD
    WGB_CHECK_LAUNCH();
  });
}

}  // extern "C"
