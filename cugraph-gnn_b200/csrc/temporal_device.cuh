// Temporal neighbour sampling: device kernels used by multihop.cu when a call carries edge times.
//
// Semantics (what the reference asks pylibcugraph for: python/cugraph-pyg/cugraph_pyg/sampler/distributed_sampler.py:
// 56-90, 808-819, 897-900 -- `*_temporal_neighbor_sample` with temporal_sampling_comparison and starting_vertex_times;
// pinned by tests/loader/test_neighbor_loader.py:943-1058 and restated on the CPU in oracle/wg_oracle.cpp:
// wgo_hetero_impl, temporal branch):
//   * every frontier row carries a time: a seed its starting time, any other vertex the time of the edge that reached
//     it FIRST (first occurrence in the hop's edge list, the same order the renumbering uses);
//   * only the edges of the row whose time compares as requested with the row's time are ELIGIBLE;
//   * the uniform one-hop algorithm (Fisher-Yates chain over positions, S1) runs over the eligible edges of the row in
//     CSR order, with the stream geometry of the plain sampler, so that an open time window reproduces plain sampling
//     bit for bit.
//
// Three kernels per (hop, edge type): count eligible edges per row, scan the clipped counts, sample (uniform: the S1 chain
// over the eligible list; biased: A-Res where only eligible edges compete).  All of them read
// the whole row (the eligibility of an edge is data), so a temporal hop costs deg(v) * 8 B of edge-time reads per
// frontier row where the plain sampler touches 16 B of row_ptr: it is bound by the edge-time stream, which is read
// coalesced (a warp / a CTA walks one row).
//
// STATUS: written after this round's GPU budget was spent; compiled for sm_100a, NOT yet run on a GPU.
#pragma once

#include "sample_device.cuh"

namespace wgb {

enum : int {
  kTimeStrictlyIncreasing      = 0,
  kTimeMonotonicallyIncreasing = 1,
  kTimeStrictlyDecreasing      = 2,
  kTimeMonotonicallyDecreasing = 3,
};

__host__ __device__ __forceinline__ bool time_ok(int cmp, long long edge_time, long long vertex_time)
{
  switch (cmp) {
    case kTimeStrictlyIncreasing: return edge_time > vertex_time;
    case kTimeMonotonicallyIncreasing: return edge_time >= vertex_time;
    case kTimeStrictlyDecreasing: return edge_time < vertex_time;
    default: return edge_time <= vertex_time;
  }
}

// one warp per frontier row: eligible[b] = #{p in row(centers[b]) : time_ok(etime[p], ftime[b])},
// clipped[b] = min(eligible[b], M) (M <= 0: not clipped)
template <bool CHUNKED>
__global__ void __launch_bounds__(256) temporal_count_kernel(ChunkRef row_ptr, unsigned long long row_ptr_off, ChunkRef etime,
                                                             unsigned long long etime_off, const long long* __restrict__ centers,
                                                             const long long* __restrict__ ftime, int M, int cmp,
                                                             int* __restrict__ eligible, int* __restrict__ clipped,
                                                             const int* __restrict__ n_dev)
{
  const int n      = *n_dev;
  const int lane   = threadIdx.x & 31;
  const int nwarps = gridDim.x * 8;
  for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < n; b += nwarps) {
    const unsigned long long node = (unsigned long long)centers[b];
    const long long start = load_i64<CHUNKED>(row_ptr, row_ptr_off + node);
    const long long end   = load_i64<CHUNKED>(row_ptr, row_ptr_off + node + 1);
    const long long tv    = ftime[b];
    int c = 0;
    for (long long p = start + lane; p < end; p += 32)
      c += time_ok(cmp, load_i64<CHUNKED>(etime, etime_off + (unsigned long long)p), tv) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) {
      eligible[b] = c;
      clipped[b]  = (M > 0 && c > M) ? M : c;
    }
  }
}

// offsets[i] = sum_{k<i} counts[k] for i <= n, *total_out = offsets[n]; same single-pass ticketed scan as
// count_scan_kernel (sample_device.cuh), the counts come from memory instead of from row_ptr
__global__ void __launch_bounds__(kScanBlock) scan_counts_kernel(const int* __restrict__ counts, int* __restrict__ offsets,
                                                                 unsigned long long* state, unsigned int* ticket,
                                                                 const int* __restrict__ n_dev, int* __restrict__ total_out)
{
  const int n = *n_dev;
  while (true) {
    const int tile = take_ticket(ticket);
    if ((long long)tile * kScanTile > n) return;
    const long long base = (long long)tile * kScanTile + (long long)threadIdx.x * kScanItems;
    unsigned int v[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      long long i = base + k;
      v[k]        = i < n ? (unsigned int)counts[i] : 0u;
    }
    unsigned int agg          = block_scan_items(v);
    unsigned long long prefix = scan_tile_prefix(state, tile, agg);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      long long i = base + k;
      if (i <= n) offsets[i] = (int)(prefix + v[k]);
      if (i == n && total_out) *total_out = (int)(prefix + v[k]);
    }
  }
}

// Block-wide exclusive prefix over one eligibility flag per thread (BLOCK threads, ballots + BLOCK/32 warp counts in
// shared memory): returns this thread's index in the chunk's eligible list, `total` = eligible edges of the chunk.
// Every thread of the block must call it; contains one __syncthreads (callers add one before s_wcnt is reused).
template <int BLOCK>
__device__ __forceinline__ int chunk_eligible_index(bool f, int* s_wcnt, int& total)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned int bal = __ballot_sync(0xffffffffu, f);
  const int within       = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) s_wcnt[wid] = __popc(bal);
  __syncthreads();
  int wbase = 0;
  total     = 0;
#pragma unroll
  for (int w = 0; w < BLOCK / 32; w++) {
    int s = s_wcnt[w];
    if (w < wid) wbase += s;
    total += s;
  }
  return wbase + within;
}

// every eligible edge of one row, in CSR order, to out[off ..): the row is walked in chunks of BLOCK positions
template <typename ColT, bool CHUNKED, int BLOCK>
__device__ __forceinline__ void copy_eligible_row(const ChunkRef& col, unsigned long long col_off, const ChunkRef& etime,
                                                  unsigned long long etime_off, long long start, long long end, long long tv, int cmp,
                                                  int off, int b, ColT* __restrict__ out, int* __restrict__ lid,
                                                  long long* __restrict__ gid, int* s_wcnt)
{
  int running = 0;
  for (long long chunk = start; chunk < end; chunk += BLOCK) {
    const long long p = chunk + threadIdx.x;
    const bool f      = p < end && time_ok(cmp, load_i64<CHUNKED>(etime, etime_off + (unsigned long long)p), tv);
    int total;
    const int idx = chunk_eligible_index<BLOCK>(f, s_wcnt, total);
    if (f) {
      const int o = off + running + idx;
      out[o]      = load_elt<ColT, CHUNKED>(col, col_off + (unsigned long long)p);
      if (lid) lid[o] = b;
      if (gid) gid[o] = p;
    }
    running += total;
    __syncthreads();  // s_wcnt is rewritten by the next chunk
  }
}

// One CTA per frontier row.  N = eligible[b].  N <= M (or M <= 0): every eligible edge, in CSR order.  Otherwise the
// chain of uniform_general_kernel picks M indices into the row's eligible list (same draws: thread j of row b uses
// stream b * T + j, T and the draws per thread from the fan-out tables -- for M <= 32 that is stream 32 b + j, one draw,
// the geometry of uniform_small_kernel), and the row is walked once more to turn indices into edge positions.
template <typename ColT, bool CHUNKED>
__global__ void __launch_bounds__(kGeneralBlock) temporal_uniform_kernel(
  ChunkRef row_ptr, unsigned long long row_ptr_off, ChunkRef col, unsigned long long col_off, ChunkRef etime,
  unsigned long long etime_off, const long long* __restrict__ centers, const long long* __restrict__ ftime,
  const int* __restrict__ eligible, int M, int cmp, unsigned long long seed, const int* __restrict__ offsets,
  ColT* __restrict__ out, int* __restrict__ lid, long long* __restrict__ gid, const Affine* __restrict__ tab,
  const int* __restrict__ n_dev)
{
  const int n = *n_dev;
  __shared__ unsigned long long keys[1024];
  __shared__ int xs[1024];
  __shared__ int prev[1024];
  __shared__ int ptr[2][1024];
  __shared__ int sel[1024];                   // sel[i]: index of the i-th pick in the row's eligible list
  __shared__ long long s_pos[kGeneralBlock];  // edge positions of the eligible edges of the chunk being walked
  __shared__ int s_wcnt[kGeneralBlock / 32];
  const int tid = threadIdx.x;
  int T = 32, ipt = 1, P = 32, P2 = 64;
  if (M > 0) {
    const int func_idx = (M - 1) / 32;
    T                  = kWarpCountTab[func_idx] * 32;
    ipt                = kItemsTab[func_idx];
    P                  = T * ipt;
    while (P2 < P)
      P2 <<= 1;
  }
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    const int N = eligible[b];
    if (N <= 0) continue;  // block-uniform
    const unsigned long long node = (unsigned long long)centers[b];
    const long long start = load_i64<CHUNKED>(row_ptr, row_ptr_off + node);
    const long long end   = load_i64<CHUNKED>(row_ptr, row_ptr_off + node + 1);
    const long long tv    = ftime[b];
    const int off         = offsets[b];
    const bool take_all   = M <= 0 || N <= M;
    if (!take_all) {
      for (int j = tid; j < T; j += blockDim.x) {
        Pcg rng;
        rng.init_tab(seed, (unsigned long long)b * (unsigned long long)T + (unsigned long long)j, tab);
        for (int k = 0; k < ipt; k++) {
          int id = k * T + j;
          int xr = rng.next_i32();  // always drawn
          int x  = id < M ? xr % (N - id) : N;
          xs[id]   = x;
          keys[id] = ((unsigned long long)(unsigned int)x << 32) | (unsigned int)id;
        }
      }
      for (int id = P + tid; id < P2; id += blockDim.x)
        keys[id] = ~0ULL;
      for (int id = tid; id < 1024; id += blockDim.x) {
        prev[id]   = -1;
        ptr[0][id] = -1;
      }
      __syncthreads();
      bitonic_sort_smem(keys, P2, false);
      // prev[i]: the latest step j < i that read the same position (sorted (x, step) pairs are adjacent)
      for (int k = tid + 1; k < P2; k += blockDim.x) {
        unsigned long long kk = keys[k], kp = keys[k - 1];
        unsigned int i = (unsigned int)kk;
        if (kk != ~0ULL && i < (unsigned int)M && (kk >> 32) == (kp >> 32)) prev[i] = (int)(unsigned int)kp;
      }
      // ptr[0][t]: the latest step that wrote the tail position N-1-t before step t reads it
      for (int i = tid; i < M; i += blockDim.x) {
        int x = xs[i];
        if (x >= N - M) {
          int t = N - 1 - x;
          if (t != i) atomicMax(&ptr[0][t], i);
        }
      }
      __syncthreads();
      for (int i = tid; i < M; i += blockDim.x) {
        int f     = ptr[0][i];
        ptr[1][i] = f >= 0 ? f : i;
      }
      __syncthreads();
      int cur = 1;
      for (int s = 1; s < M; s <<= 1) {
        for (int i = tid; i < M; i += blockDim.x)
          ptr[cur ^ 1][i] = ptr[cur][ptr[cur][i]];
        __syncthreads();
        cur ^= 1;
      }
      for (int i = tid; i < M; i += blockDim.x) {
        int pj = prev[i];
        sel[i] = pj >= 0 ? N - 1 - ptr[cur][pj] : xs[i];
      }
      __syncthreads();
    }
    if (take_all) {
      copy_eligible_row<ColT, CHUNKED, kGeneralBlock>(col, col_off, etime, etime_off, start, end, tv, cmp, off, b, out, lid, gid, s_wcnt);
      continue;
    }
    // walk the row once more: the chunk-wide prefix of the eligibility flags is the index in the eligible list
    int running = 0;
    for (long long chunk = start; chunk < end; chunk += blockDim.x) {
      const long long p = chunk + tid;
      const bool f      = p < end && time_ok(cmp, load_i64<CHUNKED>(etime, etime_off + (unsigned long long)p), tv);
      int total;
      const int idx = chunk_eligible_index<kGeneralBlock>(f, s_wcnt, total);
      if (f) s_pos[idx] = p;
      __syncthreads();
      for (int i = tid; i < M; i += blockDim.x) {
        const int a = sel[i] - running;
        if (a >= 0 && a < total) {
          const long long q = s_pos[a];
          out[off + i]      = load_elt<ColT, CHUNKED>(col, col_off + (unsigned long long)q);
          if (lid) lid[off + i] = b;
          if (gid) gid[off + i] = q;
        }
      }
      running += total;
      __syncthreads();  // s_wcnt / s_pos are rewritten by the next chunk (and sel, keys, ... by the next row)
    }
  }
}

// Biased form: the A-Res selection of weighted_kernel (sample_device.cuh: thread j of BLOCK owns positions j, j + BLOCK, ...
// of the row, stream b * BLOCK + j, key = log2(u) / w, the M largest keys win) where only ELIGIBLE positions draw a key and
// compete.  Rows with at most M eligible edges return all of them in CSR order.  With every edge eligible this is
// weighted_kernel (same draws, same candidates).
template <typename ColT, typename WT, int BLOCK, bool CHUNKED>
__global__ void __launch_bounds__(BLOCK) temporal_weighted_kernel(
  ChunkRef row_ptr, unsigned long long row_ptr_off, ChunkRef col, unsigned long long col_off, ChunkRef wgt, unsigned long long wgt_off,
  ChunkRef etime, unsigned long long etime_off, const long long* __restrict__ centers, const long long* __restrict__ ftime,
  const int* __restrict__ eligible, int M, int cmp, unsigned long long seed, const int* __restrict__ offsets, ColT* __restrict__ out,
  int* __restrict__ lid, long long* __restrict__ gid, const Affine* __restrict__ tab, const int* __restrict__ n_dev)
{
  const int n = *n_dev;
  constexpr int C = 2048;
  __shared__ unsigned long long cand[C];
  __shared__ int s_cnt;
  __shared__ unsigned int s_thr;
  __shared__ int s_wcnt[BLOCK / 32];
  const int tid = threadIdx.x;
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    const int Ne = eligible[b];
    if (Ne <= 0) continue;  // block-uniform
    const unsigned long long node = (unsigned long long)centers[b];
    const long long start = load_i64<CHUNKED>(row_ptr, row_ptr_off + node);
    const long long end   = load_i64<CHUNKED>(row_ptr, row_ptr_off + node + 1);
    const long long tv    = ftime[b];
    const int N           = (int)(end - start);
    const int off         = offsets[b];
    if (M <= 0 || Ne <= M) {
      copy_eligible_row<ColT, CHUNKED, BLOCK>(col, col_off, etime, etime_off, start, end, tv, cmp, off, b, out, lid, gid, s_wcnt);
      continue;
    }
    if (tid == 0) {
      s_cnt = 0;
      s_thr = 0u;
    }
    __syncthreads();
    Pcg rng;
    rng.init_tab(seed, (unsigned long long)b * BLOCK + (unsigned long long)tid, tab);
    int ub           = 0;  // block-uniform upper bound of s_cnt
    const int rounds = (N + BLOCK - 1) / BLOCK;
    for (int r = 0; r <= rounds; r++) {
      const bool last = r == rounds;
      if (last || ub + BLOCK > C) {
        // compact: keep the M largest
        int cnt = s_cnt;
        __syncthreads();
        for (int i = cnt + tid; i < C; i += BLOCK)
          cand[i] = 0ULL;
        __syncthreads();
        bitonic_sort_smem(cand, C, true);
        if (tid == 0) {
          s_cnt = min(cnt, M);
          s_thr = cnt >= M ? (unsigned int)(cand[M - 1] >> 32) : 0u;
        }
        ub = M;
        __syncthreads();
        if (last) break;
      }
      const int id = r * BLOCK + tid;
      if (id < N && time_ok(cmp, load_i64<CHUNKED>(etime, etime_off + (unsigned long long)(start + id)), tv)) {
        WT w             = load_elt<WT, CHUNKED>(wgt, wgt_off + (unsigned long long)(start + id));
        float key        = gen_key_from_weight_dev<WT>(w, rng);
        unsigned int enc = float_order_bits(key);
        if (enc >= s_thr) {
          int pos   = atomicAdd(&s_cnt, 1);
          cand[pos] = ((unsigned long long)enc << 32) | (unsigned long long)(0xffffffffu - (unsigned int)id);
        }
      }
      ub += BLOCK;
      __syncthreads();
    }
    for (int i = tid; i < M; i += BLOCK) {
      const int id = (int)(0xffffffffu - (unsigned int)(cand[i] & 0xffffffffULL));
      out[off + i] = load_elt<ColT, CHUNKED>(col, col_off + (unsigned long long)(start + id));
      if (lid) lid[off + i] = b;
      if (gid) gid[off + i] = start + id;
    }
    __syncthreads();
  }
}

}  // namespace wgb
