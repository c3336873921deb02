// S0: multi-hop, multi-label neighbour sampling with renumbering, as ONE native call (sm_100a).
//
// Replaces, behind the call cugraph-pyg makes (python/cugraph-pyg/cugraph_pyg/sampler/distributed_sampler.py:
// 784-819, 888-902: pylibcugraph.homogeneous_{uniform,biased}_neighbor_sample with renumber=True,
// retain_seeds=True, deduplicate_sources=True, prior_sources_behavior="exclude", return_hops=True), the
// per-hop "sample -> append_unique" loop of the in-repo path
// (python/pylibwholegraph/pylibwholegraph/torch/graph_structure.py:136-196), which costs the reference 3
// host syncs, ~10 launches and several Python round trips PER HOP.
//
// B200-first design:
//  * no host synchronisation until every hop is done: frontier sizes and edge counts stay on the device,
//    kernels are launched over host-known upper bounds (|frontier| * fanout) and read the true sizes from
//    device memory; a single 32-byte D2H copy at the end sizes the outputs.
//  * per hop: count+scan (single pass) -> sample (sub-warp per row) -> hash insert with atomicMin of the
//    first edge position -> single-pass flag+scan+compact that emits the next frontier in first-occurrence
//    order and assigns local ids.  The next frontier of a label is exactly the vertices new in this hop.
//  * one open-addressing table keyed (label, vertex) for the whole call group, persistent across calls in
//    the sampler object; slots are invalidated by an 8-bit epoch in the key instead of a memset.
//  * the final pass scatters (major, minor, edge id) from hop-major scratch to the label-major / hop-minor
//    layout the decoders expect (sampler/sampler.py:525-740) and turns hash slots into local ids, so
//    renumbering costs no extra pass over the edges.
//
// Random numbers: hop h draws with seed hop_seed(random_state, h) = random_state + h * 0x9E3779B97F4A7C15
// and the S1/S2 stream geometry over the label-major concatenated frontier, which is what oracle/
// wg_oracle.cpp:wgo_multihop_sample restates on the CPU.

#include "wm_common.cuh"
#include "pcg.cuh"
#include "sample_device.cuh"

#include <wholememory/b200_ops.h>

#include <algorithm>

namespace wgb {

constexpr int kMaxHops = 16;

struct MhSlot {
  unsigned long long key;  // epoch(8) | label * V + vertex (56)
  unsigned long long aux;  // (255 - epoch)(8) | t(23) | index(32) | is_tag(1): see tag()/fin()
};

__host__ __device__ __forceinline__ unsigned long long mh_tag(unsigned int t, unsigned int e)
{
  return ((unsigned long long)t << 33) | ((unsigned long long)e << 1) | 1ULL;
}
__host__ __device__ __forceinline__ unsigned long long mh_fin(unsigned int t, unsigned int rank)
{
  return ((unsigned long long)t << 33) | ((unsigned long long)rank << 1);
}
constexpr unsigned long long kAuxMask = (1ULL << 56) - 1;

__device__ __forceinline__ unsigned long long mh_mix(unsigned long long x)
{
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// claim-or-find the slot of `item` (= label * V + vertex) in the current epoch
__device__ __forceinline__ unsigned int mh_find_or_insert(MhSlot* table, unsigned int mask, unsigned long long item,
                                                          unsigned long long epoch)
{
  const unsigned long long key = (epoch << 56) | item;
  unsigned int slot            = (unsigned int)mh_mix(item) & mask;
  while (true) {
    unsigned long long cur = ld_relaxed_u64(&table[slot].key);
    if (cur == key) return slot;
    if ((cur >> 56) != epoch) {  // stale or never used: try to claim
      unsigned long long prev = atomicCAS(&table[slot].key, cur, key);
      if (prev == cur || prev == key) return slot;
      continue;  // somebody else claimed it for another key: re-inspect this slot
    }
    slot = (slot + 1) & mask;
  }
}

// label of every seed: slabel[s] = l such that label_offsets[l] <= s < label_offsets[l+1]
__global__ void __launch_bounds__(256) mh_seed_label_kernel(const long long* __restrict__ label_offsets, int B, int S,
                                                            int* __restrict__ slabel)
{
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    int lo = 0, hi = B;  // find the last l with label_offsets[l] <= s
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (label_offsets[mid] <= s) lo = mid;
      else hi = mid;
    }
    slabel[s] = lo;
  }
}

// K3: insert the endpoints discovered in this hop; remember their slot and race for "first position"
// SEEDS=true : item e is seed e, label from slabel;  SEEDS=false: item e is an edge, label = flabel[erow[e]]
template <typename VT, bool SEEDS>
__global__ void __launch_bounds__(256) mh_insert_kernel(MhSlot* table, unsigned int mask, unsigned long long epoch,
                                                        unsigned long long V, unsigned int t,
                                                        const VT* __restrict__ vertices, const int* __restrict__ n_dev,
                                                        const int* __restrict__ erow, const int* __restrict__ flabel,
                                                        unsigned int* __restrict__ slot_of)
{
  const int n = *n_dev;
  const unsigned long long inv_epoch = (255ULL - epoch) << 56;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    int label               = SEEDS ? flabel[e] : flabel[erow[e]];
    unsigned long long item = (unsigned long long)label * V + (unsigned long long)vertices[e];
    unsigned int slot       = mh_find_or_insert(table, mask, item, epoch);
    atomicMin(&table[slot].aux, inv_epoch | mh_tag(t, (unsigned int)e));
    slot_of[e] = slot;
  }
}

// K4: flag first occurrences, scan, compact into the next frontier (first-occurrence order), assign ranks
template <typename VT, bool SEEDS>
__global__ void __launch_bounds__(kScanBlock) mh_compact_kernel(MhSlot* table, unsigned long long epoch, unsigned int t,
                                                                const VT* __restrict__ vertices,
                                                                const int* __restrict__ n_dev,
                                                                const int* __restrict__ erow,
                                                                const int* __restrict__ flabel,
                                                                const unsigned int* __restrict__ slot_of,
                                                                long long* __restrict__ next_frontier,
                                                                int* __restrict__ next_flabel, int* __restrict__ next_n,
                                                                unsigned long long* state, unsigned int* ticket)
{
  const int n          = *n_dev;
  const int tile       = take_ticket(ticket);
  const long long base = (long long)tile * kScanTile + (long long)threadIdx.x * kScanItems;
  const unsigned long long inv_epoch = (255ULL - epoch) << 56;
  unsigned int v[kScanItems];
  unsigned int slot[kScanItems];
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    long long e = base + k;
    v[k]        = 0;
    if (e < n) {
      slot[k] = slot_of[e];
      v[k]    = (table[slot[k]].aux & kAuxMask) == mh_tag(t, (unsigned int)e) ? 1u : 0u;
    }
  }
  unsigned int flags = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++)
    flags |= v[k] << k;
  unsigned int agg          = block_scan_items(v);
  unsigned long long prefix = scan_tile_prefix(state, tile, agg);
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    long long e = base + k;
    if (e < n && ((flags >> k) & 1u)) {
      unsigned int rank   = (unsigned int)(prefix + v[k]);
      next_frontier[rank] = (long long)vertices[e];
      next_flabel[rank]   = SEEDS ? flabel[e] : flabel[erow[e]];
      table[slot[k]].aux  = inv_epoch | mh_fin(t, rank);
    }
    if (e == n) *next_n = (int)(prefix + v[k]);
  }
}

// fr_off[l] = first frontier row of label l (flabel is non-decreasing), fr_off[B] = n
__global__ void __launch_bounds__(256) mh_label_bounds_kernel(const int* __restrict__ flabel, const int* __restrict__ n_dev, int B,
                                                              int* __restrict__ fr_off)
{
  const int n = *n_dev;
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l <= B; l += gridDim.x * blockDim.x) {
    int lo = 0, hi = n;  // first index with flabel[idx] >= l
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (flabel[mid] < l) lo = mid + 1;
      else hi = mid;
    }
    fr_off[l] = lo;
  }
}

// per-label bookkeeping after the last hop
struct MhMeta {
  const int* fr_off[kMaxHops + 1];  // [t][B+1]
  const int* off[kMaxHops];         // [h][ub_frontier+1] sample offsets (null when the hop is skipped)
  int L;
  int B;
};

// counts[0 .. B*L)           edges of (label, hop)
// counts[B*L .. B*L+B)       nodes of label
// counts[B*L+B .. B*L+2B)    source rows of label (CSR major rows)
// base[t*B + l]              local id of the first vertex label l discovered at step t
__global__ void __launch_bounds__(256) mh_meta_kernel(MhMeta m, long long* __restrict__ counts, int* __restrict__ base)
{
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < m.B; l += gridDim.x * blockDim.x) {
    int acc = 0, rows = 0;
    for (int t = 0; t <= m.L; t++) {
      base[t * m.B + l] = acc;
      acc += m.fr_off[t][l + 1] - m.fr_off[t][l];
      if (t == m.L - 1) rows = acc;
    }
    if (m.L == 0) rows = 0;
    for (int h = 0; h < m.L; h++) {
      long long e = 0;
      if (m.off[h]) e = (long long)m.off[h][m.fr_off[h][l + 1]] - (long long)m.off[h][m.fr_off[h][l]];
      counts[(long long)l * m.L + h] = e;
    }
    counts[(long long)m.B * m.L + l]       = acc;
    counts[(long long)m.B * m.L + m.B + l] = rows;
  }
}

// single-block exclusive scan of `n` int64 values (n is small: B*L, B); writes n+1 outputs
__global__ void __launch_bounds__(1024) mh_scan_i64_kernel(const long long* __restrict__ in, long long n,
                                                           long long* __restrict__ out, long long* __restrict__ total_out)
{
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (long long base = 0; base < n; base += blockDim.x) {
    long long i = base + threadIdx.x;
    long long x = i < n ? in[i] : 0;
    long long inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    long long wbase = 0;
    for (int w = 0; w < wid; w++)
      wbase += s_warp[w];
    long long carry = s_carry;
    if (i < n) out[i] = carry + wbase + inc - x;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = carry + wbase + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[n] = s_carry;
    if (total_out) *total_out = s_carry;
  }
}

// final pass over the edges of one hop: scratch (hop-major) -> outputs (label-major, hop-minor)
template <typename OutT, bool CHUNKED>
__global__ void __launch_bounds__(256) mh_emit_edges_kernel(const MhSlot* __restrict__ table, int h, int L, int B,
                                                            const int* __restrict__ n_edges_dev,
                                                            const int* __restrict__ off_h, const int* __restrict__ erow,
                                                            const unsigned int* __restrict__ slot_of,
                                                            const long long* __restrict__ gid,
                                                            const int* __restrict__ flabel_h, MhMeta m,
                                                            const int* __restrict__ base, const long long* __restrict__ lho,
                                                            ChunkRef edge_id_ref, unsigned long long edge_id_off, bool has_edge_id,
                                                            OutT* __restrict__ majors, OutT* __restrict__ minors,
                                                            long long* __restrict__ edge_id_out)
{
  const int n = *n_edges_dev;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int f       = erow[e];
    const int l       = flabel_h[f];
    const int f0      = m.fr_off[h][l];
    const long long p = lho[(long long)l * L + h] + (long long)(e - off_h[f0]);
    unsigned long long aux = table[slot_of[e]].aux & kAuxMask;
    const unsigned int t   = (unsigned int)(aux >> 33);
    const unsigned int rk  = (unsigned int)(aux >> 1);
    if (majors) majors[p] = (OutT)(base[h * B + l] + (f - f0));
    minors[p] = (OutT)(base[t * B + l] + (int)(rk - (unsigned int)m.fr_off[t][l]));
    long long g = gid[e];
    if (has_edge_id) g = load_i64<CHUNKED>(edge_id_ref, edge_id_off + (unsigned long long)g);
    edge_id_out[p] = g;
  }
}

// final pass over the frontier rows of step t: renumber map (+ CSR major offsets for t < L)
__global__ void __launch_bounds__(256) mh_emit_rows_kernel(int t, int L, int B, const int* __restrict__ n_rows_dev,
                                                           const long long* __restrict__ frontier,
                                                           const int* __restrict__ flabel, MhMeta m,
                                                           const int* __restrict__ base, const long long* __restrict__ rmo,
                                                           const long long* __restrict__ lho, const long long* __restrict__ rbase,
                                                           long long* __restrict__ map_out, long long* __restrict__ major_offsets)
{
  const int n = *n_rows_dev;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
    const int l     = flabel[f];
    const int f0    = m.fr_off[t][l];
    const int local = base[t * B + l] + (f - f0);
    map_out[rmo[l] + local] = frontier[f];
    if (major_offsets && t < L) {
      long long eo = lho[(long long)l * L + t];
      if (m.off[t]) eo += (long long)(m.off[t][f] - m.off[t][f0]);
      major_offsets[rbase[l] + local] = eo;
    }
  }
}

// label_hop_offsets in CSR mode index major_offsets: rbase[l] + base[h][l]
__global__ void __launch_bounds__(256) mh_csr_label_hop_kernel(int L, int B, const int* __restrict__ base,
                                                               const long long* __restrict__ rbase,
                                                               const long long* __restrict__ lho,
                                                               long long* __restrict__ label_hop_offsets,
                                                               long long* __restrict__ major_offsets)
{
  for (long long i = blockIdx.x * blockDim.x + threadIdx.x; i <= (long long)B * L; i += (long long)gridDim.x * blockDim.x) {
    if (i == (long long)B * L) {
      label_hop_offsets[i]    = rbase[B];
      major_offsets[rbase[B]] = lho[(long long)B * L];
    } else {
      int l = (int)(i / L), h = (int)(i % L);
      label_hop_offsets[i] = rbase[l] + base[h * B + l];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) mh_seeds_to_i64_kernel(const T* __restrict__ in, int n, long long* __restrict__ out)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = (long long)in[i];
}

// re-insert already numbered vertices after the table had to grow (take-all hops only)
__global__ void __launch_bounds__(256) mh_reinsert_kernel(MhSlot* table, unsigned int mask, unsigned long long epoch,
                                                          unsigned long long V, unsigned int t,
                                                          const long long* __restrict__ frontier,
                                                          const int* __restrict__ flabel, const int* __restrict__ n_dev)
{
  const int n = *n_dev;
  const unsigned long long inv_epoch = (255ULL - epoch) << 56;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
    unsigned long long item = (unsigned long long)flabel[f] * V + (unsigned long long)frontier[f];
    unsigned int slot       = mh_find_or_insert(table, mask, item, epoch);
    table[slot].aux         = inv_epoch | mh_fin(t, (unsigned int)f);
  }
}

}  // namespace wgb

// ---------------------------------------------------------------------------------------------------
// sampler object: persistent scratch + hash table
// ---------------------------------------------------------------------------------------------------
struct wholegraph_multihop_sampler_ {
  struct Buf {
    void* p  = nullptr;
    size_t n = 0;
  };
  Buf table;
  unsigned int table_slots = 0;
  int epoch                = 0;
  Buf slabel, scan_state, small_i32, small_i64, counts;
  Buf frontier[wgb::kMaxHops + 1], flabel[wgb::kMaxHops + 1], fr_off[wgb::kMaxHops + 1];
  Buf off[wgb::kMaxHops], dest[wgb::kMaxHops], erow[wgb::kMaxHops], gid[wgb::kMaxHops], slot[wgb::kMaxHops];
  Buf base;
  long long* h_totals = nullptr;  // pinned
  int device          = -1;
};

namespace wgb {

static void* ensure(wholegraph_multihop_sampler_::Buf& b, size_t bytes)
{
  if (bytes < 256) bytes = 256;
  if (b.n < bytes) {
    if (b.p) WGB_CUDA_TRY(cudaFree(b.p));
    b.p = nullptr;
    b.n = 0;
    size_t want = bytes + bytes / 8;  // a little slack so that slowly growing call groups do not realloc every call
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc(&b.p, bytes);
      want = bytes;
      if (e != cudaSuccess) {
        cudaGetLastError();
        throw std::bad_alloc();
      }
    }
    b.n = want;
  }
  return b.p;
}

static int grid_over(long long n, int sms) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)sms * 8)); }

struct MhCall {
  wholegraph_multihop_sampler_* sp;
  ChunkRef row_ptr, col, wgt, eid;
  unsigned long long row_ptr_off, col_off, wgt_off, eid_off;
  wholememory_dtype_t col_dtype, wgt_dtype, seed_dtype;
  bool weighted, has_eid, chunked;
  const void* seeds;
  const long long* label_offsets;  // device
  int S, B, L;
  int fanout[kMaxHops];
  unsigned long long V;
  unsigned long long random_state;
  int flags;
  void *ctx_majors, *ctx_minors, *ctx_edge_id, *ctx_lho, *ctx_map, *ctx_rmo, *ctx_major_offsets;
  wholememory_env_func_t* env;
  cudaStream_t stream;
};

template <typename ColT, bool CHUNKED>
static void launch_hop_sample(const MhCall& c, const long long* frontier, const int* n_dev, long long ub_rows, int M,
                              unsigned long long seed, const int* off, ColT* dest, int* erow, long long* gid)
{
  int sms           = num_sms();
  const Affine* tab = skip_table_device();
  int n_ub          = (int)ub_rows;
  if (M <= 0) {
    int grid = std::max(1, std::min((n_ub + 7) / 8, sms * 8));
    sample_all_kernel<long long, ColT, CHUNKED><<<grid, 256, 0, c.stream>>>(c.row_ptr, c.row_ptr_off, c.col, c.col_off, frontier, n_ub, off, dest, erow, gid, n_dev);
  } else if (c.weighted) {
    int grid = std::max(1, std::min(n_ub, sms * 8));
    if (c.wgt_dtype == WHOLEMEMORY_DT_FLOAT) {
      if (M <= 256)
        weighted_kernel<long long, ColT, float, 128, CHUNKED><<<grid, 128, 0, c.stream>>>(c.row_ptr, c.row_ptr_off, c.col, c.col_off, c.wgt, c.wgt_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
      else
        weighted_kernel<long long, ColT, float, 256, CHUNKED><<<grid, 256, 0, c.stream>>>(c.row_ptr, c.row_ptr_off, c.col, c.col_off, c.wgt, c.wgt_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
    } else {
      if (M <= 256)
        weighted_kernel<long long, ColT, double, 128, CHUNKED><<<grid, 128, 0, c.stream>>>(c.row_ptr, c.row_ptr_off, c.col, c.col_off, c.wgt, c.wgt_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
      else
        weighted_kernel<long long, ColT, double, 256, CHUNKED><<<grid, 256, 0, c.stream>>>(c.row_ptr, c.row_ptr_off, c.col, c.col_off, c.wgt, c.wgt_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
    }
  } else if (M <= 32) {
    int batches = (n_ub + 31) / 32;
    int grid    = std::max(1, std::min((batches + 7) / 8, sms * 8));
    if (M <= 8)
      uniform_small_kernel<long long, ColT, 8, CHUNKED><<<grid, 256, 0, c.stream>>>(c.row_ptr, c.row_ptr_off, c.col, c.col_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
    else if (M <= 16)
      uniform_small_kernel<long long, ColT, 16, CHUNKED><<<grid, 256, 0, c.stream>>>(c.row_ptr, c.row_ptr_off, c.col, c.col_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
    else
      uniform_small_kernel<long long, ColT, 32, CHUNKED><<<grid, 256, 0, c.stream>>>(c.row_ptr, c.row_ptr_off, c.col, c.col_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
  } else {
    int grid = std::max(1, std::min(n_ub, sms * 16));
    uniform_general_kernel<long long, ColT, CHUNKED><<<grid, kGeneralBlock, 0, c.stream>>>(c.row_ptr, c.row_ptr_off, c.col, c.col_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
  }
  WGB_CHECK_LAUNCH();
}

template <typename ColT, bool CHUNKED>
static void multihop_run(MhCall& c)
{
  auto* sp        = c.sp;
  const int sms   = num_sms();
  const int B     = c.B, L = c.L, S = c.S;
  cudaStream_t st = c.stream;

  // ---- host-side bounds ---------------------------------------------------------------------------
  long long ub_rows[kMaxHops + 1];
  long long ub_edges[kMaxHops];
  bool bounded = true;
  ub_rows[0]   = S;
  for (int h = 0; h < L; h++) {
    if (c.fanout[h] > 0 && bounded) {
      ub_edges[h] = ub_rows[h] * c.fanout[h];
    } else if (c.fanout[h] == 0) {
      ub_edges[h] = 0;
    } else {
      bounded     = false;
      ub_edges[h] = -1;  // known only after the hop's count (host sync)
    }
    ub_rows[h + 1] = ub_edges[h];
  }

  // small per-call arrays
  int* small_i32 = static_cast<int*>(ensure(sp->small_i32, sizeof(int) * (size_t)(2 * (kMaxHops + 2))));
  int* n_rows_dev  = small_i32;                   // [L+1] frontier sizes
  int* n_edges_dev = small_i32 + (kMaxHops + 1);  // [L]   edge counts
  int* slabel = static_cast<int*>(ensure(sp->slabel, sizeof(int) * (size_t)std::max(S, 1)));
  for (int t = 0; t <= L; t++)
    ensure(sp->fr_off[t], sizeof(int) * (size_t)(B + 1));
  ensure(sp->base, sizeof(int) * (size_t)(L + 1) * (size_t)std::max(B, 1));

  // scan states: one slice per single-pass scan of the call (seed compaction + 2 per hop)
  auto tiles_for = [](long long n) { return (int)((n + kScanTile) / kScanTile); };
  auto grow_table = [&](long long items_needed, int upto_t) {
    unsigned long long want = 1024;
    while ((long long)want < 2 * items_needed)
      want <<= 1;
    if (want > (1ULL << 31)) throw logic_error("call group too large for one multi-hop call (hash table > 2^31 slots); split the seeds");
    if (sp->table_slots >= want && sp->epoch >= 1 && sp->epoch < 254 && upto_t < 0) {
      sp->epoch++;
      return;
    }
    if (sp->table_slots < want) {
      ensure(sp->table, (size_t)want * sizeof(MhSlot));
      sp->table_slots = (unsigned int)want;
      sp->epoch       = 0;
    }
    if (sp->epoch == 0 || sp->epoch >= 254 || upto_t >= 0) {
      // fresh table (or epoch wrap, or growth in the middle of a call): all-ones = "stale key, worst aux"
      WGB_CUDA_TRY(cudaMemsetAsync(sp->table.p, 0xFF, (size_t)sp->table_slots * sizeof(MhSlot), st));
      if (upto_t < 0) sp->epoch = 1;
      else if (sp->epoch == 0 || sp->epoch >= 254) sp->epoch = 1;
      for (int t = 0; t <= upto_t; t++) {  // re-insert what is already numbered
        mh_reinsert_kernel<<<grid_over(ub_rows[t], sms), 256, 0, st>>>(
          static_cast<MhSlot*>(sp->table.p), sp->table_slots - 1, (unsigned long long)sp->epoch, c.V, (unsigned int)t,
          static_cast<long long*>(sp->frontier[t].p), static_cast<int*>(sp->flabel[t].p), n_rows_dev + t);
        WGB_CHECK_LAUNCH();
      }
    }
  };

  long long known_items = S;  // seeds + every hop whose edge bound is known before the call starts
  for (int h = 0; h < L && ub_edges[h] >= 0; h++)
    known_items += ub_edges[h];
  grow_table(known_items, -1);
  MhSlot* table = static_cast<MhSlot*>(sp->table.p);
  unsigned long long epoch = (unsigned long long)sp->epoch;

  // ---- step 0: seeds -> frontier_0 (dedup per label, first occurrence keeps the id) --------------------
  long long* seeds64 = static_cast<long long*>(ensure(sp->dest[kMaxHops - 1], sizeof(long long) * (size_t)std::max(S, 1)));
  if (c.seed_dtype == WHOLEMEMORY_DT_INT)
    mh_seeds_to_i64_kernel<int><<<grid_over(S, sms), 256, 0, st>>>(static_cast<const int*>(c.seeds), S, seeds64);
  else
    mh_seeds_to_i64_kernel<long long><<<grid_over(S, sms), 256, 0, st>>>(static_cast<const long long*>(c.seeds), S, seeds64);
  WGB_CHECK_LAUNCH();
  mh_seed_label_kernel<<<grid_over(S, sms), 256, 0, st>>>(c.label_offsets, B, S, slabel);
  WGB_CHECK_LAUNCH();
  // scan-state arena: sized generously per use, zeroed once per use (tiny)
  auto scan_slice = [&](long long n_items) {
    int tiles     = tiles_for(n_items);
    size_t bytes  = scan_state_bytes(tiles);
    void* p       = ensure(sp->scan_state, bytes);
    WGB_CUDA_TRY(cudaMemsetAsync(p, 0, bytes, st));
    return std::make_pair(static_cast<unsigned long long*>(p), tiles);
  };
  int* n_seeds_dev = small_i32 + 2 * (kMaxHops + 1);
  WGB_CUDA_TRY(cudaMemcpyAsync(n_seeds_dev, &S, sizeof(int), cudaMemcpyHostToDevice, st));
  unsigned int* slot0 = static_cast<unsigned int*>(ensure(sp->slot[kMaxHops - 1], sizeof(unsigned int) * (size_t)std::max(S, 1)));
  long long* frontier0 = static_cast<long long*>(ensure(sp->frontier[0], sizeof(long long) * (size_t)std::max(S, 1)));
  int* flabel0         = static_cast<int*>(ensure(sp->flabel[0], sizeof(int) * (size_t)std::max(S, 1)));
  mh_insert_kernel<long long, true><<<grid_over(S, sms), 256, 0, st>>>(table, sp->table_slots - 1, epoch, c.V, 0u, seeds64, n_seeds_dev, nullptr, slabel, slot0);
  WGB_CHECK_LAUNCH();
  {
    auto ss = scan_slice(S);
    mh_compact_kernel<long long, true><<<ss.second, kScanBlock, 0, st>>>(table, epoch, 0u, seeds64, n_seeds_dev, nullptr, slabel, slot0, frontier0, flabel0, n_rows_dev + 0, ss.first, reinterpret_cast<unsigned int*>(ss.first + ss.second));
    WGB_CHECK_LAUNCH();
  }
  mh_label_bounds_kernel<<<grid_over(B + 1, sms), 256, 0, st>>>(flabel0, n_rows_dev + 0, B, static_cast<int*>(sp->fr_off[0].p));
  WGB_CHECK_LAUNCH();

  // ---- hops ---------------------------------------------------------------------------------------------
  MhMeta meta;
  memset(&meta, 0, sizeof(meta));
  meta.L = L;
  meta.B = B;
  meta.fr_off[0] = static_cast<int*>(sp->fr_off[0].p);
  for (int h = 0; h < L; h++) {
    const long long rows_ub = ub_rows[h];
    const int M             = c.fanout[h];
    long long* frontier     = static_cast<long long*>(sp->frontier[h].p);
    int* flabel             = static_cast<int*>(sp->flabel[h].p);
    int* off                = static_cast<int*>(ensure(sp->off[h], sizeof(int) * (size_t)(rows_ub + 1 + kScanTile)));
    meta.off[h]             = nullptr;
    long long edges_ub      = ub_edges[h];
    if (M != 0 && rows_ub > 0) {
      // K1: counts + scan over the frontier
      auto ss = scan_slice(rows_ub);
      count_scan_kernel<long long, CHUNKED><<<ss.second, kScanBlock, 0, st>>>(c.row_ptr, c.row_ptr_off, frontier, (int)rows_ub, M, off, ss.first, reinterpret_cast<unsigned int*>(ss.first + ss.second), n_rows_dev + h, n_edges_dev + h);
      WGB_CHECK_LAUNCH();
      meta.off[h] = off;
      if (edges_ub < 0) {
        // take-all hop: the edge count is data dependent -> one extra host sync to size the scratch
        int e_host = 0, r_host = 0;
        WGB_CUDA_TRY(cudaMemcpyAsync(&e_host, n_edges_dev + h, sizeof(int), cudaMemcpyDeviceToHost, st));
        WGB_CUDA_TRY(cudaMemcpyAsync(&r_host, n_rows_dev + h, sizeof(int), cudaMemcpyDeviceToHost, st));
        WGB_CUDA_TRY(cudaStreamSynchronize(st));
        WGB_EXPECTS(e_host >= 0, "edge count of a take-all hop overflowed int32");
        edges_ub = e_host;
        known_items += e_host;
        ub_edges[h]    = edges_ub;
        ub_rows[h + 1] = edges_ub;
        for (int hh = h + 1; hh < L; hh++) {  // later bounded hops can now be bounded too
          if (c.fanout[hh] > 0) ub_edges[hh] = ub_rows[hh] * c.fanout[hh];
          else if (c.fanout[hh] == 0) ub_edges[hh] = 0;
          else break;
          ub_rows[hh + 1] = ub_edges[hh];
          known_items += ub_edges[hh];
        }
        if ((unsigned long long)(2 * known_items) > sp->table_slots) {
          grow_table(known_items, h);
          table = static_cast<MhSlot*>(sp->table.p);
          epoch = (unsigned long long)sp->epoch;
        }
      }
    } else {
      WGB_CUDA_TRY(cudaMemsetAsync(n_edges_dev + h, 0, sizeof(int), st));
      edges_ub       = 0;
      ub_edges[h]    = 0;
      ub_rows[h + 1] = 0;
    }
    WGB_EXPECTS(edges_ub < (1LL << 31) - kScanTile, "too many edges in one hop for one call; split the seeds into more calls");
    ColT* dest     = static_cast<ColT*>(ensure(sp->dest[h], sizeof(ColT) * (size_t)std::max<long long>(edges_ub, 1)));
    int* erow      = static_cast<int*>(ensure(sp->erow[h], sizeof(int) * (size_t)std::max<long long>(edges_ub, 1)));
    long long* gid = static_cast<long long*>(ensure(sp->gid[h], sizeof(long long) * (size_t)std::max<long long>(edges_ub, 1)));
    unsigned int* slot = static_cast<unsigned int*>(ensure(sp->slot[h], sizeof(unsigned int) * (size_t)std::max<long long>(edges_ub, 1)));
    long long* next_frontier = static_cast<long long*>(ensure(sp->frontier[h + 1], sizeof(long long) * (size_t)std::max<long long>(edges_ub, 1)));
    int* next_flabel         = static_cast<int*>(ensure(sp->flabel[h + 1], sizeof(int) * (size_t)std::max<long long>(edges_ub, 1)));
    if (edges_ub > 0) {
      // K2: sample
      launch_hop_sample<ColT, CHUNKED>(c, frontier, n_rows_dev + h, rows_ub, M, c.random_state + (unsigned long long)h * 0x9E3779B97F4A7C15ULL, off, dest, erow, gid);
      // K3: insert (label, neighbour)
      mh_insert_kernel<ColT, false><<<grid_over(edges_ub, sms), 256, 0, st>>>(table, sp->table_slots - 1, epoch, c.V, (unsigned int)(h + 1), dest, n_edges_dev + h, erow, flabel, slot);
      WGB_CHECK_LAUNCH();
      // K4: first occurrences -> next frontier
      auto ss = scan_slice(edges_ub);
      mh_compact_kernel<ColT, false><<<ss.second, kScanBlock, 0, st>>>(table, epoch, (unsigned int)(h + 1), dest, n_edges_dev + h, erow, flabel, slot, next_frontier, next_flabel, n_rows_dev + h + 1, ss.first, reinterpret_cast<unsigned int*>(ss.first + ss.second));
      WGB_CHECK_LAUNCH();
    } else {
      WGB_CUDA_TRY(cudaMemsetAsync(n_rows_dev + h + 1, 0, sizeof(int), st));
    }
    mh_label_bounds_kernel<<<grid_over(B + 1, sms), 256, 0, st>>>(next_flabel, n_rows_dev + h + 1, B, static_cast<int*>(sp->fr_off[h + 1].p));
    WGB_CHECK_LAUNCH();
    meta.fr_off[h + 1] = static_cast<int*>(sp->fr_off[h + 1].p);
  }

  // ---- per-label bookkeeping + offsets ---------------------------------------------------------------------
  const long long n_counts = (long long)B * L + 2LL * B;
  long long* counts = static_cast<long long*>(ensure(sp->counts, sizeof(long long) * (size_t)(n_counts + 1)));
  long long* scans  = static_cast<long long*>(ensure(sp->small_i64, sizeof(long long) * (size_t)(n_counts + 8)));
  long long* lho    = scans;                        // B*L + 1
  long long* rmo    = scans + (long long)B * L + 1;  // B + 1
  long long* rbase  = rmo + B + 1;                   // B + 1
  long long* totals = rbase + B + 1;                 // 3
  int* base         = static_cast<int*>(sp->base.p);
  mh_meta_kernel<<<grid_over(B, sms), 256, 0, st>>>(meta, counts, base);
  WGB_CHECK_LAUNCH();
  mh_scan_i64_kernel<<<1, 1024, 0, st>>>(counts, (long long)B * L, lho, totals + 0);
  WGB_CHECK_LAUNCH();
  mh_scan_i64_kernel<<<1, 1024, 0, st>>>(counts + (long long)B * L, B, rmo, totals + 1);
  WGB_CHECK_LAUNCH();
  mh_scan_i64_kernel<<<1, 1024, 0, st>>>(counts + (long long)B * L + B, B, rbase, totals + 2);
  WGB_CHECK_LAUNCH();
  // ---- the one host sync of the call: output sizes ------------------------------------------------------------
  WGB_CUDA_TRY(cudaMemcpyAsync(sp->h_totals, totals, 3 * sizeof(long long), cudaMemcpyDeviceToHost, st));
  WGB_CUDA_TRY(cudaStreamSynchronize(st));
  const long long n_edges = sp->h_totals[0], n_nodes = sp->h_totals[1], n_srcrows = sp->h_totals[2];

  const bool csr    = (c.flags & WHOLEGRAPH_MULTIHOP_CSR) != 0;
  const bool idx64  = (c.flags & WHOLEGRAPH_MULTIHOP_INT64_IDS) != 0;
  const wholememory_dtype_t idx_dt = idx64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT;
  void* out_minors  = output_alloc(c.env, c.ctx_minors, n_edges, idx_dt);
  void* out_majors  = (!csr && c.ctx_majors) ? output_alloc(c.env, c.ctx_majors, n_edges, idx_dt) : nullptr;
  long long* out_eid = static_cast<long long*>(output_alloc(c.env, c.ctx_edge_id, n_edges, WHOLEMEMORY_DT_INT64));
  long long* out_lho = static_cast<long long*>(output_alloc(c.env, c.ctx_lho, (long long)B * L + 1, WHOLEMEMORY_DT_INT64));
  long long* out_map = static_cast<long long*>(output_alloc(c.env, c.ctx_map, n_nodes, WHOLEMEMORY_DT_INT64));
  long long* out_rmo = static_cast<long long*>(output_alloc(c.env, c.ctx_rmo, B + 1, WHOLEMEMORY_DT_INT64));
  long long* out_moff = nullptr;
  if (csr) {
    WGB_EXPECTS(c.ctx_major_offsets != nullptr, "CSR compression needs a major_offsets output context");
    out_moff = static_cast<long long*>(output_alloc(c.env, c.ctx_major_offsets, n_srcrows + 1, WHOLEMEMORY_DT_INT64));
  }
  WGB_CUDA_TRY(cudaMemcpyAsync(out_rmo, rmo, sizeof(long long) * (size_t)(B + 1), cudaMemcpyDeviceToDevice, st));
  if (!csr) {
    WGB_CUDA_TRY(cudaMemcpyAsync(out_lho, lho, sizeof(long long) * (size_t)((long long)B * L + 1), cudaMemcpyDeviceToDevice, st));
  } else {
    mh_csr_label_hop_kernel<<<grid_over((long long)B * L + 1, sms), 256, 0, st>>>(L, B, base, rbase, lho, out_lho, out_moff);
    WGB_CHECK_LAUNCH();
  }
  // ---- final pass: edges and rows to their label-major places ---------------------------------------------------
  for (int h = 0; h < L; h++) {
    if (ub_edges[h] <= 0) continue;
    int grid = grid_over(ub_edges[h], sms);
    if (idx64)
      mh_emit_edges_kernel<long long, CHUNKED><<<grid, 256, 0, st>>>(table, h, L, B, n_edges_dev + h, static_cast<int*>(sp->off[h].p), static_cast<int*>(sp->erow[h].p), static_cast<unsigned int*>(sp->slot[h].p), static_cast<long long*>(sp->gid[h].p), static_cast<int*>(sp->flabel[h].p), meta, base, lho, c.eid, c.eid_off, c.has_eid, static_cast<long long*>(out_majors), static_cast<long long*>(out_minors), out_eid);
    else
      mh_emit_edges_kernel<int, CHUNKED><<<grid, 256, 0, st>>>(table, h, L, B, n_edges_dev + h, static_cast<int*>(sp->off[h].p), static_cast<int*>(sp->erow[h].p), static_cast<unsigned int*>(sp->slot[h].p), static_cast<long long*>(sp->gid[h].p), static_cast<int*>(sp->flabel[h].p), meta, base, lho, c.eid, c.eid_off, c.has_eid, static_cast<int*>(out_majors), static_cast<int*>(out_minors), out_eid);
    WGB_CHECK_LAUNCH();
  }
  for (int t = 0; t <= L; t++) {
    if (ub_rows[t] <= 0) continue;
    mh_emit_rows_kernel<<<grid_over(ub_rows[t], sms), 256, 0, st>>>(t, L, B, n_rows_dev + t, static_cast<long long*>(sp->frontier[t].p), static_cast<int*>(sp->flabel[t].p), meta, base, rmo, lho, rbase, out_map, out_moff);
    WGB_CHECK_LAUNCH();
  }
}

}  // namespace wgb

extern "C" {

wholememory_error_code_t wholegraph_create_multihop_sampler(wholegraph_multihop_sampler_t* sampler)
{
  if (!sampler) return WHOLEMEMORY_INVALID_INPUT;
  return wgb::guarded("wholegraph_create_multihop_sampler", [&] {
    auto* s = new wholegraph_multihop_sampler_();
    WGB_CUDA_TRY(cudaGetDevice(&s->device));
    WGB_CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&s->h_totals), 8 * sizeof(long long)));
    *sampler = s;
  });
}

wholememory_error_code_t wholegraph_destroy_multihop_sampler(wholegraph_multihop_sampler_t s)
{
  if (!s) return WHOLEMEMORY_INVALID_INPUT;
  auto drop = [](wholegraph_multihop_sampler_::Buf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
  };
  cudaDeviceSynchronize();
  drop(s->table); drop(s->slabel); drop(s->scan_state); drop(s->small_i32); drop(s->small_i64); drop(s->counts); drop(s->base);
  for (int i = 0; i <= wgb::kMaxHops; i++) {
    drop(s->frontier[i]); drop(s->flabel[i]); drop(s->fr_off[i]);
  }
  for (int i = 0; i < wgb::kMaxHops; i++) {
    drop(s->off[i]); drop(s->dest[i]); drop(s->erow[i]); drop(s->gid[i]); drop(s->slot[i]);
  }
  if (s->h_totals) cudaFreeHost(s->h_totals);
  cudaGetLastError();
  delete s;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholegraph_multihop_neighbor_sample(
  wholegraph_multihop_sampler_t sampler, wholememory_tensor_t csr_row_ptr, wholememory_tensor_t csr_col,
  wholememory_tensor_t csr_weight, wholememory_tensor_t csr_edge_id, wholememory_tensor_t seeds,
  wholememory_tensor_t label_offsets, const int* fanout, int num_hops, unsigned long long random_state, int flags,
  void* out_majors_ctx, void* out_minors_ctx, void* out_edge_id_ctx, void* out_label_hop_offsets_ctx,
  void* out_renumber_map_ctx, void* out_renumber_map_offsets_ctx, void* out_major_offsets_ctx,
  wholememory_env_func_t* p_env_fns, void* stream)
{
  using namespace wgb;
  if (!sampler || !csr_row_ptr || !csr_col || !seeds || !label_offsets || !fanout || !p_env_fns) return WHOLEMEMORY_INVALID_INPUT;
  if (num_hops < 1 || num_hops >= kMaxHops - 1) return WHOLEMEMORY_INVALID_INPUT;
  if (!out_minors_ctx || !out_edge_id_ctx || !out_label_hop_offsets_ctx || !out_renumber_map_ctx || !out_renumber_map_offsets_ctx) return WHOLEMEMORY_INVALID_INPUT;
  auto* rd = wholememory_tensor_get_tensor_description(csr_row_ptr);
  auto* cd = wholememory_tensor_get_tensor_description(csr_col);
  auto* sd = wholememory_tensor_get_tensor_description(seeds);
  auto* ld = wholememory_tensor_get_tensor_description(label_offsets);
  if (rd->dim != 1 || cd->dim != 1 || sd->dim != 1 || ld->dim != 1) return WHOLEMEMORY_INVALID_INPUT;
  if (rd->dtype != WHOLEMEMORY_DT_INT64 || ld->dtype != WHOLEMEMORY_DT_INT64) return WHOLEMEMORY_INVALID_INPUT;
  if (cd->dtype != WHOLEMEMORY_DT_INT && cd->dtype != WHOLEMEMORY_DT_INT64) return WHOLEMEMORY_INVALID_INPUT;
  if (sd->dtype != WHOLEMEMORY_DT_INT && sd->dtype != WHOLEMEMORY_DT_INT64) return WHOLEMEMORY_INVALID_INPUT;
  if (ld->sizes[0] < 1 || rd->sizes[0] < 1) return WHOLEMEMORY_INVALID_INPUT;
  for (int h = 0; h < num_hops; h++)
    if (fanout[h] > 1024) return WHOLEMEMORY_NOT_IMPLEMENTED;
  if (csr_weight) {
    auto* wd = wholememory_tensor_get_tensor_description(csr_weight);
    if (wd->dim != 1 || wd->sizes[0] != cd->sizes[0] || (wd->dtype != WHOLEMEMORY_DT_FLOAT && wd->dtype != WHOLEMEMORY_DT_DOUBLE)) return WHOLEMEMORY_INVALID_INPUT;
  }
  if (csr_edge_id) {
    auto* ed = wholememory_tensor_get_tensor_description(csr_edge_id);
    if (ed->dim != 1 || ed->sizes[0] != cd->sizes[0] || ed->dtype != WHOLEMEMORY_DT_INT64) return WHOLEMEMORY_INVALID_INPUT;
  }
  return guarded("wholegraph_multihop_neighbor_sample", [&] {
    WGB_EXPECTS(sd->sizes[0] < (1LL << 31) - kScanTile, "too many seeds for one call");
    WGB_EXPECTS(ld->sizes[0] - 1 < (1LL << 23), "too many labels for one call");
    MhCall c;
    memset(&c.wgt, 0, sizeof(c.wgt));
    memset(&c.eid, 0, sizeof(c.eid));
    c.sp          = sampler;
    c.row_ptr     = make_chunk_ref(csr_row_ptr);
    c.row_ptr_off = (unsigned long long)rd->storage_offset;
    c.col         = make_chunk_ref(csr_col);
    c.col_off     = (unsigned long long)cd->storage_offset;
    c.col_dtype   = cd->dtype;
    c.weighted    = csr_weight != nullptr;
    c.wgt_off     = 0;
    c.wgt_dtype   = WHOLEMEMORY_DT_FLOAT;
    if (csr_weight) {
      c.wgt       = make_chunk_ref(csr_weight);
      c.wgt_off   = (unsigned long long)wholememory_tensor_get_tensor_description(csr_weight)->storage_offset;
      c.wgt_dtype = wholememory_tensor_get_tensor_description(csr_weight)->dtype;
    }
    c.has_eid = csr_edge_id != nullptr;
    c.eid_off = 0;
    if (csr_edge_id) {
      c.eid     = make_chunk_ref(csr_edge_id);
      c.eid_off = (unsigned long long)wholememory_tensor_get_tensor_description(csr_edge_id)->storage_offset;
    }
    c.chunked = c.row_ptr.world > 1 || c.col.world > 1 || (c.weighted && c.wgt.world > 1) || (c.has_eid && c.eid.world > 1);
    c.seeds         = wholememory_tensor_get_data_pointer(seeds);
    c.seed_dtype    = sd->dtype;
    c.label_offsets = static_cast<const long long*>(wholememory_tensor_get_data_pointer(label_offsets));
    c.S             = (int)sd->sizes[0];
    c.B             = (int)ld->sizes[0] - 1;
    c.L             = num_hops;
    for (int h = 0; h < num_hops; h++)
      c.fanout[h] = fanout[h];
    c.V = (unsigned long long)(rd->sizes[0] - 1);
    WGB_EXPECTS((long double)c.V * (long double)std::max(c.B, 1) < 72057594037927936.0L, "labels x vertices must stay below 2^56");
    c.random_state      = random_state;
    c.flags             = flags;
    c.ctx_majors        = out_majors_ctx;
    c.ctx_minors        = out_minors_ctx;
    c.ctx_edge_id       = out_edge_id_ctx;
    c.ctx_lho           = out_label_hop_offsets_ctx;
    c.ctx_map           = out_renumber_map_ctx;
    c.ctx_rmo           = out_renumber_map_offsets_ctx;
    c.ctx_major_offsets = out_major_offsets_ctx;
    c.env               = p_env_fns;
    c.stream            = as_stream(stream);
    if (cd->dtype == WHOLEMEMORY_DT_INT) {
      if (c.chunked) multihop_run<int, true>(c);
      else multihop_run<int, false>(c);
    } else {
      if (c.chunked) multihop_run<long long, true>(c);
      else multihop_run<long long, false>(c);
    }
  });
}

}  // extern "C"
