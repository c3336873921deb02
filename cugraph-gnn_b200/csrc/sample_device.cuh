// Device kernels of the one-hop samplers (shared by sample.cu and multihop.cu).
#pragma once

#include "wm_common.cuh"
#include "pcg.cuh"

namespace wgb {

// ---- small helpers ------------------------------------------------------------------------------
template <bool CHUNKED>
__device__ __forceinline__ long long load_i64(const ChunkRef& ref, unsigned long long elt)
{
  return __ldg(reinterpret_cast<const long long*>(ref.at<CHUNKED>(elt * 8ULL)));
}
template <typename T, bool CHUNKED>
__device__ __forceinline__ T load_elt(const ChunkRef& ref, unsigned long long elt)
{
  return __ldg(reinterpret_cast<const T*>(ref.at<CHUNKED>(elt * sizeof(T))));
}

#ifndef WGB_COPY_W
#define WGB_COPY_W 4  // reads in flight per lane in the copy-row path (passes of 12 cover a fan-out-10 row in one trip but spill: no faster)
#endif
// read an element with an L2 cache policy (createpolicy...): single-use random reads of the graph should not push the
// sampler's tables out of L2.  policy == 0: the plain read-only load.
template <typename T, bool CHUNKED>
__device__ __forceinline__ T load_elt_policy(const ChunkRef& ref, unsigned long long elt, unsigned long long policy)
{
#ifndef WGB_HOST_EMULATION
  if (policy != 0ULL) {
    const T* p = reinterpret_cast<const T*>(ref.at<CHUNKED>(elt * sizeof(T)));
    T v;
    if constexpr (sizeof(T) == 4) {
      unsigned int r;
      asm volatile("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(policy));
      v = (T)r;
    } else {
      unsigned long long r;
      asm volatile("ld.global.nc.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(policy));
      v = (T)r;
    }
    return v;
  }
#endif
  (void)policy;
  return load_elt<T, CHUNKED>(ref, elt);
}

// request the sector of an element into L2 (no register, no wait): used where the value is only needed much later
template <typename T, bool CHUNKED>
__device__ __forceinline__ void prefetch_l2(const ChunkRef& ref, unsigned long long elt)
{
#ifndef WGB_HOST_EMULATION
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ref.at<CHUNKED>(elt * sizeof(T))));
#else
  (void)ref;
  (void)elt;
#endif
}

#ifdef WGB_HOST_EMULATION
// tests/emu compiles this header with g++ to check kernel LOGIC on the CPU (test infrastructure, never the product)
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
#else
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

// ---- single-pass exclusive scan (decoupled look-back) ----------------------------------------------
// A tile publishes one 64-bit word: flag(2 bits) | value(62 bits); flag 1 = tile aggregate,
// flag 2 = inclusive prefix.  Tiles take a ticket, so a predecessor is always resident or done.
constexpr int kScanBlock = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile  = kScanBlock * kScanItems;
constexpr unsigned long long kFlagAgg    = 1ULL << 62;
constexpr unsigned long long kFlagPrefix = 2ULL << 62;
constexpr unsigned long long kValueMask  = (1ULL << 62) - 1;

inline size_t scan_state_bytes(int tiles) { return (size_t)tiles * 8 + 16; }

// Returns the exclusive prefix of this tile; must be called by every thread of the block.
// `agg` is the tile aggregate (valid in every thread), `tile` the ticket.
__device__ __forceinline__ unsigned long long scan_tile_prefix(unsigned long long* state, int tile, unsigned long long agg)
{
  __shared__ unsigned long long s_prefix;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned long long exclusive = 0;
    if (tile == 0) {
      if (lane == 0) st_relaxed_u64(&state[0], kFlagPrefix | agg);
    } else {
      if (lane == 0) st_relaxed_u64(&state[tile], kFlagAgg | agg);
      int idx = tile - 1;
      while (true) {
        int t = idx - lane;
        unsigned long long w;
        do {
          w = (t >= 0) ? ld_relaxed_u64(&state[t]) : kFlagPrefix;
        } while (__any_sync(0xffffffffu, (w >> 62) == 0));
        unsigned int pm = __ballot_sync(0xffffffffu, (w >> 62) == 2);
        unsigned long long val = w & kValueMask;
        if (pm) {
          int first = __ffs(pm) - 1;
          if (lane > first) val = 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          val += __shfl_xor_sync(0xffffffffu, val, o);
        exclusive += val;
        if (pm) break;
        idx -= 32;
      }
      if (lane == 0) st_relaxed_u64(&state[tile], kFlagPrefix | ((exclusive + agg) & kValueMask));
    }
    if (lane == 0) s_prefix = exclusive;
  }
  __syncthreads();
  unsigned long long r = s_prefix;
  __syncthreads();
  return r;
}

// block-wide exclusive scan of kScanItems values per thread (blocked arrangement);
// returns the block aggregate, rewrites v[] with exclusive prefixes inside the tile.
template <int ITEMS>
__device__ __forceinline__ unsigned int block_scan_items(unsigned int (&v)[ITEMS])
{
  __shared__ unsigned int s_warp[kScanBlock / 32];
  unsigned int tsum = 0;
#pragma unroll
  for (int k = 0; k < ITEMS; k++) {
    unsigned int x = v[k];
    v[k]           = tsum;
    tsum += x;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned int inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned int y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  unsigned int wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kScanBlock / 32; w++) {
    unsigned int s = s_warp[w];
    if (w < wid) wbase += s;
    total += s;
  }
  __syncthreads();
  unsigned int tbase = wbase + inc - tsum;
#pragma unroll
  for (int k = 0; k < ITEMS; k++)
    v[k] += tbase;
  return total;
}

__device__ __forceinline__ int take_ticket(unsigned int* ticket)
{
  __shared__ int s_tile;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  int t = s_tile;
  __syncthreads();
  return t;
}

// counts + exclusive scan in one pass: offsets[i] = sum_{k<i} min(deg(centers[k]), M), offsets[n] = total
// (reference: get_sample_count_without_replacement_kernel + thrust::exclusive_scan,
//  unweighted_sample_without_replacement_func.cuh:29-48, 323-327)
template <typename IdT, bool CHUNKED>
__global__ void __launch_bounds__(kScanBlock) count_scan_kernel(ChunkRef row_ptr, unsigned long long row_ptr_off,
                                                                 const IdT* __restrict__ centers, int n, int M,
                                                                 int* __restrict__ offsets,
                                                                 unsigned long long* state, unsigned int* ticket,
                                                                 const int* __restrict__ n_dev = nullptr,
                                                                 int* __restrict__ total_out = nullptr)
{
  if (n_dev) n = *n_dev;  // frontier size produced on the device by the previous hop
  // persistent CTAs take tiles by ticket: tiles start in increasing order, so every predecessor a tile looks back
  // at is resident or finished, and the grid can be sized from the resident CTA count instead of from n (which
  // may only be known as a host-side upper bound when it lives on the device).
  while (true) {
    const int tile = take_ticket(ticket);
    if ((long long)tile * kScanTile > n) return;
    const long long base = (long long)tile * kScanTile + (long long)threadIdx.x * kScanItems;
    unsigned int v[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      long long i = base + k;
      unsigned int c = 0;
      if (i < n) {
        unsigned long long node = (unsigned long long)centers[i];
        long long s = load_i64<CHUNKED>(row_ptr, row_ptr_off + node);
        long long e = load_i64<CHUNKED>(row_ptr, row_ptr_off + node + 1);
        long long d = e - s;
        if (d < 0) d = 0;
        if (M > 0 && d > M) d = M;
        c = (unsigned int)d;
      }
      v[k] = c;
    }
    unsigned int agg = block_scan_items(v);
    unsigned long long prefix = scan_tile_prefix(state, tile, agg);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      long long i = base + k;
      if (i <= n) offsets[i] = (int)(prefix + v[k]);
      if (i == n && total_out) *total_out = (int)(prefix + v[k]);
    }
  }
}

// ---- fan-out <= 0: copy the whole adjacency (reference: sample_comm.cuh:14-48) ------------------------
template <typename IdT, typename ColT, bool CHUNKED>
__global__ void __launch_bounds__(256) sample_all_kernel(ChunkRef row_ptr, unsigned long long row_ptr_off, ChunkRef col,
                                                         unsigned long long col_off, const IdT* __restrict__ centers,
                                                         int n, const int* __restrict__ offsets, ColT* __restrict__ out,
                                                         int* __restrict__ lid, long long* __restrict__ gid,
                                                         const int* __restrict__ n_dev = nullptr)
{
  if (n_dev) n = *n_dev;
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * 8;
  for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < n; b += nwarps) {
    unsigned long long node = (unsigned long long)centers[b];
    long long start = load_i64<CHUNKED>(row_ptr, row_ptr_off + node);
    long long end   = load_i64<CHUNKED>(row_ptr, row_ptr_off + node + 1);
    long long N     = end - start;
    int off         = offsets[b];
    for (long long j = lane; j < N; j += 32) {
      out[off + j] = load_elt<ColT, CHUNKED>(col, col_off + start + j);
      if (lid) lid[off + j] = b;
      if (gid) gid[off + j] = start + j;
    }
  }
}

// ---- the Fisher-Yates chain, resolved by a (sub-)warp ---------------------------------------------------
// Sequential definition (cpp/tests/wholegraph_ops/graph_sampling_test_utils.cu:295-310):
//     Q = [0..N);  for i in 0..M-1:  a[i] = Q[r[i]];  Q[r[i]] = Q[N-i-1]
// Parallel form used here: step i reads position x_i = r[i] and copies position y_i = N-1-i into it.
//   V_i  := value of position y_i at time i.  Position y_i can only have been written by a step j < i
//           with x_j == y_i; let f(i) be the LATEST such j.  Then V_i = V_f(i), or y_i if there is none
//           -> V_i = y_root(i) after pointer jumping over f.
//   a_i  = V_j* where j* is the latest j < i with x_j == x_i, or x_i if there is none.
// Lane g of a G-lane group owns step i = g.  `Wg` is G ints of shared memory private to the group.
// Every lane of the warp calls this together (groups that have nothing to resolve pass valid = false): all collectives
// run under the full mask -- a sub-warp mask costs a WARPSYNC + reconvergence sequence around every shuffle / vote / match
// (19 % of the instructions of the sampling loop before this form) -- and groups are told apart by the sub-group id in
// the match key and by the shuffle width.
template <int G>
__device__ __forceinline__ int resolve_chain_group(int x, bool valid, int g, int lane, int sub, unsigned int gmask,
                                                   int N, int M, int* Wg)
{
  const unsigned long long key = valid ? (((unsigned long long)(unsigned int)sub << 32) | (unsigned int)x)
                                       : (0x8000000000000000ULL | (unsigned long long)lane);
  unsigned int m   = __match_any_sync(0xffffffffu, key) & gmask;
  unsigned int low = m & ((1u << lane) - 1u);
  bool has_prev    = valid && low != 0;
  // fast path (the common case when N >> M): no two steps read the same position and no step reads one of
  // the M tail positions that get copied -> Q is still the identity wherever it is read, a[i] = r[i].
  const unsigned int slow = __ballot_sync(0xffffffffu, has_prev || (valid && x >= N - M));
  if (slow == 0u) return x;  // warp-uniform
  int jstar        = has_prev ? (31 - __clz(low)) - sub * G : 0;
  Wg[g]            = -1;
  __syncwarp();
  if (valid && x >= N - M) {
    int t = N - 1 - x;
    if (t != g) atomicMax(&Wg[t], g);
  }
  __syncwarp();
  int f = Wg[g];
  __syncwarp();
  int p = f >= 0 ? f : g;
  for (int s = 1; s < G; s <<= 1)  // log2(G) >= log2(M) rounds: extra rounds leave the roots where they are
    p = __shfl_sync(0xffffffffu, p, p, G);
  int V  = N - 1 - p;
  int Vj = __shfl_sync(0xffffffffu, V, jstar, G);
  return has_prev ? Vj : x;
}

// position of the (n + 1)-th set bit of m (n = 0: the lowest), 32 if m has fewer
__device__ __forceinline__ unsigned int nth_set_bit(unsigned int m, int n)
{
#pragma unroll 1
  for (int i = 0; i < n; i++)
    m &= m - 1u;
  return m ? (unsigned int)(__ffs((int)m) - 1) : 32u;
}

// fan-out <= 32: G = 8/16/32 lanes per seed row.
// RNG geometry of the reference for M <= 32: block = 32 threads, 1 draw per thread, thread j of seed
// row b uses subsequence b*32 + j (func.cuh:412-447) -> lane g draws once from stream 32*b + g.
//
// One warp, 32 consecutive frontier rows (lane = row in phase 1, G lanes per row in phase 2).  Shared by the one-hop /
// per-hop kernel below and by the fused per-label sampler (multihop_fused.cuh): `b_own` is the row's index in the
// label-major concatenated frontier (it fixes the random stream, nothing else), `tag_own` what the sink records as the
// edge's source row, `off_own` where the row's edges start in the sink's arrays.
//   sink.ids(pos, tag, edge_position_in_csr)   and   sink.val(pos, neighbour)
template <typename ColT, int G, bool CHUNKED, typename Sink>
__device__ __forceinline__ void uniform_small_rows32(const ChunkRef& col, unsigned long long col_off, int M, unsigned long long seed,
                                                     const Affine* __restrict__ tab, const Affine& lane_skip, int* Wg, int lane,
                                                     long long b_own, int tag_own, long long start_own, int N_own, int off_own,
                                                     Sink& sink, unsigned long long col_policy = 0ULL)
{
  constexpr int GPW = 32 / G;
  const int g = lane & (G - 1), sub = lane / G;
  const unsigned int gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (sub * G));
  // Rows with N <= M are copied whole and draw nothing: on power-law graphs they are most of a frontier (papers100M shape,
  // hop 2: 2.9 edges per row against a fan-out of 10), so they take a path of their own -- the lane that owns the row walks
  // it, four independent reads in flight, no shuffles, no chain -- and only the rows that really sample go through the
  // G-lane steps below, packed GPW at a time.  (Before: every row paid a full step, ~175 warp instructions per GPW rows;
  // the sampling loop was 42 % of the instructions of the fused sampler.)
  const bool sampled_own = N_own > M;
  const int n_copy       = sampled_own ? 0 : N_own;
  // The copy rows of the batch are walked edge by edge ACROSS the warp, not row by row per lane.  The fused sampler turned out
  // to be bound by the number of uncoalesced memory requests an SM can issue (its insert rate per SM is the same, ~0.18 G/s,
  // whether 16 or 148 labels are in flight: profiles/r2ag_*), and the lane-per-row form cost one request per edge for the col
  // read and three more for the maj / gid / dest stores.  Here a lane owns an EDGE of the batch's concatenated copy rows: the
  // stores of a pass are three coalesced requests, the col reads one request per row touched.
  {
    int vbeg = n_copy;  // exclusive prefix of the copy counts over the warp's 32 rows
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, vbeg, o);
      if (lane >= o) vbeg += y;
    }
    const int total = __shfl_sync(0xffffffffu, vbeg, 31);
    vbeg -= n_copy;
    constexpr int kCopyW = WGB_COPY_W;
#pragma unroll 1
    for (int q0 = 0; q0 < total; q0 += 32 * kCopyW) {
      ColT v[kCopyW];
      int pos[kCopyW], tag[kCopyW];
      long long epos[kCopyW];
#pragma unroll
      for (int u = 0; u < kCopyW; u++) {
        const int q = q0 + 32 * u + lane;
        // the row that holds virtual position q: the largest r with vbeg_r <= q (rows before it with the same prefix are empty)
        int r = 0;
#pragma unroll
        for (int st = 16; st > 0; st >>= 1) {
          const int cand = r + st;
          const int bc   = __shfl_sync(0xffffffffu, vbeg, cand & 31);
          if (cand < 32 && bc <= q) r = cand;
        }
        const int vb           = __shfl_sync(0xffffffffu, vbeg, r);
        const long long startr = __shfl_sync(0xffffffffu, start_own, r);
        const int offr         = __shfl_sync(0xffffffffu, off_own, r);
        const int tagr         = __shfl_sync(0xffffffffu, tag_own, r);
        pos[u]  = -1;
        if (q < total) {
          const int k = q - vb;
          pos[u]      = offr + k;
          tag[u]      = tagr;
          epos[u]     = startr + k;
          v[u]        = load_elt_policy<ColT, CHUNKED>(col, col_off + (unsigned long long)epos[u], col_policy);
        }
      }
#pragma unroll
      for (int u = 0; u < kCopyW; u++)
        if (pos[u] >= 0) {
          sink.ids(pos[u], tag[u], epos[u]);
          sink.val(pos[u], v[u]);
        }
    }
  }
  unsigned int rem = __ballot_sync(0xffffffffu, sampled_own);
  if (rem == 0u) return;  // warp-uniform
  Affine skip_own{1ULL, 0ULL};
  if (sampled_own) skip_own = affine_skip_tab(tab, 32ULL * (unsigned long long)b_own);
  // GPW sampled rows per step.  The random `col` read of a step is its only long-latency operation; steps run in
  // chunks of kChunk whose loads are all issued before the first dependent store, so a lane keeps kChunk DRAM
  // reads in flight instead of one (the chain resolution in between is register/shuffle work only).
  constexpr int kChunk = G < 8 ? G : 8;
#pragma unroll 1
  while (rem != 0u) {
    ColT val[kChunk];
    int pos[kChunk];
    unsigned int wmask = 0;
#pragma unroll
    for (int k = 0; k < kChunk; k++) {
      if (k > 0 && rem == 0u) break;  // warp-uniform: no sampled row left for this step
      // this group's row of the step: the (sub + 1)-th of the remaining sampled rows (none: the group idles through the
      // collectives with valid = false)
      const unsigned int pick = nth_set_bit(rem, sub);
      const bool have = pick < 32u;
      const int src   = have ? (int)pick : 0;
#pragma unroll
      for (int q = 0; q < GPW; q++)
        rem &= rem - 1u;
      const int N     = __shfl_sync(0xffffffffu, N_own, src);
      const long long start = __shfl_sync(0xffffffffu, start_own, src);
      const int off   = __shfl_sync(0xffffffffu, off_own, src);
      const int tag   = __shfl_sync(0xffffffffu, tag_own, src);
      Affine row_skip;
      row_skip.g = __shfl_sync(0xffffffffu, skip_own.g, src);
      row_skip.s = __shfl_sync(0xffffffffu, skip_own.s, src);
      const long long b = __shfl_sync(0xffffffffu, b_own, src);
      pos[k] = off + g;
      const bool valid = have && g < M;
      int x            = -1;
      if (have) {
        Pcg rng;
        rng.init_with_skip(seed, 32ULL * (unsigned long long)b + (unsigned long long)g, affine_then(row_skip, lane_skip));
        const int xr = rng.next_i32();
        if (valid) x = xr % (N - g);
      }
      const int a = resolve_chain_group<G>(x, valid, g, lane, sub, gmask, N, M, Wg);
      if (valid) {
        val[k] = load_elt_policy<ColT, CHUNKED>(col, col_off + (unsigned long long)(start + a), col_policy);
        sink.ids(off + g, tag, start + a);
        wmask |= 1u << k;
      }
      __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < kChunk; k++)
      if ((wmask >> k) & 1u) sink.val(pos[k], val[k]);
  }
}

template <typename ColT>
struct OneHopSink {
  ColT* __restrict__ out;
  int* __restrict__ lid;
  long long* __restrict__ gid;
  __device__ __forceinline__ void ids(int pos, int tag, long long edge_pos) const
  {
    if (lid) lid[pos] = tag;
    if (gid) gid[pos] = edge_pos;
  }
  __device__ __forceinline__ void val(int pos, ColT v) const { out[pos] = v; }
};

template <typename IdT, typename ColT, int G, bool CHUNKED>
__global__ void __launch_bounds__(256) uniform_small_kernel(ChunkRef row_ptr, unsigned long long row_ptr_off,
                                                            ChunkRef col, unsigned long long col_off,
                                                            const IdT* __restrict__ centers, int n, int M,
                                                            unsigned long long seed, const int* __restrict__ offsets,
                                                            ColT* __restrict__ out, int* __restrict__ lid,
                                                            long long* __restrict__ gid, const Affine* __restrict__ tab,
                                                            const int* __restrict__ n_dev = nullptr)
{
  if (n_dev) n = *n_dev;
  __shared__ int W[8][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = lane & (G - 1), sub = lane / G;
  const Affine lane_skip   = affine_skip_loop((unsigned long long)g);
  int* Wg                  = &W[wib][sub * G];
  const long long nwarps   = (long long)gridDim.x * 8;
  OneHopSink<ColT> sink{out, lid, gid};
  for (long long batch = (long long)blockIdx.x * 8 + wib; batch * 32 < n; batch += nwarps) {
    // phase 1: one lane per seed row -- 32 independent row_ptr reads in flight per warp
    const long long b_own = batch * 32 + lane;
    long long start_own   = 0;
    int N_own = 0, off_own = 0;
    if (b_own < n) {
      unsigned long long node = (unsigned long long)centers[b_own];
      start_own     = load_i64<CHUNKED>(row_ptr, row_ptr_off + node);
      long long end = load_i64<CHUNKED>(row_ptr, row_ptr_off + node + 1);
      N_own         = (int)(end - start_own);
      off_own       = offsets[b_own];
    }
    // phase 2: G lanes per row
    uniform_small_rows32<ColT, G, CHUNKED>(col, col_off, M, seed, tab, lane_skip, Wg, lane, b_own, (int)b_own, start_own, N_own, off_own, sink);
  }
}

// 32 < fan-out <= 1024: one CTA per seed row, chain resolved in shared memory
// (sorted (x, i) pairs give j*, atomicMax gives f, double-buffered pointer jumping gives the roots).
__device__ const int kWarpCountTab[32] = {1, 1, 1, 2, 2, 2, 4, 4, 4, 4, 4, 4, 8, 8, 8, 8,
                                          8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8};
__device__ const int kItemsTab[32]     = {1, 2, 3, 2, 3, 3, 2, 2, 3, 3, 3, 3, 2, 2, 2, 2,
                                          3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4};
constexpr int kGeneralBlock = 128;

__device__ __forceinline__ void bitonic_sort_smem(unsigned long long* keys, int P2, bool descending)
{
  for (int k = 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          bool up = ((i & k) == 0) != descending;
          unsigned long long a = keys[i], b = keys[ixj];
          if ((a > b) == up) {
            keys[i]   = b;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

template <typename IdT, typename ColT, bool CHUNKED>
__global__ void __launch_bounds__(kGeneralBlock) uniform_general_kernel(
  ChunkRef row_ptr, unsigned long long row_ptr_off, ChunkRef col, unsigned long long col_off,
  const IdT* __restrict__ centers, int n, int M, unsigned long long seed, const int* __restrict__ offsets,
  ColT* __restrict__ out, int* __restrict__ lid, long long* __restrict__ gid, const Affine* __restrict__ tab,
  const int* __restrict__ n_dev = nullptr)
{
  if (n_dev) n = *n_dev;
  __shared__ unsigned long long keys[1024];
  __shared__ int xs[1024];
  __shared__ int prev[1024];
  __shared__ int ptr[2][1024];
  const int func_idx = (M - 1) / 32;
  const int T        = kWarpCountTab[func_idx] * 32;
  const int ipt      = kItemsTab[func_idx];
  const int P        = T * ipt;
  int P2             = 64;
  while (P2 < P)
    P2 <<= 1;
  const int tid = threadIdx.x;
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    unsigned long long node = (unsigned long long)centers[b];
    long long start = load_i64<CHUNKED>(row_ptr, row_ptr_off + node);
    long long end   = load_i64<CHUNKED>(row_ptr, row_ptr_off + node + 1);
    int N           = (int)(end - start);
    int off         = offsets[b];
    if (N <= 0) continue;
    if (N <= M) {
      for (int j = tid; j < N; j += blockDim.x) {
        out[off + j] = load_elt<ColT, CHUNKED>(col, col_off + (unsigned long long)(start + j));
        if (lid) lid[off + j] = b;
        if (gid) gid[off + j] = start + j;
      }
      continue;
    }
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
    for (int k = tid + 1; k < P2; k += blockDim.x) {
      unsigned long long kk = keys[k], kp = keys[k - 1];
      unsigned int i = (unsigned int)kk;
      if (kk != ~0ULL && i < (unsigned int)M && (kk >> 32) == (kp >> 32)) prev[i] = (int)(unsigned int)kp;
    }
    for (int i = tid; i < M; i += blockDim.x) {
      int x = xs[i];
      if (x >= N - M) {
        int t = N - 1 - x;
        if (t != i) atomicMax(&ptr[0][t], i);
      }
    }
    __syncthreads();
    for (int i = tid; i < M; i += blockDim.x) {
      int f = ptr[0][i];
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
      int a  = pj >= 0 ? N - 1 - ptr[cur][pj] : xs[i];
      out[off + i] = load_elt<ColT, CHUNKED>(col, col_off + (unsigned long long)(start + a));
      if (lid) lid[off + i] = b;
      if (gid) gid[off + i] = start + a;
    }
    __syncthreads();
  }
}

// ---- weighted (A-Res) sampling -------------------------------------------------------------------------
// key = log2(u)/w, keep the M largest (weighted_sample_without_replacement_func.cuh:32-51, 208-291).
__device__ __forceinline__ unsigned int float_order_bits(float f)
{
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <typename WT>
__device__ __forceinline__ float gen_key_from_weight_dev(WT weight, Pcg& rng)
{
  float u = rng.next_float();
  u       = -(0.5f + 0.5f * u);
  unsigned long long r2;
  int extra = -1;
  do {
    r2 = rng.next_u64();
    extra++;
  } while (!r2);
  int one_bit = __clzll((long long)r2) + extra * 64;
  u           = scalbnf(u, -one_bit);
  return (log1pf(u) / logf(2.0f)) * (1.0f / (float)weight);
}

// One CTA of BLOCK threads per seed row; BLOCK equals the reference's block size (128, or 256 when
// M > 256) so that thread j IS the reference's thread j: stream b*BLOCK + j, elements j, j+BLOCK, ...
// Top-M selection: streaming filter against the current M-th largest key, candidates compacted by
// an in-shared-memory bitonic sort whenever the buffer could overflow (replaces RAFT block warpsort).
template <typename IdT, typename ColT, typename WT, int BLOCK, bool CHUNKED>
__global__ void __launch_bounds__(BLOCK) weighted_kernel(ChunkRef row_ptr, unsigned long long row_ptr_off, ChunkRef col,
                                                         unsigned long long col_off, ChunkRef wgt,
                                                         unsigned long long wgt_off, const IdT* __restrict__ centers,
                                                         int n, int M, unsigned long long seed,
                                                         const int* __restrict__ offsets, ColT* __restrict__ out,
                                                         int* __restrict__ lid, long long* __restrict__ gid,
                                                         const Affine* __restrict__ tab,
                                                         const int* __restrict__ n_dev = nullptr)
{
  if (n_dev) n = *n_dev;
  constexpr int C = 2048;
  __shared__ unsigned long long cand[C];
  __shared__ int s_cnt;
  __shared__ unsigned int s_thr;
  const int tid = threadIdx.x;
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    unsigned long long node = (unsigned long long)centers[b];
    long long start = load_i64<CHUNKED>(row_ptr, row_ptr_off + node);
    long long end   = load_i64<CHUNKED>(row_ptr, row_ptr_off + node + 1);
    int N           = (int)(end - start);
    int off         = offsets[b];
    if (N <= 0) continue;
    if (N <= M) {
      for (int j = tid; j < N; j += BLOCK) {
        out[off + j] = load_elt<ColT, CHUNKED>(col, col_off + (unsigned long long)(start + j));
        if (lid) lid[off + j] = b;
        if (gid) gid[off + j] = start + j;
      }
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
      int id = r * BLOCK + tid;
      if (id < N) {
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
      int id       = (int)(0xffffffffu - (unsigned int)(cand[i] & 0xffffffffULL));
      out[off + i] = load_elt<ColT, CHUNKED>(col, col_off + (unsigned long long)(start + id));
      if (lid) lid[off + i] = b;
      if (gid) gid[off + i] = start + id;
    }
    __syncthreads();
  }
}

}  // namespace wgb
