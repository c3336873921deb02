// S0, fused form: every hop of a call group in ONE persistent kernel, one thread-block cluster per label (mini-batch).
// Included by multihop.cu (needs its sampler object); same results, bit for bit, as the per-hop kernel chain there.
//
// Why: the kernel chain of multihop.cu (count+scan -> plan -> reinsert -> sample -> insert -> compact per hop, ~16
// launches) streams every intermediate of the call group (2.5 M edges x ~40 B) through DRAM between kernels and keeps
// one 16 B/slot hash table for all labels (94 MB at the bench shape: beyond L2) -- measured 7x the algorithmic DRAM
// traffic.  A label (1024 seeds -> <= 282 k vertices at fan-out [25,10], ~40 k in practice) is an independent unit of
// work; everything it produces between its hops fits in L2 next to the other labels in flight.  So:
//   * a cluster of CL CTAs (CL = 1..8, chosen so that labels-in-flight x CL covers the 148 SMs; call groups of >= 74
//     labels run one CTA per label) takes labels by ticket and runs seeds -> [count -> sample -> dedup -> compact] x hops
//     for its label, phases separated by cluster barriers (CL > 1; per-CTA partial sums travel through distributed shared
//     memory) or by the CTA barrier (CL = 1) instead of kernel boundaries;
//   * the label's hash table is private (no label in the key): 8-byte slots {vertex:32 | aux:32}, buckets of four = one
//     32-byte sector read by one 256-bit load, sized from the label's true counts every hop and memset by the cluster
//     itself (measured at 148 labels in flight: the call group's 325 MB of scratch does NOT stay in the 126 MB L2 -- 40 %
//     sector hit rate -- see profiles/r2_sampler_latency_experiments.txt for what that does and does not cost);
//   * all per-label scratch is laid out label-major (label l's vertices at lo[l] * fstride, edges at lo[l] * estride):
//     the label's vertex array IS its renumber map (seeds first, then every hop's new vertices in first-occurrence
//     order), its edge arrays ARE the output segments -- _finish is a segmented copy;
//   * the only cross-label dependency is the random stream geometry (row b of the label-major concatenated frontier
//     draws from subsequence 32 b + lane): the base row index of a label at a hop is a decoupled look-back over the
//     labels before it (flag | count words, labels are taken in ticket order so every predecessor is resident or done).
// Eligibility (multihop_begin decides): homogeneous, uniform, not temporal, graph on this GPU (world 1), every fan-out
// in [1, 32], |V| < 2^32 - 1.  Everything else takes the kernel chain.
// The file is included twice by multihop.cu: first for the device half (before the sampler object, whose pending-call state
// holds an FzArgs), then with WGB_FZ_HOST_HALF defined for the host half (after MhCall / MhOutCtx exist).
#ifndef WGB_HOST_EMULATION
#ifndef WGB_FZ_HOST_HALF

#include <cooperative_groups.h>

namespace wgb {

namespace cg = cooperative_groups;

#ifndef WGB_FZ_THREADS
#define WGB_FZ_THREADS 1024
#endif
#ifndef WGB_FZ_ILP
#define WGB_FZ_ILP 1
#endif
#ifndef WGB_FZ_BOUND
#define WGB_FZ_BOUND 1024  // launch bound: the register cap of the kernel is 65536 / WGB_FZ_BOUND
#endif
constexpr int kFzThreads            = WGB_FZ_THREADS;  // 32 warps x 64 registers = the SM's register file.  Measured on C4 (profiles/
                                                        // run_r2p.sh): 1024 threads 0.340 ms per 64-label call group, 896 0.357, 768 0.382, 640 0.385
                                                        // (wider register budgets do not pay for the lost warps: the kernel is latency bound)
constexpr unsigned int kFzPending   = 0x80000000u;
constexpr unsigned long long kFzEmpty = ~0ULL;
constexpr unsigned long long kFzFlagAgg = 1ULL << 62, kFzFlagPrefix = 2ULL << 62, kFzValMask = (1ULL << 62) - 1;

struct FzArgs {
  ChunkRef row_ptr, col;  // world 1
  unsigned long long row_ptr_off, col_off;
  const void* seeds;
  int seed_is64;
  const long long* label_offsets;
  int B, L;
  int fanout[kMaxHops];
  long long fstride, estride;  // upper bounds per seed: vertices (1 + estride), edges
  unsigned long long random_state;
  // label-major scratch (label l: vertices at lo[l] * fstride, edges at lo[l] * estride, table at fz_table_base(l))
  long long* F;           // vertices = renumber map segments
  int* Orow;              // per source row: label-local offset of its first edge
  long long* Rstart;      // per source row: first position of its adjacency in the CSR  } written by the count phase, read by the
  int* Rdeg;              // per source row: its degree                                   } sampling phase (coalesced; saves the
                          //                                                               second random row_ptr read of every row)
  int* maj;               // per edge: local id of its source row
  int* mnr;               // per edge: local id of its endpoint
  long long* gid;         // per edge: position in the CSR
  void* dest;             // per edge: endpoint (global id, ColT)
  unsigned int* aux;      // per edge: slot index, then the slot's aux word
  unsigned int* rank_of;  // per edge that is a first occurrence: rank among the hop's new vertices
  unsigned long long* table;
  int* seed_local;          // [S] local id of every input seed
  int* n_step;              // [B][L+1] vertices discovered at step t
  int* e_hop;               // [B][L] edges of hop h
  unsigned long long* pub;  // [L][B] look-back words: frontier rows of (hop, label)
  unsigned int* ticket;
  long long* bad_seed;   // set when a seed is outside [0, V): the call fails at _finish
  unsigned long long V;
  long long* dbg;        // WGB_MH_TIMING: [B][kFzDbgSlots] global-timer stamps of the label's phases (else null)
  const Affine* tab;
};

// buckets (of four slots) for a table that receives at most `keys` distinct keys: two slots per key (load <= 0.5).  Measured
// with 1.5 slots per key (WGB_FZ_SLOTS_X8 = 12, a quarter less table): clear 9.5 -> 7.1 us but insert 107 -> 118 us per label
// (longer bucket walks), 0.558 against 0.550 ms per call group -- the smaller footprint does not pay for the extra probes.
#ifndef WGB_FZ_SLOTS_X8
#define WGB_FZ_SLOTS_X8 16  // slots per key, in eighths
#endif
__device__ __forceinline__ unsigned int fz_buckets(unsigned int keys)
{
  return (unsigned int)(((unsigned long long)keys * WGB_FZ_SLOTS_X8) >> 5) + 8u;
}

__device__ __forceinline__ long long fz_table_base(long long lo_l, long long fstride, int l)
{
  return 4 * ((lo_l * fstride + 1) / 2 + 16LL * l);  // in slots; a multiple of 4 (bucket = 32-byte sector)
}

__device__ __forceinline__ unsigned int fz_home(unsigned int v, unsigned int nb)
{
  unsigned int h = v * 0x9E3779B1u;
  h ^= h >> 15;
  h *= 0x85EBCA77u;
  h ^= h >> 13;
  return (unsigned int)(((unsigned long long)h * nb) >> 32);
}

struct FzProbe {
  unsigned long long w[4];  // the key's home bucket as it was when the probe was issued
  unsigned int v;
};

// L2 residency.  With one label per SM the call group's scratch (~2 MB per label: tables, vertex and edge arrays) is 300 MB,
// more than the 126 MB L2, and the random table traffic (probe, CAS, first-occurrence reads) ran at the rate of random DRAM
// sectors: making the loops latency tolerant changed nothing (profiles/r2y_*).  The tables are what is touched at random and
// repeatedly (0.46 MB per label at hop 2, 69 MB per call group), so every table access carries an evict-last policy and the
// single-use random reads of the graph (col_idx) an evict-first one; the streamed scratch arrays keep the default.
struct FzPolicy {
  unsigned long long keep, stream;
};
__device__ __forceinline__ FzPolicy fz_make_policy()
{
  FzPolicy p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.stream));
  return p;
}

__device__ __forceinline__ void fz_load_bucket(const unsigned long long* __restrict__ tbl, unsigned int b, unsigned long long (&w)[4],
                                               unsigned long long pol)
{
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
               : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3])
               : "l"(tbl + 4ULL * b), "l"(pol)
               : "memory");
}
__device__ __forceinline__ unsigned long long fz_load_slot(const unsigned long long* p, unsigned long long pol)
{
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void fz_store_slot(unsigned long long* p, unsigned long long v, unsigned long long pol)
{
  asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
// (atom.cas takes no cache hint -- ptxas: "Illegal modifier '.L2::cache_hint' for instruction 'atom'" -- unlike atom.add / red; the
// sector it hits was brought in by the probe with the evict-last policy a moment earlier)
__device__ __forceinline__ unsigned long long fz_cas_slot(unsigned long long* p, unsigned long long expect, unsigned long long desired,
                                                          unsigned long long /*pol*/)
{
  return atomicCAS(p, expect, desired);
}
__device__ __forceinline__ void fz_min_aux(unsigned long long* slot, unsigned int v, unsigned long long pol)
{
  // little endian: the aux word is the low half of the slot
  asm volatile("red.relaxed.gpu.global.min.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(reinterpret_cast<unsigned int*>(slot)), "r"(v), "l"(pol) : "memory");
}

__device__ __forceinline__ FzProbe fz_probe(const unsigned long long* __restrict__ tbl, unsigned int nb, unsigned int v, unsigned long long pol)
{
  FzProbe p;
  p.v = v;
  fz_load_bucket(tbl, fz_home(v, nb), p.w, pol);
  return p;
}

// claim-or-find `p.v` in a table of nb buckets, starting from its already loaded home bucket; the slot's aux after this
// thread's visit is at most `mine`.  Returns the slot.
__device__ __forceinline__ unsigned int fz_upsert(unsigned long long* __restrict__ tbl, unsigned int nb, FzProbe p, unsigned int mine,
                                                  unsigned long long pol)
{
  const unsigned int v = p.v;
  unsigned int b       = fz_home(v, nb);
  const unsigned long long fresh = ((unsigned long long)v << 32) | mine;
  while (true) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      unsigned long long cur = p.w[j];
      if (cur == kFzEmpty) {
        cur = fz_cas_slot(&tbl[4ULL * b + j], kFzEmpty, fresh, pol);
        if (cur == kFzEmpty) return 4u * b + j;
      }
      if ((unsigned int)(cur >> 32) == v) {
        if ((unsigned int)cur > mine) fz_min_aux(&tbl[4ULL * b + j], mine, pol);
        return 4u * b + j;
      }
    }
    b = b + 1 == nb ? 0u : b + 1;
    fz_load_bucket(tbl, b, p.w, pol);
  }
}

struct FzShared {
  unsigned int warp[34];
  unsigned long long xchg[2][2];  // [parity][0] = this CTA's partial sum of the current exchange
  long long rowbase;              // CTA 0: first row of this label in the label-major frontier of the next hop
  int label;
  int W[kFzThreads / 32][32];
};

constexpr int kFzDbgSlots = 32;
constexpr int kFzIlp = WGB_FZ_ILP;  // items a thread keeps in flight per stage of a latency-bound loop
#ifndef WGB_FZ_ILP_HASH
#define WGB_FZ_ILP_HASH 1
#endif
constexpr int kFzIlpHash = WGB_FZ_ILP_HASH;  // the same, for the three per-hop loops over the hash table (re-insert, insert, local ids)
#ifndef WGB_FZ_ILP_SCAN
#define WGB_FZ_ILP_SCAN 1
#endif
constexpr int kFzIlpScan = WGB_FZ_ILP_SCAN;  // 32-item chunks a warp keeps in flight in both passes of the ordered scans
#ifndef WGB_FZ_INSERT_STREAM
#define WGB_FZ_INSERT_STREAM 1  // 1: the three insert loops run lane-decoupled in a function of their own (fz_insert_stream); 0: FZ_HASH_LOOP
#endif
#ifndef WGB_FZ_SAMPLE_INLINE
#define WGB_FZ_SAMPLE_INLINE 0
#endif
#if WGB_FZ_SAMPLE_INLINE
#define FZ_SAMPLE_FN __forceinline__
#else
#define FZ_SAMPLE_FN __noinline__
#endif
#ifndef WGB_FZ_PIPE
#define WGB_FZ_PIPE 1  // 1: those loops run software-pipelined (fz_for_each_pipe); 0: staged with kFzIlpHash items per thread
#endif
#if WGB_FZ_PIPE
#define FZ_HASH_LOOP fz_for_each_pipe
#else
#define FZ_HASH_LOOP fz_for_each<kFzIlpHash>
#endif

struct FzCluster {
  FzShared* sh;
  unsigned int rank, size;
  unsigned int xround;

  // `mine` = this CTA's partial sum (same value in every thread); returns the sum over the CTAs before this one, `total`
  // over all.  One cluster barrier.
  __device__ __forceinline__ unsigned long long exchange(unsigned long long mine, unsigned long long& total)
  {
    const unsigned int par = xround & 1u;
    xround++;
    if (threadIdx.x == 0) sh->xchg[par][0] = mine;
    sync();
    unsigned long long before = 0;
    total                     = 0;
    for (unsigned int r = 0; r < size; r++) {
      const unsigned long long t = size == 1 ? sh->xchg[par][0] : cg::this_cluster().map_shared_rank(&sh->xchg[par][0], r)[0];
      if (r < rank) before += t;
      total += t;
    }
    return before;
  }
  // a one-CTA cluster (call groups of >= 74 labels: one label per SM) synchronises with the CTA barrier: no cluster-scope
  // release (MEMBAR.ALL.GPU + barrier.cluster were 25 % of the kernel's stall samples at cluster size 2)
  __device__ __forceinline__ void sync()
  {
    if (size == 1) __syncthreads();
    else cg::this_cluster().sync();
  }
};

// Ordered exclusive scan over n items spread over the cluster.  CTA c owns the contiguous chunk [c * chunk, (c+1) * chunk),
// warp w of it a contiguous 1/32 of that; a warp walks its part 32 items at a time (coalesced), so the whole scan costs ONE
// block-wide scan (of the 32 warp totals) and one cluster barrier, whatever n is.
// Pass 1 runs in three stages, kFzIlp items per thread per stage, so that the long-latency loads of a stage are all in
// flight before the first dependent instruction:
//   a(i) -> A            first (address-generating) load
//   b(i, A) -> Bv        the long-latency load(s), nothing but loads
//   c(i, Bv) -> value    arithmetic + side effects (called once per item)
// Pass 2:  again(i) -> the same value (cheap form);  out(i, exclusive_prefix, value).
// Returns the total.
template <typename StA, typename StB, typename StC, typename Again, typename Out>
__device__ __forceinline__ unsigned int fz_ordered_scan(FzCluster& c, long long n, StA sa, StB sb, StC sc, Again again, Out out)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long chunk = ((n + c.size - 1) / c.size + kFzThreads - 1) / kFzThreads * kFzThreads;
  const long long beg   = (long long)c.rank * chunk < n ? (long long)c.rank * chunk : n;
  const long long end   = beg + chunk < n ? beg + chunk : n;
  const long long wlen  = chunk / (kFzThreads / 32);  // a multiple of 32
  const long long wbeg  = beg + wid * wlen < end ? beg + wid * wlen : end;
  const long long wend  = wbeg + wlen < end ? wbeg + wlen : end;
  unsigned int mine = 0;
#pragma unroll 1
  for (long long i0 = wbeg + lane; i0 < wend; i0 += 32 * kFzIlpScan) {
    decltype(sa(0LL)) A[kFzIlpScan];
    decltype(sb(0LL, A[0])) Bv[kFzIlpScan];
#pragma unroll
    for (int u = 0; u < kFzIlpScan; u++)
      if (i0 + 32 * u < wend) A[u] = sa(i0 + 32 * u);
#pragma unroll
    for (int u = 0; u < kFzIlpScan; u++)
      if (i0 + 32 * u < wend) Bv[u] = sb(i0 + 32 * u, A[u]);
#pragma unroll
    for (int u = 0; u < kFzIlpScan; u++)
      if (i0 + 32 * u < wend) mine += sc(i0 + 32 * u, Bv[u]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if (lane == 0) c.sh->warp[wid] = mine;
  __syncthreads();
  if (wid == 0) {
    const unsigned int w = lane < kFzThreads / 32 ? c.sh->warp[lane] : 0u;
    unsigned int wi      = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned int y = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += y;
    }
    c.sh->warp[lane] = wi - w;
    if (lane == 31) c.sh->warp[32] = wi;
  }
  __syncthreads();
  const unsigned int wbase = c.sh->warp[wid];
  unsigned long long total;
  unsigned int carry = (unsigned int)c.exchange((unsigned long long)c.sh->warp[32], total) + wbase;  // (barrier: warp[] is free again)
#pragma unroll 1
  for (long long i0 = wbeg; i0 < wend; i0 += 32 * kFzIlpScan) {
    unsigned int vs[kFzIlpScan];
#pragma unroll
    for (int u = 0; u < kFzIlpScan; u++) {  // the reads of kFzIlpScan chunks are in flight together
      const long long i = i0 + 32 * u + lane;
      vs[u]             = i < wend ? again(i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kFzIlpScan; u++) {
      const long long i    = i0 + 32 * u + lane;
      const unsigned int v = vs[u];
      unsigned int inc     = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
      }
      if (i < wend) out(i, carry + inc - v, v);
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  return (unsigned int)total;
}

// a plain loop over n items of the cluster in the same staged form (no ordering between items)
template <int ILP = kFzIlp, typename StA, typename StB, typename StC>
__device__ __forceinline__ void fz_for_each(const FzCluster& c, int n, StA sa, StB sb, StC sc)
{
  const int CT = (int)c.size * kFzThreads, cti = (int)c.rank * kFzThreads + (int)threadIdx.x;
#pragma unroll 1
  for (int i0 = cti; i0 < n; i0 += CT * ILP) {
    decltype(sa(0)) A[ILP];
    decltype(sb(0, A[0])) Bv[ILP];
#pragma unroll
    for (int u = 0; u < ILP; u++)
      if (i0 + CT * u < n) A[u] = sa(i0 + CT * u);
#pragma unroll
    for (int u = 0; u < ILP; u++)
      if (i0 + CT * u < n) Bv[u] = sb(i0 + CT * u, A[u]);
#pragma unroll
    for (int u = 0; u < ILP; u++)
      if (i0 + CT * u < n) sc(i0 + CT * u, Bv[u]);
  }
}

// The same loop, software-pipelined: while item i runs its last (longest) stage -- the compare-and-swap chain of an insert --
// the bucket of item i + CT is already being read and the value of item i + 2 CT is on its way, so an iteration exposes one
// round trip instead of three.  (Running two items in lock step instead, WGB_FZ_ILP_HASH = 2, did not pay: 0.358 against
// 0.338 ms per call group; the CAS chains still ran one after the other.)  A bucket read ahead of the thread's own previous
// insert can be stale; fz_upsert tolerates stale views by construction (a slot only ever goes from empty to one key, and a
// failed CAS returns the slot's current contents).
template <typename StA, typename StB, typename StC>
__device__ __forceinline__ void fz_for_each_pipe(const FzCluster& c, int n, StA sa, StB sb, StC sc)
{
  const int CT = (int)c.size * kFzThreads;
  int i        = (int)c.rank * kFzThreads + (int)threadIdx.x;
  if (i >= n) return;
  auto A2 = sa(i);
  auto B0 = sb(i, A2);
  if (i + CT < n) A2 = sa(i + CT);
#pragma unroll 1
  for (; i < n; i += CT) {
    auto B1 = B0;
    if (i + CT < n) B1 = sb(i + CT, A2);
    if (i + 2 * CT < n) A2 = sa(i + 2 * CT);
    sc(i, B0);
    B0 = B1;
  }
}

// Inserts, lane-decoupled.  A warp that runs 32 inserts in lock step pays, per round, for its SLOWEST lane -- the one whose home
// bucket is full and who walks two or three buckets (probe + CAS round trips each) -- and at 32 lanes some lane nearly always
// does: the insert loops spent ~6 dependent round trips per item although an item needs ~2 (phase clock: 112 us for 20 items per
// thread; neither more items in flight, nor prefetching, nor L2 residency hints changed that, profiles/r2y_*, r2aa_*).  Here a
// lane that has finished its item starts its next one while its neighbours are still probing: every trip of the loop is ONE
// bucket read (+ one CAS where a slot is claimed) for every lane that still has work.
// Not inlined, on purpose: fz_label_kernel holds ~40 values live across its phases and is capped at 64 registers, so inside the
// inlined form of this loop every trip began with three spill reloads (LDL) in front of the bucket address -- local memory is
// 0.7 MB per SM here, far beyond L1, so those reloads were L2 round trips on the critical path of EVERY trip, which is why no
// latency-hiding change to the loops moved the phase times (profiles/r2y_*, r2aa_*, r2ab_*).  As a function of its own the loop's
// state is its arguments.
//   keys[i] -> vertex;  aux word to install / lower to = pending | i;  slot_out[i] = slot (when slot_out != nullptr)
template <typename KeyT>
__device__ __noinline__ void fz_insert_stream(unsigned long long* __restrict__ tbl, unsigned int nb, unsigned long long pol,
                                              const KeyT* __restrict__ keys, int n, unsigned int pending, unsigned int* __restrict__ slot_out,
                                              int first, int stride)
{
  int i        = first;
  bool live    = i < n;
  unsigned int v = 0, b = 0, vn = 0;
  if (live) {
    v = (unsigned int)keys[i];
    b = fz_home(v, nb);
    if (i + stride < n) vn = (unsigned int)keys[i + stride];  // the next item's key is always one item ahead
  }
  while (__any_sync(0xffffffffu, live)) {
    if (live) {
      unsigned long long w[4];
      fz_load_bucket(tbl, b, w, pol);
      const unsigned int m           = pending | (unsigned int)i;
      const unsigned long long fresh = ((unsigned long long)v << 32) | m;
      int slot                       = -1;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (slot < 0) {
          unsigned long long cur = w[j];
          if (cur == kFzEmpty) {
            cur = atomicCAS(&tbl[4ULL * b + j], kFzEmpty, fresh);
            if (cur == kFzEmpty) slot = (int)(4u * b + j);
          }
          if (slot < 0 && (unsigned int)(cur >> 32) == v) {
            if ((unsigned int)cur > m) fz_min_aux(&tbl[4ULL * b + j], m, pol);
            slot = (int)(4u * b + j);
          }
        }
      }
      if (slot >= 0) {
        if (slot_out) slot_out[i] = (unsigned int)slot;
        i += stride;
        live = i < n;
        if (live) {
          v = vn;
          b = fz_home(v, nb);
          if (i + stride < n) vn = (unsigned int)keys[i + stride];
        }
      } else {
        b = b + 1 == nb ? 0u : b + 1;
      }
    }
  }
}

// decoupled look-back over the labels before `l` (one warp): publishes this label's count, returns the sum of the counts
// of labels 0..l-1
__device__ __forceinline__ unsigned long long fz_lookback(unsigned long long* state, int l, unsigned long long agg)
{
  const int lane               = threadIdx.x & 31;
  unsigned long long exclusive = 0;
  if (l == 0) {
    if (lane == 0) st_relaxed_u64(&state[0], kFzFlagPrefix | agg);
  } else {
    if (lane == 0) st_relaxed_u64(&state[l], kFzFlagAgg | agg);
    int idx = l - 1;
    while (true) {
      const int t = idx - lane;
      unsigned long long w;
      do {
        w = (t >= 0) ? ld_relaxed_u64(&state[t]) : kFzFlagPrefix;
      } while (__any_sync(0xffffffffu, (w >> 62) == 0));
      const unsigned int pm  = __ballot_sync(0xffffffffu, (w >> 62) == 2);
      unsigned long long val = w & kFzValMask;
      if (pm) {
        const int first = __ffs(pm) - 1;
        if (lane > first) val = 0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        val += __shfl_xor_sync(0xffffffffu, val, o);
      exclusive += val;
      if (pm) break;
      idx -= 32;
    }
    if (lane == 0) st_relaxed_u64(&state[l], kFzFlagPrefix | ((exclusive + agg) & kFzValMask));
  }
  return exclusive;
}

template <typename ColT>
struct FzSink {
  ColT* __restrict__ dest;
  int* __restrict__ maj;
  long long* __restrict__ gid;
  __device__ __forceinline__ void ids(int pos, int tag, long long edge_pos) const
  {
    maj[pos] = tag;
    gid[pos] = edge_pos;
  }
  __device__ __forceinline__ void val(int pos, ColT v) const { dest[pos] = v; }
};

// (not inlined, for the same reason as fz_insert_stream: the sampling loop gets a register allocation of its own)
template <typename ColT, int G>
__device__ FZ_SAMPLE_FN void fz_sample_rows(const FzArgs& a, const FzCluster& c, const long long* __restrict__ Rs, const int* __restrict__ Rd,
                                            const int* __restrict__ Ol,
                                               int nbase, int n_rows, int ebase, long long rowbase, int M, unsigned long long seed,
                                               FzSink<ColT>& sink, unsigned long long col_policy)
{
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = lane & (G - 1), sub = lane / G;
  const Affine lane_skip = affine_skip_loop((unsigned long long)g);
  int* Wg                = &c.sh->W[wib][sub * G];
  const int warps        = (int)c.size * (kFzThreads / 32);
  // the extents of batch k + 1 are read (three coalesced loads) before batch k is sampled: with the sampling phase alone in its
  // barrier interval the chain extents -> col read -> stores is what a warp waits for, batch after batch
  auto load_rows = [&](int batch, long long& start, int& N, int& off) {
    const int r = batch * 32 + lane;
    start       = 0;
    N = off = 0;
    if ((long long)batch * 32 < n_rows && r < n_rows) {
      start = Rs[nbase + r];
      N     = Rd[nbase + r];
      off   = Ol[nbase + r] - ebase;  // relative to the hop's first edge; the sink's arrays start there
    }
  };
  int batch = (int)c.rank * (kFzThreads / 32) + wib;
  long long start_nxt;
  int N_nxt, off_nxt;
  load_rows(batch, start_nxt, N_nxt, off_nxt);
  for (; (long long)batch * 32 < n_rows; batch += warps) {
    const int r               = batch * 32 + lane;
    const long long start_own = start_nxt;
    const int N_own = N_nxt, off_own = off_nxt;
    load_rows(batch + warps, start_nxt, N_nxt, off_nxt);
    uniform_small_rows32<ColT, G, false>(a.col, a.col_off, M, seed, a.tab, lane_skip, Wg, lane, rowbase + r, nbase + r, start_own, N_own,
                                         off_own, sink, col_policy);
  }
}

// launch bound 1024 (the CTA is kFzThreads = 896 wide): caps the kernel at 64 registers per thread, so that 28 warps take
// 56 K of the SM's 64 K registers and the bulk-copy gather CTA (4 warps x 64) fits beside it
template <typename ColT>
__global__ void __launch_bounds__(WGB_FZ_BOUND, 1) fz_label_kernel(const __grid_constant__ FzArgs a)
{
  __shared__ FzShared sh;
  FzCluster c;
  c.sh     = &sh;
  c.rank   = cg::this_cluster().block_rank();
  c.size   = cg::this_cluster().num_blocks();
  c.xround = 0;
  const int tid          = threadIdx.x;
  const unsigned int CT  = c.size * kFzThreads;        // threads of the cluster
  const unsigned int cti = c.rank * kFzThreads + tid;  // this thread's index in the cluster
  const int L            = a.L, B = a.B;
  const FzPolicy pol     = fz_make_policy();

  while (true) {
    // ---- next label (tickets are handed out in label order) ------------------------------------------------
    if (c.rank == 0 && tid == 0) sh.label = (int)atomicAdd(a.ticket, 1u);
    c.sync();
    const int l = c.size == 1 ? sh.label : *cg::this_cluster().map_shared_rank(&sh.label, 0);
    c.sync();  // everybody has read it before CTA 0 takes the next ticket
    if (l >= B) return;

    const long long lo_l = a.label_offsets[l];
    const int n_seeds    = (int)(a.label_offsets[l + 1] - lo_l);
    // the label's slices of the scratch arrays: base pointers come from the parameter block every time they are used (two
    // 64-bit offsets stay live across the label instead of eleven pointers -- the kernel is capped at 64 registers)
    const long long vbeg = lo_l * a.fstride, ebeg = lo_l * a.estride;
#define Fl (a.F + vbeg)
#define Ol (a.Orow + vbeg)
#define majl (a.maj + ebeg)
#define mnrl (a.mnr + ebeg)
#define gidl (a.gid + ebeg)
#define destl (static_cast<ColT*>(a.dest) + ebeg)
#define auxl (a.aux + ebeg)
#define rankl (a.rank_of + ebeg)
#define tb (a.table + tbeg)
#define n_step (a.n_step + (long long)l * (L + 1))
#define e_hop (a.e_hop + (long long)l * L)
    const long long tbeg = fz_table_base(lo_l, a.fstride, l);

    // optional phase clock (WGB_MH_TIMING): thread 0 of the cluster stamps the global timer after every phase
#define FZ_T(p)                                                                            \
  do {                                                                                     \
    if (a.dbg && cti == 0 && (p) < kFzDbgSlots) {                                          \
      unsigned long long t_;                                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                               \
      a.dbg[(long long)l * kFzDbgSlots + (p)] = (long long)t_;                             \
    }                                                                                      \
  } while (0)
    FZ_T(0);
    // ---- step 0: the label's distinct seeds, first occurrence first ----------------------------------------
    unsigned int nb = fz_buckets((unsigned int)n_seeds);
    for (unsigned int i = cti; i < nb * 4u; i += CT)
      fz_store_slot(&tb[i], kFzEmpty, pol.keep);
    c.sync();
    FZ_T(1);
    // sanitised seeds into the label's edge scratch (free until hop 0 writes it): a seed outside [0, V) is flagged and replaced
    // by vertex 0 so that nothing is read out of bounds
    for (int sd = (int)cti; sd < n_seeds; sd += (int)CT) {
      long long v = a.seed_is64 ? static_cast<const long long*>(a.seeds)[lo_l + sd] : (long long)static_cast<const int*>(a.seeds)[lo_l + sd];
      if (v < 0 || (unsigned long long)v >= a.V) {
        *a.bad_seed = 1;
        v           = 0;
      }
      gidl[sd] = v;
    }
#if WGB_FZ_INSERT_STREAM
    fz_insert_stream<long long>(tb, nb, pol.keep, gidl, n_seeds, kFzPending, auxl, (int)cti, (int)CT);  // (each thread reads back its own writes)
#else
    FZ_HASH_LOOP(
      c, n_seeds, [&](int s) -> long long { return gidl[s]; }, [&](int, long long v) -> FzProbe { return fz_probe(tb, nb, (unsigned int)v, pol.keep); },
      [&](int s, const FzProbe& pr) { auxl[s] = fz_upsert(tb, nb, pr, kFzPending | (unsigned int)s, pol.keep); });
#endif
    c.sync();
    FZ_T(2);
    int known = (int)fz_ordered_scan(
      c, n_seeds, [&](long long s) -> unsigned int { return auxl[s]; },
      [&](long long, unsigned int slot) -> unsigned long long { return fz_load_slot(&tb[slot], pol.keep); },
      [&](long long s, unsigned long long w) -> unsigned int {
        auxl[s] = (unsigned int)w;
        return (unsigned int)w == (kFzPending | (unsigned int)s) ? 1u : 0u;
      },
      [&](long long s) -> unsigned int { return auxl[s] == (kFzPending | (unsigned int)s) ? 1u : 0u; },
      [&](long long s, unsigned int rank, unsigned int f) {
        if (f) {
          Fl[rank] = gidl[s];
          rankl[s] = rank;
        }
      });
    if (c.rank == 0 && tid < 32) {
      const unsigned long long rb = fz_lookback(a.pub, l, (unsigned long long)known);
      if (tid == 0) {
        sh.rowbase = (long long)rb;
        n_step[0]  = known;
      }
    }
    c.sync();  // ranks of the first occurrences are visible
    FZ_T(3);
    for (int s = (int)cti; s < n_seeds; s += (int)CT)
      a.seed_local[lo_l + s] = (int)rankl[auxl[s] & ~kFzPending];
    FZ_T(4);

    // ---- hops ----------------------------------------------------------------------------------------------------
    int nbase = 0, n_rows = known, ebase = 0;  // frontier of hop h = Fl[nbase, nbase + n_rows)
    for (int h = 0; h < L; h++) {
      const int M = a.fanout[h];
      const int T0 = 5 + 6 * h;
      // P1: min(deg, M) per frontier row, exclusive scan in row order -> the row's first edge
      struct Row2 {
        long long s0, s1;
      };
      const int e_h = (int)fz_ordered_scan(
        c, n_rows, [&](long long r) -> long long { return Fl[nbase + r]; },
        [&](long long, long long node) -> Row2 {
          Row2 x;
          x.s0 = load_i64<false>(a.row_ptr, a.row_ptr_off + (unsigned long long)node);
          x.s1 = load_i64<false>(a.row_ptr, a.row_ptr_off + (unsigned long long)node + 1);
          return x;
        },
        [&](long long r, const Row2& x) -> unsigned int {
          long long d   = x.s1 - x.s0;
          (a.Rstart + vbeg)[nbase + r] = x.s0;
          (a.Rdeg + vbeg)[nbase + r]   = (int)d;
          d             = d < 0 ? 0 : (d > M ? M : d);
          Ol[nbase + r] = (int)d;
          return (unsigned int)d;
        },
        [&](long long r) -> unsigned int { return (unsigned int)Ol[nbase + r]; },
        [&](long long r, unsigned int prefix, unsigned int) { Ol[nbase + r] = ebase + (int)prefix; });
      // (the scan's barrier also ordered CTA 0's rowbase before everybody's read below)
      const long long rowbase = c.size == 1 ? sh.rowbase : *cg::this_cluster().map_shared_rank(&sh.rowbase, 0);
      if (c.rank == 0 && tid == 0) e_hop[h] = e_h;
      known = nbase + n_rows;
      FZ_T(T0);
      c.sync();  // row offsets and extents visible
      // P2: the hop's rows are sampled.  The table is not involved: it is cleared only AFTER this phase, because sampling streams
      // ~1.6 MB of graph sectors per label through L2 (240 MB per 148-label call group) -- a table cleared before it was out in DRAM
      // again by the time the inserts came, and they ran at the DRAM rate of random read-modify-writes (27 G/s chip-wide, what
      // profiles/r2ay_atomic_probe.txt measures for a 283 MB table) instead of the L2 rate (40-45 G/s for tables that fit).
      if (e_h > 0) {
        FzSink<ColT> sink{destl + ebase, majl + ebase, gidl + ebase};
        const unsigned long long hop_seed = a.random_state + (unsigned long long)h * 0x9E3779B97F4A7C15ULL;
        if (M <= 8) fz_sample_rows<ColT, 8>(a, c, a.Rstart + vbeg, a.Rdeg + vbeg, Ol, nbase, n_rows, ebase, rowbase, M, hop_seed, sink, pol.stream);
        else if (M <= 16) fz_sample_rows<ColT, 16>(a, c, a.Rstart + vbeg, a.Rdeg + vbeg, Ol, nbase, n_rows, ebase, rowbase, M, hop_seed, sink, pol.stream);
        else fz_sample_rows<ColT, 32>(a, c, a.Rstart + vbeg, a.Rdeg + vbeg, Ol, nbase, n_rows, ebase, rowbase, M, hop_seed, sink, pol.stream);
      }
      FZ_T(T0 + 1);
      // P3: a fresh table for (everything numbered so far + this hop's edges), written by every thread as it leaves the sampling
      // loop (no barrier in between: nothing reads the table before the one below)
      nb = fz_buckets((unsigned int)known + (unsigned int)e_h);
      for (unsigned int i = cti; i < nb * 4u; i += CT)
        fz_store_slot(&tb[i], kFzEmpty, pol.keep);
      c.sync();  // table cleared, edges written
      FZ_T(T0 + 2);
      // P4: numbered vertices enter the table with their local id, endpoints with their edge index (the smallest wins a new
      // vertex); the two kinds of insert commute (the slot keeps the minimum), so they share a phase
#if WGB_FZ_INSERT_STREAM
      fz_insert_stream<long long>(tb, nb, pol.keep, Fl, known, 0u, nullptr, (int)cti, (int)CT);
      fz_insert_stream<ColT>(tb, nb, pol.keep, destl + ebase, e_h, kFzPending, auxl + ebase, (int)cti, (int)CT);
#else
      FZ_HASH_LOOP(
        c, known, [&](int j) -> unsigned int { return (unsigned int)Fl[j]; },
        [&](int, unsigned int v) -> FzProbe { return fz_probe(tb, nb, v, pol.keep); },
        [&](int j, const FzProbe& pr) { fz_upsert(tb, nb, pr, (unsigned int)j, pol.keep); });
      FZ_HASH_LOOP(
        c, e_h, [&](int i) -> unsigned int { return (unsigned int)destl[ebase + i]; },
        [&](int, unsigned int v) -> FzProbe { return fz_probe(tb, nb, v, pol.keep); },
        [&](int i, const FzProbe& pr) { auxl[ebase + i] = fz_upsert(tb, nb, pr, kFzPending | (unsigned int)i, pol.keep); });
#endif
      c.sync();
      FZ_T(T0 + 3);
      // P5: first occurrences in edge order -> the next frontier, appended to the label's vertex array
      const int n_new = (int)fz_ordered_scan(
        c, e_h, [&](long long i) -> unsigned int { return auxl[ebase + i]; },
        [&](long long, unsigned int slot) -> unsigned long long { return fz_load_slot(&tb[slot], pol.keep); },
        [&](long long i, unsigned long long w) -> unsigned int {
          auxl[ebase + i] = (unsigned int)w;
          return (unsigned int)w == (kFzPending | (unsigned int)i) ? 1u : 0u;
        },
        [&](long long i) -> unsigned int { return auxl[ebase + i] == (kFzPending | (unsigned int)i) ? 1u : 0u; },
        [&](long long i, unsigned int rank, unsigned int f) {
          if (f) {
            Fl[known + rank] = (long long)destl[ebase + i];
            rankl[ebase + i] = rank;
          }
        });
      if (c.rank == 0 && tid < 32) {
        unsigned long long rb = 0;
        if (h + 1 < L) rb = fz_lookback(a.pub + (long long)(h + 1) * B, l, (unsigned long long)n_new);
        if (tid == 0) {
          sh.rowbase    = (long long)rb;
          n_step[h + 1] = n_new;
        }
      }
      c.sync();  // ranks visible
      FZ_T(T0 + 4);
      // P6: endpoints -> local ids
      FZ_HASH_LOOP(
        c, e_h, [&](int i) -> unsigned int { return auxl[ebase + i]; },
        [&](int, unsigned int ax) -> unsigned int { return (ax & kFzPending) ? rankl[ebase + (ax & ~kFzPending)] : 0u; },
        [&](int i, unsigned int rk) {
          const unsigned int ax = auxl[ebase + i];
          mnrl[ebase + i]       = (ax & kFzPending) ? known + (int)rk : (int)ax;
        });
      FZ_T(T0 + 5);
      nbase  = known;
      n_rows = n_new;
      ebase += e_h;
    }
#undef FZ_T
#undef Fl
#undef Ol
#undef majl
#undef mnrl
#undef gidl
#undef destl
#undef auxl
#undef rankl
#undef tb
#undef n_step
#undef e_hop
    // the next label's first barrier orders this label's last reads before its scratch is touched again (it is not:
    // scratch is per label), and sh.* before it is rewritten
  }
}

// ---- after the labels: offsets in the layout of multihop_finish -----------------------------------------------------
// counts[0 .. B*L) edges of (label, hop); counts[B*L .. +B) vertices of label; counts[B*L+B .. +B) source rows of label;
// base[t*B + l] local id of the first vertex label l discovered at step t
__global__ void __launch_bounds__(256) fz_meta_kernel(int B, int L, const int* __restrict__ n_step, const int* __restrict__ e_hop,
                                                      long long* __restrict__ counts, int* __restrict__ base)
{
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < B; l += gridDim.x * blockDim.x) {
    int acc = 0, rows = 0;
    for (int t = 0; t <= L; t++) {
      base[t * B + l] = acc;
      acc += n_step[(long long)l * (L + 1) + t];
      if (t == L - 1) rows = acc;
    }
    for (int h = 0; h < L; h++)
      counts[(long long)l * L + h] = e_hop[(long long)l * L + h];
    counts[(long long)B * L + l]     = acc;
    counts[(long long)B * L + B + l] = rows;
  }
}

// label-major scratch -> outputs: blockIdx.y = label
template <typename OutT, bool CHUNKED>
__global__ void __launch_bounds__(256) fz_emit_edges_kernel(int L, const long long* __restrict__ label_offsets, long long estride,
                                                            const long long* __restrict__ lho, const int* __restrict__ maj,
                                                            const int* __restrict__ mnr, const long long* __restrict__ gid,
                                                            ChunkRef edge_id_ref, unsigned long long edge_id_off, bool has_edge_id,
                                                            OutT* __restrict__ majors, OutT* __restrict__ minors,
                                                            long long* __restrict__ edge_id_out)
{
  const int l        = blockIdx.y;
  const long long p0 = lho[(long long)l * L], n = lho[(long long)(l + 1) * L] - p0;
  const long long s0 = label_offsets[l] * estride;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (majors) majors[p0 + i] = (OutT)maj[s0 + i];
    minors[p0 + i] = (OutT)mnr[s0 + i];
    long long g    = gid[s0 + i];
    if (has_edge_id) g = load_i64<CHUNKED>(edge_id_ref, edge_id_off + (unsigned long long)g);
    edge_id_out[p0 + i] = g;
  }
}

__global__ void __launch_bounds__(256) fz_emit_rows_kernel(int L, const long long* __restrict__ label_offsets, long long fstride,
                                                           const long long* __restrict__ rmo, const long long* __restrict__ rbase,
                                                           const long long* __restrict__ lho, const long long* __restrict__ F,
                                                           const int* __restrict__ Orow, long long* __restrict__ map_out,
                                                           long long* __restrict__ major_offsets)
{
  const int l        = blockIdx.y;
  const long long m0 = rmo[l], n = rmo[l + 1] - m0;
  const long long s0 = label_offsets[l] * fstride;
  const long long r0 = major_offsets ? rbase[l] : 0, nr = major_offsets ? rbase[l + 1] - r0 : 0;
  const long long e0 = lho[(long long)l * L];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    map_out[m0 + i] = F[s0 + i];
    if (i < nr) major_offsets[r0 + i] = e0 + (long long)Orow[s0 + i];
  }
}

}  // namespace wgb

#else  // WGB_FZ_HOST_HALF
// ---- host side ---------------------------------------------------------------------------------------------------
namespace wgb {

// WGB_MH_FUSED=0 sends every call down the kernel chain (A/B measurements, tests of both paths); read per call
static bool fz_enabled()
{
  const char* e = getenv("WGB_MH_FUSED");
  return !(e && atoi(e) == 0);
}

// first half of a call on the fused path; false: not eligible, the caller runs the kernel chain
template <typename ColT>
static bool multihop_begin_fused(MhCall& c)
{
  if (!fz_enabled() || c.hetero || c.temporal || c.weighted || c.chunked || c.T != 1) return false;
  if (c.S <= 0 || c.B <= 0 || c.V >= 0xFFFFFFFFULL) return false;
  long long estride = 0, prod = 1;
  for (int h = 0; h < c.L; h++) {
    if (c.fanout[h] < 1 || c.fanout[h] > 32) return false;
    prod *= c.fanout[h];
    if ((long long)c.S * prod >= (1LL << 28)) return false;  // the kernel chain reports the limit
    estride += prod;
  }
  if ((long long)c.S / c.B > 65536) return false;  // a label is one cluster's work: keep it mini-batch sized
  auto* sp        = c.sp;
  const int B = c.B, L = c.L, S = c.S;
  cudaStream_t st = c.stream;
  const long long fstride = estride + 1;
  const size_t nv = (size_t)S * (size_t)fstride, ne = (size_t)S * (size_t)estride;
  const size_t slots = (size_t)(4 * (((long long)S * fstride + 1) / 2 + 16LL * B)) + 64;

  FzArgs a;
  memset(&a, 0, sizeof(a));
  a.row_ptr = c.csr[0].row_ptr; a.row_ptr_off = c.csr[0].row_ptr_off;
  a.col = c.csr[0].col; a.col_off = c.csr[0].col_off;
  a.seeds = c.seeds; a.seed_is64 = c.seed_dtype == WHOLEMEMORY_DT_INT64 ? 1 : 0;
  a.label_offsets = c.label_offsets;
  a.B = B; a.L = L;
  for (int h = 0; h < L; h++) a.fanout[h] = c.fanout[h];
  a.fstride = fstride; a.estride = estride;
  a.random_state = c.random_state;
  a.F       = static_cast<long long*>(ensure(sp->fz[0], nv * sizeof(long long)));
  a.Orow    = static_cast<int*>(ensure(sp->fz[1], nv * sizeof(int)));
  a.Rstart  = static_cast<long long*>(ensure(sp->fz[13], nv * sizeof(long long)));
  a.Rdeg    = static_cast<int*>(ensure(sp->fz[14], nv * sizeof(int)));
  a.maj     = static_cast<int*>(ensure(sp->fz[2], ne * sizeof(int)));
  a.mnr     = static_cast<int*>(ensure(sp->fz[3], ne * sizeof(int)));
  a.gid     = static_cast<long long*>(ensure(sp->fz[4], ne * sizeof(long long)));
  // a 4-byte endpoint array shares the storage of the local-id array it is turned into (phase 6 writes mnr[i] after every read
  // of dest has happened, two barriers earlier): 4 bytes less scratch per edge
  a.dest    = sizeof(ColT) == sizeof(int) ? static_cast<void*>(a.mnr) : ensure(sp->fz[5], ne * sizeof(ColT));
  a.aux     = static_cast<unsigned int*>(ensure(sp->fz[6], ne * sizeof(unsigned int)));
  a.rank_of = static_cast<unsigned int*>(ensure(sp->fz[7], ne * sizeof(unsigned int)));
  a.table   = static_cast<unsigned long long*>(ensure(sp->fz[8], slots * sizeof(unsigned long long)));
  a.seed_local = static_cast<int*>(ensure(sp->fz[9], (size_t)S * sizeof(int)));
  int* counters = static_cast<int*>(ensure(sp->fz[10], sizeof(int) * (size_t)B * (size_t)(2 * L + 1)));
  a.n_step = counters;
  a.e_hop  = counters + (size_t)B * (size_t)(L + 1);
  const size_t pub_bytes = sizeof(unsigned long long) * ((size_t)L * (size_t)B + 2);
  unsigned long long* pub = static_cast<unsigned long long*>(ensure(sp->fz[11], pub_bytes));
  WGB_CUDA_TRY(cudaMemsetAsync(pub, 0, pub_bytes, st));
  a.pub      = pub;
  a.ticket   = reinterpret_cast<unsigned int*>(pub + (size_t)L * (size_t)B);
  a.bad_seed = reinterpret_cast<long long*>(pub + (size_t)L * (size_t)B + 1);
  a.V        = c.V;
  a.dbg      = nullptr;
  if (sp->timing) {
    a.dbg = static_cast<long long*>(ensure(sp->fz[12], sizeof(long long) * (size_t)B * kFzDbgSlots));
    WGB_CUDA_TRY(cudaMemsetAsync(a.dbg, 0, sizeof(long long) * (size_t)B * kFzDbgSlots, st));
  }
  a.tab    = skip_table_device();

  sp->pending.active   = false;
  sp->pending.finished = false;
  sp->marks_used       = 0;
  mh_mark(sp, "start", st);

  // one cluster per label in flight; CL CTAs per cluster so that the labels in flight cover the SMs
  const int sms = num_sms();
  int CL = 1;
  while (CL < 8 && (long long)B * (CL * 2) <= sms) CL *= 2;
  static int max_clusters[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};  // [ColT is 64-bit][log2 CL]
  int& mc = max_clusters[sizeof(ColT) == 8 ? 1 : 0][CL == 1 ? 0 : CL == 2 ? 1 : CL == 4 ? 2 : 3];
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id               = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned int)CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim         = dim3(kFzThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream           = st;
  cfg.attrs            = attr;
  cfg.numAttrs         = 1;
  if (mc == 0) {
    // WGB_MH_CARVEOUT=<percent of 228 KB>: ask for a larger shared-memory carve-out than the kernel needs (< 4 KB), so that a
    // bulk-copy gather CTA (gather_scatter.cu, 96 KB of tile rings) can join the SM -- the L1 / shared split is fixed while
    // a CTA is resident.  Off by default: measured on C4 (profiles/r2k_*.json), a gather that co-resides takes HBM
    // bandwidth from this latency-bound kernel and the pipelined step gets no shorter than running the two back to back
    // (0.77 ms), while a gather that only fills the gaps this kernel leaves gives 0.64 ms.
    if (const char* e = getenv("WGB_MH_CARVEOUT"))
      if (atoi(e) > 0) WGB_CUDA_TRY(cudaFuncSetAttribute(fz_label_kernel<ColT>, cudaFuncAttributePreferredSharedMemoryCarveout, std::min(100, atoi(e))));
    cfg.gridDim = dim3((unsigned int)(CL * std::max(1, sms / CL)), 1, 1);
    int n       = 0;
    WGB_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, fz_label_kernel<ColT>, &cfg));
    WGB_EXPECTS(n > 0, "the fused sampler kernel does not fit on this device");
    mc = n;
  }
  int clusters = std::min(B, mc);
  // WGB_MH_SMS=<n>: at most n SMs for this kernel (labels are taken by ticket, so fewer CTAs simply take more labels each).  A
  // sampler CTA fills an SM's register file, so the SMs it leaves alone are where the feature gather of the previous call group
  // runs at the same time (spatial split of the GPU between the two stages of a loader's pipeline).
  if (const char* e = getenv("WGB_MH_SMS"))
    if (atoi(e) > 0) clusters = std::max(1, std::min(clusters, atoi(e) / CL));
  cfg.gridDim        = dim3((unsigned int)(clusters * CL), 1, 1);
  WGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, fz_label_kernel<ColT>, a));
  WGB_CHECK_LAUNCH();
  mh_mark(sp, "fused labels", st);

  // ---- offsets, in the arrays multihop_finish reads ----------------------------------------------------------------
  const long long n_groups = (long long)B * L, n_counts = n_groups + 2LL * B;
  long long* counts = static_cast<long long*>(ensure(sp->counts, sizeof(long long) * (size_t)(n_counts + 1)));
  long long* scans  = static_cast<long long*>(ensure(sp->small_i64, sizeof(long long) * (size_t)(n_counts + 8)));
  long long* lho    = scans;
  long long* rmo    = scans + n_groups + 1;
  long long* rbase  = rmo + B + 1;
  long long* totals = rbase + B + 1;
  int* base = static_cast<int*>(ensure(sp->base, sizeof(int) * (size_t)(L + 1) * (size_t)B));
  fz_meta_kernel<<<grid_over(B, sms), 256, 0, st>>>(B, L, a.n_step, a.e_hop, counts, base);
  WGB_CHECK_LAUNCH();
  MhScan3 sc;
  sc.in[0] = counts;                sc.out[0] = lho;   sc.n[0] = n_groups;
  sc.in[1] = counts + n_groups;     sc.out[1] = rmo;   sc.n[1] = B;
  sc.in[2] = counts + n_groups + B; sc.out[2] = rbase; sc.n[2] = B;
  sc.totals = totals;
  mh_scan3_kernel<<<3, 1024, 0, st>>>(sc);
  WGB_CHECK_LAUNCH();
  WGB_CUDA_TRY(cudaMemcpyAsync(sp->h_totals, totals, 3 * sizeof(long long), cudaMemcpyDeviceToHost, st));
  WGB_CUDA_TRY(cudaMemcpyAsync(sp->h_totals + 3, a.bad_seed, sizeof(long long), cudaMemcpyDeviceToHost, st));
  WGB_CUDA_TRY(cudaEventRecord(sp->ready, st));
  mh_mark(sp, "meta+scan3", st);

  auto& pd = sp->pending;
  pd.fused = true;
  pd.fz    = a;
  pd.B = B; pd.L = L; pd.flags = c.flags; pd.has_eid = c.csr[0].has_eid; pd.chunked = false;
  pd.eid = c.csr[0].eid; pd.eid_off = c.csr[0].eid_off;
  pd.hetero = false; pd.T = 1; pd.Vt = 1; pd.tbase = nullptr;
  pd.S = S;
  pd.lho = lho; pd.rmo = rmo; pd.rbase = rbase; pd.base = base;
  pd.active = true;
  return true;
}

// second half: wait for the sizes, allocate the outputs, copy the label segments
static void multihop_finish_fused(wholegraph_multihop_sampler_* sp, const MhOutCtx& c)
{
  auto& pd        = sp->pending;
  pd.active       = false;
  pd.finished     = true;
  const int sms   = num_sms();
  const int B = pd.B, L = pd.L;
  cudaStream_t st = c.stream;
  const FzArgs& a = pd.fz;
  WGB_CUDA_TRY(cudaEventSynchronize(sp->ready));
  WGB_CUDA_TRY(cudaStreamWaitEvent(st, sp->ready, 0));
  const long long n_edges = sp->h_totals[0], n_nodes = sp->h_totals[1], n_srcrows = sp->h_totals[2];
  if (sp->h_totals[3] != 0) throw invalid_input("a seed vertex id is outside [0, number of vertices)");
  const bool csr    = (pd.flags & WHOLEGRAPH_MULTIHOP_CSR) != 0;
  const bool idx64  = (pd.flags & WHOLEGRAPH_MULTIHOP_INT64_IDS) != 0;
  const wholememory_dtype_t idx_dt = idx64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT;
  void* out_minors   = output_alloc(c.env, c.minors, n_edges, idx_dt);
  void* out_majors   = (!csr && c.majors) ? output_alloc(c.env, c.majors, n_edges, idx_dt) : nullptr;
  long long* out_eid = static_cast<long long*>(output_alloc(c.env, c.edge_id, n_edges, WHOLEMEMORY_DT_INT64));
  long long* out_lho = static_cast<long long*>(output_alloc(c.env, c.lho, (long long)B * L + 1, WHOLEMEMORY_DT_INT64));
  long long* out_map = static_cast<long long*>(output_alloc(c.env, c.map, n_nodes, WHOLEMEMORY_DT_INT64));
  long long* out_rmo = static_cast<long long*>(output_alloc(c.env, c.rmo, B + 1, WHOLEMEMORY_DT_INT64));
  long long* out_moff = nullptr;
  if (csr) {
    WGB_EXPECTS(c.major_offsets != nullptr, "CSR compression needs a major_offsets output context");
    out_moff = static_cast<long long*>(output_alloc(c.env, c.major_offsets, n_srcrows + 1, WHOLEMEMORY_DT_INT64));
  }
  if (c.step_counts) {
    int* out_sc = static_cast<int*>(output_alloc(c.env, c.step_counts, (long long)(L + 1) * B, WHOLEMEMORY_DT_INT));
    WGB_CUDA_TRY(cudaMemcpyAsync(out_sc, pd.base, sizeof(int) * (size_t)(L + 1) * (size_t)B, cudaMemcpyDeviceToDevice, st));
  }
  WGB_CUDA_TRY(cudaMemcpyAsync(out_rmo, pd.rmo, sizeof(long long) * (size_t)(B + 1), cudaMemcpyDeviceToDevice, st));
  if (!csr) {
    WGB_CUDA_TRY(cudaMemcpyAsync(out_lho, pd.lho, sizeof(long long) * (size_t)((long long)B * L + 1), cudaMemcpyDeviceToDevice, st));
  } else {
    mh_csr_label_hop_kernel<<<grid_over((long long)B * L + 1, sms), 256, 0, st>>>(L, B, pd.base, pd.rbase, pd.lho, out_lho, out_moff);
    WGB_CHECK_LAUNCH();
  }
  // grid: blockIdx.y = label, enough CTAs per label to cover its segment with ~4 items per thread
  auto gx = [&](long long total) { return (unsigned int)std::max<long long>(1, std::min<long long>(64, (total / std::max(B, 1) + 1023) / 1024)); };
  if (n_edges > 0) {
    dim3 grid(gx(n_edges), (unsigned int)B);
    if (idx64)
      fz_emit_edges_kernel<long long, false><<<grid, 256, 0, st>>>(L, a.label_offsets, a.estride, pd.lho, a.maj, a.mnr, a.gid, pd.eid, pd.eid_off, pd.has_eid,
                                                                   static_cast<long long*>(out_majors), static_cast<long long*>(out_minors), out_eid);
    else
      fz_emit_edges_kernel<int, false><<<grid, 256, 0, st>>>(L, a.label_offsets, a.estride, pd.lho, a.maj, a.mnr, a.gid, pd.eid, pd.eid_off, pd.has_eid,
                                                             static_cast<int*>(out_majors), static_cast<int*>(out_minors), out_eid);
    WGB_CHECK_LAUNCH();
  }
  if (n_nodes > 0) {
    fz_emit_rows_kernel<<<dim3(gx(n_nodes), (unsigned int)B), 256, 0, st>>>(L, a.label_offsets, a.fstride, pd.rmo, pd.rbase, pd.lho, a.F, a.Orow, out_map, out_moff);
    WGB_CHECK_LAUNCH();
  }
  mh_mark(sp, "emit", st);
  mh_collect_marks(sp);
  if (a.dbg) {  // WGB_MH_TIMING: per-phase time of a label, averaged over labels and calls
    std::vector<long long> t((size_t)B * kFzDbgSlots);
    WGB_CUDA_TRY(cudaMemcpy(t.data(), a.dbg, t.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    long long first = 0, last = 0;
    for (int l = 0; l < B; l++) {
      const long long* tl = t.data() + (size_t)l * kFzDbgSlots;
      long long prev = tl[0];
      if (l == 0 || tl[0] < first) first = tl[0];
      for (int p = 1; p < kFzDbgSlots; p++) {
        if (tl[p] == 0) continue;
        sp->fz_phase_ns[p] += (double)(tl[p] - prev);
        prev = tl[p];
        if (tl[p] > last) last = tl[p];
      }
    }
    sp->fz_phase_labels += B;
    sp->fz_span_ns += (double)(last - first);
    sp->fz_calls++;
  }
}

static void fz_print_phases(wholegraph_multihop_sampler_* sp)
{
  if (!sp->timing || sp->fz_phase_labels == 0) return;
  static const char* seed_names[5] = {"", "seeds: clear table", "seeds: insert", "seeds: first occurrences", "seeds: local ids"};
  static const char* hop_names[6]  = {"count + scan", "sample", "clear table", "insert known + endpoints", "first occurrences", "local ids"};
  fprintf(stderr, "[wgb multihop fused] mean time of a label per phase (%lld labels), kernel span %.1f us per call\n", sp->fz_phase_labels,
          1e-3 * sp->fz_span_ns / (double)std::max<long long>(1, sp->fz_calls));
  double tot = 0;
  for (int p = 1; p < kFzDbgSlots; p++) {
    if (sp->fz_phase_ns[p] == 0) continue;
    char name[64];
    if (p < 5) snprintf(name, sizeof(name), "%s", seed_names[p]);
    else snprintf(name, sizeof(name), "hop%d: %s", (p - 5) / 6, hop_names[(p - 5) % 6]);
    fprintf(stderr, "  %-36s %9.2f us\n", name, 1e-3 * sp->fz_phase_ns[p] / (double)sp->fz_phase_labels);
    tot += sp->fz_phase_ns[p];
  }
  fprintf(stderr, "  %-36s %9.2f us\n", "label total", 1e-3 * tot / (double)sp->fz_phase_labels);
}

}  // namespace wgb

#endif  // WGB_FZ_HOST_HALF
#endif  // WGB_HOST_EMULATION
