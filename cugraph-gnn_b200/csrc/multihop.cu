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
//  * a call is two halves, _begin (enqueue every hop, no host wait) and _finish (wait for ONE event that marks the
//    output sizes, allocate through the callbacks, scatter scratch -> outputs), so loaders keep call group k+1 running
//    while they finish call group k.  Frontier sizes and edge counts stay on the device; kernels are launched over
//    host-known upper bounds (|frontier| * fanout) and read the true sizes from device memory.
//  * per hop: count+scan (single pass, persistent ticketed tiles) -> sample (sub-warp per row) -> hash insert that
//    records the first edge position of every (label, vertex) -> single-pass flag+scan+compact that emits the next
//    frontier in first-occurrence order.  The next frontier of a label is exactly the vertices new in this hop.
//  * one open-addressing table keyed (label, vertex), buckets of two 16-byte slots = one 32-byte sector, read with one
//    256-bit load and claimed with one 128-bit CAS (what bounds dedup is the rate of random requests, not bytes).
//    Its allocation covers the worst case, but every hop only uses the first `nslots` slots, chosen ON THE DEVICE from
//    the true counts (2 x (known vertices + this hop's edges)).  Each hop runs in a new 8-bit epoch (no memset): the
//    vertices numbered so far are re-inserted (cheap: earlier hops are an order of magnitude smaller).
//  * compact only READS the table: an edge's slot becomes a reference step(4) | index(28) that is either final (endpoint
//    numbered in an earlier step) or "pending, first seen at edge e'"; the final pass resolves pending references
//    through the dense rank_of array while it scatters (major, minor, edge id) from hop-major scratch to the
//    label-major / hop-minor layout the decoders expect (sampler/sampler.py:525-740).
//  * heterogeneous graphs (T edge types = T CSRs over one id space, Vt vertex types = id ranges): count+scan and sample
//    run once per edge type, a small kernel interleaves the per-type offsets so that the hop's edge list stays
//    label-major, insert / compact are unchanged; typed local ids and the [label][edge type][hop] grouping are
//    bookkeeping after the hops.
//
// Random numbers: hop h draws with seed hop_seed(random_state, h) = random_state + h * 0x9E3779B97F4A7C15
// and the S1/S2 stream geometry over the label-major concatenated frontier, which is what oracle/
// wg_oracle.cpp:wgo_multihop_sample restates on the CPU.

#include "wm_common.cuh"
#include "pcg.cuh"
#include "sample_device.cuh"
#include "temporal_device.cuh"

#include <wholememory/b200_ops.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

namespace wgb {

constexpr int kMaxHops = 16;

struct MhSlot {
  unsigned long long key;  // epoch(8) | label * V + vertex (56)
  unsigned long long aux;  // (255 - epoch)(8) | t(23) | index(32) | is_tag(1): see mh_tag()/mh_fin()
};

// "first position" race: tag(t, e) while hop t-1 is being inserted, replaced by fin(t, rank) once the
// first occurrence has been ranked.  fin(t, .) < tag(t', .) for t < t', so a vertex numbered in an earlier
// step always wins the atomicMin against edges of later hops.
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

// Open addressing over BUCKETS of two 16-byte slots = one 32-byte sector.  What bounds these kernels is the number
// of L2 REQUESTS per edge (random 8..32-byte accesses cost one request each, ~10^11 requests/s chip-wide), not
// bytes, so the table is built to need as few as possible:
//   * one 256-bit load (LDG.256) returns both keys AND both aux words of a bucket;
//   * a new key is installed together with its first position by ONE 128-bit CAS (ATOMG.CAS.128) on the slot;
//   * an edge that finds its vertex already has the current aux from the bucket load and only sends an atomicMin
//     if it can still lower it.
// => ~2 requests for a new vertex, ~1 for a repeated one (the previous 64-bit design: 4 key loads + CAS + aux
// load + atomicMin).  nslots is always a multiple of 4.
constexpr unsigned int kBucket = 2;

__device__ __forceinline__ unsigned int mh_home(unsigned long long item, unsigned int nslots)
{
  // fast range over the buckets (any bucket count), returns the first slot of the home bucket
  return (unsigned int)(((mh_mix(item) >> 32) * (unsigned long long)(nslots / kBucket)) >> 32) * kBucket;
}

struct MhBucket {
  unsigned long long k[kBucket], a[kBucket];
};
#ifdef WGB_HOST_EMULATION
// tests/emu compiles this file with g++ to check its LOGIC on the CPU (test infrastructure, never the product): the two
// PTX accesses become an unordered read of the four words and a 128-bit CAS serialised by a lock shared with nothing else
// (aux words are only ever lowered by atomicMin on slots whose key is already installed, so the lock is enough).
__device__ __forceinline__ MhBucket mh_load_bucket(const MhSlot* table, unsigned int g)
{
  MhBucket r;
  for (unsigned int j = 0; j < kBucket; j++) {
    r.k[j] = __atomic_load_n(&table[g + j].key, __ATOMIC_RELAXED);
    r.a[j] = __atomic_load_n(&table[g + j].aux, __ATOMIC_RELAXED);
  }
  return r;
}
__device__ __forceinline__ void mh_cas_slot(MhSlot* p, unsigned long long ek, unsigned long long ea, unsigned long long nk,
                                            unsigned long long na, unsigned long long& ok, unsigned long long& oa)
{
  static int lock = 0;
  while (__atomic_exchange_n(&lock, 1, __ATOMIC_ACQUIRE)) {}
  ok = __atomic_load_n(&p->key, __ATOMIC_RELAXED);
  oa = __atomic_load_n(&p->aux, __ATOMIC_RELAXED);
  if (ok == ek && oa == ea) {
    __atomic_store_n(&p->aux, na, __ATOMIC_RELAXED);
    __atomic_store_n(&p->key, nk, __ATOMIC_RELEASE);
  }
  __atomic_store_n(&lock, 0, __ATOMIC_RELEASE);
}
#else
__device__ __forceinline__ MhBucket mh_load_bucket(const MhSlot* table, unsigned int g)
{
  MhBucket r;
  asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(r.k[0]), "=l"(r.a[0]), "=l"(r.k[1]), "=l"(r.a[1])
               : "l"(table + g)
               : "memory");
  return r;
}

// 128-bit compare-and-swap of a whole slot; returns the previous contents
__device__ __forceinline__ void mh_cas_slot(MhSlot* p, unsigned long long ek, unsigned long long ea, unsigned long long nk,
                                            unsigned long long na, unsigned long long& ok, unsigned long long& oa)
{
  asm volatile(
    "{\n .reg .b128 c, s, d;\n mov.b128 c, {%2, %3};\n mov.b128 s, {%4, %5};\n"
    " atom.relaxed.gpu.global.cas.b128 d, [%6], c, s;\n mov.b128 {%0, %1}, d;\n}"
    : "=l"(ok), "=l"(oa)
    : "l"(ek), "l"(ea), "l"(nk), "l"(na), "l"(p)
    : "memory");
}
#endif

// Claim-or-find the slot of `item` (= label * V + vertex) in the current epoch, starting at bucket `g` whose
// contents have already been loaded into `cur` (lets callers batch the first, random, access of several items).
// A new key is installed with aux = `mine`; *seen = the aux this thread knows the slot to hold afterwards
// (`mine` if it installed the key, else the value read) -- callers race with atomicMin only if *seen > mine.
__device__ __forceinline__ unsigned int mh_upsert_from(MhSlot* table, unsigned int nslots, unsigned long long item,
                                                       unsigned long long epoch, unsigned long long mine, unsigned int g,
                                                       MhBucket cur, unsigned long long* seen)
{
  const unsigned long long key = (epoch << 56) | item;
  while (true) {
#pragma unroll
    for (unsigned int j = 0; j < kBucket; j++) {
      unsigned long long ck = cur.k[j], ca = cur.a[j];
      if (ck != key && (ck >> 56) != epoch) {  // stale or never used: try to claim it, key and aux at once
        unsigned long long ok, oa;
        mh_cas_slot(&table[g + j], ck, ca, key, mine, ok, oa);
        if (ok == ck && oa == ca) {
          *seen = mine;
          return g + j;
        }
        ck = ok;  // somebody else got there first: look at what the slot holds now
        ca = oa;
      }
      if (ck == key) {
        *seen = ca;
        return g + j;
      }
      // another key of this epoch: move on
    }
    g   = g + kBucket >= nslots ? 0u : g + kBucket;
    cur = mh_load_bucket(table, g);
  }
}

__device__ __forceinline__ void mh_race_first(MhSlot* table, unsigned int slot, unsigned long long seen, unsigned long long mine)
{
  // hub vertices are reached by many edges of one label: only an edge that can still lower the first position
  // pays for the same-address atomic
  if (seen > mine) atomicMin(&table[slot].aux, mine);
}

// ---- step 0: seeds ------------------------------------------------------------------------------------
// label of seed s by binary search in label_offsets, insert (label, seed), race for the first position
template <typename SeedT>
__global__ void __launch_bounds__(256) mh_seed_insert_kernel(MhSlot* table, const unsigned int* __restrict__ nslots_dev,
                                                             unsigned long long epoch, unsigned long long V,
                                                             const SeedT* __restrict__ seeds, int S,
                                                             const long long* __restrict__ label_offsets, int B,
                                                             int* __restrict__ slabel, unsigned int* __restrict__ slot_of,
                                                             long long* __restrict__ clean, long long* __restrict__ bad_seed)
{
  const unsigned int nslots          = *nslots_dev;
  const unsigned long long inv_epoch = (255ULL - epoch) << 56;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    int lo = 0, hi = B;  // last l with label_offsets[l] <= s
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (label_offsets[mid] <= s) lo = mid;
      else hi = mid;
    }
    slabel[s]                     = lo;
    // a seed outside [0, V) would alias another label's key and index row_ptr out of bounds: flag it (the call fails
    // at _finish with WHOLEMEMORY_INVALID_INPUT) and carry on with vertex 0 so that nothing is read out of bounds
    long long sv = (long long)seeds[s];
    if (sv < 0 || (unsigned long long)sv >= V) {
      *bad_seed = 1;
      sv        = 0;
    }
    clean[s]                      = sv;  // what every later kernel of the call reads as "the seeds"
    const unsigned long long item = (unsigned long long)lo * V + (unsigned long long)sv;
    const unsigned long long mine = inv_epoch | mh_tag(0u, (unsigned int)s);
    const unsigned int home       = mh_home(item, nslots);
    unsigned long long seen;
    const unsigned int slot = mh_upsert_from(table, nslots, item, epoch, mine, home, mh_load_bucket(table, home), &seen);
    mh_race_first(table, slot, seen, mine);
    slot_of[s] = slot;
  }
}

// ---- K3: insert the neighbours sampled in this hop -------------------------------------------------------
// kInsertIlp edges per thread: the bucket loads of all of them are issued before the first dependent CAS, so
// the random round trips of the edges overlap.
constexpr int kInsertIlp = 1;  // measured (profiles/micro/insert_probe.cu): the CAS chains serialise per thread, more warps beat more ILP
template <typename VT>
__global__ void __launch_bounds__(256) mh_insert_kernel(MhSlot* table, const unsigned int* __restrict__ nslots_dev,
                                                        unsigned long long epoch, unsigned long long V, unsigned int t,
                                                        const VT* __restrict__ vertices, const int* __restrict__ n_dev,
                                                        const int* __restrict__ erow, const int* __restrict__ flabel,
                                                        unsigned int* __restrict__ slot_of)
{
  const int n                        = *n_dev;
  const unsigned int nslots          = *nslots_dev;
  const unsigned long long inv_epoch = (255ULL - epoch) << 56;
  const int stride                   = gridDim.x * blockDim.x;
  for (int base = blockIdx.x * blockDim.x + threadIdx.x; base < n; base += stride * kInsertIlp) {
    int row[kInsertIlp];
    unsigned long long item[kInsertIlp];
    unsigned int home[kInsertIlp];
    MhBucket cur[kInsertIlp];
#pragma unroll
    for (int k = 0; k < kInsertIlp; k++) {
      int e  = base + k * stride;
      row[k] = e < n ? erow[e] : 0;
    }
#pragma unroll
    for (int k = 0; k < kInsertIlp; k++) {
      int e   = base + k * stride;
      item[k] = e < n ? (unsigned long long)flabel[row[k]] * V + (unsigned long long)vertices[e] : 0ULL;
    }
#pragma unroll
    for (int k = 0; k < kInsertIlp; k++) {
      int e   = base + k * stride;
      home[k] = mh_home(item[k], nslots);
      if (e < n) cur[k] = mh_load_bucket(table, home[k]);
    }
    unsigned int slot[kInsertIlp];
    unsigned long long seen[kInsertIlp];
#pragma unroll
    for (int k = 0; k < kInsertIlp; k++) {
      int e = base + k * stride;
      if (e < n) slot[k] = mh_upsert_from(table, nslots, item[k], epoch, inv_epoch | mh_tag(t, (unsigned int)e), home[k], cur[k], &seen[k]);
    }
#pragma unroll
    for (int k = 0; k < kInsertIlp; k++) {
      int e = base + k * stride;
      if (e < n) {
        mh_race_first(table, slot[k], seen[k], inv_epoch | mh_tag(t, (unsigned int)e));
        slot_of[e] = slot[k];
      }
    }
  }
}

// ---- K4: flag first occurrences, scan, compact into the next frontier, assign ranks ------------------------
// The table is only READ here (one random 8-byte access per edge): every edge turns its slot index into a
// reference to its endpoint, packed step(4) | index(28):
//   * endpoint numbered in an earlier step t' < t:   t'  | rank          (final)
//   * endpoint new in this step, first seen at e':    15  | e'            (pending), and rank_of[e'] = its rank
// Pending references are resolved by the final pass through rank_of (a dense array a few MB large), so neither a
// random write-back of the ranks into the table nor a second random pass over it is needed; the table itself is
// re-hashed by the next hop anyway.
constexpr unsigned int kRefPending = 15u;

// Striped arrangement: item k of thread x is edge tile_base + k * 256 + x, so every global access of a warp is
// coalesced (a blocked arrangement makes lanes stride by 4 * items bytes).  Flags are single bits, so the
// tile-wide exclusive scan is ballots + popcounts per (stripe, warp) and ONE 64-entry scan in shared memory.
constexpr int kCompactItems = 8;
constexpr int kCompactTile  = kScanBlock * kCompactItems;

template <typename VT, bool SEEDS>
__global__ void __launch_bounds__(kScanBlock) mh_compact_kernel(const MhSlot* __restrict__ table, unsigned int t,
                                                                const VT* __restrict__ vertices,
                                                                const int* __restrict__ n_dev, int n_host,
                                                                const int* __restrict__ erow,
                                                                const int* __restrict__ flabel,
                                                                unsigned int* __restrict__ slot_to_ref,
                                                                unsigned int* __restrict__ rank_of,
                                                                long long* __restrict__ next_frontier,
                                                                int* __restrict__ next_flabel, int* __restrict__ next_n,
                                                                unsigned long long* state, unsigned int* ticket)
{
  constexpr int kWarps = kScanBlock / 32;
  __shared__ unsigned int s_cnt[kCompactItems * kWarps];  // [stripe][warp] counts -> exclusive prefixes
  __shared__ unsigned int s_total;
  const int n    = n_dev ? *n_dev : n_host;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // persistent CTAs take tiles by ticket (in order, so the look-back never waits for a tile that has not started);
  // the grid is sized from the resident CTA count, not from the host-side upper bound of n
  while (true) {
    const int tile = take_ticket(ticket);
    if ((long long)tile * kCompactTile > n) return;
    const long long base = (long long)tile * kCompactTile + threadIdx.x;
    unsigned int slot[kCompactItems];
#pragma unroll
    for (int k = 0; k < kCompactItems; k++) {
      long long e = base + k * kScanBlock;
      slot[k]     = e < n ? slot_to_ref[e] : 0u;
    }
    unsigned long long aux[kCompactItems];
#pragma unroll
    for (int k = 0; k < kCompactItems; k++) {
      long long e = base + k * kScanBlock;
      aux[k]      = e < n ? (ld_relaxed_u64(&table[slot[k]].aux) & kAuxMask) : 0ULL;
    }
    unsigned int flags = 0;
    unsigned int within[kCompactItems];  // flagged items before this one in its (stripe, warp)
#pragma unroll
    for (int k = 0; k < kCompactItems; k++) {
      long long e          = base + k * kScanBlock;
      const bool f         = e < n && aux[k] == mh_tag(t, (unsigned int)e);
      const unsigned int b = __ballot_sync(0xffffffffu, f);
      within[k]            = __popc(b & ((1u << lane) - 1u));
      flags |= (f ? 1u : 0u) << k;
      if (lane == 0) s_cnt[k * kWarps + wid] = __popc(b);
    }
    if (!SEEDS || rank_of) {
      // references do not depend on the scan: written while the tile prefix is being looked up
      // tag(t, e'): bit 0 set, e' in bits 1..32;  fin(t', rank): bit 0 clear
#pragma unroll
      for (int k = 0; k < kCompactItems; k++) {
        long long e            = base + k * kScanBlock;
        const unsigned int idx = (unsigned int)(aux[k] >> 1) & 0x0FFFFFFFu;
        const unsigned int st  = (aux[k] & 1ULL) ? kRefPending : (unsigned int)(aux[k] >> 33);
        if (e < n) slot_to_ref[e] = (st << 28) | idx;
      }
    }
    __syncthreads();
    if (wid == 0) {  // exclusive scan of the kCompactItems * kWarps (= 64) counts, two per lane
      static_assert(kCompactItems * kWarps == 64, "scan below assumes 64 counters");
      unsigned int a = s_cnt[2 * lane], b = s_cnt[2 * lane + 1];
      unsigned int inc = a + b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
      }
      s_cnt[2 * lane]     = inc - a - b;
      s_cnt[2 * lane + 1] = inc - b;
      if (lane == 31) s_total = inc;
    }
    __syncthreads();
    const unsigned int agg    = s_total;
    unsigned long long prefix = scan_tile_prefix(state, tile, agg);
#pragma unroll
    for (int k = 0; k < kCompactItems; k++) {
      long long e = base + k * kScanBlock;
      if ((flags >> k) & 1u) {
        unsigned int rank   = (unsigned int)prefix + s_cnt[k * kWarps + wid] + within[k];
        next_frontier[rank] = (long long)vertices[e];
        next_flabel[rank]   = SEEDS ? flabel[e] : flabel[erow[e]];
        if (!SEEDS || rank_of) rank_of[e] = rank;
      }
    }
    if (n >= (long long)tile * kCompactTile && n < (long long)(tile + 1) * kCompactTile && threadIdx.x == 0)
      *next_n = (int)(prefix + agg);  // the tile that holds index n is the last one: every flagged item precedes n
    __syncthreads();  // s_cnt is rewritten by the next tile
  }
}

// fr_off[t][l] = first frontier row of label l at step t (flabel is non-decreasing), fr_off[t][B] = n_t.
// One thread per frontier row; a row writes the offsets of every label that starts at it.  Nothing on the hop
// chain reads these, so one launch (blockIdx.y = step) after the last hop does all steps.
struct MhLabelBounds {
  const int* flabel[kMaxHops + 1];
  int* fr_off[kMaxHops + 1];
};
__global__ void __launch_bounds__(256) mh_label_bounds_kernel(MhLabelBounds a, const int* __restrict__ n_rows, int B)
{
  const int t                    = blockIdx.y;
  const int n                    = n_rows[t];
  const int* __restrict__ flabel = a.flabel[t];
  int* __restrict__ fr_off       = a.fr_off[t];
  if (n == 0) {
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l <= B; l += gridDim.x * blockDim.x)
      fr_off[l] = 0;
    return;
  }
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const int l    = flabel[r];
    const int prev = r > 0 ? flabel[r - 1] : -1;
    for (int x = prev + 1; x <= l; x++)
      fr_off[x] = r;
    if (r == n - 1)
      for (int x = l + 1; x <= B; x++)
        fr_off[x] = n;
  }
}

// slots used by the hop that is about to be inserted: 2 x (vertices numbered so far + its edges)
__global__ void mh_plan_hop_kernel(const int* __restrict__ n_rows, int h, const int* __restrict__ n_edges_h,
                                   unsigned int capacity, unsigned int* __restrict__ nslots_out)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long items = (unsigned long long)*n_edges_h;
    for (int t = 0; t <= h; t++)
      items += (unsigned long long)n_rows[t];
    unsigned long long want = 2 * items + 1024;  // load factor <= 0.5
    if (want > capacity) want = capacity;
    *nslots_out = (unsigned int)(want & ~3ULL);
  }
}

// vertices numbered in steps 0..h enter the new epoch with their final (step, rank)
struct MhFrontiers {
  const long long* frontier[kMaxHops + 1];
  const int* flabel[kMaxHops + 1];
};
__global__ void __launch_bounds__(256) mh_reinsert_kernel(MhSlot* table, const unsigned int* __restrict__ nslots_dev,
                                                          unsigned long long epoch, unsigned long long V,
                                                          MhFrontiers fr, const int* __restrict__ n_rows)
{
  const unsigned int t               = blockIdx.y;
  const int n                        = n_rows[t];
  const unsigned int nslots          = *nslots_dev;
  const unsigned long long inv_epoch = (255ULL - epoch) << 56;
  const long long* __restrict__ frontier = fr.frontier[t];
  const int* __restrict__ flabel         = fr.flabel[t];
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
    unsigned long long item = (unsigned long long)flabel[f] * V + (unsigned long long)frontier[f];
    unsigned int home       = mh_home(item, nslots);
    unsigned long long seen;
    // (label, vertex) pairs of the frontiers are distinct: always a fresh claim, installed with its final aux
    mh_upsert_from(table, nslots, item, epoch, inv_epoch | mh_fin(t, (unsigned int)f), home, mh_load_bucket(table, home), &seen);
  }
}

// ---- per-label bookkeeping after the last hop ------------------------------------------------------------
struct MhMeta {
  const int* fr_off[kMaxHops + 1];  // [t][B+1]
  const int* off[kMaxHops];         // [h] sample offsets over the frontier rows (null: hop skipped)
  int L;
  int B;
};

// counts[0 .. B*L)           edges of (label, hop)
// counts[B*L .. B*L+B)       vertices of label
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
    for (int h = 0; h < m.L; h++) {
      long long e = 0;
      if (m.off[h]) e = (long long)m.off[h][m.fr_off[h][l + 1]] - (long long)m.off[h][m.fr_off[h][l]];
      counts[(long long)l * m.L + h] = e;
    }
    counts[(long long)m.B * m.L + l]       = acc;
    counts[(long long)m.B * m.L + m.B + l] = rows;
  }
}

// three exclusive scans (blockIdx.x selects the array), each by one block; writes n+1 outputs + the total
struct MhScan3 {
  const long long* in[3];
  long long* out[3];
  long long n[3];
  long long* totals;
};
__global__ void __launch_bounds__(1024) mh_scan3_kernel(MhScan3 a)
{
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  const long long* __restrict__ in = a.in[blockIdx.x];
  long long* __restrict__ out      = a.out[blockIdx.x];
  const long long n                = a.n[blockIdx.x];
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (long long base = 0; base < n; base += blockDim.x) {
    long long i   = base + threadIdx.x;
    long long x   = i < n ? in[i] : 0;
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
    out[n]               = s_carry;
    a.totals[blockIdx.x] = s_carry;
  }
}

// local id of every INPUT seed (duplicates included): the link-prediction loaders index the label's nodes with it
// (edge_label_index; the reference sorts + unique_consecutive's every batch in python for this,
// sampler/distributed_sampler.py:487-533).  typed_local != null: heterogeneous ids (per label and vertex type).
__global__ void __launch_bounds__(256) mh_seed_local_kernel(const unsigned int* __restrict__ seed_ref,
                                                            const unsigned int* __restrict__ seed_rank,
                                                            const int* __restrict__ slabel, const int* __restrict__ fr_off0,
                                                            const int* __restrict__ typed_local0, int S, int* __restrict__ out)
{
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    const unsigned int first = seed_ref[s] & 0x0FFFFFFFu;  // position of the first occurrence of this seed's vertex
    const unsigned int rank  = seed_rank[first];           // its row in frontier 0
    out[s] = typed_local0 ? typed_local0[rank] : (int)rank - fr_off0[slabel[s]];
  }
}

// ---- temporal calls: the time every frontier row carries (temporal_device.cuh) ------------------------------------
// step 0: the time of the FIRST occurrence of a (label, seed) pair is the one its frontier row gets
__global__ void __launch_bounds__(256) mh_seed_time_kernel(const unsigned int* __restrict__ seed_ref,
                                                           const unsigned int* __restrict__ seed_rank,
                                                           const long long* __restrict__ seed_times, int S,
                                                           long long* __restrict__ ftime0)
{
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x)
    if ((seed_ref[s] & 0x0FFFFFFFu) == (unsigned int)s) ftime0[seed_rank[s]] = seed_times[s];
}

// after the compaction of a hop: a vertex that is new in this hop takes the time of the edge that reached it first
// (the edge whose reference is "pending, first seen at myself").  Edge times per edge type live in device memory.
struct MhTemporalDesc {
  int T;
  ChunkRef etime[16];
  unsigned long long etime_off[16];
};
struct MhTypePos {
  const int* pos[16];  // [t][f]: first edge of (row f, type t) in the hop's edge list; only read when T > 1
};
template <bool CHUNKED>
__global__ void __launch_bounds__(256) mh_next_time_kernel(const MhTemporalDesc* __restrict__ d, MhTypePos tp,
                                                           const int* __restrict__ n_edges,
                                                           const unsigned int* __restrict__ slot_to_ref,
                                                           const unsigned int* __restrict__ rank_of,
                                                           const int* __restrict__ erow, const long long* __restrict__ gid,
                                                           long long* __restrict__ next_ftime)
{
  const int n = *n_edges;
  const int T = d->T;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const unsigned int ref = slot_to_ref[e];
    if ((ref >> 28) != kRefPending || (ref & 0x0FFFFFFFu) != (unsigned int)e) continue;
    int t = 0;
    if (T > 1) {
      const int f = erow[e];
      for (int i = 1; i < T; i++)
        t += tp.pos[i][f] <= e ? 1 : 0;  // pos is non-decreasing in the type for a fixed row
    }
    next_ftime[rank_of[e]] = load_i64<CHUNKED>(d->etime[t], d->etime_off[t] + (unsigned long long)gid[e]);
  }
}

// ---- final pass -------------------------------------------------------------------------------------------
struct MhHopBufs {
  const int* off[kMaxHops];
  const int* erow[kMaxHops];
  const unsigned int* slot[kMaxHops];     // per edge: reference to its endpoint, see mh_compact_kernel
  const unsigned int* rank_of[kMaxHops];  // per edge that is a first occurrence: rank of its endpoint in the next frontier
  const long long* gid[kMaxHops];
};

// edges of hop blockIdx.y: scratch (hop-major) -> outputs (label-major, hop-minor), (step, rank) -> local ids
template <typename OutT, bool CHUNKED>
__global__ void __launch_bounds__(256) mh_emit_edges_kernel(int L, int B,
                                                            const int* __restrict__ n_edges, MhHopBufs hb, MhFrontiers fr,
                                                            MhMeta m, const int* __restrict__ base,
                                                            const long long* __restrict__ lho, ChunkRef edge_id_ref,
                                                            unsigned long long edge_id_off, bool has_edge_id,
                                                            OutT* __restrict__ majors, OutT* __restrict__ minors,
                                                            long long* __restrict__ edge_id_out)
{
  const int h = blockIdx.y;
  const int n = n_edges[h];
  const int* __restrict__ off_h            = hb.off[h];
  const int* __restrict__ erow             = hb.erow[h];
  const unsigned int* __restrict__ slot_of = hb.slot[h];
  const long long* __restrict__ gid        = hb.gid[h];
  const int* __restrict__ flabel_h         = fr.flabel[h];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    unsigned int ref             = slot_of[e];  // step(4) | index(28), see mh_compact_kernel
    if ((ref >> 28) == kRefPending) ref = ((unsigned int)(h + 1) << 28) | hb.rank_of[h][ref & 0x0FFFFFFFu];
    const int f                  = erow[e];
    long long g                  = gid[e];
    const int l                  = flabel_h[f];
    const int f0                 = m.fr_off[h][l];
    const long long p            = lho[(long long)l * L + h] + (long long)(e - off_h[f0]);
    const unsigned int t         = ref >> 28;
    const unsigned int rk        = ref & 0x0FFFFFFFu;
    if (majors) majors[p] = (OutT)(base[h * B + l] + (f - f0));
    minors[p] = (OutT)(base[t * B + l] + (int)(rk - (unsigned int)m.fr_off[t][l]));
    if (has_edge_id) g = load_i64<CHUNKED>(edge_id_ref, edge_id_off + (unsigned long long)g);
    edge_id_out[p] = g;
  }
}

// frontier rows of step blockIdx.y: renumber map (+ CSR major offsets for steps < L)
__global__ void __launch_bounds__(256) mh_emit_rows_kernel(int L, int B, const int* __restrict__ n_rows, MhFrontiers fr,
                                                           MhMeta m, const int* __restrict__ base,
                                                           const long long* __restrict__ rmo, const long long* __restrict__ lho,
                                                           const long long* __restrict__ rbase,
                                                           long long* __restrict__ map_out, long long* __restrict__ major_offsets)
{
  const int t = blockIdx.y;
  const int n = n_rows[t];
  const long long* __restrict__ frontier = fr.frontier[t];
  const int* __restrict__ flabel         = fr.flabel[t];
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


// ---- heterogeneous graphs -------------------------------------------------------------------------------------
// T edge types, each with its own CSR over the GLOBAL vertex id space; Vt vertex types own contiguous id ranges.
// A hop samples the whole frontier once per edge type (fan-out [hop * T + etype]); the hop's edge list is ordered
// (frontier row, edge type, slot), so it stays label-major and insert / compact above run unchanged over it.
// Everything type-specific is bookkeeping around them: where (row, type) starts in the hop's edge list, and local ids
// that restart per (label, vertex type).  The descriptor lives in device memory (too large for a parameter block).
constexpr int kMaxEdgeTypes   = 16;
constexpr int kMaxVertexTypes = 16;
static_assert(kMaxEdgeTypes == 16, "MhTemporalDesc / MhTypePos are sized for 16 edge types");

struct MhHeteroDesc {
  int T, Vt, L, B;
  long long vto[kMaxVertexTypes + 1];              // vertex type vt owns global ids [vto[vt], vto[vt+1])
  const int* off[kMaxHops][kMaxEdgeTypes];         // [h][t][f]: exclusive scan over frontier rows of the type's counts (null: skipped)
  const int* pos[kMaxHops][kMaxEdgeTypes];         // [h][t][f]: first edge of (row f, type t) in the hop's edge list
  const int* tcnt[kMaxHops + 1][kMaxVertexTypes];  // [s][vt][r]: rows r' < r of step s whose vertex has type vt
  int* typed_local[kMaxHops + 1];                  // [s][r]: local id of frontier row r inside (label, vertex type)
  ChunkRef eid[kMaxEdgeTypes];
  unsigned long long eid_off[kMaxEdgeTypes];
  int has_eid[kMaxEdgeTypes];
};

__device__ __forceinline__ int mh_vtype_of(const MhHeteroDesc* __restrict__ d, long long v)
{
  int vt = 0;
  for (int i = 1; i < d->Vt; i++)
    vt += v >= d->vto[i] ? 1 : 0;
  return vt;
}

// pos[t][f] = sum_t' off[t'][f] + sum_{t' < t} (off[t'][f+1] - off[t'][f]);  *n_edges = sum_t off[t][n]
struct MhCombine {
  const int* off[kMaxEdgeTypes];
  int* pos[kMaxEdgeTypes];
  int T;
};
__global__ void __launch_bounds__(256) mh_combine_types_kernel(MhCombine a, const int* __restrict__ n_dev, int* __restrict__ n_edges)
{
  const int n = *n_dev;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f <= n; f += gridDim.x * blockDim.x) {
    int run = 0;
    for (int t = 0; t < a.T; t++)
      run += a.off[t] ? a.off[t][f] : 0;
    if (f == n) *n_edges = run;
    for (int t = 0; t < a.T; t++) {
      a.pos[t][f] = run;
      if (f < n && a.off[t]) run += a.off[t][f + 1] - a.off[t][f];
    }
  }
}

// tcnt[s][vt][r] for every (step, vertex type) slice = blockIdx.y; persistent ticketed scan per slice
struct MhVtypeScan {
  int* tcnt[(kMaxHops + 1)];  // [s] -> Vt arrays of (cap_s + 1) ints, back to back
  long long cap[(kMaxHops + 1)];
  unsigned long long* state;  // per slice: tiles_max words + 2 (ticket)
  long long state_stride;
};
__global__ void __launch_bounds__(kScanBlock) mh_vtype_scan_kernel(const MhHeteroDesc* __restrict__ d, MhFrontiers fr,
                                                                   const int* __restrict__ n_rows, MhVtypeScan a)
{
  const int s = blockIdx.y / d->Vt, vt = blockIdx.y % d->Vt;
  const int n = n_rows[s];
  const long long* __restrict__ frontier = fr.frontier[s];
  int* __restrict__ out                  = a.tcnt[s] + (long long)vt * (a.cap[s] + 1);
  unsigned long long* state              = a.state + (long long)blockIdx.y * a.state_stride;
  unsigned int* ticket                   = reinterpret_cast<unsigned int*>(state + a.state_stride - 2);
  const long long lo = d->vto[vt], hi = d->vto[vt + 1];
  while (true) {
    const int tile = take_ticket(ticket);
    if ((long long)tile * kScanTile > n) return;
    const long long base = (long long)tile * kScanTile + (long long)threadIdx.x * kScanItems;
    unsigned int v[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      long long r = base + k;
      long long x = r < n ? frontier[r] : -1;
      v[k]        = (x >= lo && x < hi) ? 1u : 0u;
    }
    unsigned int agg          = block_scan_items(v);
    unsigned long long prefix = scan_tile_prefix(state, tile, agg);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      long long r = base + k;
      if (r <= n) out[r] = (int)(prefix + v[k]);
    }
  }
}

// per label: first typed local id of every (step, vertex type), node counts per (label, vtype), edge counts per
// (label, edge type, hop)
__global__ void __launch_bounds__(256) mh_hetero_meta_kernel(const MhHeteroDesc* __restrict__ d, MhMeta m,
                                                             long long* __restrict__ edge_counts,
                                                             long long* __restrict__ node_counts, int* __restrict__ tbase)
{
  const int T = d->T, Vt = d->Vt, L = d->L, B = d->B;
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < B; l += gridDim.x * blockDim.x) {
    for (int vt = 0; vt < Vt; vt++) {
      int acc = 0;
      for (int s = 0; s <= L; s++) {
        tbase[((long long)s * Vt + vt) * B + l] = acc;
        const int* c = d->tcnt[s][vt];
        acc += c[m.fr_off[s][l + 1]] - c[m.fr_off[s][l]];
      }
      node_counts[(long long)l * Vt + vt] = acc;
    }
    for (int t = 0; t < T; t++)
      for (int h = 0; h < L; h++) {
        const int* o = d->off[h][t];
        edge_counts[((long long)l * T + t) * L + h] = o ? (long long)o[m.fr_off[h][l + 1]] - (long long)o[m.fr_off[h][l]] : 0;
      }
  }
}

// frontier rows of step blockIdx.y: typed local id + renumber map
__global__ void __launch_bounds__(256) mh_hetero_rows_kernel(const MhHeteroDesc* __restrict__ d, MhFrontiers fr, MhMeta m,
                                                             const int* __restrict__ n_rows, const int* __restrict__ tbase,
                                                             const long long* __restrict__ rmo, long long* __restrict__ map_out)
{
  const int s = blockIdx.y;
  const int n = n_rows[s];
  const int Vt = d->Vt, B = d->B;
  const long long* __restrict__ frontier = fr.frontier[s];
  const int* __restrict__ flabel         = fr.flabel[s];
  int* __restrict__ typed                = d->typed_local[s];
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const long long v = frontier[r];
    const int l       = flabel[r];
    const int vt      = mh_vtype_of(d, v);
    const int* c      = d->tcnt[s][vt];
    const int tl      = tbase[((long long)s * Vt + vt) * B + l] + c[r] - c[m.fr_off[s][l]];
    typed[r]          = tl;
    map_out[rmo[(long long)l * Vt + vt] + tl] = v;
  }
}

// edges of hop blockIdx.y -> [label][edge type][hop] groups
template <typename OutT, bool CHUNKED>
__global__ void __launch_bounds__(256) mh_hetero_emit_edges_kernel(const MhHeteroDesc* __restrict__ d,
                                                                   const int* __restrict__ n_edges, MhHopBufs hb, MhFrontiers fr,
                                                                   MhMeta m, const long long* __restrict__ lto,
                                                                   OutT* __restrict__ majors, OutT* __restrict__ minors,
                                                                   long long* __restrict__ edge_id, int* __restrict__ edge_type,
                                                                   long long* __restrict__ edge_renumber_map)
{
  const int h = blockIdx.y;
  const int n = n_edges[h];
  const int T = d->T, L = d->L;
  const int* __restrict__ erow             = hb.erow[h];
  const unsigned int* __restrict__ slot_of = hb.slot[h];
  const long long* __restrict__ gid        = hb.gid[h];
  const int* __restrict__ flabel_h         = fr.flabel[h];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    unsigned int ref = slot_of[e];
    if ((ref >> 28) == kRefPending) ref = ((unsigned int)(h + 1) << 28) | hb.rank_of[h][ref & 0x0FFFFFFFu];
    const int f  = erow[e];
    const int l  = flabel_h[f];
    const int f0 = m.fr_off[h][l];
    int t        = 0;
    for (int i = 1; i < T; i++)
      t += d->pos[h][i][f] <= e ? 1 : 0;  // pos is non-decreasing in the type for a fixed row
    const long long group = lto[((long long)l * T + t) * L];
    const long long p     = lto[((long long)l * T + t) * L + h] + (long long)(e - d->pos[h][t][f]) +
                        (long long)(d->off[h][t][f] - d->off[h][t][f0]);
    majors[p]    = (OutT)d->typed_local[h][f];
    minors[p]    = (OutT)d->typed_local[ref >> 28][ref & 0x0FFFFFFFu];
    edge_type[p] = t;
    edge_id[p]   = p - group;
    long long g  = gid[e];
    if (d->has_eid[t]) g = load_i64<CHUNKED>(d->eid[t], d->eid_off[t] + (unsigned long long)g);
    edge_renumber_map[p] = g;
  }
}

}  // namespace wgb

#include "multihop_fused.cuh"  // device half of the fused per-label path (FzArgs, kernels)

// ---------------------------------------------------------------------------------------------------
// sampler object: persistent scratch + hash table
// ---------------------------------------------------------------------------------------------------
struct wholegraph_multihop_sampler_ {
  struct Buf {
    void* p  = nullptr;
    size_t n = 0;
  };
  Buf table;
  unsigned int table_slots = 0;  // allocated capacity
  int epoch                = 0;  // last epoch handed out (0: table must be initialised)
  Buf slabel, scan_state, small_i32, small_i64, counts, seed_slot, seed_rank, seed_clean;
  long long* bad_seed_dev = nullptr;  // device flag of the call in flight: a seed was outside [0, V)
  Buf frontier[wgb::kMaxHops + 1], flabel[wgb::kMaxHops + 1], fr_off[wgb::kMaxHops + 1];
  Buf off[wgb::kMaxHops], dest[wgb::kMaxHops], erow[wgb::kMaxHops], gid[wgb::kMaxHops], slot[wgb::kMaxHops], rank_of[wgb::kMaxHops];
  Buf base;
  // heterogeneous calls only
  Buf pos[wgb::kMaxHops], tcnt[wgb::kMaxHops + 1], typed[wgb::kMaxHops + 1], vscan_state, tbase, desc_dev;
  // temporal calls only
  Buf ftime[wgb::kMaxHops + 1], eligible[wgb::kMaxHops], clipped, tdesc_dev;
  // fused per-label path (multihop_fused.cuh): label-major scratch
  Buf fz[15];
  double fz_phase_ns[32] = {};  // WGB_MH_TIMING: in-kernel phase clock of the fused path
  double fz_span_ns = 0;
  long long fz_phase_labels = 0, fz_calls = 0;
  wgb::MhTemporalDesc tdesc_host;
  wgb::MhHeteroDesc desc_host;
  long long* h_totals = nullptr;  // pinned
  int device          = -1;
  // a call between _begin and _finish
  struct Pending {
    bool active = false;
    int B = 0, L = 0, flags = 0;
    bool has_eid = false, chunked = false;
    wgb::ChunkRef eid;
    unsigned long long eid_off = 0;
    long long ub_rows[wgb::kMaxHops + 1];
    long long ub_edges[wgb::kMaxHops];
    wgb::MhFrontiers fr;
    wgb::MhMeta meta;
    wgb::MhHopBufs hb;
    long long *lho = nullptr, *rmo = nullptr, *rbase = nullptr;
    int *base = nullptr, *n_rows_dev = nullptr, *n_edges_dev = nullptr;
    int S = 0;
    const unsigned int *seed_ref = nullptr, *seed_rank = nullptr;  // per input seed: reference to its first occurrence
    const int* slabel = nullptr;
    bool finished = false;  // outputs of the last call exist; seed ids may be asked for until the next _begin
    bool hetero = false;  // typed outputs ([label][edge type][hop], ids per (label, vertex type))
    int T = 1, Vt = 1;
    int* tbase = nullptr;
#ifndef WGB_HOST_EMULATION
    bool fused = false;  // begun by multihop_begin_fused: finish copies label segments
    wgb::FzArgs fz;
#endif
  } pending;
  cudaEvent_t ready = nullptr;  // recorded after the output sizes have been copied to h_totals
  // WGB_MH_TIMING=1: per-stage device times (cudaEvents between launches), averaged, printed at destroy
  bool timing = false;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
  size_t marks_used = 0;
  std::vector<std::pair<std::string, std::pair<double, long long>>> stage_stats;
  long long timed_calls = 0;
  int warm_calls        = 0;
};

namespace wgb {

static void* ensure(wholegraph_multihop_sampler_::Buf& b, size_t bytes)
{
  if (bytes < 256) bytes = 256;
  if (b.n < bytes) {
    if (b.p) WGB_CUDA_TRY(cudaFree(b.p));
    b.p = nullptr;
    b.n = 0;
    size_t want = bytes + bytes / 8;  // slack so that slowly growing call groups do not realloc every call
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e    = cudaMalloc(&b.p, bytes);
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

static void mh_mark(wholegraph_multihop_sampler_* sp, const char* name, cudaStream_t st)
{
  if (!sp->timing) return;
  if (sp->marks_used == sp->marks.size()) {
    cudaEvent_t e;
    WGB_CUDA_TRY(cudaEventCreate(&e));
    sp->marks.emplace_back(name, e);
  }
  sp->marks[sp->marks_used].first = name;
  WGB_CUDA_TRY(cudaEventRecord(sp->marks[sp->marks_used].second, st));
  sp->marks_used++;
}

static const char* hop_stage(wholegraph_multihop_sampler_* sp, int h, const char* stage)
{
  if (!sp->timing) return stage;
  static std::vector<std::string*> pool;  // leaked on purpose: names must outlive the marks
  std::string want = "hop" + std::to_string(h) + " " + stage;
  for (auto* p : pool)
    if (*p == want) return p->c_str();
  pool.push_back(new std::string(want));
  return pool.back()->c_str();
}

// called when the stream is known to be idle past the last mark
static void mh_collect_marks(wholegraph_multihop_sampler_* sp)
{
  if (!sp->timing || sp->marks_used < 2) return;
  WGB_CUDA_TRY(cudaEventSynchronize(sp->marks[sp->marks_used - 1].second));
  if (sp->warm_calls++ < 3) {  // the first calls allocate scratch while the device waits
    sp->marks_used = 0;
    return;
  }
  for (size_t i = 1; i < sp->marks_used; i++) {
    float ms = 0.f;
    WGB_CUDA_TRY(cudaEventElapsedTime(&ms, sp->marks[i - 1].second, sp->marks[i].second));
    std::string name = sp->marks[i].first;
    size_t k = 0;
    for (; k < sp->stage_stats.size(); k++)
      if (sp->stage_stats[k].first == name) break;
    if (k == sp->stage_stats.size()) sp->stage_stats.push_back({name, {0.0, 0}});
    sp->stage_stats[k].second.first += ms;
    sp->stage_stats[k].second.second++;
  }
  sp->timed_calls++;
  sp->marks_used = 0;
}

static void mh_print_marks(wholegraph_multihop_sampler_* sp)
{
  if (!sp->timing || sp->timed_calls == 0) return;
  double total = 0;
  fprintf(stderr, "[wgb multihop] per-call device time by stage, %lld calls\n", sp->timed_calls);
  for (auto& s : sp->stage_stats) {
    fprintf(stderr, "  %-28s %9.2f us  (%lld marks)\n", s.first.c_str(), 1e3 * s.second.first / (double)sp->timed_calls, s.second.second);
    total += s.second.first;
  }
  fprintf(stderr, "  %-28s %9.2f us\n", "total", 1e3 * total / (double)sp->timed_calls);
}

static int grid_over(long long n, int sms) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)sms * 8)); }

struct MhTypeCsr {
  ChunkRef row_ptr, col, wgt, eid, etime;
  unsigned long long row_ptr_off = 0, col_off = 0, wgt_off = 0, eid_off = 0, etime_off = 0;
  bool has_eid = false;
};

struct MhCall {
  wholegraph_multihop_sampler_* sp;
  int T  = 1;                        // edge types (one CSR each)
  int Vt = 1;                        // vertex types
  bool hetero = false;               // typed outputs
  MhTypeCsr csr[kMaxEdgeTypes];
  long long vto[kMaxVertexTypes + 1];
  wholememory_dtype_t col_dtype, wgt_dtype, seed_dtype;
  bool weighted, chunked;
  const void* seeds;
  const long long* label_offsets;  // device
  int S, B, L;
  int fanout[kMaxHops * kMaxEdgeTypes];  // [h * T + t]
  unsigned long long V;
  unsigned long long random_state;
  int flags;
  cudaStream_t stream;
  // temporal calls (edge times per type in csr[t].etime)
  bool temporal                = false;
  int time_cmp                 = 0;
  const long long* seed_times  = nullptr;  // device, [S]
};

template <typename ColT, bool CHUNKED>
static void launch_hop_sample(const MhCall& c, const MhTypeCsr& g, const long long* frontier, const int* n_dev, long long ub_rows,
                              int M, unsigned long long seed, const int* off, ColT* dest, int* erow, long long* gid)
{
  int sms           = num_sms();
  const Affine* tab = skip_table_device();
  int n_ub          = (int)ub_rows;
  if (M <= 0) {
    int grid = std::max(1, std::min((n_ub + 7) / 8, sms * 8));
    sample_all_kernel<long long, ColT, CHUNKED><<<grid, 256, 0, c.stream>>>(g.row_ptr, g.row_ptr_off, g.col, g.col_off, frontier, n_ub, off, dest, erow, gid, n_dev);
  } else if (c.weighted) {
    int grid = std::max(1, std::min(n_ub, sms * 8));
    if (c.wgt_dtype == WHOLEMEMORY_DT_FLOAT) {
      if (M <= 256)
        weighted_kernel<long long, ColT, float, 128, CHUNKED><<<grid, 128, 0, c.stream>>>(g.row_ptr, g.row_ptr_off, g.col, g.col_off, g.wgt, g.wgt_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
      else
        weighted_kernel<long long, ColT, float, 256, CHUNKED><<<grid, 256, 0, c.stream>>>(g.row_ptr, g.row_ptr_off, g.col, g.col_off, g.wgt, g.wgt_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
    } else {
      if (M <= 256)
        weighted_kernel<long long, ColT, double, 128, CHUNKED><<<grid, 128, 0, c.stream>>>(g.row_ptr, g.row_ptr_off, g.col, g.col_off, g.wgt, g.wgt_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
      else
        weighted_kernel<long long, ColT, double, 256, CHUNKED><<<grid, 256, 0, c.stream>>>(g.row_ptr, g.row_ptr_off, g.col, g.col_off, g.wgt, g.wgt_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
    }
  } else if (M <= 32) {
    int batches = (n_ub + 31) / 32;
    int grid    = std::max(1, std::min((batches + 7) / 8, sms * 8));
    if (M <= 8)
      uniform_small_kernel<long long, ColT, 8, CHUNKED><<<grid, 256, 0, c.stream>>>(g.row_ptr, g.row_ptr_off, g.col, g.col_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
    else if (M <= 16)
      uniform_small_kernel<long long, ColT, 16, CHUNKED><<<grid, 256, 0, c.stream>>>(g.row_ptr, g.row_ptr_off, g.col, g.col_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
    else
      uniform_small_kernel<long long, ColT, 32, CHUNKED><<<grid, 256, 0, c.stream>>>(g.row_ptr, g.row_ptr_off, g.col, g.col_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
  } else {
    int grid = std::max(1, std::min(n_ub, sms * 16));
    uniform_general_kernel<long long, ColT, CHUNKED><<<grid, kGeneralBlock, 0, c.stream>>>(g.row_ptr, g.row_ptr_off, g.col, g.col_off, frontier, n_ub, M, seed, off, dest, erow, gid, tab, n_dev);
  }
  WGB_CHECK_LAUNCH();
}

#ifndef WGB_HOST_EMULATION
template <typename ColT>
static bool multihop_begin_fused(MhCall& c);
#endif

template <typename ColT, bool CHUNKED>
static void multihop_begin(MhCall& c)
{
#ifndef WGB_HOST_EMULATION
  c.sp->pending.fused = false;
  if (!CHUNKED && multihop_begin_fused<ColT>(c)) return;
#endif
  auto* sp        = c.sp;
  sp->pending.active   = false;  // a call that was begun but never finished is abandoned: its scratch is reused below
  sp->pending.finished = false;
  sp->marks_used = 0;
  const int sms   = num_sms();
  const int B     = c.B, L = c.L, S = c.S;
  cudaStream_t st = c.stream;

  // ---- host-side bounds ---------------------------------------------------------------------------
  long long ub_rows[kMaxHops + 1];
  long long ub_edges[kMaxHops];
  const int T = c.T;
  // per-hop sum of the positive fan-outs over the edge types; -1: some type takes all neighbours
  auto hop_fanout_sum = [&](int h) -> long long {
    long long sum = 0;
    for (int t = 0; t < T; t++) {
      if (c.fanout[h * T + t] < 0) return -1;
      sum += c.fanout[h * T + t];
    }
    return sum;
  };
  bool bounded = true;
  ub_rows[0]   = S;
  for (int h = 0; h < L; h++) {
    const long long fs = hop_fanout_sum(h);
    if (fs == 0) {
      ub_edges[h] = 0;
    } else if (fs > 0 && bounded) {
      ub_edges[h] = ub_rows[h] * fs;
    } else {
      bounded     = false;
      ub_edges[h] = -1;  // known only after the hop's count (host sync)
    }
    ub_rows[h + 1] = ub_edges[h];
  }
  WGB_EXPECTS(S < (1 << 28), "too many seeds for one call (2^28)");
  long long known_items = S;  // seeds + every hop whose edge bound is known before the call starts
  for (int h = 0; h < L && ub_edges[h] >= 0; h++)
    known_items += ub_edges[h];

  // ---- table capacity and epochs ---------------------------------------------------------------------
  auto ensure_capacity = [&](long long items) {
    unsigned long long want = (unsigned long long)std::max<long long>(2 * items, 2048);
    if (want > 0xFFFFFFF0ULL) throw logic_error("call group too large for one multi-hop call; split the seeds into more calls");
    if (sp->table_slots < want) {
      ensure(sp->table, (size_t)want * sizeof(MhSlot));
      sp->table_slots = (unsigned int)(std::min<unsigned long long>(sp->table.n / sizeof(MhSlot), 0xFFFFFFF0ULL) & ~3ULL);
      sp->epoch       = 0;
    }
  };
  auto next_epoch = [&]() -> unsigned long long {
    if (sp->epoch == 0 || sp->epoch >= 253) {
      // all-ones = "stale key, worst possible aux"; everything alive is re-inserted by the caller afterwards
      WGB_CUDA_TRY(cudaMemsetAsync(sp->table.p, 0xFF, (size_t)sp->table_slots * sizeof(MhSlot), st));
      sp->epoch = 0;
    }
    return (unsigned long long)(++sp->epoch);
  };
  ensure_capacity(known_items);
  MhSlot* table = static_cast<MhSlot*>(sp->table.p);

  // ---- small per-call device arrays -------------------------------------------------------------------
  int* small_i32     = static_cast<int*>(ensure(sp->small_i32, sizeof(int) * (size_t)(4 * (kMaxHops + 2) + kMaxHops * kMaxEdgeTypes)));
  int* n_edges_t_dev = small_i32 + 4 * (kMaxHops + 2);  // [L][T] edge counts per type (T > 1)
  int* n_rows_dev    = small_i32;                                                          // [L+1] frontier sizes
  int* n_edges_dev   = small_i32 + (kMaxHops + 1);                                         // [L]   edge counts
  unsigned int* nslots_dev = reinterpret_cast<unsigned int*>(small_i32 + 2 * (kMaxHops + 1));  // [L+1] slots in use per step
  int* slabel = static_cast<int*>(ensure(sp->slabel, sizeof(int) * (size_t)std::max(S, 1)));
  for (int t = 0; t <= L; t++)
    ensure(sp->fr_off[t], sizeof(int) * (size_t)(B + 1));
  ensure(sp->base, sizeof(int) * (size_t)(L + 1) * (size_t)std::max(B, 1));
  WGB_CUDA_TRY(cudaMemsetAsync(n_edges_dev, 0, sizeof(int) * kMaxHops, st));
  MhHeteroDesc& desc = sp->desc_host;
  if (c.hetero) {
    memset(&desc, 0, sizeof(desc));
    desc.T = T; desc.Vt = c.Vt; desc.L = L; desc.B = B;
    for (int v = 0; v <= c.Vt; v++) desc.vto[v] = c.vto[v];
    for (int t = 0; t < T; t++) {
      desc.eid[t]     = c.csr[t].eid;
      desc.eid_off[t] = c.csr[t].eid_off;
      desc.has_eid[t] = c.csr[t].has_eid ? 1 : 0;
    }
  }

  const MhTemporalDesc* tdesc_dev = nullptr;
  if (c.temporal) {
    MhTemporalDesc& td = sp->tdesc_host;
    memset(&td, 0, sizeof(td));
    td.T = T;
    for (int t = 0; t < T; t++) {
      td.etime[t]     = c.csr[t].etime;
      td.etime_off[t] = c.csr[t].etime_off;
    }
    void* p = ensure(sp->tdesc_dev, sizeof(MhTemporalDesc));
    WGB_CUDA_TRY(cudaMemcpyAsync(p, &td, sizeof(MhTemporalDesc), cudaMemcpyHostToDevice, st));
    tdesc_dev = static_cast<const MhTemporalDesc*>(p);
  }

  auto scan_slice = [&](long long n_items, int tile_items = kScanTile) {
    int tiles    = (int)((n_items + tile_items) / tile_items);
    size_t bytes = scan_state_bytes(tiles);
    void* p      = ensure(sp->scan_state, bytes);
    WGB_CUDA_TRY(cudaMemsetAsync(p, 0, bytes, st));
    return std::make_pair(static_cast<unsigned long long*>(p), tiles);
  };
  auto scan_grid = [&](int tiles) { return std::min(tiles, sms * 8); };
  auto ticket_of = [](std::pair<unsigned long long*, int> ss) { return reinterpret_cast<unsigned int*>(ss.first + ss.second); };

  MhFrontiers fr;
  MhMeta meta;
  MhHopBufs hb;
  memset(&fr, 0, sizeof(fr));
  memset(&meta, 0, sizeof(meta));
  memset(&hb, 0, sizeof(hb));
  meta.L = L;
  meta.B = B;

  mh_mark(sp, "start", st);
  // ---- step 0: seeds -> frontier_0 (dedup per label, first occurrence keeps the id) --------------------
  {
    unsigned int nslots0 = (unsigned int)(std::min<unsigned long long>(sp->table_slots, (unsigned long long)S * 2 + 1024) & ~3ULL);
    WGB_CUDA_TRY(cudaMemcpyAsync(nslots_dev, &nslots0, sizeof(unsigned int), cudaMemcpyHostToDevice, st));
    unsigned long long epoch = next_epoch();
    unsigned int* slot0  = static_cast<unsigned int*>(ensure(sp->seed_slot, sizeof(unsigned int) * (size_t)std::max(S, 1)));
    unsigned int* rank0  = static_cast<unsigned int*>(ensure(sp->seed_rank, sizeof(unsigned int) * (size_t)std::max(S, 1)));
    long long* frontier0 = static_cast<long long*>(ensure(sp->frontier[0], sizeof(long long) * (size_t)std::max(S, 1)));
    int* flabel0         = static_cast<int*>(ensure(sp->flabel[0], sizeof(int) * (size_t)std::max(S, 1)));
    auto ss              = scan_slice(S, kCompactTile);
    // seeds are range-checked on the device (no host copy of them exists): bad_seed travels to the host with the sizes
    long long* clean    = static_cast<long long*>(ensure(sp->seed_clean, sizeof(long long) * (size_t)std::max(S, 1) + 16));
    long long* bad_seed = clean + std::max(S, 1);
    WGB_CUDA_TRY(cudaMemsetAsync(bad_seed, 0, sizeof(long long), st));
    sp->bad_seed_dev = bad_seed;
    if (c.seed_dtype == WHOLEMEMORY_DT_INT)
      mh_seed_insert_kernel<int><<<grid_over(S, sms), 256, 0, st>>>(table, nslots_dev, epoch, c.V, static_cast<const int*>(c.seeds), S, c.label_offsets, B, slabel, slot0, clean, bad_seed);
    else
      mh_seed_insert_kernel<long long><<<grid_over(S, sms), 256, 0, st>>>(table, nslots_dev, epoch, c.V, static_cast<const long long*>(c.seeds), S, c.label_offsets, B, slabel, slot0, clean, bad_seed);
    WGB_CHECK_LAUNCH();
    mh_compact_kernel<long long, true><<<scan_grid(ss.second), kScanBlock, 0, st>>>(table, 0u, clean, nullptr, S, nullptr, slabel, slot0, rank0, frontier0, flabel0, n_rows_dev, ss.first, ticket_of(ss));
    WGB_CHECK_LAUNCH();
    if (c.temporal) {
      long long* ftime0 = static_cast<long long*>(ensure(sp->ftime[0], sizeof(long long) * (size_t)std::max(S, 1)));
      mh_seed_time_kernel<<<grid_over(S, sms), 256, 0, st>>>(slot0, rank0, c.seed_times, S, ftime0);
      WGB_CHECK_LAUNCH();
    }
    mh_mark(sp, "seeds", st);
    fr.frontier[0]  = frontier0;
    fr.flabel[0]    = flabel0;
    meta.fr_off[0]  = static_cast<int*>(sp->fr_off[0].p);
  }

  // ---- hops ---------------------------------------------------------------------------------------------
  for (int h = 0; h < L; h++) {
    const long long rows_ub = ub_rows[h];
    long long* frontier     = static_cast<long long*>(sp->frontier[h].p);
    int* flabel             = static_cast<int*>(sp->flabel[h].p);
    const size_t off_stride = (size_t)(std::max<long long>(rows_ub, 0) + 1 + kScanTile);
    int* off_all            = static_cast<int*>(ensure(sp->off[h], sizeof(int) * off_stride * (size_t)T));
    int* pos_all            = T > 1 ? static_cast<int*>(ensure(sp->pos[h], sizeof(int) * off_stride * (size_t)T)) : nullptr;
    long long edges_ub      = ub_edges[h];
    const long long fs      = hop_fanout_sum(h);
    meta.off[h]             = nullptr;
    const int* off_t[kMaxEdgeTypes];  // null: the type is skipped in this hop
    const int* pos_t[kMaxEdgeTypes];
    for (int t = 0; t < T; t++)
      off_t[t] = pos_t[t] = nullptr;
    if (c.temporal) {
      ensure(sp->eligible[h], sizeof(int) * off_stride * (size_t)T);
      ensure(sp->clipped, sizeof(int) * off_stride);
    }
    if (fs != 0 && rows_ub > 0) {
      // K1: counts + scan over the frontier, once per edge type
      for (int t = 0; t < T; t++) {
        const int M = c.fanout[h * T + t];
        if (M == 0) continue;
        int* off = off_all + (size_t)t * off_stride;
        auto ss  = scan_slice(rows_ub);
        int* total_out = T > 1 ? n_edges_t_dev + h * T + t : n_edges_dev + h;
        if (c.temporal) {
          // eligible edges per row (the whole row is read), clipped to the fan-out, then the same single-pass scan
          int* eligible = static_cast<int*>(sp->eligible[h].p) + (size_t)t * off_stride;
          int* clipped  = static_cast<int*>(sp->clipped.p);
          temporal_count_kernel<CHUNKED><<<std::max(1, (int)std::min<long long>((rows_ub + 7) / 8, (long long)sms * 8)), 256, 0, st>>>(
            c.csr[t].row_ptr, c.csr[t].row_ptr_off, c.csr[t].etime, c.csr[t].etime_off, frontier, static_cast<const long long*>(sp->ftime[h].p), M,
            c.time_cmp, eligible, clipped, n_rows_dev + h);
          WGB_CHECK_LAUNCH();
          scan_counts_kernel<<<scan_grid(ss.second), kScanBlock, 0, st>>>(clipped, off, ss.first, ticket_of(ss), n_rows_dev + h, total_out);
        } else {
          count_scan_kernel<long long, CHUNKED><<<scan_grid(ss.second), kScanBlock, 0, st>>>(c.csr[t].row_ptr, c.csr[t].row_ptr_off, frontier, (int)rows_ub, M, off, ss.first, ticket_of(ss), n_rows_dev + h, total_out);
        }
        WGB_CHECK_LAUNCH();
        off_t[t] = off;
      }
      if (T > 1) {
        MhCombine cb;
        cb.T = T;
        for (int t = 0; t < T; t++) {
          cb.off[t] = off_t[t];
          cb.pos[t] = pos_all + (size_t)t * off_stride;
          pos_t[t]  = cb.pos[t];
        }
        mh_combine_types_kernel<<<grid_over(rows_ub + 1, sms), 256, 0, st>>>(cb, n_rows_dev + h, n_edges_dev + h);
        WGB_CHECK_LAUNCH();
      } else {
        pos_t[0] = off_t[0];
      }
      mh_mark(sp, hop_stage(sp, h, "count+scan"), st);
      meta.off[h] = off_t[0];
      if (edges_ub < 0) {
        // take-all hop: the edge count is data dependent -> one extra host sync to size the scratch
        int e_host = 0;
        WGB_CUDA_TRY(cudaMemcpyAsync(&e_host, n_edges_dev + h, sizeof(int), cudaMemcpyDeviceToHost, st));
        WGB_CUDA_TRY(cudaStreamSynchronize(st));
        WGB_EXPECTS(e_host >= 0, "edge count of a take-all hop overflowed int32");
        edges_ub = e_host;
        known_items += e_host;
        ub_edges[h]    = edges_ub;
        ub_rows[h + 1] = edges_ub;
        for (int hh = h + 1; hh < L; hh++) {  // later bounded hops can now be bounded too
          const long long fs2 = hop_fanout_sum(hh);
          if (fs2 < 0) break;
          ub_edges[hh]    = ub_rows[hh] * fs2;
          ub_rows[hh + 1] = ub_edges[hh];
          known_items += ub_edges[hh];
        }
        ensure_capacity(known_items);  // may reallocate: everything alive is re-inserted below anyway
        table = static_cast<MhSlot*>(sp->table.p);
      }
    } else {
      edges_ub       = 0;
      ub_edges[h]    = 0;
      ub_rows[h + 1] = 0;
    }
    if (c.hetero)
      for (int t = 0; t < T; t++) {
        desc.off[h][t] = off_t[t];
        desc.pos[h][t] = pos_t[t];
      }
    WGB_EXPECTS(edges_ub < (1LL << 28), "too many edges in one hop for one call (2^28); split the seeds into more calls");
    const size_t ecap = (size_t)std::max<long long>(edges_ub, 1);
    ColT* dest               = static_cast<ColT*>(ensure(sp->dest[h], sizeof(ColT) * ecap));
    int* erow                = static_cast<int*>(ensure(sp->erow[h], sizeof(int) * ecap));
    long long* gid           = static_cast<long long*>(ensure(sp->gid[h], sizeof(long long) * ecap));
    unsigned int* slot       = static_cast<unsigned int*>(ensure(sp->slot[h], sizeof(unsigned int) * ecap));
    unsigned int* rank_of    = static_cast<unsigned int*>(ensure(sp->rank_of[h], sizeof(unsigned int) * ecap));
    long long* next_frontier = static_cast<long long*>(ensure(sp->frontier[h + 1], sizeof(long long) * ecap));
    int* next_flabel         = static_cast<int*>(ensure(sp->flabel[h + 1], sizeof(int) * ecap));
    hb.off[h]  = off_t[0];
    hb.erow[h] = erow;
    hb.slot[h] = slot;
    hb.rank_of[h] = rank_of;
    hb.gid[h]  = gid;
    if (edges_ub > 0) {
      // new epoch for this hop: pick the slot count on the device, re-insert what is numbered so far
      unsigned long long epoch = next_epoch();
      mh_plan_hop_kernel<<<1, 32, 0, st>>>(n_rows_dev, h, n_edges_dev + h, sp->table_slots, nslots_dev + h + 1);
      WGB_CHECK_LAUNCH();
      long long max_rows = 1;
      for (int t = 0; t <= h; t++)
        max_rows = std::max(max_rows, ub_rows[t]);
      mh_reinsert_kernel<<<dim3(grid_over(max_rows, sms), h + 1), 256, 0, st>>>(table, nslots_dev + h + 1, epoch, c.V, fr, n_rows_dev);
      WGB_CHECK_LAUNCH();
      mh_mark(sp, hop_stage(sp, h, "plan+reinsert"), st);
      // K2: sample, once per edge type; type t writes its edges of row f at pos_t[f]
      const unsigned long long hop_seed = c.random_state + (unsigned long long)h * 0x9E3779B97F4A7C15ULL;
      for (int t = 0; t < T; t++) {
        if (!off_t[t]) continue;
        const unsigned long long type_seed = hop_seed + (unsigned long long)t * 0xD1B54A32D192ED03ULL;
        if (c.temporal) {
          const int* eligible    = static_cast<const int*>(sp->eligible[h].p) + (size_t)t * off_stride;
          const long long* ftime = static_cast<const long long*>(sp->ftime[h].p);
          const MhTypeCsr& g     = c.csr[t];
          const int M            = c.fanout[h * T + t];
          const Affine* tab      = skip_table_device();
          if (c.weighted) {
            const int grid = std::max(1, (int)std::min<long long>(rows_ub, (long long)sms * 8));
            auto launch = [&](auto wt_tag, auto block_tag) {
              using WT             = decltype(wt_tag);
              constexpr int kBlock = decltype(block_tag)::value;
              temporal_weighted_kernel<ColT, WT, kBlock, CHUNKED><<<grid, kBlock, 0, st>>>(
                g.row_ptr, g.row_ptr_off, g.col, g.col_off, g.wgt, g.wgt_off, g.etime, g.etime_off, frontier, ftime, eligible, M, c.time_cmp,
                type_seed, pos_t[t], dest, erow, gid, tab, n_rows_dev + h);
            };
            // BLOCK is part of the random-stream geometry (thread j of BLOCK owns positions j, j + BLOCK, ...), as in launch_hop_sample
            if (c.wgt_dtype == WHOLEMEMORY_DT_FLOAT) {
              if (M <= 256) launch(0.f, std::integral_constant<int, 128>{});
              else launch(0.f, std::integral_constant<int, 256>{});
            } else {
              if (M <= 256) launch(0.0, std::integral_constant<int, 128>{});
              else launch(0.0, std::integral_constant<int, 256>{});
            }
          } else {
            temporal_uniform_kernel<ColT, CHUNKED><<<std::max(1, (int)std::min<long long>(rows_ub, (long long)sms * 16)), kGeneralBlock, 0, st>>>(
              g.row_ptr, g.row_ptr_off, g.col, g.col_off, g.etime, g.etime_off, frontier, ftime, eligible, M, c.time_cmp, type_seed, pos_t[t],
              dest, erow, gid, tab, n_rows_dev + h);
          }
          WGB_CHECK_LAUNCH();
        } else {
          launch_hop_sample<ColT, CHUNKED>(c, c.csr[t], frontier, n_rows_dev + h, rows_ub, c.fanout[h * T + t], type_seed, pos_t[t], dest, erow, gid);
        }
      }
      mh_mark(sp, hop_stage(sp, h, "sample"), st);
      // K3: insert (label, neighbour)
      mh_insert_kernel<ColT><<<std::max(1, (int)std::min<long long>((edges_ub + 255) / 256, (long long)sms * 16)), 256, 0, st>>>(table, nslots_dev + h + 1, epoch, c.V, (unsigned int)(h + 1), dest, n_edges_dev + h, erow, flabel, slot);
      WGB_CHECK_LAUNCH();
      mh_mark(sp, hop_stage(sp, h, "insert"), st);
      // K4: first occurrences -> next frontier
      auto ss = scan_slice(edges_ub, kCompactTile);
      mh_compact_kernel<ColT, false><<<scan_grid(ss.second), kScanBlock, 0, st>>>(table, (unsigned int)(h + 1), dest, n_edges_dev + h, 0, erow, flabel, slot, rank_of, next_frontier, next_flabel, n_rows_dev + h + 1, ss.first, ticket_of(ss));
      WGB_CHECK_LAUNCH();
      if (c.temporal) {
        long long* next_ftime = static_cast<long long*>(ensure(sp->ftime[h + 1], sizeof(long long) * ecap));
        MhTypePos tp;
        for (int t = 0; t < kMaxEdgeTypes; t++)
          tp.pos[t] = t < T ? pos_t[t] : nullptr;
        mh_next_time_kernel<CHUNKED><<<grid_over(edges_ub, sms), 256, 0, st>>>(tdesc_dev, tp, n_edges_dev + h, slot, rank_of, erow, gid, next_ftime);
        WGB_CHECK_LAUNCH();
      }
      mh_mark(sp, hop_stage(sp, h, "compact"), st);
    } else {
      WGB_CUDA_TRY(cudaMemsetAsync(n_rows_dev + h + 1, 0, sizeof(int), st));
    }
    fr.frontier[h + 1] = next_frontier;
    fr.flabel[h + 1]   = next_flabel;
    meta.fr_off[h + 1] = static_cast<int*>(sp->fr_off[h + 1].p);
  }

  // ---- per-label bookkeeping + offsets ---------------------------------------------------------------------
  const int Vt             = c.Vt;
  const long long n_groups = c.hetero ? (long long)B * T * L : (long long)B * L;  // edge groups
  const long long n_maps   = c.hetero ? (long long)B * Vt : (long long)B;          // renumber map segments
  const long long n_counts = n_groups + n_maps + B;
  long long* counts = static_cast<long long*>(ensure(sp->counts, sizeof(long long) * (size_t)(n_counts + 1)));
  long long* scans  = static_cast<long long*>(ensure(sp->small_i64, sizeof(long long) * (size_t)(n_counts + 8)));
  long long* lho    = scans;                 // n_groups + 1
  long long* rmo    = scans + n_groups + 1;  // n_maps + 1
  long long* rbase  = rmo + n_maps + 1;      // B + 1 (homogeneous CSR only)
  long long* totals = rbase + B + 1;         // 3
  int* base         = static_cast<int*>(sp->base.p);
  int* tbase        = nullptr;
  {
    MhLabelBounds lb;
    long long most = B + 1;
    for (int t = 0; t <= L; t++) {
      lb.flabel[t] = fr.flabel[t];
      lb.fr_off[t] = static_cast<int*>(sp->fr_off[t].p);
      most         = std::max(most, ub_rows[t]);
    }
    mh_label_bounds_kernel<<<dim3(grid_over(most, sms), L + 1), 256, 0, st>>>(lb, n_rows_dev, B);
    WGB_CHECK_LAUNCH();
  }
  MhScan3 sc;
  if (!c.hetero) {
    mh_meta_kernel<<<grid_over(B, sms), 256, 0, st>>>(meta, counts, base);
    WGB_CHECK_LAUNCH();
    sc.in[0] = counts;                           sc.out[0] = lho;   sc.n[0] = (long long)B * L;
    sc.in[1] = counts + (long long)B * L;        sc.out[1] = rmo;   sc.n[1] = B;
    sc.in[2] = counts + (long long)B * L + B;    sc.out[2] = rbase; sc.n[2] = B;
  } else {
    // per (step, vertex type): exclusive counts over the frontier rows, one launch for all slices
    MhVtypeScan vs;
    long long tiles_max = 1;
    for (int t = 0; t <= L; t++) {
      const long long cap = std::max<long long>(ub_rows[t], 0);
      vs.cap[t]           = cap;
      vs.tcnt[t]          = static_cast<int*>(ensure(sp->tcnt[t], sizeof(int) * (size_t)(cap + 1) * (size_t)Vt));
      desc.typed_local[t] = static_cast<int*>(ensure(sp->typed[t], sizeof(int) * (size_t)std::max<long long>(cap, 1)));
      for (int vt = 0; vt < Vt; vt++)
        desc.tcnt[t][vt] = vs.tcnt[t] + (long long)vt * (cap + 1);
      tiles_max = std::max(tiles_max, (cap + kScanTile) / kScanTile);
    }
    vs.state_stride   = tiles_max + 2;
    const size_t vbytes = sizeof(unsigned long long) * (size_t)vs.state_stride * (size_t)((L + 1) * Vt);
    vs.state          = static_cast<unsigned long long*>(ensure(sp->vscan_state, vbytes));
    WGB_CUDA_TRY(cudaMemsetAsync(vs.state, 0, vbytes, st));
    MhHeteroDesc* desc_dev = static_cast<MhHeteroDesc*>(ensure(sp->desc_dev, sizeof(MhHeteroDesc)));
    WGB_CUDA_TRY(cudaMemcpyAsync(desc_dev, &desc, sizeof(MhHeteroDesc), cudaMemcpyHostToDevice, st));
    mh_vtype_scan_kernel<<<dim3((unsigned int)std::min<long long>(tiles_max, (long long)sms * 4), (L + 1) * Vt), kScanBlock, 0, st>>>(desc_dev, fr, n_rows_dev, vs);
    WGB_CHECK_LAUNCH();
    tbase = static_cast<int*>(ensure(sp->tbase, sizeof(int) * (size_t)(L + 1) * (size_t)Vt * (size_t)std::max(B, 1)));
    mh_hetero_meta_kernel<<<grid_over(B, sms), 256, 0, st>>>(desc_dev, meta, counts, counts + n_groups, tbase);
    WGB_CHECK_LAUNCH();
    sc.in[0] = counts;            sc.out[0] = lho;   sc.n[0] = n_groups;
    sc.in[1] = counts + n_groups; sc.out[1] = rmo;   sc.n[1] = n_maps;
    sc.in[2] = counts;            sc.out[2] = rbase; sc.n[2] = 0;
  }
  sc.totals = totals;
  mh_scan3_kernel<<<3, 1024, 0, st>>>(sc);
  WGB_CHECK_LAUNCH();
  // ---- output sizes travel to pinned memory; _finish waits for `ready`, not for the stream ----------------------
  WGB_CUDA_TRY(cudaMemcpyAsync(sp->h_totals, totals, 3 * sizeof(long long), cudaMemcpyDeviceToHost, st));
  WGB_CUDA_TRY(cudaMemcpyAsync(sp->h_totals + 3, sp->bad_seed_dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
  WGB_CUDA_TRY(cudaEventRecord(sp->ready, st));
  mh_mark(sp, "meta+scan3", st);
  auto& pd = sp->pending;
  pd.B = B; pd.L = L; pd.flags = c.flags; pd.has_eid = c.csr[0].has_eid; pd.chunked = CHUNKED;
  pd.eid = c.csr[0].eid; pd.eid_off = c.csr[0].eid_off;
  pd.hetero = c.hetero; pd.T = T; pd.Vt = Vt; pd.tbase = tbase;
  pd.S = S; pd.seed_ref = static_cast<const unsigned int*>(sp->seed_slot.p); pd.seed_rank = static_cast<const unsigned int*>(sp->seed_rank.p);
  pd.slabel = slabel;
  for (int t = 0; t <= L; t++) pd.ub_rows[t] = ub_rows[t];
  for (int h = 0; h < L; h++) pd.ub_edges[h] = ub_edges[h];
  pd.fr = fr; pd.meta = meta; pd.hb = hb;
  pd.lho = lho; pd.rmo = rmo; pd.rbase = rbase; pd.base = base;
  pd.n_rows_dev = n_rows_dev; pd.n_edges_dev = n_edges_dev;
  pd.active = true;
}

struct MhOutCtx {
  void *majors, *minors, *edge_id, *lho, *map, *rmo, *major_offsets, *step_counts;
  void *edge_type = nullptr, *edge_renumber_map = nullptr, *edge_renumber_map_offsets = nullptr;  // heterogeneous calls
  wholememory_env_func_t* env;
  cudaStream_t stream;
};

}  // namespace wgb
#define WGB_FZ_HOST_HALF
#include "multihop_fused.cuh"  // host half: multihop_begin_fused / multihop_finish_fused
namespace wgb {

// second half of a call: wait for the sizes, allocate the outputs, scatter scratch -> outputs
static void multihop_finish(wholegraph_multihop_sampler_* sp, const MhOutCtx& c)
{
  WGB_EXPECTS(sp->pending.active, "no call in flight on this sampler object");
#ifndef WGB_HOST_EMULATION
  if (sp->pending.fused) {
    multihop_finish_fused(sp, c);
    return;
  }
#endif
  auto& pd        = sp->pending;
  pd.active       = false;
  pd.finished     = true;
  const int sms   = num_sms();
  const int B = pd.B, L = pd.L;
  cudaStream_t st = c.stream;
  const MhFrontiers& fr = pd.fr;
  const MhMeta& meta    = pd.meta;
  const MhHopBufs& hb   = pd.hb;
  long long *lho = pd.lho, *rmo = pd.rmo, *rbase = pd.rbase;
  int *base = pd.base, *n_rows_dev = pd.n_rows_dev, *n_edges_dev = pd.n_edges_dev;
  const long long* ub_rows  = pd.ub_rows;
  const long long* ub_edges = pd.ub_edges;
  // ---- the one host wait of the call ------------------------------------------------------------------------------
  WGB_CUDA_TRY(cudaEventSynchronize(sp->ready));
  WGB_CUDA_TRY(cudaStreamWaitEvent(st, sp->ready, 0));  // no-op on the stream the call was begun on
  const long long n_edges = sp->h_totals[0], n_nodes = sp->h_totals[1], n_srcrows = sp->h_totals[2];
  if (sp->h_totals[3] != 0) throw invalid_input("a seed vertex id is outside [0, number of vertices)");

  const bool csr    = (pd.flags & WHOLEGRAPH_MULTIHOP_CSR) != 0;
  const bool idx64  = (pd.flags & WHOLEGRAPH_MULTIHOP_INT64_IDS) != 0;
  const wholememory_dtype_t idx_dt = idx64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT;
  if (pd.hetero) {
    const int T = pd.T, Vt = pd.Vt;
    WGB_EXPECTS(!csr, "CSR compression is not defined for heterogeneous sampling (as in the reference)");
    WGB_EXPECTS(c.majors && c.edge_type && c.edge_renumber_map && c.edge_renumber_map_offsets, "heterogeneous call finished through the homogeneous entry point");
    void* out_majors    = output_alloc(c.env, c.majors, n_edges, idx_dt);
    void* out_minors    = output_alloc(c.env, c.minors, n_edges, idx_dt);
    long long* out_eid  = static_cast<long long*>(output_alloc(c.env, c.edge_id, n_edges, WHOLEMEMORY_DT_INT64));
    int* out_etype      = static_cast<int*>(output_alloc(c.env, c.edge_type, n_edges, WHOLEMEMORY_DT_INT));
    long long* out_lto  = static_cast<long long*>(output_alloc(c.env, c.lho, (long long)B * T * L + 1, WHOLEMEMORY_DT_INT64));
    long long* out_map  = static_cast<long long*>(output_alloc(c.env, c.map, n_nodes, WHOLEMEMORY_DT_INT64));
    long long* out_rmo  = static_cast<long long*>(output_alloc(c.env, c.rmo, (long long)B * Vt + 1, WHOLEMEMORY_DT_INT64));
    long long* out_emap = static_cast<long long*>(output_alloc(c.env, c.edge_renumber_map, n_edges, WHOLEMEMORY_DT_INT64));
    long long* out_ermo = static_cast<long long*>(output_alloc(c.env, c.edge_renumber_map_offsets, (long long)B * T + 1, WHOLEMEMORY_DT_INT64));
    if (c.step_counts) {
      int* out_sc = static_cast<int*>(output_alloc(c.env, c.step_counts, (long long)(L + 1) * Vt * B, WHOLEMEMORY_DT_INT));
      WGB_CUDA_TRY(cudaMemcpyAsync(out_sc, pd.tbase, sizeof(int) * (size_t)(L + 1) * (size_t)Vt * (size_t)B, cudaMemcpyDeviceToDevice, st));
    }
    WGB_CUDA_TRY(cudaMemcpyAsync(out_rmo, rmo, sizeof(long long) * (size_t)((long long)B * Vt + 1), cudaMemcpyDeviceToDevice, st));
    WGB_CUDA_TRY(cudaMemcpyAsync(out_lto, lho, sizeof(long long) * (size_t)((long long)B * T * L + 1), cudaMemcpyDeviceToDevice, st));
    // edge_renumber_map_offsets[l*T + t] = label_type_hop_offsets[(l*T + t) * L]
    WGB_CUDA_TRY(cudaMemcpy2DAsync(out_ermo, sizeof(long long), lho, sizeof(long long) * (size_t)L, sizeof(long long), (size_t)((long long)B * T + 1), cudaMemcpyDeviceToDevice, st));
    const MhHeteroDesc* desc_dev = static_cast<const MhHeteroDesc*>(sp->desc_dev.p);
    long long max_edges = 0, max_rows = 1;
    for (int h = 0; h < L; h++)
      max_edges = std::max(max_edges, ub_edges[h]);
    for (int t = 0; t <= L; t++)
      max_rows = std::max(max_rows, ub_rows[t]);
    if (n_nodes > 0) {
      mh_hetero_rows_kernel<<<dim3(grid_over(std::min(max_rows, n_nodes), sms), L + 1), 256, 0, st>>>(desc_dev, fr, meta, n_rows_dev, pd.tbase, rmo, out_map);
      WGB_CHECK_LAUNCH();
    }
    if (n_edges > 0 && max_edges > 0) {
      dim3 grid(grid_over(std::min(max_edges, n_edges), sms), L);
      auto emit = [&](auto out_tag, auto chunk_tag) {
        using OutT                = decltype(out_tag);
        constexpr bool kChunkedId = decltype(chunk_tag)::value;
        mh_hetero_emit_edges_kernel<OutT, kChunkedId><<<grid, 256, 0, st>>>(desc_dev, n_edges_dev, hb, fr, meta, lho, static_cast<OutT*>(out_majors), static_cast<OutT*>(out_minors), out_eid, out_etype, out_emap);
      };
      if (idx64) {
        if (pd.chunked) emit((long long)0, std::true_type{});
        else emit((long long)0, std::false_type{});
      } else {
        if (pd.chunked) emit((int)0, std::true_type{});
        else emit((int)0, std::false_type{});
      }
      WGB_CHECK_LAUNCH();
    }
    mh_mark(sp, "emit", st);
    mh_collect_marks(sp);
    return;
  }
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
    // vertices label l discovered at step t (t = 0: its seeds): base[(t+1)*B + l] - base[t*B + l]; handed out as
    // the [L+1, B] table of first local ids plus the per-label totals so that readers never have to reduce
    // over the edge arrays (the reference's decoders do, with a host sync per hop: sampler.py:570-575)
    int* out_sc = static_cast<int*>(output_alloc(c.env, c.step_counts, (long long)(L + 1) * B, WHOLEMEMORY_DT_INT));
    WGB_CUDA_TRY(cudaMemcpyAsync(out_sc, base, sizeof(int) * (size_t)(L + 1) * (size_t)B, cudaMemcpyDeviceToDevice, st));
  }
  WGB_CUDA_TRY(cudaMemcpyAsync(out_rmo, rmo, sizeof(long long) * (size_t)(B + 1), cudaMemcpyDeviceToDevice, st));
  if (!csr) {
    WGB_CUDA_TRY(cudaMemcpyAsync(out_lho, lho, sizeof(long long) * (size_t)((long long)B * L + 1), cudaMemcpyDeviceToDevice, st));
  } else {
    mh_csr_label_hop_kernel<<<grid_over((long long)B * L + 1, sms), 256, 0, st>>>(L, B, base, rbase, lho, out_lho, out_moff);
    WGB_CHECK_LAUNCH();
  }
  // ---- final pass: edges and rows to their label-major places ---------------------------------------------------
  long long max_edges = 0, max_rows = 1;
  for (int h = 0; h < L; h++)
    max_edges = std::max(max_edges, ub_edges[h]);
  for (int t = 0; t <= L; t++)
    max_rows = std::max(max_rows, ub_rows[t]);
  if (n_edges > 0 && max_edges > 0) {
    // tighter than the upper bound: no hop has more edges than the call has in total
    dim3 grid(grid_over(std::min(max_edges, n_edges), sms), L);
    auto emit = [&](auto out_tag, auto chunk_tag) {
      using OutT                = decltype(out_tag);
      constexpr bool kChunkedId = decltype(chunk_tag)::value;
      mh_emit_edges_kernel<OutT, kChunkedId><<<grid, 256, 0, st>>>(L, B, n_edges_dev, hb, fr, meta, base, lho, pd.eid, pd.eid_off, pd.has_eid, static_cast<OutT*>(out_majors), static_cast<OutT*>(out_minors), out_eid);
    };
    if (idx64) {
      if (pd.chunked) emit((long long)0, std::true_type{});
      else emit((long long)0, std::false_type{});
    } else {
      if (pd.chunked) emit((int)0, std::true_type{});
      else emit((int)0, std::false_type{});
    }
    WGB_CHECK_LAUNCH();
  }
  if (n_nodes > 0) {
    mh_emit_rows_kernel<<<dim3(grid_over(std::min(max_rows, n_nodes), sms), L + 1), 256, 0, st>>>(L, B, n_rows_dev, fr, meta, base, rmo, lho, rbase, out_map, out_moff);
    WGB_CHECK_LAUNCH();
  }
  mh_mark(sp, "emit", st);
  mh_collect_marks(sp);
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
    WGB_CUDA_TRY(cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming));
    const char* t = getenv("WGB_MH_TIMING");
    s->timing     = t && atoi(t) > 0;
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
  drop(s->table); drop(s->seed_slot); drop(s->seed_rank); drop(s->seed_clean); drop(s->slabel); drop(s->scan_state); drop(s->small_i32); drop(s->small_i64); drop(s->counts); drop(s->base);
  for (int i = 0; i <= wgb::kMaxHops; i++) {
    drop(s->frontier[i]); drop(s->flabel[i]); drop(s->fr_off[i]); drop(s->tcnt[i]); drop(s->typed[i]);
  }
  drop(s->vscan_state); drop(s->tbase); drop(s->desc_dev); drop(s->clipped); drop(s->tdesc_dev);
  for (auto& b : s->fz)
    drop(b);
  for (int i = 0; i <= wgb::kMaxHops; i++)
    drop(s->ftime[i]);
  for (int i = 0; i < wgb::kMaxHops; i++)
    drop(s->eligible[i]);
  for (int i = 0; i < wgb::kMaxHops; i++) {
    drop(s->off[i]); drop(s->dest[i]); drop(s->erow[i]); drop(s->gid[i]); drop(s->slot[i]); drop(s->rank_of[i]); drop(s->pos[i]);
  }
  if (s->h_totals) cudaFreeHost(s->h_totals);
  if (s->ready) cudaEventDestroy(s->ready);
  wgb::mh_print_marks(s);
#ifndef WGB_HOST_EMULATION
  wgb::fz_print_phases(s);
#endif
  for (auto& m : s->marks)
    cudaEventDestroy(m.second);
  cudaGetLastError();
  delete s;
  return WHOLEMEMORY_SUCCESS;
}

namespace wgb {

// shared by the homogeneous and the heterogeneous entry points: validate, fill the call, dispatch on (col dtype, chunked)
static wholememory_error_code_t multihop_begin_entry(const char* what, wholegraph_multihop_sampler_t sampler, int T,
                                                     const wholememory_tensor_t* csr_row_ptr, const wholememory_tensor_t* csr_col,
                                                     const wholememory_tensor_t* csr_weight, const wholememory_tensor_t* csr_edge_id,
                                                     const long long* vertex_type_offsets, int Vt, bool hetero,
                                                     wholememory_tensor_t seeds, wholememory_tensor_t label_offsets,
                                                     const int* fanout, int num_hops, unsigned long long random_state, int flags,
                                                     void* stream, const wholememory_tensor_t* csr_edge_time = nullptr,
                                                     wholememory_tensor_t seed_times = nullptr, int time_cmp = 0)
{
  if (!sampler || !csr_row_ptr || !csr_col || !seeds || !label_offsets || !fanout) return WHOLEMEMORY_INVALID_INPUT;
  const bool temporal = csr_edge_time != nullptr;
  if (temporal) {
    if (!seed_times || time_cmp < kTimeStrictlyIncreasing || time_cmp > kTimeMonotonicallyDecreasing) return WHOLEMEMORY_INVALID_INPUT;
    auto* td = wholememory_tensor_get_tensor_description(seed_times);
    auto* s0 = wholememory_tensor_get_tensor_description(seeds);
    if (td->dim != 1 || td->dtype != WHOLEMEMORY_DT_INT64 || td->sizes[0] != s0->sizes[0]) return WHOLEMEMORY_INVALID_INPUT;
  }
  if (T < 1 || T > kMaxEdgeTypes || Vt < 1 || Vt > kMaxVertexTypes) return WHOLEMEMORY_INVALID_INPUT;
  if (num_hops < 1 || num_hops >= kMaxHops - 1) return WHOLEMEMORY_INVALID_INPUT;
  auto* sd = wholememory_tensor_get_tensor_description(seeds);
  auto* ld = wholememory_tensor_get_tensor_description(label_offsets);
  if (sd->dim != 1 || ld->dim != 1 || ld->dtype != WHOLEMEMORY_DT_INT64 || ld->sizes[0] < 1) return WHOLEMEMORY_INVALID_INPUT;
  if (sd->dtype != WHOLEMEMORY_DT_INT && sd->dtype != WHOLEMEMORY_DT_INT64) return WHOLEMEMORY_INVALID_INPUT;
  bool weighted = csr_weight != nullptr && csr_weight[0] != nullptr;
  for (int t = 0; t < T; t++) {
    if (!csr_row_ptr[t] || !csr_col[t]) return WHOLEMEMORY_INVALID_INPUT;
    auto* rd = wholememory_tensor_get_tensor_description(csr_row_ptr[t]);
    auto* cd = wholememory_tensor_get_tensor_description(csr_col[t]);
    auto* r0 = wholememory_tensor_get_tensor_description(csr_row_ptr[0]);
    auto* c0 = wholememory_tensor_get_tensor_description(csr_col[0]);
    if (rd->dim != 1 || cd->dim != 1 || rd->dtype != WHOLEMEMORY_DT_INT64 || rd->sizes[0] < 1) return WHOLEMEMORY_INVALID_INPUT;
    if (cd->dtype != WHOLEMEMORY_DT_INT && cd->dtype != WHOLEMEMORY_DT_INT64) return WHOLEMEMORY_INVALID_INPUT;
    if (rd->sizes[0] != r0->sizes[0] || cd->dtype != c0->dtype) return WHOLEMEMORY_INVALID_INPUT;  // one id space, one id type
    if (weighted != (csr_weight != nullptr && csr_weight[t] != nullptr)) return WHOLEMEMORY_INVALID_INPUT;
    if (weighted) {
      auto* wd = wholememory_tensor_get_tensor_description(csr_weight[t]);
      auto* w0 = wholememory_tensor_get_tensor_description(csr_weight[0]);
      if (wd->dim != 1 || wd->sizes[0] != cd->sizes[0] || (wd->dtype != WHOLEMEMORY_DT_FLOAT && wd->dtype != WHOLEMEMORY_DT_DOUBLE) || wd->dtype != w0->dtype) return WHOLEMEMORY_INVALID_INPUT;
    }
    if (csr_edge_id && csr_edge_id[t]) {
      auto* ed = wholememory_tensor_get_tensor_description(csr_edge_id[t]);
      if (ed->dim != 1 || ed->sizes[0] != cd->sizes[0] || ed->dtype != WHOLEMEMORY_DT_INT64) return WHOLEMEMORY_INVALID_INPUT;
    }
    if (temporal) {
      if (!csr_edge_time[t]) return WHOLEMEMORY_INVALID_INPUT;
      auto* et = wholememory_tensor_get_tensor_description(csr_edge_time[t]);
      if (et->dim != 1 || et->sizes[0] != cd->sizes[0] || et->dtype != WHOLEMEMORY_DT_INT64) return WHOLEMEMORY_INVALID_INPUT;
    }
  }
  for (int i = 0; i < num_hops * T; i++)
    if (fanout[i] > 1024) return WHOLEMEMORY_NOT_IMPLEMENTED;
  if (hetero) {
    if (!vertex_type_offsets || vertex_type_offsets[0] != 0) return WHOLEMEMORY_INVALID_INPUT;
    for (int v = 0; v < Vt; v++)
      if (vertex_type_offsets[v + 1] < vertex_type_offsets[v]) return WHOLEMEMORY_INVALID_INPUT;
  }
  return guarded(what, [&] {
    WGB_EXPECTS(sd->sizes[0] < (1LL << 31) - kScanTile, "too many seeds for one call");
    WGB_CHECK_INPUT(sd->sizes[0] == 0 || wholememory_tensor_get_tensor_description(csr_row_ptr[0])->sizes[0] > 1, "seeds were given for a graph without vertices");
    WGB_EXPECTS(ld->sizes[0] - 1 < (1LL << 23), "too many labels for one call");
    MhCall c;
    c.sp        = sampler;
    c.T         = T;
    c.Vt        = Vt;
    c.hetero    = hetero;
    c.weighted  = weighted;
    c.wgt_dtype = weighted ? wholememory_tensor_get_tensor_description(csr_weight[0])->dtype : WHOLEMEMORY_DT_FLOAT;
    c.col_dtype = wholememory_tensor_get_tensor_description(csr_col[0])->dtype;
    c.chunked   = false;
    for (int t = 0; t < T; t++) {
      MhTypeCsr& g = c.csr[t];
      memset(&g.wgt, 0, sizeof(g.wgt));
      memset(&g.eid, 0, sizeof(g.eid));
      g.row_ptr     = make_chunk_ref(csr_row_ptr[t]);
      g.row_ptr_off = (unsigned long long)wholememory_tensor_get_tensor_description(csr_row_ptr[t])->storage_offset;
      g.col         = make_chunk_ref(csr_col[t]);
      g.col_off     = (unsigned long long)wholememory_tensor_get_tensor_description(csr_col[t])->storage_offset;
      c.chunked     = c.chunked || g.row_ptr.world > 1 || g.col.world > 1;
      if (weighted) {
        g.wgt     = make_chunk_ref(csr_weight[t]);
        g.wgt_off = (unsigned long long)wholememory_tensor_get_tensor_description(csr_weight[t])->storage_offset;
        c.chunked = c.chunked || g.wgt.world > 1;
      }
      memset(&g.etime, 0, sizeof(g.etime));
      if (temporal) {
        g.etime     = make_chunk_ref(csr_edge_time[t]);
        g.etime_off = (unsigned long long)wholememory_tensor_get_tensor_description(csr_edge_time[t])->storage_offset;
        c.chunked   = c.chunked || g.etime.world > 1;
      }
      g.has_eid = csr_edge_id && csr_edge_id[t];
      if (g.has_eid) {
        g.eid     = make_chunk_ref(csr_edge_id[t]);
        g.eid_off = (unsigned long long)wholememory_tensor_get_tensor_description(csr_edge_id[t])->storage_offset;
        c.chunked = c.chunked || g.eid.world > 1;
      }
    }
    c.seeds         = wholememory_tensor_get_data_pointer(seeds);
    c.seed_dtype    = sd->dtype;
    c.label_offsets = static_cast<const long long*>(wholememory_tensor_get_data_pointer(label_offsets));
    c.S             = (int)sd->sizes[0];
    c.B             = (int)ld->sizes[0] - 1;
    c.L             = num_hops;
    for (int i = 0; i < num_hops * T; i++)
      c.fanout[i] = fanout[i];
    c.V = (unsigned long long)(wholememory_tensor_get_tensor_description(csr_row_ptr[0])->sizes[0] - 1);
    for (int v = 0; v <= Vt; v++)
      c.vto[v] = hetero ? vertex_type_offsets[v] : (v == 0 ? 0 : (long long)c.V);
    if (hetero) WGB_EXPECTS((unsigned long long)c.vto[Vt] <= c.V, "vertex_type_offsets exceed the vertex id space of the CSRs");
    WGB_EXPECTS((long double)c.V * (long double)std::max(c.B, 1) < 72057594037927936.0L, "labels x vertices must stay below 2^56");
    c.random_state = random_state;
    c.flags        = flags;
    c.stream       = as_stream(stream);
    c.temporal     = temporal;
    c.time_cmp     = time_cmp;
    c.seed_times   = temporal ? static_cast<const long long*>(wholememory_tensor_get_data_pointer(seed_times)) : nullptr;
    if (c.col_dtype == WHOLEMEMORY_DT_INT) {
      if (c.chunked) multihop_begin<int, true>(c);
      else multihop_begin<int, false>(c);
    } else {
      if (c.chunked) multihop_begin<long long, true>(c);
      else multihop_begin<long long, false>(c);
    }
  });
}

}  // namespace wgb

wholememory_error_code_t wholegraph_multihop_neighbor_sample_begin(
  wholegraph_multihop_sampler_t sampler, wholememory_tensor_t csr_row_ptr, wholememory_tensor_t csr_col,
  wholememory_tensor_t csr_weight, wholememory_tensor_t csr_edge_id, wholememory_tensor_t seeds,
  wholememory_tensor_t label_offsets, const int* fanout, int num_hops, unsigned long long random_state, int flags,
  void* stream)
{
  return wgb::multihop_begin_entry("wholegraph_multihop_neighbor_sample_begin", sampler, 1, &csr_row_ptr, &csr_col, &csr_weight,
                                   &csr_edge_id, nullptr, 1, false, seeds, label_offsets, fanout, num_hops, random_state, flags, stream);
}

wholememory_error_code_t wholegraph_hetero_multihop_neighbor_sample_begin(
  wholegraph_multihop_sampler_t sampler, int num_edge_types, const wholememory_tensor_t* csr_row_ptr,
  const wholememory_tensor_t* csr_col, const wholememory_tensor_t* csr_weight, const wholememory_tensor_t* csr_edge_id,
  const long long* vertex_type_offsets, int num_vertex_types, wholememory_tensor_t seeds, wholememory_tensor_t label_offsets,
  const int* fanout, int num_hops, unsigned long long random_state, int flags, void* stream)
{
  if (flags & WHOLEGRAPH_MULTIHOP_CSR) return WHOLEMEMORY_INVALID_INPUT;
  return wgb::multihop_begin_entry("wholegraph_hetero_multihop_neighbor_sample_begin", sampler, num_edge_types, csr_row_ptr, csr_col,
                                   csr_weight, csr_edge_id, vertex_type_offsets, num_vertex_types, true, seeds, label_offsets, fanout,
                                   num_hops, random_state, flags, stream);
}

wholememory_error_code_t wholegraph_temporal_multihop_neighbor_sample_begin(
  wholegraph_multihop_sampler_t sampler, int num_edge_types, const wholememory_tensor_t* csr_row_ptr,
  const wholememory_tensor_t* csr_col, const wholememory_tensor_t* csr_weight, const wholememory_tensor_t* csr_edge_time,
  const wholememory_tensor_t* csr_edge_id, const long long* vertex_type_offsets, int num_vertex_types, int heterogeneous,
  wholememory_tensor_t seeds, wholememory_tensor_t seed_times, wholememory_tensor_t label_offsets, const int* fanout, int num_hops,
  unsigned long long random_state, int time_comparison, int flags, void* stream)
{
  if (!csr_edge_time || !seed_times) return WHOLEMEMORY_INVALID_INPUT;
  if (heterogeneous) {
    if (flags & WHOLEGRAPH_MULTIHOP_CSR) return WHOLEMEMORY_INVALID_INPUT;
  } else if (num_edge_types != 1) {
    return WHOLEMEMORY_INVALID_INPUT;
  }
  return wgb::multihop_begin_entry("wholegraph_temporal_multihop_neighbor_sample_begin", sampler, num_edge_types, csr_row_ptr, csr_col,
                                   csr_weight, csr_edge_id, heterogeneous ? vertex_type_offsets : nullptr, heterogeneous ? num_vertex_types : 1,
                                   heterogeneous != 0, seeds, label_offsets, fanout, num_hops, random_state, flags, stream, csr_edge_time,
                                   seed_times, time_comparison);
}

wholememory_error_code_t wholegraph_hetero_multihop_neighbor_sample_finish(
  wholegraph_multihop_sampler_t sampler, void* out_majors_ctx, void* out_minors_ctx, void* out_edge_id_ctx,
  void* out_edge_type_ctx, void* out_label_type_hop_offsets_ctx, void* out_renumber_map_ctx,
  void* out_renumber_map_offsets_ctx, void* out_edge_renumber_map_ctx, void* out_edge_renumber_map_offsets_ctx,
  void* out_label_type_step_base_ctx, wholememory_env_func_t* p_env_fns, void* stream)
{
  using namespace wgb;
  if (!sampler || !p_env_fns) return WHOLEMEMORY_INVALID_INPUT;
  if (!out_majors_ctx || !out_minors_ctx || !out_edge_id_ctx || !out_edge_type_ctx || !out_label_type_hop_offsets_ctx ||
      !out_renumber_map_ctx || !out_renumber_map_offsets_ctx || !out_edge_renumber_map_ctx || !out_edge_renumber_map_offsets_ctx)
    return WHOLEMEMORY_INVALID_INPUT;
  return guarded("wholegraph_hetero_multihop_neighbor_sample_finish", [&] {
    WGB_EXPECTS(sampler->pending.active && sampler->pending.hetero, "no heterogeneous call in flight on this sampler object");
    MhOutCtx o;
    o.majors                    = out_majors_ctx;
    o.minors                    = out_minors_ctx;
    o.edge_id                   = out_edge_id_ctx;
    o.edge_type                 = out_edge_type_ctx;
    o.lho                       = out_label_type_hop_offsets_ctx;
    o.map                       = out_renumber_map_ctx;
    o.rmo                       = out_renumber_map_offsets_ctx;
    o.edge_renumber_map         = out_edge_renumber_map_ctx;
    o.edge_renumber_map_offsets = out_edge_renumber_map_offsets_ctx;
    o.major_offsets             = nullptr;
    o.step_counts               = out_label_type_step_base_ctx;
    o.env                       = p_env_fns;
    o.stream                    = as_stream(stream);
    multihop_finish(sampler, o);
  });
}

wholememory_error_code_t wholegraph_multihop_neighbor_sample_finish(
  wholegraph_multihop_sampler_t sampler, void* out_majors_ctx, void* out_minors_ctx, void* out_edge_id_ctx,
  void* out_label_hop_offsets_ctx, void* out_renumber_map_ctx, void* out_renumber_map_offsets_ctx,
  void* out_major_offsets_ctx, void* out_label_step_base_ctx, wholememory_env_func_t* p_env_fns, void* stream)
{
  using namespace wgb;
  if (!sampler || !p_env_fns) return WHOLEMEMORY_INVALID_INPUT;
  if (!out_minors_ctx || !out_edge_id_ctx || !out_label_hop_offsets_ctx || !out_renumber_map_ctx || !out_renumber_map_offsets_ctx) return WHOLEMEMORY_INVALID_INPUT;
  return guarded("wholegraph_multihop_neighbor_sample_finish", [&] {
    WGB_EXPECTS(!(sampler->pending.active && sampler->pending.hetero), "a heterogeneous call must be finished with wholegraph_hetero_multihop_neighbor_sample_finish");
    MhOutCtx o;
    o.majors        = out_majors_ctx;
    o.minors        = out_minors_ctx;
    o.edge_id       = out_edge_id_ctx;
    o.lho           = out_label_hop_offsets_ctx;
    o.map           = out_renumber_map_ctx;
    o.rmo           = out_renumber_map_offsets_ctx;
    o.major_offsets = out_major_offsets_ctx;
    o.step_counts   = out_label_step_base_ctx;
    o.env           = p_env_fns;
    o.stream        = as_stream(stream);
    multihop_finish(sampler, o);
  });
}

wholememory_error_code_t wholegraph_multihop_seed_local_ids(wholegraph_multihop_sampler_t sampler, void* out_seed_local_id_ctx,
                                                            wholememory_env_func_t* p_env_fns, void* stream)
{
  using namespace wgb;
  if (!sampler || !out_seed_local_id_ctx || !p_env_fns) return WHOLEMEMORY_INVALID_INPUT;
  return guarded("wholegraph_multihop_seed_local_ids", [&] {
    auto& pd = sampler->pending;
    WGB_EXPECTS(pd.finished && !pd.active, "seed ids are available between _finish of a call and the next _begin on the same sampler object");
    cudaStream_t st = as_stream(stream);
    int* out        = static_cast<int*>(output_alloc(p_env_fns, out_seed_local_id_ctx, pd.S, WHOLEMEMORY_DT_INT));
#ifndef WGB_HOST_EMULATION
    if (pd.fused) {
      if (pd.S > 0) WGB_CUDA_TRY(cudaMemcpyAsync(out, pd.fz.seed_local, sizeof(int) * (size_t)pd.S, cudaMemcpyDeviceToDevice, st));
      return;
    }
#endif
    if (pd.S > 0) {
      mh_seed_local_kernel<<<grid_over(pd.S, num_sms()), 256, 0, st>>>(pd.seed_ref, pd.seed_rank, pd.slabel, pd.meta.fr_off[0],
                                                                       pd.hetero ? sampler->desc_host.typed_local[0] : nullptr, pd.S, out);
      WGB_CHECK_LAUNCH();
    }
  });
}

wholememory_error_code_t wholegraph_multihop_neighbor_sample(
  wholegraph_multihop_sampler_t sampler, wholememory_tensor_t csr_row_ptr, wholememory_tensor_t csr_col,
  wholememory_tensor_t csr_weight, wholememory_tensor_t csr_edge_id, wholememory_tensor_t seeds,
  wholememory_tensor_t label_offsets, const int* fanout, int num_hops, unsigned long long random_state, int flags,
  void* out_majors_ctx, void* out_minors_ctx, void* out_edge_id_ctx, void* out_label_hop_offsets_ctx,
  void* out_renumber_map_ctx, void* out_renumber_map_offsets_ctx, void* out_major_offsets_ctx,
  void* out_label_step_base_ctx,
  wholememory_env_func_t* p_env_fns, void* stream)
{
  if (!p_env_fns) return WHOLEMEMORY_INVALID_INPUT;
  if (!out_minors_ctx || !out_edge_id_ctx || !out_label_hop_offsets_ctx || !out_renumber_map_ctx || !out_renumber_map_offsets_ctx) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_error_code_t e = wholegraph_multihop_neighbor_sample_begin(sampler, csr_row_ptr, csr_col, csr_weight, csr_edge_id, seeds,
                                                                         label_offsets, fanout, num_hops, random_state, flags, stream);
  if (e != WHOLEMEMORY_SUCCESS) return e;
  return wholegraph_multihop_neighbor_sample_finish(sampler, out_majors_ctx, out_minors_ctx, out_edge_id_ctx, out_label_hop_offsets_ctx,
                                                    out_renumber_map_ctx, out_renumber_map_offsets_ctx, out_major_offsets_ctx,
                                                    out_label_step_base_ctx, p_env_fns, stream);
}

}  // extern "C"
