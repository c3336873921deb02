// G1 / G2: row gather and scatter over a WholeMemory table (sm_100a).
//
// Replaces the reference's gather_func_kernel / gather_func_sub_warp_kernel / scatter_func_kernel
// (cpp/src/wholememory_ops/functions/gather_scatter_func.cuh:243-365, 509-587) and, for tables
// striped over the GPUs of one NVSwitch box, the whole NCCL bucket / all-to-all pipeline of
// cpp/src/wholememory_ops/gather_op_impl_nccl.cu:23-171: remote rows are read (or written) with
// plain ld.global / st.global on peer-mapped pointers from inside this kernel.
//
// Design: the dense side (gather output / scatter input) is viewed as a flat
// sequence of 16-byte vectors; thread t moves vectors t, t+T, t+2T, t+3T of a 4*T window, all four
// loads issued before the first store, so every warp keeps 4 x 512 B in flight and both sides are
// fully coalesced for any row width (one warp-wide access = 512 contiguous bytes of one or more
// rows).  No shared-memory bounce, no per-access divide: the row of a vector comes from a
// multiply-shift, the owning rank from a compare chain over <= 7 chunk boundaries that live in
// the kernel parameter block.

#include "wm_common.cuh"

namespace wgb {

template <int BYTES>
struct vec_of;
template <>
struct vec_of<1> { using type = unsigned char; };
template <>
struct vec_of<2> { using type = unsigned short; };
template <>
struct vec_of<4> { using type = unsigned int; };
template <>
struct vec_of<8> { using type = uint2; };
template <>
struct vec_of<16> { using type = uint4; };

// streaming loads/stores: every byte is touched once per call, keep L1 for the index vector.
template <typename V>
__device__ __forceinline__ V ld_stream(const V* p)
{
  return *p;
}
#ifndef WGB_HOST_EMULATION  // tests/emu compiles this file with g++ (logic check on the CPU): plain loads there
template <>
__device__ __forceinline__ uint4 ld_stream<uint4>(const uint4* p)
{
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
template <>
__device__ __forceinline__ uint2 ld_stream<uint2>(const uint2* p)
{
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
#endif

struct RowDiv {  // q = n / d for n < 2^31 via multiply-shift
  unsigned int mul;
  unsigned int shift;
  unsigned int d;
  __device__ __forceinline__ unsigned int div(unsigned int n) const
  {
    return (unsigned int)(((unsigned long long)n * mul) >> shift);
  }
};

static RowDiv make_row_div(unsigned int d)
{
  RowDiv r;
  r.d   = d;
  int l = 0;
  while ((1u << l) < d)
    l++;
  r.shift = 31 + l;
  r.mul   = (unsigned int)(((1ULL << r.shift) + d - 1) / d);
  return r;
}

constexpr int kBlock  = 256;
constexpr int kUnroll = 4;

// ---- same-dtype path: pure byte movement ---------------------------------------------------------
// SCATTER=false: dense[i, :] = table[idx[i], :]      SCATTER=true: table[idx[i], :] = dense[i, :]
// HOT: rows replicated on this GPU (wgb_hot_rows) are read from the replica instead of the owning rank -- on a power-law
// graph the hottest 10 % of the rows are 84 % of what a call group gathers (profiles/hot_rows_probe.py), which takes
// that share of the traffic off NVLink.
template <typename IdxT, int VEC, bool CHUNKED, bool SCATTER, bool HOT = false>
__global__ void __launch_bounds__(kBlock) rows_copy_kernel(ChunkRef table,
                                                           unsigned long long table_off_bytes,
                                                           unsigned long long row_stride_bytes,
                                                           const IdxT* __restrict__ idx,
                                                           unsigned int total_vecs,
                                                           RowDiv vpr,
                                                           char* __restrict__ dense,
                                                           unsigned long long dense_stride_bytes,
                                                           wgb_hot_rows hot = wgb_hot_rows())
{
  using V                   = typename vec_of<VEC>::type;
  const unsigned int window = kBlock * kUnroll;
  for (unsigned int base = blockIdx.x * window; base < total_vecs; base += gridDim.x * window) {
    V val[kUnroll];
    char* tptr[kUnroll];
    char* dptr[kUnroll];
    bool ok[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      unsigned int j   = base + u * kBlock + threadIdx.x;
      ok[u]            = j < total_vecs;
      unsigned int row = vpr.div(ok[u] ? j : 0);
      unsigned int c   = (ok[u] ? j : 0) - row * vpr.d;
      long long r      = ok[u] ? (long long)idx[row] : -1;
      ok[u]            = ok[u] && r >= 0;
      unsigned long long off = table_off_bytes + (unsigned long long)(r < 0 ? 0 : r) * row_stride_bytes;
      int hs                 = -1;
      if (HOT && ok[u]) hs = __ldg(hot.slot + r);
      tptr[u] = (HOT && hs >= 0) ? const_cast<char*>(hot.rows) + (unsigned long long)hs * hot.stride_bytes + (unsigned long long)c * VEC
                                 : table.at<CHUNKED>(off) + (unsigned long long)c * VEC;
      dptr[u] = dense + (unsigned long long)row * dense_stride_bytes + (unsigned long long)c * VEC;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      if (ok[u]) val[u] = ld_stream(reinterpret_cast<const V*>(SCATTER ? dptr[u] : tptr[u]));
    }
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      if (ok[u]) *reinterpret_cast<V*>(SCATTER ? tptr[u] : dptr[u]) = val[u];
    }
  }
}

#ifndef WGB_HOST_EMULATION  // the copy-engine path is PTX (cp.async.bulk, mbarrier): not part of the CPU logic check of tests/emu
// ---- same-dtype gather as bulk-async copies (TMA) ---------------------------------------------------------
// The register path above needs every thread slot and every register of an SM to keep ~64 KB of loads in flight
// (256 threads x 8 CTAs x 4 x 16 B); while it runs, nothing else fits on the SM, so the sampler of the next call group
// cannot execute underneath the gather although the two bind different resources (measured: step = sum, not max).
// Here the bytes in flight live in SHARED memory, moved by the copy engine: a warp owns a ring of `stages` tiles of 32
// rows; lane j issues ONE cp.async.bulk (global -> shared, completion counted on the tile's mbarrier) for row j of a
// tile, and a whole tile leaves with ONE bulk store (shared -> global) because gathered rows are consecutive in the
// output.  The data never passes through registers.  A CTA is 2-8 warps with a handful of registers each: it co-resides
// with the fused sampler CTA (multihop_fused.cuh: 30 warps, no shared memory to speak of) on every SM.
// Rows whose index is negative are skipped (output untouched): tiles that contain one, or the ragged last tile, leave
// row by row.  CHUNKED: the source of a row is the owning GPU's memory (peer-mapped, the copy crosses NVLink).
constexpr int kBulkRows     = 32;
constexpr int kBulkMaxWarps = 4;  // 128 threads x <= 64 registers: fits beside a 28-warp sampler CTA in the register file

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

template <typename IdxT, bool CHUNKED, bool HOT>
__global__ void __launch_bounds__(kBulkMaxWarps * 32) rows_bulk_gather_kernel(ChunkRef table, unsigned long long table_off_bytes,
                                                               unsigned long long row_stride_bytes, const IdxT* __restrict__ idx,
                                                               long long n_rows, unsigned int row_bytes, char* __restrict__ dense,
                                                               unsigned long long dense_stride_bytes, int stages, wgb_hot_rows hot,
                                                               unsigned int* __restrict__ ticket)
{
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const unsigned int tile_bytes = kBulkRows * row_bytes;
  unsigned char* ring           = bulk_smem + (size_t)wib * stages * tile_bytes;
  unsigned long long* bars      = reinterpret_cast<unsigned long long*>(bulk_smem + (size_t)warps * stages * tile_bytes) + wib * stages;
  if (lane == 0) {
    for (int s = 0; s < stages; s++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  const long long n_tiles = (n_rows + kBulkRows - 1) / kBulkRows;
  const long long gw = (long long)blockIdx.x * warps + wib, nw = (long long)gridDim.x * warps;
  // gathered rows are touched once: evict-first in L2, so that the stream does not push out the sampler's tables
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));

  // Where a row comes from is a chain of dependent reads -- idx[i], then (HOT) the replica slot of that row, a random 4-byte
  // read of a table-sized array -- before the copy can be issued.  Resolved at issue time, that chain (1.5-2 us) throttled every
  // warp to one tile per chain: 1.06 ms against 0.76 for the same rows without the slot lookup (two GPUs, profiles/r2w_*).
  // So the chain is software-pipelined over the ring's iterations: the tile taken by ticket in iteration k has its index read
  // in k, its slot read in k + 1, and its copies issued in k + 2.
  auto read_index = [&](long long tile) -> long long {
    const long long i = tile * kBulkRows + lane;
    return (tile < n_tiles && i < n_rows) ? (long long)idx[i] : -1;
  };
  auto read_slot = [&](long long r) -> int { return (HOT && r >= 0) ? __ldg(hot.slot + r) : -1; };
  auto source_of = [&](long long r, int hs) -> const char* {
    if (HOT && hs >= 0) return hot.rows + (unsigned long long)hs * hot.stride_bytes;
    return table.at<CHUNKED>(table_off_bytes + (unsigned long long)r * row_stride_bytes);
  };
  auto issue_loads = [&](long long r, int hs, int s) {
    const bool ok           = r >= 0;
    const unsigned int mask = __ballot_sync(0xffffffffu, ok);
    const unsigned int bar  = smem_u32(&bars[s]);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(__popc(mask) * row_bytes) : "memory");
    __syncwarp();
    if (ok)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(ring + (size_t)s * tile_bytes + (size_t)lane * row_bytes)),
                   "l"(source_of(r, hs)), "r"(row_bytes), "r"(bar), "l"(pol)
                   : "memory");
  };

  // Tiles are handed out by ticket, one per request: the grid may hold more CTAs than are resident at once (two per SM
  // when the gather has the SM to itself, one when a sampler CTA lives there) and a CTA that starts late simply finds
  // fewer tiles left.
  auto next_tile = [&]() -> long long {
    unsigned int t = 0;
    if (lane == 0) t = atomicAdd(ticket, 1u);
    return (long long)__shfl_sync(0xffffffffu, t, 0);
  };
  // the ring remembers which tile sits in which stage (registers: stages <= 8, indexed by unrolled compare chains)
  long long in_stage[8];
#pragma unroll
  for (int k = 0; k < 8; k++)
    in_stage[k] = n_tiles;
  auto set_stage = [&](int s, long long t) {
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (k == s) in_stage[k] = t;
  };
  auto get_stage = [&](int s) -> long long {
    long long t = n_tiles;
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (k == s) t = in_stage[k];
    return t;
  };
  // prologue: stages - 1 tiles in flight (resolved on the spot), then the two look-ahead slots of the pipeline
  for (int k = 0; k < stages - 1; k++) {
    const long long t = next_tile();
    set_stage(k, t);
    if (t < n_tiles) {
      const long long r = read_index(t);
      issue_loads(r, read_slot(r), k);
    }
  }
  long long t1 = next_tile();            // its slot is read at the top of the first iteration
  long long r1 = read_index(t1);
  long long t2 = next_tile();            // its index is in flight
  long long r2 = read_index(t2);
  int s = 0;
  unsigned int parity = 0;
  while (true) {
    const long long tile = get_stage(s);
    if (tile >= n_tiles) break;  // tickets are monotonic: every later stage holds nothing either
    const int hs1 = read_slot(r1);  // in flight while this iteration waits for its tile
    // the tile's rows have landed
    const unsigned int bar = smem_u32(&bars[s]);
    asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
    const long long i       = tile * kBulkRows + lane;
    const bool ok           = i < n_rows && (long long)idx[i] >= 0;
    const unsigned int mask = __ballot_sync(0xffffffffu, ok);
    const unsigned char* st = ring + (size_t)s * tile_bytes;
    if (mask == 0xffffffffu && dense_stride_bytes == row_bytes) {
      if (lane == 0)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dense + (unsigned long long)tile * kBulkRows * dense_stride_bytes),
                     "r"(smem_u32(st)), "r"(tile_bytes), "l"(pol)
                     : "memory");
    } else if (ok) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dense + (unsigned long long)i * dense_stride_bytes),
                   "r"(smem_u32(st + (size_t)lane * row_bytes)), "r"(row_bytes), "l"(pol)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    // refill the stage the PREVIOUS tile left from: its store must have finished reading shared memory
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    const int sp = s == 0 ? stages - 1 : s - 1;
    set_stage(sp, t1);
    if (t1 < n_tiles) issue_loads(r1, hs1, sp);
    // rotate the look-ahead pipeline and take the next ticket
    t1 = t2;
    r1 = r2;
    t2 = next_tile();
    r2 = read_index(t2);
    if (++s == stages) {
      s = 0;
      parity ^= 1u;
    }
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

#endif  // WGB_HOST_EMULATION

// ---- converting path -----------------------------------------------------------------------------
// element conversion as the reference does it (gather_scatter_func.cuh:150-197): half / bf16 go
// through float, everything else is a static_cast.
template <typename T>
struct via { using type = T; };
template <>
struct via<__half> { using type = float; };
template <>
struct via<__nv_bfloat16> { using type = float; };

template <typename From, typename To>
__device__ __forceinline__ To convert_elt(From x)
{
  typename via<From>::type a = static_cast<typename via<From>::type>(x);
  typename via<To>::type b   = static_cast<typename via<To>::type>(a);
  return static_cast<To>(b);
}

template <typename T, int N>
struct alignas(sizeof(T) * N) elt_pack {
  T v[N];
};

template <typename TableT, typename DenseT, typename IdxT, int ALIGN, bool SCATTER>
__global__ void __launch_bounds__(kBlock) rows_convert_kernel(ChunkRef table,
                                                              unsigned long long table_off_elts,
                                                              unsigned long long row_stride_elts,
                                                              const IdxT* __restrict__ idx,
                                                              unsigned int total_vecs,
                                                              RowDiv vpr,
                                                              DenseT* __restrict__ dense,
                                                              unsigned long long dense_stride_elts)
{
  using TP = elt_pack<TableT, ALIGN>;
  using DP = elt_pack<DenseT, ALIGN>;
  for (unsigned int j = blockIdx.x * kBlock + threadIdx.x; j < total_vecs; j += gridDim.x * kBlock) {
    unsigned int row = vpr.div(j);
    unsigned int c   = j - row * vpr.d;
    long long r      = (long long)idx[row];
    if (r < 0) continue;
    unsigned long long off = (table_off_elts + (unsigned long long)r * row_stride_elts) * sizeof(TableT);
    TP* tp = reinterpret_cast<TP*>(table.at<true>(off)) + c;
    DP* dp = reinterpret_cast<DP*>(dense + (unsigned long long)row * dense_stride_elts) + c;
    if (!SCATTER) {
      TP in = *tp;
      DP out;
#pragma unroll
      for (int k = 0; k < ALIGN; k++)
        out.v[k] = convert_elt<TableT, DenseT>(in.v[k]);
      *dp = out;
    } else {
      DP in = *dp;
      TP out;
#pragma unroll
      for (int k = 0; k < ALIGN; k++)
        out.v[k] = convert_elt<DenseT, TableT>(in.v[k]);
      *tp = out;
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------
struct RowsOpArgs {
  ChunkRef table;
  wholememory_matrix_description_t table_desc;
  const void* idx;
  wholememory_dtype_t idx_dtype;
  int64_t n;
  void* dense;  // already offset-free base pointer
  wholememory_matrix_description_t dense_desc;
  wgb_hot_rows hot;  // slot == nullptr: none
  bool scatter;
  int sms;
  cudaStream_t stream;
};

static int largest_pow2_alignment(std::initializer_list<unsigned long long> values, int cap)
{
  int a = cap;
  for (; a > 1; a /= 2) {
    bool ok = true;
    for (auto v : values)
      if (v % (unsigned long long)a != 0) ok = false;
    if (ok) break;
  }
  return a;
}

static int grid_for(unsigned long long work_items, int per_block, int sms_limit)
{
  unsigned long long blocks = (work_items + per_block - 1) / per_block;
  int sms                   = num_sms();
  unsigned long long cap    = (unsigned long long)(sms_limit > 0 ? std::min(sms_limit, sms) : sms) * 8ULL;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <typename IdxT, int VEC, bool SCATTER>
static void launch_copy(const RowsOpArgs& a, int64_t row0, int64_t rows, unsigned int vpr)
{
  size_t elt                    = dtype_size(a.table_desc.dtype);
  unsigned long long total      = (unsigned long long)rows * vpr;
  RowDiv rd                     = make_row_div(vpr);
  const IdxT* idx               = static_cast<const IdxT*>(a.idx) + row0;
  unsigned long long dstride    = (unsigned long long)a.dense_desc.stride * elt;
  char* dense                   = static_cast<char*>(a.dense) + (unsigned long long)a.dense_desc.storage_offset * elt + (unsigned long long)row0 * dstride;
  unsigned long long toff       = (unsigned long long)a.table_desc.storage_offset * elt;
  unsigned long long tstride    = (unsigned long long)a.table_desc.stride * elt;
  int grid                      = grid_for(total, kBlock * kUnroll, a.sms);
  if (!SCATTER && a.hot.slot != nullptr && a.table.world > 1 && a.hot.stride_bytes % VEC == 0 &&
      reinterpret_cast<unsigned long long>(a.hot.rows) % VEC == 0)
    rows_copy_kernel<IdxT, VEC, true, false, true><<<grid, kBlock, 0, a.stream>>>(a.table, toff, tstride, idx, (unsigned int)total, rd, dense, dstride, a.hot);
  else if (a.table.world > 1)
    rows_copy_kernel<IdxT, VEC, true, SCATTER><<<grid, kBlock, 0, a.stream>>>(a.table, toff, tstride, idx, (unsigned int)total, rd, dense, dstride);
  else
    rows_copy_kernel<IdxT, VEC, false, SCATTER><<<grid, kBlock, 0, a.stream>>>(a.table, toff, tstride, idx, (unsigned int)total, rd, dense, dstride);
  WGB_CHECK_LAUNCH();
}

#ifndef WGB_HOST_EMULATION
// Which same-dtype gathers go through the copy engine: all that qualify (16-byte aligned rows of at most 2 KB, >= 4096 rows),
// unless WGB_GATHER_BULK=0 (read per call) sends them down the register path.
// Measured on C4, 148-label call groups (4.85 M rows of 512 B per step):
//   one GPU, local table     gather alone 0.758 ms (1.0 of the HBM peak) against 0.853 for the register kernel; pipelined step
//                            1.327 against 1.334 ms, end to end 1.411 against 1.443                   (profiles/r2s_bench_c4_bulk*.json)
//   two GPUs, 10 % replica   gather alone 0.986 against 1.265 ms; step 1.548 against 1.666, end to end 1.642 against 1.816
//   two GPUs, no replica     0.83 against 0.78 of the NVLink bound                                   (profiles/r2x_bench_n2_bulk*.json)
// (Round-2 history: with 64-label call groups and the index -> slot chain resolved at issue time the register kernel was the
// better neighbour of the sampler on two GPUs, 0.811 against 0.900 ms per step, profiles/r2l_*.json; the software-pipelined
// chain and one-CTA-per-label sampler calls reversed that.)
static bool bulk_enabled(int /*world*/)
{
  const char* e = getenv("WGB_GATHER_BULK");
  return !(e && *e && atoi(e) == 0);
}

// Shared memory of the tile rings per CTA: 96 KB (2 warps x 3 tiles of 16 KB at 512-byte rows).  An SM that runs nothing
// else takes two such CTAs (128 KB of row loads in flight); an SM that runs a sampler CTA takes one -- the L1 / shared
// split of an SM cannot change while CTAs are resident, so the gather CTA only fits because the sampler kernel asks for a
// 132 KB carve-out (WGB_MH_CARVEOUT in multihop_fused.cuh).  WGB_GATHER_BULK_KB overrides the ring size.
static size_t bulk_smem_budget()
{
  if (const char* e = getenv("WGB_GATHER_BULK_KB"))
    if (atoi(e) >= 16 && atoi(e) <= 200) return (size_t)atoi(e) * 1024;
  return 96 * 1024;
}
constexpr long long kBulkMinRows = 4096;

// tile tickets of the launches in flight: a small ring of device counters, one per launch, zeroed on the launch's stream
static unsigned int* bulk_ticket_slot(cudaStream_t st)
{
  constexpr int kSlots = 256;
  static thread_local unsigned int* slots[kMaxWorld * 2] = {};
  static thread_local unsigned int next[kMaxWorld * 2]  = {};
  int dev = 0;
  WGB_CUDA_TRY(cudaGetDevice(&dev));
  WGB_EXPECTS(dev < kMaxWorld * 2, "device ordinal out of range");
  if (!slots[dev]) WGB_CUDA_TRY(cudaMalloc(&slots[dev], kSlots * sizeof(unsigned int)));
  unsigned int* p = slots[dev] + (next[dev]++ % kSlots);
  WGB_CUDA_TRY(cudaMemsetAsync(p, 0, sizeof(unsigned int), st));
  return p;
}

// same-dtype gather through the copy engine when shapes allow it (everything 16-byte aligned, a tile ring fits)
template <typename IdxT>
static bool try_bulk_gather(const RowsOpArgs& a)
{
  if (!bulk_enabled(a.table.world) || a.n < kBulkMinRows) return false;
  const size_t elt                   = dtype_size(a.table_desc.dtype);
  const unsigned long long row_bytes = (unsigned long long)a.table_desc.sizes[1] * elt;
  const unsigned long long tstride   = (unsigned long long)a.table_desc.stride * elt;
  const unsigned long long toff      = (unsigned long long)a.table_desc.storage_offset * elt;
  const unsigned long long dstride   = (unsigned long long)a.dense_desc.stride * elt;
  char* dense                        = static_cast<char*>(a.dense) + (unsigned long long)a.dense_desc.storage_offset * elt;
  unsigned long long align_or        = row_bytes | tstride | toff | dstride | reinterpret_cast<unsigned long long>(dense);
  for (int r = 0; r < a.table.world; r++)
    align_or |= reinterpret_cast<unsigned long long>(a.table.base[r]) | (r > 0 ? a.table.start[r] : 0ULL);
  const bool hot = a.hot.slot != nullptr && a.table.world > 1;
  if (hot) align_or |= a.hot.stride_bytes | reinterpret_cast<unsigned long long>(a.hot.rows);
  if (align_or % 16 != 0 || row_bytes == 0) return false;
  const size_t tile_bytes = (size_t)kBulkRows * row_bytes;
  const size_t kBulkSmemBudget = bulk_smem_budget();
  // bytes in flight per SM = warps x (stages - 1) x tile: deep rings on few warps (the copy engine does the work)
  if (kBulkSmemBudget / tile_bytes < 2) return false;  // rows above 2 KB: the register path
  const int warps  = (int)std::max<size_t>(1, std::min<size_t>(kBulkMaxWarps, kBulkSmemBudget / (3 * tile_bytes)));
  const int stages = (int)std::min<size_t>(8, kBulkSmemBudget / ((size_t)warps * tile_bytes));  // >= 2
  const size_t smem = (size_t)warps * stages * tile_bytes + (size_t)warps * stages * sizeof(unsigned long long);
  const int sms     = num_sms();
  const long long tiles = (a.n + kBulkRows - 1) / kBulkRows;
  // as many CTAs as fit an SM that has nothing else on it (at most 2): tiles are taken by ticket, so CTAs that only start
  // when the sampler kernel leaves cost nothing
  const int per_sm  = (int)std::max<size_t>(1, std::min<size_t>(2, (200 * 1024) / (smem + 1024)));
  int grid          = (int)std::min<long long>((tiles + warps - 1) / warps, (long long)(a.sms > 0 ? std::min(a.sms, sms) : sms) * per_sm);
  unsigned int* ticket = bulk_ticket_slot(a.stream);
  auto launch = [&](auto kernel) {
    static thread_local const void* configured[8] = {};
    bool known = false;
    for (auto* k : configured) known = known || k == reinterpret_cast<const void*>(kernel);
    if (!known) {
      WGB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
      for (auto& k : configured)
        if (!k) {
          k = reinterpret_cast<const void*>(kernel);
          break;
        }
    }
    kernel<<<grid, warps * 32, smem, a.stream>>>(a.table, toff, tstride, static_cast<const IdxT*>(a.idx), (long long)a.n, (unsigned int)row_bytes, dense,
                                                 dstride, stages, a.hot, ticket);
    WGB_CHECK_LAUNCH();
  };
  if (hot) launch(rows_bulk_gather_kernel<IdxT, true, true>);
  else if (a.table.world > 1) launch(rows_bulk_gather_kernel<IdxT, true, false>);
  else launch(rows_bulk_gather_kernel<IdxT, false, false>);
  return true;
}

#endif  // WGB_HOST_EMULATION

template <typename IdxT, bool SCATTER>
static void run_copy(const RowsOpArgs& a)
{
#ifndef WGB_HOST_EMULATION
  if (!SCATTER && try_bulk_gather<IdxT>(a)) return;
#endif
  size_t elt = dtype_size(a.table_desc.dtype);
  unsigned long long row_bytes = (unsigned long long)a.table_desc.sizes[1] * elt;
  unsigned long long dense_addr = reinterpret_cast<unsigned long long>(a.dense) + (unsigned long long)a.dense_desc.storage_offset * elt;
  unsigned long long base_or = 0;
  for (int r = 0; r < a.table.world; r++)
    base_or |= reinterpret_cast<unsigned long long>(a.table.base[r]) | (r > 0 ? a.table.start[r] : 0ULL);
  int vec = largest_pow2_alignment({row_bytes, (unsigned long long)a.table_desc.stride * elt,
                                    (unsigned long long)a.table_desc.storage_offset * elt, dense_addr,
                                    (unsigned long long)a.dense_desc.stride * elt, base_or},
                                   16);
  if ((size_t)vec < elt) vec = (int)elt;  // element alignment is guaranteed by construction
  unsigned int vpr = (unsigned int)(row_bytes / vec);
  // keep every launch below 2^31 vectors so the kernel can use 32-bit index math
  int64_t rows_per_launch = std::max<int64_t>(1, (int64_t)((1ULL << 31) - kBlock * kUnroll) / vpr);
  for (int64_t row0 = 0; row0 < a.n; row0 += rows_per_launch) {
    int64_t rows = std::min(rows_per_launch, a.n - row0);
    switch (vec) {
      case 16: launch_copy<IdxT, 16, SCATTER>(a, row0, rows, vpr); break;
      case 8: launch_copy<IdxT, 8, SCATTER>(a, row0, rows, vpr); break;
      case 4: launch_copy<IdxT, 4, SCATTER>(a, row0, rows, vpr); break;
      case 2: launch_copy<IdxT, 2, SCATTER>(a, row0, rows, vpr); break;
      default: launch_copy<IdxT, 1, SCATTER>(a, row0, rows, vpr); break;
    }
  }
}

template <typename TableT, typename DenseT, typename IdxT, bool SCATTER>
static void run_convert(const RowsOpArgs& a)
{
  unsigned long long dense_addr = reinterpret_cast<unsigned long long>(a.dense) / sizeof(DenseT) + (unsigned long long)a.dense_desc.storage_offset;
  unsigned long long tbase_or = 0;
  for (int r = 0; r < a.table.world; r++)
    tbase_or |= reinterpret_cast<unsigned long long>(a.table.base[r]) / sizeof(TableT) | (r > 0 ? a.table.start[r] / sizeof(TableT) : 0ULL);
  int align = largest_pow2_alignment({(unsigned long long)a.table_desc.sizes[1], (unsigned long long)a.table_desc.stride,
                                      (unsigned long long)a.table_desc.storage_offset, dense_addr,
                                      (unsigned long long)a.dense_desc.stride, tbase_or},
                                     4);
  if (align == 2) align = 1;
  unsigned int vpr        = (unsigned int)(a.table_desc.sizes[1] / align);
  int64_t rows_per_launch = std::max<int64_t>(1, (int64_t)((1ULL << 31) - kBlock) / vpr);
  RowDiv rd               = make_row_div(vpr);
  for (int64_t row0 = 0; row0 < a.n; row0 += rows_per_launch) {
    int64_t rows             = std::min(rows_per_launch, a.n - row0);
    unsigned long long total = (unsigned long long)rows * vpr;
    const IdxT* idx          = static_cast<const IdxT*>(a.idx) + row0;
    DenseT* dense            = static_cast<DenseT*>(a.dense) + a.dense_desc.storage_offset + row0 * a.dense_desc.stride;
    int grid                 = grid_for(total, kBlock, a.sms);
    if (align == 4)
      rows_convert_kernel<TableT, DenseT, IdxT, 4, SCATTER><<<grid, kBlock, 0, a.stream>>>(
        a.table, (unsigned long long)a.table_desc.storage_offset, (unsigned long long)a.table_desc.stride, idx,
        (unsigned int)total, rd, dense, (unsigned long long)a.dense_desc.stride);
    else
      rows_convert_kernel<TableT, DenseT, IdxT, 1, SCATTER><<<grid, kBlock, 0, a.stream>>>(
        a.table, (unsigned long long)a.table_desc.storage_offset, (unsigned long long)a.table_desc.stride, idx,
        (unsigned int)total, rd, dense, (unsigned long long)a.dense_desc.stride);
    WGB_CHECK_LAUNCH();
  }
}

template <typename F>
static void for_float_type(wholememory_dtype_t dt, F&& f)
{
  switch (dt) {
    case WHOLEMEMORY_DT_FLOAT: f(float{}); break;
    case WHOLEMEMORY_DT_HALF: f(__half{}); break;
    case WHOLEMEMORY_DT_DOUBLE: f(double{}); break;
    case WHOLEMEMORY_DT_BF16: f(__nv_bfloat16{}); break;
    default: throw logic_error("not a floating dtype");
  }
}
template <typename F>
static void for_int_type(wholememory_dtype_t dt, F&& f)
{
  switch (dt) {
    case WHOLEMEMORY_DT_INT: f(int32_t{}); break;
    case WHOLEMEMORY_DT_INT64: f(int64_t{}); break;
    case WHOLEMEMORY_DT_INT16: f(int16_t{}); break;
    case WHOLEMEMORY_DT_INT8: f(int8_t{}); break;
    default: throw logic_error("not an integer dtype");
  }
}

template <typename IdxT, bool SCATTER>
static void run_rows_op_typed(const RowsOpArgs& a)
{
  if (a.table_desc.dtype == a.dense_desc.dtype) {
    run_copy<IdxT, SCATTER>(a);
    return;
  }
  if (wholememory_dtype_is_floating_number(a.table_desc.dtype)) {
    for_float_type(a.table_desc.dtype, [&](auto t) {
      for_float_type(a.dense_desc.dtype, [&](auto d) {
        using T = decltype(t);
        using D = decltype(d);
        if constexpr (!std::is_same<T, D>::value) run_convert<T, D, IdxT, SCATTER>(a);
      });
    });
  } else {
    for_int_type(a.table_desc.dtype, [&](auto t) {
      for_int_type(a.dense_desc.dtype, [&](auto d) {
        using T = decltype(t);
        using D = decltype(d);
        if constexpr (!std::is_same<T, D>::value) run_convert<T, D, IdxT, SCATTER>(a);
      });
    });
  }
}

// Validation shared by gather and scatter; mirrors cpp/src/wholememory_ops/gather_op.cpp:12-70 and
// functions/gather_func.cu:53-66 (same error codes for the same mistakes).
wholememory_error_code_t rows_op(wholememory_tensor_t wm_tensor, wholememory_tensor_t indices_tensor,
                                 wholememory_tensor_t dense_tensor, void* stream, int sms, bool scatter, const wgb_hot_rows* hot)
{
  if (!wm_tensor || !indices_tensor || !dense_tensor) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_tensor_description_t td = *wholememory_tensor_get_tensor_description(wm_tensor);
  if (td.dim != 1 && td.dim != 2) {
    log_msg(LEVEL_ERROR, "wholememory_tensor should be 1D or 2D tensor.");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (td.dim == 1 && !wholememory_unsqueeze_tensor(&td, 1)) return WHOLEMEMORY_LOGIC_ERROR;
  wholememory_matrix_description_t table_desc;
  if (!wholememory_convert_tensor_desc_to_matrix(&table_desc, &td)) return WHOLEMEMORY_LOGIC_ERROR;
  wholememory_tensor_description_t* id = wholememory_tensor_get_tensor_description(indices_tensor);
  if (id->dim != 1) {
    log_msg(LEVEL_ERROR, "indices tensor should be 1D tensor");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (id->dtype != WHOLEMEMORY_DT_INT && id->dtype != WHOLEMEMORY_DT_INT64) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_tensor_description_t dd = *wholememory_tensor_get_tensor_description(dense_tensor);
  // the reference compares against the (already unsqueezed) table description, i.e. a 1-D table takes an
  // [n, 1] output (gather_op.cpp:33-52); a plain 1-D output is accepted as well.
  if (dd.dim != td.dim && !(dd.dim == 1 && wholememory_tensor_get_tensor_description(wm_tensor)->dim == 1)) {
    log_msg(LEVEL_ERROR, "%s tensor should be same dim as wholememory_tensor.", scatter ? "input" : "output");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (dd.dim == 1 && !wholememory_unsqueeze_tensor(&dd, 1)) return WHOLEMEMORY_LOGIC_ERROR;
  wholememory_matrix_description_t dense_desc;
  if (!wholememory_convert_tensor_desc_to_matrix(&dense_desc, &dd)) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_array_description_t idx_desc;
  if (!wholememory_convert_tensor_desc_to_array(&idx_desc, id)) return WHOLEMEMORY_INVALID_INPUT;

  return guarded(scatter ? "wholememory_scatter" : "wholememory_gather", [&] {
    bool tf = wholememory_dtype_is_floating_number(table_desc.dtype);
    bool df = wholememory_dtype_is_floating_number(dense_desc.dtype);
    WGB_EXPECTS(tf || wholememory_dtype_is_integer_number(table_desc.dtype), "bad embedding dtype");
    WGB_EXPECTS(df || wholememory_dtype_is_integer_number(dense_desc.dtype), "bad output dtype");
    WGB_EXPECTS(tf == df, "embedding and output should be same number type, e.g. floating number or integer number.");
    WGB_EXPECTS(dense_desc.sizes[1] == table_desc.sizes[1], "embedding dim of table and output differ");
    WGB_EXPECTS(dense_desc.sizes[0] >= idx_desc.size, "output has fewer rows than indices");
    if (idx_desc.size == 0 || table_desc.sizes[1] == 0) return;
    wholememory_tensor_t dense_root = wholememory_tensor_get_root(dense_tensor);
    WGB_EXPECTS(!dense_root->is_wholememory, "indices/output must not be WholeMemory tensors");
    wholememory_tensor_t idx_root = wholememory_tensor_get_root(indices_tensor);
    WGB_EXPECTS(!idx_root->is_wholememory, "indices must not be a WholeMemory tensor");
    RowsOpArgs a;
    a.table      = make_chunk_ref(wm_tensor);
    a.table_desc = table_desc;
    a.idx        = static_cast<const char*>(idx_root->storage_ptr) + idx_desc.storage_offset * dtype_size(idx_desc.dtype);
    a.idx_dtype  = idx_desc.dtype;
    a.n          = idx_desc.size;
    a.dense      = dense_root->storage_ptr;
    a.dense_desc = dense_desc;
    a.scatter    = scatter;
    if (hot && !scatter) a.hot = *hot;
    a.sms        = sms;
    a.stream     = as_stream(stream);
    if (idx_desc.dtype == WHOLEMEMORY_DT_INT) {
      if (scatter) run_rows_op_typed<int32_t, true>(a);
      else run_rows_op_typed<int32_t, false>(a);
    } else {
      if (scatter) run_rows_op_typed<int64_t, true>(a);
      else run_rows_op_typed<int64_t, false>(a);
    }
  });
}

}  // namespace wgb

extern "C" {

wholememory_error_code_t wholememory_gather(wholememory_tensor_t wholememory_tensor,
                                            wholememory_tensor_t indices_tensor,
                                            wholememory_tensor_t output_tensor,
                                            wholememory_env_func_t* /*p_env_fns*/, void* stream, int gather_sms)
{
  return wgb::rows_op(wholememory_tensor, indices_tensor, output_tensor, stream, gather_sms, false, nullptr);
}

wholememory_error_code_t wholememory_scatter(wholememory_tensor_t input_tensor, wholememory_tensor_t indices_tensor,
                                             wholememory_tensor_t wholememory_tensor,
                                             wholememory_env_func_t* /*p_env_fns*/, void* stream, int scatter_sms)
{
  wgb::hot_rows_invalidate_for_tensor(wholememory_tensor);  // a replica of rows that are about to change would go stale
  return wgb::rows_op(wholememory_tensor, indices_tensor, input_tensor, stream, scatter_sms, true, nullptr);
}

}  // extern "C"
