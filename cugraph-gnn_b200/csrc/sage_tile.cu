// A1, fused form: one GraphSAGE layer on a sampled CSR block with the dense part on the 5th-generation tensor cores.
//
//     out[i, :] = [ mean_{e in row i} x[indices[e], :]  ||  x[i, :] ] . W_cat^T (+ bias)        i < n_dst
//
// Consumer replaced: pylibwholegraph/torch/gnn_model.py:119-125 of the reference (layer(sub_graph, x_feat, x_target_feat):
// a scatter-add aggregation kernel, then two cuBLAS GEMMs, W_l on the aggregate and W_r on the target rows).  Here:
//   * sparse side  = gather-mean, a quarter-warp per destination row (four rows per warp at once), fp32 accumulation in registers;
//     the result never goes to HBM: it is written, as bf16, straight into the A operand tile in shared memory, in the
//     K-major 128-byte-swizzled layout tcgen05.mma reads.  The mean is kept to ~16 bits of mantissa by splitting it into
//     two bf16 terms (hi + lo), both multiplied by the same W_l tile, so the fp32 aggregation of the SIMT path survives the
//     bf16 operand format; the target rows x[i, :] are bf16 already and enter exactly;
//   * dense side   = W_cat = [W_l || W_r] (bf16, [F_out, 2 F_in], K contiguous) is brought ONCE per CTA by TMA
//     (cp.async.bulk.tensor.2d, 128B swizzle) and stays in shared memory for every 128-row tile the persistent CTA takes;
//     one elected thread issues 24 tcgen05.mma (M = 128, N = F_out, K = 16) per tile, accumulator in tensor memory
//     (128 lanes x F_out fp32 columns), completion through tcgen05.commit -> mbarrier;
//   * epilogue     = all 32 warps read the accumulator (tcgen05.ld 32x32b.x32: warp w owns lanes 32 (w % 4), columns
//     32 (w / 4)), add the bias, and store fp32 rows; tensor memory holds two accumulator stages, so the epilogue of tile
//     t - 1 runs underneath the MMAs of tile t.
// F_in = 128 (the feature width of every BASELINE config), F_out a multiple of 16 up to 256.
// The kernel is bound by the gather (nnz x 256 B of bf16 rows from HBM/L2); the tensor-core work of a tile (25 MFLOP)
// takes ~1 us.  What the fusion saves is the write + re-read of the aggregate and of the target rows and two launches.
#include "wm_common.cuh"

#include <cuda.h>
#include <wholememory/b200_ops.h>

namespace wgb {

namespace {

constexpr int kFin       = 128;
constexpr int kTileRows  = 128;
constexpr int kThreads   = 1024;
constexpr int kKBlock    = 64;                                 // bf16 elements per 128-byte swizzled row
constexpr int kABlocks   = 3 * kFin / kKBlock;                 // mean_hi | mean_lo | self
constexpr int kBBlocks   = 2 * kFin / kKBlock;                 // W_l | W_r
constexpr int kABlockBytes = kTileRows * 128;                  // 16 KB
constexpr int kTmemCols  = 512;                                // two accumulator stages of 256 fp32 columns

__device__ __forceinline__ unsigned int smem_addr(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity)
{
  asm volatile(
    "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(
      smem_addr(bar)),
    "r"(parity)
    : "memory");
}

// shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor:
// start address >> 4 in [0,14), LBO >> 4 in [16,30) (1: unused for swizzled K-major), SBO >> 4 in [32,46), version 1 in [46,48),
// layout type SWIZZLE_128B = 2 in [61,64))
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned int saddr)
{
  return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | (1ULL << 16) | ((unsigned long long)(1024 >> 4) << 32) | (1ULL << 46) | (2ULL << 61);
}

// instruction descriptor, kind::f16: D fp32 (bits 4-5 = 1), A and B bf16 (bits 7-9 = 1, 10-12 = 1), both K-major, N >> 3 in [17,23),
// M >> 4 in [24,29)
__device__ __forceinline__ unsigned int umma_idesc_bf16(int M, int N)
{
  return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned int)(N >> 3) << 17) | ((unsigned int)(M >> 4) << 24);
}

// byte offset of (row r, bf16 element k) inside an A tile made of K-blocks of 64 elements, 128-byte swizzle:
// the 16-byte chunk index of a row is XORed with the row's position in its 8-row group
__device__ __forceinline__ unsigned int a_tile_offset(int r, int k)
{
  const int kb = k / kKBlock, kk = k % kKBlock;
  const int chunk = kk >> 3;
  return (unsigned int)(kb * kABlockBytes + (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4) + (kk & 7) * 2);
}

// fp32 accumulation of eight bf16 values packed in a uint4 (a bf16 is the upper half of the fp32 with the same value)
__device__ __forceinline__ void acc_bf16x8(float (&acc)[8], const uint4& v)
{
  acc[0] += __uint_as_float(v.x << 16); acc[1] += __uint_as_float(v.x & 0xFFFF0000u);
  acc[2] += __uint_as_float(v.y << 16); acc[3] += __uint_as_float(v.y & 0xFFFF0000u);
  acc[4] += __uint_as_float(v.z << 16); acc[5] += __uint_as_float(v.z & 0xFFFF0000u);
  acc[6] += __uint_as_float(v.w << 16); acc[7] += __uint_as_float(v.w & 0xFFFF0000u);
}

// eight fp32 means -> hi (bf16 round-to-nearest) and lo (bf16 of the remainder): hi + lo carries ~16 bits of mantissa
__device__ __forceinline__ void split_bf16x8(const float (&m)[8], uint4& hi, uint4& lo)
{
  unsigned int h[4], l[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(m[2 * c], m[2 * c + 1]);
    const float2 hf         = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(m[2 * c] - hf.x, m[2 * c + 1] - hf.y);
    h[c] = *reinterpret_cast<const unsigned int*>(&hh);
    l[c] = *reinterpret_cast<const unsigned int*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Persistent CTA, 32 warps, one 128-row tile per iteration:
//   gather(t): a quarter-warp (8 lanes) per destination row, 4 rows per warp at once; lane q of the quarter owns feature
//              columns [8q, 8q + 8) and [64 + 8q, 64 + 8q + 8) (two 16-byte chunks: one per 64-column K-block, so the 8 lanes of a
//              row read 128 contiguous bytes per load instruction and write 128 contiguous swizzled bytes of the operand tile);
//              the row's first 16 neighbour ids are prefetched one tile ahead (registers), so the critical path of a tile is one
//              round of row reads, two neighbours (4 x 16 B per lane) in flight;
//   mma(t)   : one thread, 24 tcgen05.mma into accumulator stage t & 1 (tensor memory holds two stages of 256 columns);
//   epilogue(t - 1): runs while mma(t) executes -- the accumulator of the previous tile is read (tcgen05.ld), biased and stored.
template <typename IdxT, typename PtrT>
__global__ void __launch_bounds__(kThreads, 1) sage_tile_kernel(const __grid_constant__ CUtensorMap w_map, const PtrT* __restrict__ indptr,
                                                                const IdxT* __restrict__ indices, const __nv_bfloat16* __restrict__ x,
                                                                long long x_stride, long long n_dst, int f_out,
                                                                const float* __restrict__ bias, float* __restrict__ out, long long out_stride)
{
  extern __shared__ unsigned char smem_raw[];
  // swizzled operand tiles need 1024-byte aligned bases (the host asks for 1 KB of slack)
  unsigned char* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;                                     // kABlocks x 16 KB
  unsigned char* sB = smem + kABlocks * kABlockBytes;           // kBBlocks x (f_out x 128 B)
  const unsigned int b_block_bytes = (unsigned int)f_out * 128u;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sB + kBBlocks * b_block_bytes);  // [0] W landed, [1 + s] MMAs into stage s done
  unsigned int* tmem_slot  = reinterpret_cast<unsigned int*>(bars + 3);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (wid == 1) {  // one warp owns the tensor-memory allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned int tmem = *tmem_slot;

  // W_cat -> shared memory, once: kBBlocks boxes of [f_out rows x 64 columns]
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bars[0])), "r"(kBBlocks * b_block_bytes) : "memory");
    for (int kb = 0; kb < kBBlocks; kb++)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_addr(sB + kb * b_block_bytes)),
                   "l"(reinterpret_cast<unsigned long long>(&w_map)), "r"(kb * kKBlock), "r"(0), "r"(smem_addr(&bars[0]))
                   : "memory");
  }

  const unsigned int idesc = umma_idesc_bf16(kTileRows, f_out);
  const long long n_tiles  = (n_dst + kTileRows - 1) / kTileRows;
  const int ql = lane & 7;                 // lane of the quarter-warp
  const int r  = wid * 4 + (lane >> 3);    // the quarter's row of the tile
  const __nv_bfloat16* xq = x + 8 * ql;    // this lane's two column chunks start here and 64 columns further

  // row range and first 16 neighbour ids of this quarter's row of `tile` (lane ql holds ids ql and 8 + ql; -1: none)
  auto row_range = [&](long long tile, long long& s, long long& e) {
    const long long i = tile * kTileRows + r;
    s = e = 0;
    if (tile < n_tiles && i < n_dst) {
      s = (long long)indptr[i];
      e = (long long)indptr[i + 1];
    }
  };
  auto first_ids = [&](long long s, long long e, int& i0, int& i1) {
    i0 = s + ql < e ? (int)indices[s + ql] : -1;
    i1 = s + 8 + ql < e ? (int)indices[s + 8 + ql] : -1;
  };
  auto epilogue = [&](long long tile, unsigned int stage) {
    const int lane_grp = wid & 3, col0 = (wid >> 2) * 32;
    if (col0 < f_out) {
      unsigned int v[32];
      const unsigned int taddr = tmem + ((unsigned int)(lane_grp * 32) << 16) + stage * 256u + (unsigned int)col0;
      asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // v[c] = D[row = lane][col0 + c].  A row-per-lane store would touch 32 different 128-byte lines per instruction (half of
      // the kernel's stall samples in the first version: LSU throttle).  Transpose the 32 x 32 block inside the warp (five
      // butterfly rounds of 16 shuffles), after which v[r] = D[row r][col0 + lane] and every store instruction writes one
      // 128-byte line; the bias of the lane's column rides along as a single register.
#pragma unroll
      for (int sft = 16; sft >= 1; sft >>= 1) {
        const bool upper = (lane & sft) != 0;
#pragma unroll
        for (int c = 0; c < 32; c++) {
          if ((c & sft) == 0) {
            const unsigned int send = upper ? v[c] : v[c + sft];
            const unsigned int recv = __shfl_xor_sync(0xffffffffu, send, sft);
            if (upper) v[c] = recv;
            else v[c + sft] = recv;
          }
        }
      }
      const int col = col0 + lane;
      if (col < f_out) {
        const float bv      = bias ? bias[col] : 0.f;
        const long long i0  = tile * kTileRows + lane_grp * 32;
        const int n_valid   = (int)(n_dst - i0 < 32 ? n_dst - i0 : 32);  // warp-uniform (<= 0: nothing to store)
        float* o            = out + i0 * out_stride + col;
        // streaming stores (evict-first): the output is written once and never read here, and must not push the feature rows
        // the gather re-reads out of L2 (v3 of the kernel: 33 % L2 hit rate on a block whose source rows fit L2 three times over)
#pragma unroll
        for (int rr = 0; rr < 32; rr++) {
          if (rr < n_valid) __stcs(o, __uint_as_float(v[rr]) + bv);
          o += out_stride;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  long long s_cur, e_cur;
  int id0, id1;
  row_range((long long)blockIdx.x, s_cur, e_cur);
  first_ids(s_cur, e_cur, id0, id1);
  unsigned int it = 0;  // tiles this CTA has taken
  long long prev_tile = -1;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
    const long long i = tile * kTileRows + r;
    // the next tile's row range: in flight during this tile's gather
    long long s_nxt, e_nxt;
    row_range(tile + gridDim.x, s_nxt, e_nxt);
    // the operand tile is free once the previous tile's MMAs have read it
    if (it > 0) mbar_wait(&bars[1 + ((it - 1) & 1u)], ((it - 1) >> 1) & 1u);
    // ---- sparse side: mean of the neighbour rows and the target row, as bf16, into the A tile -----------------------
    {
      float acc0[8], acc1[8];
#pragma unroll
      for (int c = 0; c < 8; c++)
        acc0[c] = acc1[c] = 0.f;
      uint4 self0 = make_uint4(0, 0, 0, 0), self1 = self0;
      const int deg = (int)(e_cur - s_cur);
      if (i < n_dst) {
        self0 = *reinterpret_cast<const uint4*>(xq + i * x_stride);
        self1 = *reinterpret_cast<const uint4*>(xq + i * x_stride + 64);
      }
      int deg_max = deg;  // uniform trip count for the warp's four rows
      deg_max = max(deg_max, __shfl_xor_sync(0xffffffffu, deg_max, 8));
      deg_max = max(deg_max, __shfl_xor_sync(0xffffffffu, deg_max, 16));
#pragma unroll 1
      for (int base = 0; base < deg_max; base += 8) {
        int mine;
        if (base == 0) mine = id0;
        else if (base == 8) mine = id1;
        else mine = base + ql < deg ? (int)indices[s_cur + base + ql] : -1;
#pragma unroll
        for (int j = 0; j < 8; j += 4) {
          if (base + j >= deg_max) break;  // warp-uniform
          int nb[4];
          uint4 lo4[4], hi4[4];
#pragma unroll
          for (int u = 0; u < 4; u++)
            nb[u] = __shfl_sync(0xffffffffu, mine, j + u, 8);
#pragma unroll
          for (int u = 0; u < 4; u++) {  // four neighbour rows (8 x 16 B per lane) in flight
            lo4[u] = hi4[u] = make_uint4(0, 0, 0, 0);
            if (nb[u] >= 0) {
              const __nv_bfloat16* p = xq + (long long)nb[u] * x_stride;
              lo4[u] = *reinterpret_cast<const uint4*>(p);
              hi4[u] = *reinterpret_cast<const uint4*>(p + 64);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            acc_bf16x8(acc0, lo4[u]);
            acc_bf16x8(acc1, hi4[u]);
          }
        }
      }
      const float inv = deg > 0 ? 1.0f / (float)deg : 0.f;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        acc0[c] *= inv;
        acc1[c] *= inv;
      }
      uint4 hi0, lo0, hi1, lo1;
      split_bf16x8(acc0, hi0, lo0);
      split_bf16x8(acc1, hi1, lo1);
      *reinterpret_cast<uint4*>(sA + a_tile_offset(r, 8 * ql))                 = hi0;
      *reinterpret_cast<uint4*>(sA + a_tile_offset(r, 64 + 8 * ql))            = hi1;
      *reinterpret_cast<uint4*>(sA + a_tile_offset(r, kFin + 8 * ql))          = lo0;
      *reinterpret_cast<uint4*>(sA + a_tile_offset(r, kFin + 64 + 8 * ql))     = lo1;
      *reinterpret_cast<uint4*>(sA + a_tile_offset(r, 2 * kFin + 8 * ql))      = self0;
      *reinterpret_cast<uint4*>(sA + a_tile_offset(r, 2 * kFin + 64 + 8 * ql)) = self1;
    }
    // the next tile's first neighbour ids: in flight during this tile's MMAs and the previous tile's epilogue
    s_cur = s_nxt;
    e_cur = e_nxt;
    first_ids(s_cur, e_cur, id0, id1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core's reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- dense side: 24 MMAs into tensor memory stage it & 1, issued by one thread ----------------------------------
    if (wid == 0) {
      if (it == 0) mbar_wait(&bars[0], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        int first = 1;
        const unsigned int d_tmem = tmem + (it & 1u) * 256u;
#pragma unroll
        for (int ab = 0; ab < kABlocks; ab++) {
          const int bb = ab < 4 ? (ab & 1) : ab - 2;  // mean_hi and mean_lo both meet W_l (B blocks 0, 1), self meets W_r (2, 3)
          const unsigned long long da = umma_desc_sw128(smem_addr(sA + ab * kABlockBytes));
          const unsigned long long db = umma_desc_sw128(smem_addr(sB + bb * b_block_bytes));
#pragma unroll
          for (int k = 0; k < kKBlock / 16; k++) {
            const unsigned int accum = first ? 0u : 1u;
            first                    = 0;
            // advancing 16 bf16 = 32 B along K inside the swizzled row: + 2 in the (address >> 4) field
            asm volatile(
              "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
              "l"(da + (unsigned long long)(2 * k)), "l"(db + (unsigned long long)(2 * k)), "r"(idesc), "r"(accum)
              : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bars[1 + (it & 1u)])) : "memory");
      }
      __syncwarp();
    }
    // ---- epilogue of the PREVIOUS tile, underneath this tile's MMAs -------------------------------------------------
    if (it > 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");  // (its MMAs were waited for before this tile's gather)
      epilogue(prev_tile, (it - 1) & 1u);
    }
    prev_tile = tile;
  }
  if (it > 0) {
    mbar_wait(&bars[1 + ((it - 1) & 1u)], ((it - 1) >> 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue(prev_tile, (it - 1) & 1u);
  }
  __syncthreads();
  if (wid == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult st;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) {
      cudaGetLastError();
      throw cuda_error("cuTensorMapEncodeTiled is not available from this driver");
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

}  // namespace wgb

extern "C" {

wholememory_error_code_t wholegraph_sage_layer_forward(wholememory_tensor_t indptr, wholememory_tensor_t indices, wholememory_tensor_t x,
                                                       wholememory_tensor_t w_cat, wholememory_tensor_t bias, wholememory_tensor_t out,
                                                       void* stream)
{
  using namespace wgb;
  if (!indptr || !indices || !x || !w_cat || !out) return WHOLEMEMORY_INVALID_INPUT;
  auto* pd = wholememory_tensor_get_tensor_description(indptr);
  auto* id = wholememory_tensor_get_tensor_description(indices);
  auto* xd = wholememory_tensor_get_tensor_description(x);
  auto* wd = wholememory_tensor_get_tensor_description(w_cat);
  auto* od = wholememory_tensor_get_tensor_description(out);
  if (pd->dim != 1 || id->dim != 1 || xd->dim != 2 || wd->dim != 2 || od->dim != 2 || pd->sizes[0] < 1) return WHOLEMEMORY_INVALID_INPUT;
  if ((pd->dtype != WHOLEMEMORY_DT_INT && pd->dtype != WHOLEMEMORY_DT_INT64) || (id->dtype != WHOLEMEMORY_DT_INT && id->dtype != WHOLEMEMORY_DT_INT64))
    return WHOLEMEMORY_INVALID_INPUT;
  if (xd->dtype != WHOLEMEMORY_DT_BF16 || wd->dtype != WHOLEMEMORY_DT_BF16 || od->dtype != WHOLEMEMORY_DT_FLOAT) return WHOLEMEMORY_INVALID_INPUT;
  const long long n_dst = pd->sizes[0] - 1;
  const int f_out       = (int)wd->sizes[0];
  if (xd->sizes[1] != kFin || wd->sizes[1] != 2 * kFin) return WHOLEMEMORY_NOT_IMPLEMENTED;  // F_in = 128
  if (f_out < 16 || f_out > 256 || f_out % 16 != 0) return WHOLEMEMORY_NOT_IMPLEMENTED;
  if (od->sizes[0] < n_dst || od->sizes[1] != f_out || xd->sizes[0] < n_dst) return WHOLEMEMORY_INVALID_INPUT;
  if (bias) {
    auto* bd = wholememory_tensor_get_tensor_description(bias);
    if (bd->dim != 1 || bd->dtype != WHOLEMEMORY_DT_FLOAT || bd->sizes[0] != f_out) return WHOLEMEMORY_INVALID_INPUT;
  }
  return guarded("wholegraph_sage_layer_forward", [&] {
    for (wholememory_tensor_t t : {indptr, indices, x, w_cat, out})
      WGB_EXPECTS(!wholememory_tensor_get_root(t)->is_wholememory, "the fused SAGE layer takes local device tensors");
    if (n_dst == 0) return;
    auto ptr_of = [](wholememory_tensor_t t) {
      auto* d = wholememory_tensor_get_tensor_description(t);
      return static_cast<char*>(wholememory_tensor_get_data_pointer(t)) + (size_t)d->storage_offset * dtype_size(d->dtype);
    };
    const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(ptr_of(x));
    char* wp                = ptr_of(w_cat);
    float* op               = reinterpret_cast<float*>(ptr_of(out));
    const float* bp         = bias ? reinterpret_cast<const float*>(ptr_of(bias)) : nullptr;
    WGB_EXPECTS(reinterpret_cast<unsigned long long>(xp) % 16 == 0 && xd->strides[0] % 8 == 0, "x rows must be 16-byte aligned");
    WGB_EXPECTS(reinterpret_cast<unsigned long long>(op) % 16 == 0 && od->strides[0] % 4 == 0, "out rows must be 16-byte aligned");
    WGB_EXPECTS(reinterpret_cast<unsigned long long>(wp) % 16 == 0 && wd->strides[0] == 2 * kFin, "W_cat must be contiguous and 16-byte aligned");
    CUtensorMap map;
    const cuuint64_t gdim[2]    = {(cuuint64_t)(2 * kFin), (cuuint64_t)f_out};
    const cuuint64_t gstride[1] = {(cuuint64_t)(2 * kFin) * 2};
    const cuuint32_t box[2]     = {(cuuint32_t)kKBlock, (cuuint32_t)f_out};
    const cuuint32_t estr[2]    = {1, 1};
    CUresult r = encode_tiled()(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wp, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw cuda_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    const size_t smem = (size_t)kABlocks * kABlockBytes + (size_t)kBBlocks * f_out * 128 + 64 + 1024;
    WGB_EXPECTS(xd->sizes[0] <= 0x7FFFFFFFLL, "the fused SAGE layer takes blocks of fewer than 2^31 source rows");
    const long long n_tiles = (n_dst + kTileRows - 1) / kTileRows;
    const int grid = (int)std::min<long long>(n_tiles, num_sms());
    cudaStream_t st = as_stream(stream);
    auto launch = [&](auto kernel, auto ip, auto ix) {
      WGB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kernel<<<grid, kThreads, smem, st>>>(map, ip, ix, xp, (long long)xd->strides[0], n_dst, f_out, bp, op, (long long)od->strides[0]);
      WGB_CHECK_LAUNCH();
    };
    const void* ipv = ptr_of(indptr);
    const void* ixv = ptr_of(indices);
    if (id->dtype == WHOLEMEMORY_DT_INT) {
      if (pd->dtype == WHOLEMEMORY_DT_INT) launch(sage_tile_kernel<int, int>, static_cast<const int*>(ipv), static_cast<const int*>(ixv));
      else launch(sage_tile_kernel<int, long long>, static_cast<const long long*>(ipv), static_cast<const int*>(ixv));
    } else {
      if (pd->dtype == WHOLEMEMORY_DT_INT) launch(sage_tile_kernel<long long, int>, static_cast<const int*>(ipv), static_cast<const long long*>(ixv));
      else launch(sage_tile_kernel<long long, long long>, static_cast<const long long*>(ipv), static_cast<const long long*>(ixv));
    }
  });
}

}  // extern "C"
