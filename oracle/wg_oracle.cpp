// TEST INFRASTRUCTURE ONLY -- CPU oracle for the cugraph-gnn hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.  The product path (cugraph-gnn_b200/) never does.
//
// This is a plain C++17 (+OpenMP) restatement of the reference's OWN host reference
// algorithms (the CPU code its gtests compare the GPU kernels against).  Every function
// cites the reference file:line it follows (paths relative to /root/reference).
//
// PARITY STATUS
//   * PCG32 core (XSH-RR 64/32): pinned against the published pcg32 known-answer vector
//     (pcg-c-basic demo, seed 42 / seq 54) in tests/test_oracle_cpu.py.
//   * RAFT PCGenerator wrapping (raft 26.10, un-vendored: cpp/cmake/thirdparty/get_raft.cmake:20-31):
//     the (seed, subsequence) -> stream mapping, in particular the skip-ahead by `subsequence`
//     performed by PCGenerator(DeviceState, subsequence), is restated from the published RAFT
//     source from memory.  The reference holds no literal RNG value anywhere
//     (its tests call the same RAFT code on the host, cpp/src/wholegraph_ops/raft_random_gen.cu:15-54)
//     => "parity unpinned" for the random stream itself.  Everything that is deterministic
//     (offsets, take-all paths, append-unique, gather/scatter, fanout -1 multi-hop) is pinned
//     against the reference's own test expectations in tests/.
//
// Build: make -C oracle   (g++ -O3 -fopenmp -shared -fPIC)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <unordered_map>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// wholememory_dtype_t values (cpp/include/wholememory/tensor_description.h:19-30)
enum : int { DT_FLOAT = 1, DT_HALF = 2, DT_DOUBLE = 3, DT_BF16 = 4, DT_INT = 5, DT_INT64 = 6, DT_INT16 = 7, DT_INT8 = 8 };

// ----------------------------------------------------------------------------------------------
// RNG: RAFT raft::random::detail::PCGenerator (PCG XSH-RR 64/32), restated.
// call sites: cpp/src/wholegraph_ops/unweighted_sample_without_replacement_func.cuh:137,183-187
//             cpp/src/wholegraph_ops/raft_random_gen.cu:32-53
// ----------------------------------------------------------------------------------------------
constexpr uint64_t kPcgMult = 6364136223846793005ULL;

struct Pcg {
  uint64_t state;
  uint64_t inc;

  inline uint32_t next_u32()
  {
    uint64_t old        = state;
    state               = old * kPcgMult + inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot        = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
  }
  // F. B. Brown, "Random Number Generation with Arbitrary Strides": affine power by squaring.
  inline void skipahead(uint64_t offset)
  {
    uint64_t G = 1, h = kPcgMult, C = 0, f = inc;
    while (offset) {
      if (offset & 1) {
        G = G * h;
        C = C * h + f;
      }
      f = f * (h + 1);
      h = h * h;
      offset >>= 1;
    }
    state = state * G + C;
  }
  // == pcg32_srandom_r(seed, subsequence) followed by a skip-ahead.
  inline void init(uint64_t seed, uint64_t subsequence, uint64_t offset)
  {
    state = 0;
    inc   = (subsequence << 1u) | 1u;
    next_u32();
    state += seed;
    next_u32();
    skipahead(offset);
  }
  // PCGenerator(DeviceState{seed, base_subsequence}, subsequence): stream = base+subsequence,
  // then skip-ahead by `subsequence` draws (restated from RAFT; see header note).
  inline void init_raft(uint64_t seed, uint64_t base_subsequence, uint64_t subsequence)
  {
    init(seed, base_subsequence + subsequence, subsequence);
  }
  inline int32_t next_i32() { return (int32_t)(next_u32() & 0x7fffffffu); }
  inline uint64_t next_u64()
  {
    uint32_t a = next_u32();
    uint32_t b = next_u32();
    return (uint64_t)a | ((uint64_t)b << 32);
  }
  inline int64_t next_i64() { return (int64_t)(next_u64() & 0x7fffffffffffffffULL); }
  inline float next_float() { return (float)(next_u32() >> 8) / (float)(1u << 24); }
};

inline int clz64(uint64_t x)
{
  // cpp/tests/wholegraph_ops/graph_sampling_test_utils.cu:529-537 (count_one == leading zeros)
  int c = 0;
  while (x) {
    x >>= 1;
    c++;
  }
  return 64 - c;
}

// cpp/tests/wholegraph_ops/graph_sampling_test_utils.cu:540-557 (host_gen_key_from_weight)
template <typename W>
inline float gen_key_from_weight(W weight, Pcg& rng)
{
  float u     = rng.next_float();
  u           = (float)(-(0.5 + 0.5 * (double)u));
  uint64_t r2 = 0;
  int extra   = -1;
  do {
    r2 = rng.next_u64();
    extra++;
  } while (!r2);
  int one_bit = clz64(r2) + extra * 64;
  u *= exp2f((float)-one_bit);
  float logk = (log1pf(u) / logf(2.0f)) * (1.0f / (float)weight);
  return logk;
}

// launch-geometry tables: unweighted_sample_without_replacement_func.cuh:412-447,
// host twin graph_sampling_test_utils.cu:345-351
const int kWarpCount[32] = {1, 1, 1, 2, 2, 2, 4, 4, 4, 4, 4, 4, 8, 8, 8, 8,
                            8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8};
const int kItems[32]     = {1, 2, 3, 2, 3, 3, 2, 2, 3, 3, 3, 3, 2, 2, 2, 2,
                            3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4};

template <typename T>
inline int64_t rd_idx(const void* p, int64_t i)
{
  return (int64_t)((const T*)p)[i];
}
inline int64_t read_index(const void* p, int dtype, int64_t i)
{
  return dtype == DT_INT ? rd_idx<int32_t>(p, i) : rd_idx<int64_t>(p, i);
}
inline void write_index(void* p, int dtype, int64_t i, int64_t v)
{
  if (dtype == DT_INT)
    ((int32_t*)p)[i] = (int32_t)v;
  else
    ((int64_t*)p)[i] = v;
}

// sequential Fisher-Yates-without-replacement resolution:
// graph_sampling_test_utils.cu:295-310 (random_sample_without_replacement_cpu_base<0>)
inline void resolve_chain(const int32_t* r, int M, int N, std::vector<int>& Q, int* a)
{
  // Only positions r[i] and N-i-1 are ever touched -> keep Q sparse when N is large.
  if ((int)Q.size() < N) Q.resize(N);
  for (int i = 0; i < N; i++)
    Q[i] = i;
  for (int i = 0; i < M; i++) {
    a[i]    = Q[r[i]];
    Q[r[i]] = Q[N - i - 1];
  }
}

inline void resolve_chain_sparse(const int32_t* r, int M, int N, int* a)
{
  // O(M^2) sparse form of the same recurrence for hub vertices (N >> M).
  int keys[1024];
  int vals[1024];
  int n = 0;
  auto get = [&](int pos) {
    for (int k = n - 1; k >= 0; k--)
      if (keys[k] == pos) return vals[k];
    return pos;
  };
  for (int i = 0; i < M; i++) {
    int x   = r[i];
    a[i]    = get(x);
    int v   = get(N - i - 1);
    keys[n] = x;
    vals[n] = v;
    n++;
  }
}

// one seed row of the uniform sampler (graph_sampling_test_utils.cu:353-401)
template <typename ColT>
inline void uniform_row(const int64_t* row_ptr, const ColT* col, int64_t v, int64_t b, int M,
                        uint64_t seed, int out_off, ColT* out_dest, int32_t* out_lid,
                        int64_t* out_gid, std::vector<int>& Q, std::vector<int32_t>& r)
{
  int64_t start = row_ptr[v];
  int64_t end   = row_ptr[v + 1];
  int64_t deg   = end - start;
  if (deg <= 0) return;
  if (M <= 0 || deg <= M) {
    for (int64_t j = 0; j < deg; j++) {
      out_dest[out_off + j] = col[start + j];
      if (out_lid) out_lid[out_off + j] = (int32_t)b;
      if (out_gid) out_gid[out_off + j] = start + j;
    }
    return;
  }
  int N        = (int)deg;
  int func_idx = (M - 1) / 32;
  int T        = kWarpCount[func_idx] * 32;
  int ipt      = kItems[func_idx];
  if ((int)r.size() < T * ipt) r.resize(T * ipt);
  for (int j = 0; j < T; j++) {
    Pcg rng;
    rng.init_raft(seed, 0, (uint64_t)(b * T + j));
    for (int k = 0; k < ipt; k++) {
      int id    = k * T + j;
      int32_t x = rng.next_i32();  // always drawn, even if id >= M
      r[id]     = id < M ? x % (N - id) : N;
    }
  }
  int a[1024];
  if (N > 8 * M)
    resolve_chain_sparse(r.data(), M, N, a);
  else
    resolve_chain(r.data(), M, N, Q, a);
  for (int i = 0; i < M; i++) {
    out_dest[out_off + i] = col[start + a[i]];
    if (out_lid) out_lid[out_off + i] = (int32_t)b;
    if (out_gid) out_gid[out_off + i] = start + a[i];
  }
}

// one seed row of the weighted (A-Res) sampler (graph_sampling_test_utils.cu:559-656).
// Output order inside a row: ascending key (heap pop order), as the host reference emits it.
template <typename ColT, typename WT>
inline void weighted_row(const int64_t* row_ptr, const ColT* col, const WT* wgt, int64_t v,
                         int64_t b, int M, uint64_t seed, int out_off, ColT* out_dest,
                         int32_t* out_lid, int64_t* out_gid, float* out_key)
{
  int64_t start = row_ptr[v];
  int64_t end   = row_ptr[v + 1];
  int64_t deg   = end - start;
  if (deg <= 0) return;
  if (M <= 0 || deg <= M) {
    for (int64_t j = 0; j < deg; j++) {
      out_dest[out_off + j] = col[start + j];
      if (out_lid) out_lid[out_off + j] = (int32_t)b;
      if (out_gid) out_gid[out_off + j] = start + j;
      if (out_key) out_key[out_off + j] = 0.f;
    }
    return;
  }
  int block = M > 256 ? 256 : 128;
  // min-heap on key (smallest of the kept keys on top)
  typedef std::pair<float, int> KI;
  auto cmp = [](const KI& l, const KI& r) { return l.first > r.first; };
  std::priority_queue<KI, std::vector<KI>, decltype(cmp)> heap(cmp);
  int processed = 0;
  for (int j = 0; j < block; j++) {
    Pcg rng;
    rng.init_raft(seed, 0, (uint64_t)(b * block + j));
    for (int64_t id = j; id < deg; id += block) {
      float key = gen_key_from_weight<WT>(wgt[start + id], rng);
      processed++;
      if (processed <= M) {
        heap.push(KI(key, (int)id));
      } else if (heap.top().first < key) {
        heap.pop();
        heap.push(KI(key, (int)id));
      }
    }
  }
  for (int i = 0; i < M; i++) {
    KI t                  = heap.top();
    out_dest[out_off + i] = col[start + t.second];
    if (out_lid) out_lid[out_off + i] = (int32_t)b;
    if (out_gid) out_gid[out_off + i] = start + t.second;
    if (out_key) out_key[out_off + i] = t.first;
    heap.pop();
  }
}

// ---- fp16 / bf16 <-> fp32 (round-to-nearest-even), matching CUDA static_cast semantics ----
inline float half_to_float(uint16_t h)
{
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp  = (h >> 10) & 0x1f;
  uint32_t man  = h & 0x3ffu;
  uint32_t f;
  if (exp == 0) {
    if (man == 0) {
      f = sign;
    } else {
      int e = -1;
      do {
        e++;
        man <<= 1;
      } while (!(man & 0x400u));
      man &= 0x3ffu;
      f = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    f = sign | 0x7f800000u | (man << 13);
  } else {
    f = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float out;
  memcpy(&out, &f, 4);
  return out;
}
inline uint16_t float_to_half(float x)
{
  uint32_t f;
  memcpy(&f, &x, 4);
  uint32_t sign = (f >> 16) & 0x8000u;
  uint32_t abs  = f & 0x7fffffffu;
  if (abs > 0x7f800000u) return (uint16_t)(sign | 0x7fffu);  // NaN (CUDA canonical 0x7fff)
  if (abs >= 0x47800000u) {                                   // >= 65536 -> inf (after rounding check below)
    return (uint16_t)(sign | 0x7c00u);
  }
  if (abs >= 0x38800000u) {  // normal half
    uint32_t man   = abs & 0x7fffffu;
    uint32_t exp   = (abs >> 23) - 112;
    uint32_t h     = (exp << 10) | (man >> 13);
    uint32_t round = man & 0x1fffu;
    if (round > 0x1000u || (round == 0x1000u && (h & 1))) h++;
    return (uint16_t)(sign | h);  // may carry into inf: correct
  }
  if (abs < 0x33000000u) return (uint16_t)sign;  // < 2^-25 -> 0
  // subnormal half
  uint32_t exp   = abs >> 23;
  uint32_t man   = (abs & 0x7fffffu) | 0x800000u;
  uint32_t shift = 126 - exp;  // 14..24
  uint32_t h     = man >> shift;
  uint32_t rem   = man & ((1u << shift) - 1);
  uint32_t half  = 1u << (shift - 1);
  if (rem > half || (rem == half && (h & 1))) h++;
  return (uint16_t)(sign | h);
}
inline float bf16_to_float(uint16_t b)
{
  uint32_t f = (uint32_t)b << 16;
  float out;
  memcpy(&out, &f, 4);
  return out;
}
inline uint16_t float_to_bf16(float x)
{
  uint32_t f;
  memcpy(&f, &x, 4);
  if ((f & 0x7fffffffu) > 0x7f800000u) return 0x7fffu;  // NaN
  uint32_t lsb = (f >> 16) & 1u;
  f += 0x7fffu + lsb;
  return (uint16_t)(f >> 16);
}

inline bool is_float_dt(int dt) { return dt == DT_FLOAT || dt == DT_HALF || dt == DT_DOUBLE || dt == DT_BF16; }
inline bool is_int_dt(int dt) { return dt == DT_INT || dt == DT_INT64 || dt == DT_INT16 || dt == DT_INT8; }
inline int dt_size(int dt)
{
  switch (dt) {
    case DT_FLOAT: case DT_INT: return 4;
    case DT_HALF: case DT_BF16: case DT_INT16: return 2;
    case DT_DOUBLE: case DT_INT64: return 8;
    case DT_INT8: return 1;
  }
  return -1;
}

// element conversion following cpp/src/wholememory_ops/functions/gather_scatter_func.cuh:150-197:
// half/bf16 go through float; everything else is a static_cast.
inline void convert_elt(const void* src, int sdt, void* dst, int ddt)
{
  if (is_float_dt(sdt)) {
    double d;
    float f;
    bool via_float = true;
    switch (sdt) {
      case DT_FLOAT: f = *(const float*)src; d = f; break;
      case DT_HALF: f = half_to_float(*(const uint16_t*)src); d = f; break;
      case DT_BF16: f = bf16_to_float(*(const uint16_t*)src); d = f; break;
      default: d = *(const double*)src; f = (float)d; via_float = false; break;
    }
    switch (ddt) {
      case DT_FLOAT: *(float*)dst = via_float ? f : (float)d; break;
      case DT_DOUBLE: *(double*)dst = d; break;
      case DT_HALF: *(uint16_t*)dst = float_to_half(via_float ? f : (float)d); break;
      case DT_BF16: *(uint16_t*)dst = float_to_bf16(via_float ? f : (float)d); break;
    }
  } else {
    int64_t v = 0;
    switch (sdt) {
      case DT_INT: v = *(const int32_t*)src; break;
      case DT_INT64: v = *(const int64_t*)src; break;
      case DT_INT16: v = *(const int16_t*)src; break;
      case DT_INT8: v = *(const int8_t*)src; break;
    }
    switch (ddt) {
      case DT_INT: *(int32_t*)dst = (int32_t)v; break;
      case DT_INT64: *(int64_t*)dst = v; break;
      case DT_INT16: *(int16_t*)dst = (int16_t)v; break;
      case DT_INT8: *(int8_t*)dst = (int8_t)v; break;
    }
  }
}

}  // namespace

extern "C" {

int wgo_num_threads()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// benchmarks size the OpenMP pool themselves: torchrun exports OMP_NUM_THREADS=1 into every rank's environment
void wgo_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// ---- RNG twins -------------------------------------------------------------------------------
// published pcg32 (initstate, initseq) stream, for the known-answer pin
void wgo_pcg32_reference_stream(uint64_t initstate, uint64_t initseq, uint32_t* out, int n)
{
  Pcg g;
  g.init(initstate, initseq, 0);
  for (int i = 0; i < n; i++)
    out[i] = g.next_u32();
}
// cpp/src/wholegraph_ops/raft_random_gen.cu:15-54 (generate_random_positive_int_cpu), int32 form
void wgo_generate_random_positive_int(int64_t random_seed, int64_t subsequence, int32_t* out, int64_t n)
{
  Pcg g;
  g.init_raft((uint64_t)random_seed, 0, (uint64_t)subsequence);
  for (int64_t i = 0; i < n; i++)
    out[i] = g.next_i32();
}
void wgo_generate_random_positive_int64(int64_t random_seed, int64_t subsequence, int64_t* out, int64_t n)
{
  Pcg g;
  g.init_raft((uint64_t)random_seed, 0, (uint64_t)subsequence);
  for (int64_t i = 0; i < n; i++)
    out[i] = g.next_i64();
}
// cpp/src/wholegraph_ops/raft_random_gen.cu:56-96
void wgo_generate_exponential_distribution_negative_float(int64_t random_seed, int64_t subsequence, float* out, int64_t n)
{
  Pcg g;
  g.init_raft((uint64_t)random_seed, 0, (uint64_t)subsequence);
  for (int64_t i = 0; i < n; i++) {
    float u     = g.next_float();
    u           = (float)(-(0.5 + 0.5 * (double)u));
    uint64_t r2 = 0;
    int extra   = -1;
    do {
      r2 = g.next_u64();
      extra++;
    } while (!r2);
    int one_bit = clz64(r2) + extra * 64;
    u           = (float)((double)u * pow(2.0, -one_bit));
    out[i]      = (float)(log1p((double)u) / log(2.0));
  }
}

// ---- S1/S2 sample offsets: graph_sampling_test_utils.cu:226-250 + prefix sum :180-200 ---------
int wgo_sample_offsets(const int64_t* row_ptr, const void* centers, int center_dtype, int64_t n,
                       int max_sample_count, int32_t* offsets /* n+1 */)
{
  int64_t acc = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t v   = read_index(centers, center_dtype, i);
    int64_t deg = row_ptr[v + 1] - row_ptr[v];
    if (max_sample_count > 0) deg = std::min<int64_t>(deg, max_sample_count);
    offsets[i] = (int32_t)acc;
    acc += deg;
  }
  offsets[n] = (int32_t)acc;
  return 0;
}

// ---- S1: uniform sampling without replacement -------------------------------------------------
// graph_sampling_test_utils.cu:252-293 (sample all), :312-401 (sampled), :404-487 (driver)
int wgo_unweighted_sample(const int64_t* row_ptr, const void* col, int col_dtype,
                          const void* centers, int center_dtype, int64_t n, int max_sample_count,
                          uint64_t seed, const int32_t* offsets, void* out_dest, int32_t* out_lid,
                          int64_t* out_gid)
{
  if (max_sample_count > 1024) return 2;  // no host reference exists (test_utils :468)
#pragma omp parallel
  {
    std::vector<int> Q;
    std::vector<int32_t> r;
#pragma omp for schedule(dynamic, 256)
    for (int64_t b = 0; b < n; b++) {
      int64_t v = read_index(centers, center_dtype, b);
      if (col_dtype == DT_INT)
        uniform_row<int32_t>(row_ptr, (const int32_t*)col, v, b, max_sample_count, seed,
                             offsets[b], (int32_t*)out_dest, out_lid, out_gid, Q, r);
      else
        uniform_row<int64_t>(row_ptr, (const int64_t*)col, v, b, max_sample_count, seed,
                             offsets[b], (int64_t*)out_dest, out_lid, out_gid, Q, r);
    }
  }
  return 0;
}

// ---- S2: weighted (A-Res) sampling without replacement ---------------------------------------
int wgo_weighted_sample(const int64_t* row_ptr, const void* col, int col_dtype, const void* wgt,
                        int wgt_dtype, const void* centers, int center_dtype, int64_t n,
                        int max_sample_count, uint64_t seed, const int32_t* offsets,
                        void* out_dest, int32_t* out_lid, int64_t* out_gid, float* out_key)
{
  if (max_sample_count > 1024) return 2;
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t b = 0; b < n; b++) {
    int64_t v = read_index(centers, center_dtype, b);
    if (col_dtype == DT_INT) {
      if (wgt_dtype == DT_FLOAT)
        weighted_row<int32_t, float>(row_ptr, (const int32_t*)col, (const float*)wgt, v, b, max_sample_count, seed, offsets[b], (int32_t*)out_dest, out_lid, out_gid, out_key);
      else
        weighted_row<int32_t, double>(row_ptr, (const int32_t*)col, (const double*)wgt, v, b, max_sample_count, seed, offsets[b], (int32_t*)out_dest, out_lid, out_gid, out_key);
    } else {
      if (wgt_dtype == DT_FLOAT)
        weighted_row<int64_t, float>(row_ptr, (const int64_t*)col, (const float*)wgt, v, b, max_sample_count, seed, offsets[b], (int64_t*)out_dest, out_lid, out_gid, out_key);
      else
        weighted_row<int64_t, double>(row_ptr, (const int64_t*)col, (const double*)wgt, v, b, max_sample_count, seed, offsets[b], (int64_t*)out_dest, out_lid, out_gid, out_key);
    }
  }
  return 0;
}

// keys of every edge of one row, for tolerance-aware set comparison in tests
int wgo_weighted_row_keys(const int64_t* row_ptr, const void* wgt, int wgt_dtype, int64_t v,
                          int64_t b, int max_sample_count, uint64_t seed, float* keys /* deg */)
{
  int64_t start = row_ptr[v], deg = row_ptr[v + 1] - row_ptr[v];
  int block = max_sample_count > 256 ? 256 : 128;
  for (int j = 0; j < block; j++) {
    Pcg rng;
    rng.init_raft(seed, 0, (uint64_t)(b * block + j));
    for (int64_t id = j; id < deg; id += block)
      keys[id] = wgt_dtype == DT_FLOAT ? gen_key_from_weight<float>(((const float*)wgt)[start + id], rng)
                                       : gen_key_from_weight<double>(((const double*)wgt)[start + id], rng);
  }
  return 0;
}

// ---- S3: append unique ------------------------------------------------------------------------
// cpp/tests/graph_ops/append_unique_test_utils.cu:53-119: targets keep ids 0..T-1 (first wins on
// duplicate targets), each new neighbour gets the next id in FIRST-OCCURRENCE order.
// unique_out must hold T + N entries; returns the unique count through *unique_count.
int wgo_append_unique(const void* targets, int64_t T, const void* neighbors, int64_t N, int dtype,
                      void* unique_out, int32_t* unique_count, int32_t* raw_to_unique /* N or null */)
{
  std::unordered_map<int64_t, int32_t> table;
  table.reserve((size_t)(T + N) * 2);
  for (int64_t i = 0; i < T; i++) {
    int64_t k = read_index(targets, dtype, i);
    table.insert(std::make_pair(k, (int32_t)i));
    write_index(unique_out, dtype, i, k);
  }
  int32_t cnt = (int32_t)T;
  for (int64_t i = 0; i < N; i++) {
    int64_t k = read_index(neighbors, dtype, i);
    auto it   = table.find(k);
    int32_t id;
    if (it == table.end()) {
      id = cnt++;
      table.insert(std::make_pair(k, id));
      write_index(unique_out, dtype, id, k);
    } else {
      id = it->second;
    }
    if (raw_to_unique) raw_to_unique[i] = id;
  }
  *unique_count = cnt;
  return 0;
}

// ---- G1/G2: gather / scatter -------------------------------------------------------------------
// semantics: cpp/src/wholememory_ops/functions/gather_scatter_func.cuh:243-305 (idx<0 skipped),
// 509-587; closed-form test tables cpp/tests/wholememory_ops/embedding_test_utils.cu:186-226.
// table element (row, c) lives at table[storage_offset + row*stride + c] (elements).
int wgo_gather(const void* table, int table_dtype, int64_t dim, int64_t stride, int64_t storage_offset,
               const void* indices, int idx_dtype, int64_t n, void* out, int out_dtype,
               int64_t out_stride, int64_t out_storage_offset)
{
  if (is_float_dt(table_dtype) != is_float_dt(out_dtype)) return 3;
  int ts = dt_size(table_dtype), os = dt_size(out_dtype);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    int64_t row = read_index(indices, idx_dtype, i);
    if (row < 0) continue;
    const char* src = (const char*)table + (storage_offset + row * stride) * ts;
    char* dst       = (char*)out + (out_storage_offset + i * out_stride) * os;
    if (table_dtype == out_dtype) {
      memcpy(dst, src, (size_t)dim * ts);
    } else {
      for (int64_t c = 0; c < dim; c++)
        convert_elt(src + c * ts, table_dtype, dst + c * os, out_dtype);
    }
  }
  return 0;
}

int wgo_scatter(const void* in, int in_dtype, int64_t dim, int64_t in_stride, int64_t in_storage_offset,
                const void* indices, int idx_dtype, int64_t n, void* table, int table_dtype,
                int64_t stride, int64_t storage_offset)
{
  if (is_float_dt(table_dtype) != is_float_dt(in_dtype)) return 3;
  int ts = dt_size(table_dtype), is = dt_size(in_dtype);
  // sequential on purpose: duplicate indices -> last writer wins deterministically in the oracle
  for (int64_t i = 0; i < n; i++) {
    int64_t row = read_index(indices, idx_dtype, i);
    if (row < 0) continue;
    char* dst       = (char*)table + (storage_offset + row * stride) * ts;
    const char* src = (const char*)in + (in_storage_offset + i * in_stride) * is;
    if (table_dtype == in_dtype) {
      memcpy(dst, src, (size_t)dim * ts);
    } else {
      for (int64_t c = 0; c < dim; c++)
        convert_elt(src + c * is, in_dtype, dst + c * ts, table_dtype);
    }
  }
  return 0;
}

// ---- A1: CSR mean/sum aggregation (PyG SAGEConv/GCN aggregation spec; fp64 accumulate) ----------
// out[i,:] = reduce_{e in [indptr[i], indptr[i+1])} x[indices[e],:]; mean over an empty row = 0
// (torch_geometric scatter-mean semantics; consumer python/cugraph-pyg/cugraph_pyg/examples/gcn_dist_mnmg.py:239).
int wgo_csr_aggregate(const int64_t* indptr, const int32_t* indices, int64_t n_dst, const float* x,
                      int64_t dim, int64_t x_stride, int mean, double* out /* n_dst*dim */)
{
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n_dst; i++) {
    double* o = out + i * dim;
    for (int64_t c = 0; c < dim; c++)
      o[c] = 0.0;
    int64_t s = indptr[i], e = indptr[i + 1];
    for (int64_t k = s; k < e; k++) {
      const float* xr = x + (int64_t)indices[k] * x_stride;
      for (int64_t c = 0; c < dim; c++)
        o[c] += (double)xr[c];
    }
    if (mean && e > s) {
      double inv = 1.0 / (double)(e - s);
      for (int64_t c = 0; c < dim; c++)
        o[c] *= inv;
    }
  }
  return 0;
}

// ---- S0: multi-hop, multi-label neighbour sampling with renumbering -----------------------------
// Composition (defined by this project, see DESIGN.md "S0"): per hop h the S1 (or S2) row sampler
// runs over the label-major concatenated frontier with seed hop_seed(h); per label the sampled
// neighbours are append-unique'd (first-occurrence order) onto the label's renumber map; the next
// frontier of a label is exactly the vertices that were new in this hop
// (deduplicate_sources=True, prior_sources_behavior="exclude", retain_seeds=True,
// python/cugraph-pyg/cugraph_pyg/sampler/distributed_sampler.py:784-793;
// multi-hop glue shape: python/pylibwholegraph/pylibwholegraph/torch/graph_structure.py:136-196;
// output contract: python/cugraph-pyg/cugraph_pyg/sampler/sampler.py:525-740).
//
// Outputs are returned through a handle so that callers can size their buffers.
struct wgo_multihop_result {
  std::vector<int32_t> majors, minors;       // renumbered local ids, label-major then hop
  std::vector<int64_t> edge_id;              // CSR position (or edge_ids[pos] if given)
  std::vector<int64_t> label_hop_offsets;    // B*L+1 edge offsets
  std::vector<int64_t> renumber_map;         // concatenated per label
  std::vector<int64_t> renumber_map_offsets; // B+1
};

uint64_t wgo_hop_seed(uint64_t random_state, int hop)
{
  return random_state + (uint64_t)hop * 0x9E3779B97F4A7C15ULL;
}

void* wgo_multihop_sample(const int64_t* row_ptr, const void* col, int col_dtype, const void* wgt,
                          int wgt_dtype, const int64_t* edge_ids, const int64_t* seeds,
                          const int64_t* label_offsets, int64_t num_labels, const int32_t* fanout,
                          int num_hops, uint64_t random_state)
{
  auto* res        = new wgo_multihop_result();
  const int64_t B  = num_labels;
  const int L      = num_hops;
  std::vector<std::vector<int64_t>> maps(B);  // per-label renumber maps
  std::vector<std::unordered_map<int64_t, int32_t>> tables(B);
  std::vector<int64_t> front_begin(B), front_end(B);  // local-id range of the label's frontier
  for (int64_t l = 0; l < B; l++) {
    for (int64_t s = label_offsets[l]; s < label_offsets[l + 1]; s++) {
      int64_t v = seeds[s];
      if (tables[l].find(v) == tables[l].end()) {
        tables[l].insert(std::make_pair(v, (int32_t)maps[l].size()));
        maps[l].push_back(v);
      }
    }
    front_begin[l] = 0;
    front_end[l]   = (int64_t)maps[l].size();
  }
  // per (label, hop) edge lists
  std::vector<std::vector<int32_t>> e_major(B * L), e_minor(B * L);
  std::vector<std::vector<int64_t>> e_id(B * L);
  for (int h = 0; h < L; h++) {
    // label-major concatenated frontier
    std::vector<int64_t> frontier;
    std::vector<int64_t> fr_off(B + 1, 0);
    for (int64_t l = 0; l < B; l++) {
      fr_off[l] = (int64_t)frontier.size();
      for (int64_t i = front_begin[l]; i < front_end[l]; i++)
        frontier.push_back(maps[l][i]);
    }
    fr_off[B]      = (int64_t)frontier.size();
    int64_t nf     = (int64_t)frontier.size();
    int M          = fanout[h];
    std::vector<int32_t> off(nf + 1, 0);
    std::vector<int64_t> dest;
    std::vector<int64_t> gid;
    if (M != 0 && nf > 0) {
      wgo_sample_offsets(row_ptr, frontier.data(), DT_INT64, nf, M, off.data());
      int64_t tot = off[nf];
      dest.resize(tot);
      gid.resize(tot);
      std::vector<int32_t> lid(tot);
      uint64_t hs = wgo_hop_seed(random_state, h);
      if (col_dtype == DT_INT64) {
        if (wgt)
          wgo_weighted_sample(row_ptr, col, col_dtype, wgt, wgt_dtype, frontier.data(), DT_INT64, nf, M, hs, off.data(), dest.data(), lid.data(), gid.data(), nullptr);
        else
          wgo_unweighted_sample(row_ptr, col, col_dtype, frontier.data(), DT_INT64, nf, M, hs, off.data(), dest.data(), lid.data(), gid.data());
      } else {
        std::vector<int32_t> d32(tot);
        if (wgt)
          wgo_weighted_sample(row_ptr, col, col_dtype, wgt, wgt_dtype, frontier.data(), DT_INT64, nf, M, hs, off.data(), d32.data(), lid.data(), gid.data(), nullptr);
        else
          wgo_unweighted_sample(row_ptr, col, col_dtype, frontier.data(), DT_INT64, nf, M, hs, off.data(), d32.data(), lid.data(), gid.data());
        for (int64_t i = 0; i < tot; i++)
          dest[i] = d32[i];
      }
    }
    for (int64_t l = 0; l < B; l++) {
      int64_t new_begin = (int64_t)maps[l].size();
      auto& mj = e_major[l * L + h];
      auto& mn = e_minor[l * L + h];
      auto& ei = e_id[l * L + h];
      for (int64_t f = fr_off[l]; f < fr_off[l + 1]; f++) {
        int32_t major_local = (int32_t)(front_begin[l] + (f - fr_off[l]));
        for (int32_t e = off[f]; e < off[f + 1]; e++) {
          int64_t v = dest[e];
          auto it   = tables[l].find(v);
          int32_t id;
          if (it == tables[l].end()) {
            id = (int32_t)maps[l].size();
            tables[l].insert(std::make_pair(v, id));
            maps[l].push_back(v);
          } else {
            id = it->second;
          }
          mj.push_back(major_local);
          mn.push_back(id);
          ei.push_back(edge_ids ? edge_ids[gid[e]] : gid[e]);
        }
      }
      front_begin[l] = new_begin;
      front_end[l]   = (int64_t)maps[l].size();
    }
  }
  res->label_hop_offsets.push_back(0);
  res->renumber_map_offsets.push_back(0);
  for (int64_t l = 0; l < B; l++) {
    for (int h = 0; h < L; h++) {
      auto& mj = e_major[l * L + h];
      res->majors.insert(res->majors.end(), mj.begin(), mj.end());
      res->minors.insert(res->minors.end(), e_minor[l * L + h].begin(), e_minor[l * L + h].end());
      res->edge_id.insert(res->edge_id.end(), e_id[l * L + h].begin(), e_id[l * L + h].end());
      res->label_hop_offsets.push_back((int64_t)res->majors.size());
    }
    res->renumber_map.insert(res->renumber_map.end(), maps[l].begin(), maps[l].end());
    res->renumber_map_offsets.push_back((int64_t)res->renumber_map.size());
  }
  return res;
}

int64_t wgo_multihop_num_edges(void* h) { return (int64_t)((wgo_multihop_result*)h)->majors.size(); }
int64_t wgo_multihop_num_nodes(void* h) { return (int64_t)((wgo_multihop_result*)h)->renumber_map.size(); }
void wgo_multihop_copy(void* h, int32_t* majors, int32_t* minors, int64_t* edge_id,
                       int64_t* label_hop_offsets, int64_t* renumber_map, int64_t* renumber_map_offsets)
{
  auto* r = (wgo_multihop_result*)h;
  if (majors) memcpy(majors, r->majors.data(), r->majors.size() * 4);
  if (minors) memcpy(minors, r->minors.data(), r->minors.size() * 4);
  if (edge_id) memcpy(edge_id, r->edge_id.data(), r->edge_id.size() * 8);
  if (label_hop_offsets) memcpy(label_hop_offsets, r->label_hop_offsets.data(), r->label_hop_offsets.size() * 8);
  if (renumber_map) memcpy(renumber_map, r->renumber_map.data(), r->renumber_map.size() * 8);
  if (renumber_map_offsets) memcpy(renumber_map_offsets, r->renumber_map_offsets.data(), r->renumber_map_offsets.size() * 8);
}
void wgo_multihop_free(void* h) { delete (wgo_multihop_result*)h; }


// ---- heterogeneous multi-hop sampling ---------------------------------------------------------------
// What cugraph-pyg asks of pylibcugraph.heterogeneous_{uniform,biased}_neighbor_sample
// (python/cugraph-pyg/cugraph_pyg/sampler/distributed_sampler.py:53-94, 784-819; fan-out vector laid out
// [hop * T + etype], loader/neighbor_loader.py:192-201) and what HeterogeneousSampleReader decodes
// (sampler/sampler.py:280-490); the only value-level pin in the reference is
// tests/sampler/test_distributed_sampler.py:19-150 (fan-out -1), which tests/test_oracle_cpu.py replays.
//
// Definition used by this project (libcugraph itself is not in the tree):
//   * vertices carry GLOBAL ids, vertex type vt owns [vtype_offsets[vt], vtype_offsets[vt+1]);
//     edge type t has its own CSR over the global id space (rows of other types are empty);
//   * hop h: the frontier is the label-major list of vertices new in the previous step (step 0: the
//     label's distinct seeds).  For every edge type t with fanout[h*T+t] != 0 the WHOLE frontier is sampled on
//     CSR_t with the one-hop algorithm (S1/S2), seed = type_seed(hop_seed(random_state, h), t);
//   * edge order inside a hop: frontier row, then edge type, then sample slot; vertices are deduplicated per
//     label in that order (first occurrence), over all types at once;
//   * local ids restart per (label, vertex type): the i-th vertex of type vt that label l discovered;
//   * outputs are grouped [label][edge type][hop]; edge_id is the position inside its (label, edge type)
//     group and edge_renumber_map holds the original edge ids in that order (edge[etype] = emap[edge_id],
//     sampler.py:334-341).
struct wgo_hetero_result {
  std::vector<int32_t> majors, minors, edge_type;
  std::vector<int64_t> edge_id;
  std::vector<int64_t> label_type_hop_offsets;     // B*T*L+1
  std::vector<int64_t> renumber_map;               // [label][vtype] segments, global ids
  std::vector<int64_t> renumber_map_offsets;       // B*Vt+1
  std::vector<int64_t> edge_renumber_map;          // original edge ids, output order
  std::vector<int64_t> edge_renumber_map_offsets;  // B*T+1
  std::vector<int32_t> step_base;                  // [(L+1)][Vt][B] first typed local id discovered at step s
};

uint64_t wgo_type_seed(uint64_t hop_seed, int etype) { return hop_seed + (uint64_t)etype * 0xD1B54A32D192ED03ULL; }

// comparison of an edge time with the time of the vertex it is sampled from (pylibcugraph temporal_sampling_comparison)
enum { WGO_T_STRICTLY_INCREASING = 0, WGO_T_MONOTONICALLY_INCREASING = 1, WGO_T_STRICTLY_DECREASING = 2, WGO_T_MONOTONICALLY_DECREASING = 3 };
static inline bool wgo_time_ok(int cmp, int64_t edge_time, int64_t vertex_time)
{
  switch (cmp) {
    case WGO_T_STRICTLY_INCREASING: return edge_time > vertex_time;
    case WGO_T_MONOTONICALLY_INCREASING: return edge_time >= vertex_time;
    case WGO_T_STRICTLY_DECREASING: return edge_time < vertex_time;
    default: return edge_time <= vertex_time;
  }
}

// Shared implementation.  edge_times == nullptr: plain heterogeneous sampling (above).  Otherwise TEMPORAL sampling
// (pylibcugraph.*_temporal_neighbor_sample; reference call site distributed_sampler.py:808-819, pins
// tests/loader/test_neighbor_loader.py:943-1058): a frontier vertex carries a time (seed: its starting time; otherwise
// the time of the edge that first reached it); only the edges of its row whose time compares as requested are eligible;
// the one-hop uniform algorithm (S1) runs over the ELIGIBLE positions of the row, in CSR order, with the same stream
// geometry.  Temporal + biased: the A-Res algorithm (S2) with its own geometry (thread j of `block` owns positions j,
// j + block, ... of the row) where only eligible positions draw a key and compete; rows with at most `fanout` eligible
// edges return all of them in CSR order.  Both reduce to the plain samplers when every edge is eligible.
static void* wgo_hetero_impl(int num_edge_types, const int64_t* const* row_ptr, const void* const* col, int col_dtype,
                             const void* const* wgt, int wgt_dtype, const int64_t* const* edge_ids,
                             const int64_t* vtype_offsets, int num_vertex_types, const int64_t* seeds,
                             const int64_t* label_offsets, int64_t num_labels, const int32_t* fanout, int num_hops,
                             uint64_t random_state, const int64_t* const* edge_times, const int64_t* seed_times, int time_cmp)
{
  auto* res       = new wgo_hetero_result();
  const int64_t B = num_labels;
  const int L = num_hops, T = num_edge_types, Vt = num_vertex_types;
  auto vtype_of = [&](int64_t v) {
    int vt = 0;
    while (vt + 1 < Vt && v >= vtype_offsets[vt + 1])
      vt++;
    return vt;
  };
  std::vector<std::vector<int64_t>> maps(B);  // discovery order, global ids
  std::vector<std::unordered_map<int64_t, int32_t>> tables(B);
  std::vector<std::vector<int64_t>> step_end(B, std::vector<int64_t>(L + 2, 0));  // discovery index at the end of step s
  std::vector<int64_t> front_begin(B), front_end(B);
  std::vector<std::vector<int64_t>> vtime(B);  // temporal: time of every discovered vertex (first arrival)
  for (int64_t l = 0; l < B; l++) {
    for (int64_t s = label_offsets[l]; s < label_offsets[l + 1]; s++) {
      int64_t v = seeds[s];
      if (tables[l].find(v) == tables[l].end()) {
        tables[l].insert(std::make_pair(v, (int32_t)maps[l].size()));
        maps[l].push_back(v);
        vtime[l].push_back(seed_times ? seed_times[s] : 0);
      }
    }
    front_begin[l] = 0;
    front_end[l]   = (int64_t)maps[l].size();
    step_end[l][1] = front_end[l];
  }
  struct Edge {
    int32_t major, minor;  // discovery indices
    int64_t id;
  };
  std::vector<std::vector<Edge>> edges((size_t)B * T * L);
  for (int h = 0; h < L; h++) {
    std::vector<int64_t> frontier, ftime;
    std::vector<int64_t> fr_off(B + 1, 0);
    for (int64_t l = 0; l < B; l++) {
      fr_off[l] = (int64_t)frontier.size();
      for (int64_t i = front_begin[l]; i < front_end[l]; i++) {
        frontier.push_back(maps[l][i]);
        ftime.push_back(vtime[l][i]);
      }
    }
    fr_off[B]  = (int64_t)frontier.size();
    int64_t nf = (int64_t)frontier.size();
    std::vector<std::vector<int32_t>> off(T, std::vector<int32_t>(nf + 1, 0));
    std::vector<std::vector<int64_t>> dest(T), gid(T);
    for (int t = 0; t < T; t++) {
      int M = fanout[h * T + t];
      if (M == 0 || nf == 0) continue;
      if (edge_times) {
        // temporal: S1 over the eligible positions of every row
        uint64_t hs = wgo_type_seed(wgo_hop_seed(random_state, h), t);
        std::vector<int> Q;
        std::vector<int32_t> r;
        for (int64_t f = 0; f < nf; f++) {
          off[t][f] = (int32_t)dest[t].size();
          int64_t v = frontier[f];
          std::vector<int64_t> elig;
          for (int64_t p = row_ptr[t][v]; p < row_ptr[t][v + 1]; p++)
            if (wgo_time_ok(time_cmp, edge_times[t][p], ftime[f])) elig.push_back(p);
          int N = (int)elig.size();
          auto take = [&](int64_t p) {
            dest[t].push_back(read_index(col[t], col_dtype, p));
            gid[t].push_back(p);
          };
          if (M < 0 || N <= M) {
            for (int64_t p : elig)
              take(p);
            continue;
          }
          if (wgt) {
            // S2 over the eligible positions (weighted_row above, masked); output in ascending key order like weighted_row
            const int block     = M > 256 ? 256 : 128;
            const int64_t start = row_ptr[t][v], deg = row_ptr[t][v + 1] - row_ptr[t][v];
            typedef std::pair<float, int64_t> KI;
            auto cmp = [](const KI& l, const KI& r) { return l.first > r.first; };
            std::priority_queue<KI, std::vector<KI>, decltype(cmp)> heap(cmp);
            int processed = 0;
            for (int j = 0; j < block; j++) {
              Pcg rng;
              rng.init_raft(hs, 0, (uint64_t)(f * block + j));
              for (int64_t id = j; id < deg; id += block) {
                if (!wgo_time_ok(time_cmp, edge_times[t][start + id], ftime[f])) continue;
                float key = wgt_dtype == DT_FLOAT ? gen_key_from_weight<float>(((const float*)wgt[t])[start + id], rng)
                                                  : gen_key_from_weight<double>(((const double*)wgt[t])[start + id], rng);
                processed++;
                if (processed <= M) {
                  heap.push(KI(key, start + id));
                } else if (heap.top().first < key) {
                  heap.pop();
                  heap.push(KI(key, start + id));
                }
              }
            }
            for (int i = 0; i < M; i++) {
              take(heap.top().second);
              heap.pop();
            }
            continue;
          }
          int func_idx = (M - 1) / 32;
          int TT       = kWarpCount[func_idx] * 32;
          int ipt      = kItems[func_idx];
          if ((int)r.size() < TT * ipt) r.resize(TT * ipt);
          for (int j = 0; j < TT; j++) {
            Pcg rng;
            rng.init_raft(hs, 0, (uint64_t)(f * TT + j));
            for (int k = 0; k < ipt; k++) {
              int id    = k * TT + j;
              int32_t x = rng.next_i32();
              r[id]     = id < M ? x % (N - id) : N;
            }
          }
          int a[1024];
          resolve_chain(r.data(), M, N, Q, a);
          for (int i = 0; i < M; i++)
            take(elig[a[i]]);
        }
        off[t][nf] = (int32_t)dest[t].size();
        continue;
      }
      wgo_sample_offsets(row_ptr[t], frontier.data(), DT_INT64, nf, M, off[t].data());
      int64_t tot = off[t][nf];
      dest[t].resize(tot);
      gid[t].resize(tot);
      std::vector<int32_t> lid(tot);
      uint64_t hs     = wgo_type_seed(wgo_hop_seed(random_state, h), t);
      const void* w_t = wgt ? wgt[t] : nullptr;
      if (col_dtype == DT_INT64) {
        if (w_t)
          wgo_weighted_sample(row_ptr[t], col[t], col_dtype, w_t, wgt_dtype, frontier.data(), DT_INT64, nf, M, hs, off[t].data(), dest[t].data(), lid.data(), gid[t].data(), nullptr);
        else
          wgo_unweighted_sample(row_ptr[t], col[t], col_dtype, frontier.data(), DT_INT64, nf, M, hs, off[t].data(), dest[t].data(), lid.data(), gid[t].data());
      } else {
        std::vector<int32_t> d32(tot);
        if (w_t)
          wgo_weighted_sample(row_ptr[t], col[t], col_dtype, w_t, wgt_dtype, frontier.data(), DT_INT64, nf, M, hs, off[t].data(), d32.data(), lid.data(), gid[t].data(), nullptr);
        else
          wgo_unweighted_sample(row_ptr[t], col[t], col_dtype, frontier.data(), DT_INT64, nf, M, hs, off[t].data(), d32.data(), lid.data(), gid[t].data());
        for (int64_t i = 0; i < tot; i++)
          dest[t][i] = d32[i];
      }
    }
    for (int64_t l = 0; l < B; l++) {
      int64_t new_begin = (int64_t)maps[l].size();
      for (int64_t f = fr_off[l]; f < fr_off[l + 1]; f++) {
        int32_t major = (int32_t)(front_begin[l] + (f - fr_off[l]));
        for (int t = 0; t < T; t++) {
          for (int32_t e = off[t][f]; e < off[t][f + 1]; e++) {
            int64_t v = dest[t][e];
            auto it   = tables[l].find(v);
            int32_t id;
            if (it == tables[l].end()) {
              id = (int32_t)maps[l].size();
              tables[l].insert(std::make_pair(v, id));
              maps[l].push_back(v);
              vtime[l].push_back(edge_times ? edge_times[t][gid[t][e]] : 0);
            } else {
              id = it->second;
            }
            int64_t g = gid[t][e];
            edges[((size_t)l * T + t) * L + h].push_back(Edge{major, id, (edge_ids && edge_ids[t]) ? edge_ids[t][g] : g});
          }
        }
      }
      front_begin[l]     = new_begin;
      front_end[l]       = (int64_t)maps[l].size();
      step_end[l][h + 2] = front_end[l];
    }
  }
  // typed local ids
  std::vector<std::vector<int32_t>> typed(B);
  res->step_base.assign((size_t)(L + 1) * Vt * B, 0);
  res->renumber_map_offsets.push_back(0);
  for (int64_t l = 0; l < B; l++) {
    std::vector<int32_t> counter(Vt, 0);
    std::vector<std::vector<int64_t>> per_type(Vt);
    typed[l].resize(maps[l].size());
    int step = 0;  // step_end[l][s] = number of vertices discovered before step s started
    for (size_t i = 0; i <= maps[l].size(); i++) {
      while (step <= L && (int64_t)i == step_end[l][step]) {
        for (int vt = 0; vt < Vt; vt++)
          res->step_base[((size_t)step * Vt + vt) * B + l] = counter[vt];
        step++;
      }
      if (i == maps[l].size()) break;
      int vt      = vtype_of(maps[l][i]);
      typed[l][i] = counter[vt]++;
      per_type[vt].push_back(maps[l][i]);
    }
    for (int vt = 0; vt < Vt; vt++) {
      res->renumber_map.insert(res->renumber_map.end(), per_type[vt].begin(), per_type[vt].end());
      res->renumber_map_offsets.push_back((int64_t)res->renumber_map.size());
    }
  }
  res->label_type_hop_offsets.push_back(0);
  res->edge_renumber_map_offsets.push_back(0);
  for (int64_t l = 0; l < B; l++) {
    for (int t = 0; t < T; t++) {
      int64_t in_group = 0;
      for (int h = 0; h < L; h++) {
        for (auto& e : edges[((size_t)l * T + t) * L + h]) {
          res->majors.push_back(typed[l][e.major]);
          res->minors.push_back(typed[l][e.minor]);
          res->edge_type.push_back(t);
          res->edge_id.push_back(in_group++);
          res->edge_renumber_map.push_back(e.id);
        }
        res->label_type_hop_offsets.push_back((int64_t)res->majors.size());
      }
      res->edge_renumber_map_offsets.push_back((int64_t)res->edge_renumber_map.size());
    }
  }
  return res;
}

void* wgo_hetero_multihop_sample(int num_edge_types, const int64_t* const* row_ptr, const void* const* col, int col_dtype,
                                 const void* const* wgt, int wgt_dtype, const int64_t* const* edge_ids,
                                 const int64_t* vtype_offsets, int num_vertex_types, const int64_t* seeds,
                                 const int64_t* label_offsets, int64_t num_labels, const int32_t* fanout, int num_hops,
                                 uint64_t random_state)
{
  return wgo_hetero_impl(num_edge_types, row_ptr, col, col_dtype, wgt, wgt_dtype, edge_ids, vtype_offsets, num_vertex_types, seeds,
                         label_offsets, num_labels, fanout, num_hops, random_state, nullptr, nullptr, 0);
}

void* wgo_temporal_multihop_sample(int num_edge_types, const int64_t* const* row_ptr, const void* const* col, int col_dtype,
                                   const int64_t* const* edge_ids, const int64_t* const* edge_times,
                                   const int64_t* vtype_offsets, int num_vertex_types, const int64_t* seeds,
                                   const int64_t* seed_times, const int64_t* label_offsets, int64_t num_labels,
                                   const int32_t* fanout, int num_hops, uint64_t random_state, int time_cmp)
{
  return wgo_hetero_impl(num_edge_types, row_ptr, col, col_dtype, nullptr, 0, edge_ids, vtype_offsets, num_vertex_types, seeds,
                         label_offsets, num_labels, fanout, num_hops, random_state, edge_times, seed_times, time_cmp);
}

void* wgo_temporal_biased_multihop_sample(int num_edge_types, const int64_t* const* row_ptr, const void* const* col, int col_dtype,
                                          const void* const* wgt, int wgt_dtype, const int64_t* const* edge_ids,
                                          const int64_t* const* edge_times, const int64_t* vtype_offsets, int num_vertex_types,
                                          const int64_t* seeds, const int64_t* seed_times, const int64_t* label_offsets,
                                          int64_t num_labels, const int32_t* fanout, int num_hops, uint64_t random_state, int time_cmp)
{
  return wgo_hetero_impl(num_edge_types, row_ptr, col, col_dtype, wgt, wgt_dtype, edge_ids, vtype_offsets, num_vertex_types, seeds,
                         label_offsets, num_labels, fanout, num_hops, random_state, edge_times, seed_times, time_cmp);
}

int64_t wgo_hetero_num_edges(void* h) { return (int64_t)((wgo_hetero_result*)h)->majors.size(); }
int64_t wgo_hetero_num_nodes(void* h) { return (int64_t)((wgo_hetero_result*)h)->renumber_map.size(); }
void wgo_hetero_copy(void* h, int32_t* majors, int32_t* minors, int32_t* edge_type, int64_t* edge_id, int64_t* lto,
                     int64_t* renumber_map, int64_t* rmo, int64_t* edge_renumber_map, int64_t* ermo, int32_t* step_base)
{
  auto* r = (wgo_hetero_result*)h;
  memcpy(majors, r->majors.data(), r->majors.size() * 4);
  memcpy(minors, r->minors.data(), r->minors.size() * 4);
  memcpy(edge_type, r->edge_type.data(), r->edge_type.size() * 4);
  memcpy(edge_id, r->edge_id.data(), r->edge_id.size() * 8);
  memcpy(lto, r->label_type_hop_offsets.data(), r->label_type_hop_offsets.size() * 8);
  memcpy(renumber_map, r->renumber_map.data(), r->renumber_map.size() * 8);
  memcpy(rmo, r->renumber_map_offsets.data(), r->renumber_map_offsets.size() * 8);
  memcpy(edge_renumber_map, r->edge_renumber_map.data(), r->edge_renumber_map.size() * 8);
  memcpy(ermo, r->edge_renumber_map_offsets.data(), r->edge_renumber_map_offsets.size() * 8);
  memcpy(step_base, r->step_base.data(), r->step_base.size() * 4);
}
void wgo_hetero_free(void* h) { delete (wgo_hetero_result*)h; }


// ---- sparse optimizers of trainable embeddings ------------------------------------------------------
// wholememory_embedding_gather_gradient_apply (cpp/src/wholememory/embedding.cpp:136-315): gradients of repeated
// indices are summed (dedup_indice_and_gradients), then ONE optimizer step per distinct row with the element rules of
// cpp/src/wholememory_ops/functions/embedding_optimizer_func.cu:
//   SGD      :205-213   g += wd*w; w -= lr*g
//   LazyAdam :389-420   b1t*=beta1; b2t*=beta2 (per row, start 1); adam_w ? w -= lr*wd*w : g += wd*w;
//                        m = b1*m+(1-b1)*g; v = b2*v+(1-b2)*g*g; w -= lr*(m/(1-b1t))/(sqrt(v/(1-b2t))+eps)
//   AdaGrad  :655-668   g += wd*w; s += g*g; w -= lr*g/(sqrt(s)+eps)
//   RMSProp  :865-880   g += wd*w; v = alpha*v+(1-alpha)*g*g; w -= lr*g/(sqrt(v)+eps)
// optimizer_type: 1 SGD, 2 LazyAdam, 3 RMSProp, 4 AdaGrad (wholememory_optimizer_type_t, embedding.h:45-58).
// Duplicates are summed in input order, in fp32, like the device path sums them in publish order.
int wgo_embedding_gradient_apply(int optimizer_type, float weight_decay, float epsilon, float beta1, float beta2,
                                 int adam_w, float alpha, float lr, float* emb, int64_t dim, const int64_t* indices,
                                 int64_t n, const float* grads, float* state_a, float* state_b, float* per_row)
{
  std::unordered_map<int64_t, std::vector<float>> sum;
  std::vector<int64_t> order;
  for (int64_t i = 0; i < n; i++) {
    auto it = sum.find(indices[i]);
    if (it == sum.end()) {
      it = sum.insert(std::make_pair(indices[i], std::vector<float>(dim, 0.f))).first;
      order.push_back(indices[i]);
    }
    for (int64_t d = 0; d < dim; d++)
      it->second[d] += grads[i * dim + d];
  }
  for (int64_t row : order) {
    const std::vector<float>& gs = sum[row];
    float b1t = 1.f, b2t = 1.f;
    if (optimizer_type == 2) {
      b1t = per_row[row * 2] * beta1;
      b2t = per_row[row * 2 + 1] * beta2;
    }
    for (int64_t d = 0; d < dim; d++) {
      float g = gs[d];
      float x = emb[row * dim + d];
      int64_t e = row * dim + d;
      if (optimizer_type == 1) {
        g += weight_decay * x;
        x -= lr * g;
      } else if (optimizer_type == 2) {
        if (adam_w) x -= lr * weight_decay * x;
        else g = g + weight_decay * x;
        float m = state_a[e], v = state_b[e];
        m = beta1 * m + (1 - beta1) * g;
        v = beta2 * v + (1 - beta2) * g * g;
        float mhat = m / (1 - b1t);
        float vhat = v / (1 - b2t);
        x = x - lr * mhat / (sqrtf(vhat) + epsilon);
        state_a[e] = m;
        state_b[e] = v;
      } else if (optimizer_type == 4) {
        g = g + weight_decay * x;
        float s2 = state_a[e] + g * g;
        x = x - lr * g / (sqrtf(s2) + epsilon);
        state_a[e] = s2;
      } else if (optimizer_type == 3) {
        g = g + weight_decay * x;
        float v = alpha * state_a[e] + (1 - alpha) * g * g;
        x = x - lr * g / (sqrtf(v) + epsilon);
        state_a[e] = v;
      } else {
        return 1;
      }
      emb[e] = x;
    }
    if (optimizer_type == 2) {
      per_row[row * 2]     = b1t;
      per_row[row * 2 + 1] = b2t;
    }
  }
  return 0;
}

}  // extern "C"
