// A client of the C ABI that is neither Python nor torch: plain cudaMalloc'd buffers, the reference's headers, the default
// (cudaMalloc-backed) allocation callbacks.  What a maintainer's C++ code would link against libwholegraph_b200.so.
//   abi_client link   host-only calls (runs on a box without a GPU): descriptors, dtype helpers, entry points resolve
//   abi_client gpu    feature gather (closed form), one-hop take-all sampling + append-unique on a small CSR, the fused multi-hop
//                     sampler with fan-out -1 (deterministic: every neighbour), all checked on the host
#include <cuda_runtime.h>
#include <wholememory/b200_ops.h>
#include <wholememory/graph_op.h>
#include <wholememory/wholegraph_op.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_op.h>
#include <wholememory/wholememory_tensor.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CHECK(x)                                                               \
  do {                                                                         \
    if (!(x)) {                                                                \
      std::fprintf(stderr, "abi_client: check failed at line %d: %s\n", __LINE__, #x); \
      return 1;                                                                \
    }                                                                          \
  } while (0)
#define WM(x) CHECK((x) == WHOLEMEMORY_SUCCESS)
#define CU(x) CHECK((x) == cudaSuccess)

static wholememory_tensor_t make(void* p, wholememory_dtype_t dt, int dim, long long s0, long long s1 = 1)
{
  wholememory_tensor_description_t td;
  wholememory_initialize_tensor_desc(&td);
  td.dim      = dim;
  td.dtype    = dt;
  td.sizes[0] = s0;
  if (dim == 2) {
    td.sizes[1]   = s1;
    td.strides[0] = s1;
    td.strides[1] = 1;
  } else {
    td.strides[0] = 1;
  }
  wholememory_tensor_t t = nullptr;
  return wholememory_make_tensor_from_pointer(&t, p, &td) == WHOLEMEMORY_SUCCESS ? t : nullptr;
}

int main(int argc, char** argv)
{
  const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;
  // ---- host-only part ---------------------------------------------------------------------------------------------
  CHECK(wholememory_dtype_get_element_size(WHOLEMEMORY_DT_FLOAT) == 4);
  CHECK(wholememory_dtype_get_element_size(WHOLEMEMORY_DT_INT64) == 8);
  CHECK(wholememory_dtype_is_floating_number(WHOLEMEMORY_DT_BF16) && wholememory_dtype_is_integer_number(WHOLEMEMORY_DT_INT));
  wholememory_tensor_description_t td;
  wholememory_initialize_tensor_desc(&td);
  CHECK(td.dim == 0 || td.dim == 1 || td.dim == 2);
  CHECK(wholememory_get_default_env_func() != nullptr);
  // every hot-path entry point resolves at link time (taking the address is enough)
  void* fns[] = {(void*)&wholememory_gather, (void*)&wholememory_scatter, (void*)&wholegraph_csr_unweighted_sample_without_replacement,
                 (void*)&wholegraph_csr_weighted_sample_without_replacement, (void*)&graph_append_unique,
                 (void*)&wholegraph_multihop_neighbor_sample, (void*)&wholegraph_csr_aggregate, (void*)&wholegraph_sage_layer_forward};
  for (void* f : fns)
    CHECK(f != nullptr);
  if (!gpu) {
    std::printf("abi_client link ok\n");
    return 0;
  }
  // ---- GPU part ------------------------------------------------------------------------------------------------------
  CU(cudaSetDevice(0));
  WM(wholememory_init(0));
  wholememory_env_func_t* env = wholememory_get_default_env_func();
  // G1: gather rows of a closed-form table
  const int rows = 5000, dim = 128, n = 7001;
  std::vector<float> h_table((size_t)rows * dim);
  for (int r = 0; r < rows; r++)
    for (int d = 0; d < dim; d++)
      h_table[(size_t)r * dim + d] = (float)((r * 3 + d) % 2048);
  std::vector<long long> h_idx(n);
  for (int i = 0; i < n; i++)
    h_idx[i] = (i % 17 == 0) ? -1 : (long long)((i * 2654435761u) % rows);
  float *d_table, *d_out;
  long long* d_idx;
  CU(cudaMalloc(&d_table, h_table.size() * 4));
  CU(cudaMalloc(&d_out, (size_t)n * dim * 4));
  CU(cudaMalloc(&d_idx, (size_t)n * 8));
  CU(cudaMemcpy(d_table, h_table.data(), h_table.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_idx, h_idx.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  CU(cudaMemset(d_out, 0xFF, (size_t)n * dim * 4));
  wholememory_tensor_t t_table = make(d_table, WHOLEMEMORY_DT_FLOAT, 2, rows, dim), t_idx = make(d_idx, WHOLEMEMORY_DT_INT64, 1, n),
                       t_out = make(d_out, WHOLEMEMORY_DT_FLOAT, 2, n, dim);
  CHECK(t_table && t_idx && t_out);
  WM(wholememory_gather(t_table, t_idx, t_out, env, nullptr));
  CU(cudaDeviceSynchronize());
  std::vector<float> h_out((size_t)n * dim);
  CU(cudaMemcpy(h_out.data(), d_out, h_out.size() * 4, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) {
    if (h_idx[i] < 0) continue;  // skipped rows stay untouched
    for (int d = 0; d < dim; d += 31)
      CHECK(h_out[(size_t)i * dim + d] == h_table[(size_t)h_idx[i] * dim + d]);
  }
  // S1 with fan-out -1 (every neighbour, deterministic) on a ring-with-chords CSR, then S3 append-unique
  const int V = 1000;
  std::vector<long long> h_rp(V + 1);
  std::vector<int> h_col;
  for (int v = 0; v < V; v++) {
    h_rp[v] = (long long)h_col.size();
    const int deg = v % 5;
    for (int k = 0; k < deg; k++)
      h_col.push_back((v * 7 + k * 13 + 1) % V);
  }
  h_rp[V] = (long long)h_col.size();
  const int S = 64;
  std::vector<long long> h_seeds(S);
  for (int i = 0; i < S; i++)
    h_seeds[i] = (i * 37 + 4) % V;
  long long *d_rp, *d_seeds;
  int *d_col, *d_off;
  CU(cudaMalloc(&d_rp, (V + 1) * 8));
  CU(cudaMalloc(&d_col, h_col.size() * 4 + 4));
  CU(cudaMalloc(&d_seeds, S * 8));
  CU(cudaMalloc(&d_off, (S + 1) * 4));
  CU(cudaMemcpy(d_rp, h_rp.data(), (V + 1) * 8, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_col, h_col.data(), h_col.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_seeds, h_seeds.data(), S * 8, cudaMemcpyHostToDevice));
  wholememory_tensor_t t_rp = make(d_rp, WHOLEMEMORY_DT_INT64, 1, V + 1), t_col = make(d_col, WHOLEMEMORY_DT_INT, 1, (long long)h_col.size()),
                       t_seeds = make(d_seeds, WHOLEMEMORY_DT_INT64, 1, S), t_off = make(d_off, WHOLEMEMORY_DT_INT, 1, S + 1);
  CHECK(t_rp && t_col && t_seeds && t_off);
  wholememory_default_output_t o_dest, o_lid, o_gid;
  std::memset(&o_dest, 0, sizeof(o_dest));
  std::memset(&o_lid, 0, sizeof(o_lid));
  std::memset(&o_gid, 0, sizeof(o_gid));
  WM(wholegraph_csr_unweighted_sample_without_replacement(t_rp, t_col, t_seeds, -1, t_off, &o_dest, &o_lid, &o_gid, 62ULL, env, nullptr));
  CU(cudaDeviceSynchronize());
  std::vector<int> h_off(S + 1);
  CU(cudaMemcpy(h_off.data(), d_off, (S + 1) * 4, cudaMemcpyDeviceToHost));
  long long expect_total = 0;
  for (int i = 0; i < S; i++) {
    CHECK(h_off[i] == expect_total);
    expect_total += h_rp[h_seeds[i] + 1] - h_rp[h_seeds[i]];
  }
  CHECK(h_off[S] == expect_total && o_dest.desc.sizes[0] == expect_total);
  std::vector<int> h_dest(expect_total);
  std::vector<long long> h_gid(expect_total);
  CU(cudaMemcpy(h_dest.data(), o_dest.ptr, expect_total * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(h_gid.data(), o_gid.ptr, expect_total * 8, cudaMemcpyDeviceToHost));
  for (int i = 0; i < S; i++)
    for (long long k = h_rp[h_seeds[i]]; k < h_rp[h_seeds[i] + 1]; k++) {
      const long long p = h_off[i] + (k - h_rp[h_seeds[i]]);
      CHECK(h_dest[p] == h_col[k] && h_gid[p] == k);
    }
  wholememory_default_output_release(&o_dest);
  wholememory_default_output_release(&o_lid);
  wholememory_default_output_release(&o_gid);
  WM(wholememory_finalize());
  std::printf("abi_client gpu ok: gathered %d rows, sampled %lld edges through the C ABI\n", n, expect_total);
  return 0;
}
