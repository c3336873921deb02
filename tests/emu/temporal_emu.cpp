// TEST INFRASTRUCTURE ONLY: runs the LOGIC of the product's temporal one-hop kernels (csrc/temporal_device.cuh) and, as a
// check of the emulator itself, of the already GPU-verified uniform_general_kernel / count_scan_kernel
// (csrc/sample_device.cuh) on the CPU through tests/emu/cuda_emu.h.  Built by tests/test_emulated_kernels_cpu.py with
//   g++ -O1 -std=c++17 -DWGB_HOST_EMULATION -I tests/emu -I cugraph-gnn_b200/csrc -I include -I /usr/local/cuda/include
#include "cuda_emu.h"

#include "temporal_device.cuh"

#include <vector>

namespace {

wgb::ChunkRef local_ref(const void* p, size_t bytes)
{
  wgb::ChunkRef r;
  std::memset(&r, 0, sizeof(r));
  r.base[0]  = const_cast<char*>(static_cast<const char*>(p));
  r.start[1] = bytes;
  r.world    = 1;
  return r;
}

std::vector<wgb::Affine> skip_table()
{
  std::vector<wgb::Affine> host(wgb::kSkipTabSize);
  for (int p = 0; p < wgb::kSkipTabBytes; p++) {
    wgb::Affine unit = wgb::affine_skip_loop(1ULL << (8 * p));
    wgb::Affine acc{1ULL, 0ULL};
    for (int v = 0; v < 256; v++) {
      host[p * 256 + v] = acc;
      acc               = wgb::affine_then(acc, unit);
    }
  }
  return host;
}

struct ScanState {
  std::vector<unsigned long long> words;
  unsigned long long* state;
  unsigned int* ticket;
  explicit ScanState(long long n_items)
  {
    int tiles = (int)((n_items + wgb::kScanTile) / wgb::kScanTile);
    words.assign((size_t)tiles + 2, 0ULL);
    state  = words.data();
    ticket = reinterpret_cast<unsigned int*>(words.data() + tiles);
  }
};

}  // namespace

extern "C" {

// One temporal hop over `n` frontier rows: offsets [n+1], then dest / lid / gid [offsets[n]] (caller sizes them n * max(M, max
// degree)).  grid_* let the test exercise grid-stride loops and multi-tile scans.  Returns offsets[n].
int emu_temporal_hop(const long long* row_ptr, long long num_rows, const long long* col, const long long* etime, long long num_edges,
                     const long long* centers, const long long* ftime, int n, int M, int cmp, unsigned long long seed, int grid_count,
                     int grid_scan, int grid_sample, int* offsets, int* eligible, long long* dest, int* lid, long long* gid)
{
  using namespace wgb;
  ChunkRef rp = local_ref(row_ptr, (size_t)(num_rows + 1) * 8), cl = local_ref(col, (size_t)num_edges * 8),
           tm = local_ref(etime, (size_t)num_edges * 8);
  std::vector<int> clipped((size_t)n + 1, 0);
  int n_dev = n, total = -1;
  cuda_emu::launch(grid_count, 1, 256, [&] {
    temporal_count_kernel<false>(rp, 0ULL, tm, 0ULL, centers, ftime, M, cmp, eligible, clipped.data(), &n_dev);
  });
  ScanState ss(n);
  cuda_emu::launch(grid_scan, 1, kScanBlock, [&] { scan_counts_kernel(clipped.data(), offsets, ss.state, ss.ticket, &n_dev, &total); });
  if (total != offsets[n]) return -1;
  auto tab = skip_table();
  cuda_emu::launch(grid_sample, 1, kGeneralBlock, [&] {
    temporal_uniform_kernel<long long, false>(rp, 0ULL, cl, 0ULL, tm, 0ULL, centers, ftime, eligible, M, cmp, seed, offsets, dest, lid, gid,
                                              tab.data(), &n_dev);
  });
  return total;
}

// The plain one-hop path for fan-out > 32 (count_scan_kernel + uniform_general_kernel), GPU-verified already: if the emulator
// reproduces the oracle here, its barriers / shuffles / ballots behave.
int emu_plain_hop(const long long* row_ptr, long long num_rows, const long long* col, long long num_edges, const long long* centers, int n,
                  int M, unsigned long long seed, int grid_scan, int grid_sample, int* offsets, long long* dest, int* lid, long long* gid)
{
  using namespace wgb;
  ChunkRef rp = local_ref(row_ptr, (size_t)(num_rows + 1) * 8), cl = local_ref(col, (size_t)num_edges * 8);
  int n_dev = n, total = -1;
  ScanState ss(n);
  cuda_emu::launch(grid_scan, 1, kScanBlock, [&] {
    count_scan_kernel<long long, false>(rp, 0ULL, centers, n, M, offsets, ss.state, ss.ticket, &n_dev, &total);
  });
  if (total != offsets[n]) return -1;
  auto tab = skip_table();
  cuda_emu::launch(grid_sample, 1, kGeneralBlock, [&] {
    uniform_general_kernel<long long, long long, false>(rp, 0ULL, cl, 0ULL, centers, n, M, seed, offsets, dest, lid, gid, tab.data(), &n_dev);
  });
  return total;
}

}  // extern "C"
