// TEST INFRASTRUCTURE ONLY (never part of the product, never linked into libwholegraph_b200.so).
//
// A tiny SIMT emulator that lets g++ compile the product's device kernels (csrc/*.cuh) and run their LOGIC on the CPU:
// one thread block at a time, every CUDA thread a cooperative fiber (ucontext) on ONE OS thread, __syncthreads() and the
// warp intrinsics as fiber barriers (deterministic; a barrier that cannot complete is reported as a deadlock).
// It exists because a round can run out of GPU minutes before a newly written kernel has been on hardware: the kernel's
// control flow, indexing and barrier placement can still be checked against the oracle here.  It says nothing about
// memory ordering, occupancy or speed; the GPU parity tests stay the gate.
//
// Supported: threadIdx/blockIdx/blockDim/gridDim (.x/.y), __shared__ (function-scope static: one block runs at a time),
// __syncthreads, __syncwarp, __ballot_sync, __any_sync, __shfl_sync/_up/_xor (int, unsigned, long long, unsigned long
// long), __match_any_sync, atomicAdd/atomicMax/atomicMin, __ldg, __popc, __clz, __ffs, __clzll, __float_as_uint.
// A warp intrinsic is a rendezvous of exactly the lanes named in its mask (one barrier per distinct mask and warp), so
// sub-warp groups that diverge from each other around partial-mask intrinsics (uniform_small_kernel) work; the lanes of
// one mask must all arrive, as on the hardware.
// Host side: the CUDA runtime calls the product's host code makes are defined in emu_runtime.cpp (malloc / memcpy /
// no-op streams and events), kernel launches are rewritten by emu_preprocess.py into launch_dim() below.
#pragma once

#include <cuda_runtime.h>  // types (dim3, uint3) and the host-side definitions of __host__/__device__ (empty under g++)

#include <ucontext.h>
#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>  // fiber switch annotations: the emulator can run under AddressSanitizer
#endif

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __global__
#define __global__
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __forceinline__
#define __forceinline__ inline

namespace cuda_emu {

struct Idx {
  unsigned int x = 0, y = 0, z = 0;
};

// A barrier for cooperative fibers: the last arrival opens the next generation, the others yield until it has.
struct Bar {
  unsigned int count   = 0;
  unsigned int arrived = 0;
  unsigned long gen    = 0;
};

struct Warp {
  unsigned int lanes = 32;  // threads of the block that live in this warp
  unsigned long long v64[32];
  std::map<unsigned int, Bar> bars;  // one per mask in use
};

struct Fiber {
  ucontext_t ctx;
  Idx idx;
  bool done = false;
};

struct Block {
  Bar bar;
  std::unique_ptr<Warp[]> warps;
  std::vector<Fiber> fibers;
  ucontext_t sched;
  unsigned int cur       = 0;
  unsigned long progress = 0;  // barrier releases + fiber exits: no change over a whole sweep = deadlock
};

inline Idx g_blockIdx, g_blockDim, g_gridDim;
inline Block* g_block                     = nullptr;
inline const std::function<void()>* g_body = nullptr;
inline std::vector<char*> g_stacks;  // reused across launches
constexpr size_t kStackBytes = 256 << 10;

inline Idx& cur_idx() { return g_block->fibers[g_block->cur].idx; }
inline Warp& my_warp() { return g_block->warps[cur_idx().x >> 5]; }
inline int my_lane() { return (int)(cur_idx().x & 31u); }

// AddressSanitizer has to be told about every stack switch (no-ops otherwise)
inline const void* g_sched_stack_bottom = nullptr;
inline size_t g_sched_stack_size        = 0;
inline void switch_begin(void** fake, const void* bottom, size_t size)
{
#if defined(__SANITIZE_ADDRESS__)
  __sanitizer_start_switch_fiber(fake, bottom, size);
#else
  (void)fake; (void)bottom; (void)size;
#endif
}
inline void switch_end(void* fake, const void** old_bottom, size_t* old_size)
{
#if defined(__SANITIZE_ADDRESS__)
  __sanitizer_finish_switch_fiber(fake, old_bottom, old_size);
#else
  (void)fake; (void)old_bottom; (void)old_size;
#endif
}

inline void yield()
{
  Block* b   = g_block;
  void* fake = nullptr;
  switch_begin(&fake, g_sched_stack_bottom, g_sched_stack_size);
  swapcontext(&b->fibers[b->cur].ctx, &b->sched);
  switch_end(fake, nullptr, nullptr);
}

inline void wait(Bar& bar)
{
  const unsigned long gen = bar.gen;
  if (++bar.arrived == bar.count) {
    bar.arrived = 0;
    bar.gen++;
    g_block->progress++;
  } else {
    while (bar.gen == gen)
      yield();
  }
}

inline void warp_rendezvous(unsigned int mask)
{
  Warp& w = my_warp();
  if (w.lanes < 32) mask &= (1u << w.lanes) - 1u;
  Bar& bar = w.bars[mask];
  if (bar.count == 0) bar.count = (unsigned int)__builtin_popcount(mask);
  wait(bar);
}

// every lane of `mask` publishes `v`, then reads what it needs through `read(all 32 values)` (only the slots of the
// lanes in `mask` are meaningful)
template <typename F>
inline auto exchange(unsigned int mask, unsigned long long v, F&& read)
{
  Warp& w = my_warp();
  w.v64[my_lane()] = v;
  warp_rendezvous(mask);
  auto r = read(w.v64);
  warp_rendezvous(mask);
  return r;
}

inline void fiber_entry()
{
  switch_end(nullptr, &g_sched_stack_bottom, &g_sched_stack_size);  // first entry: learn the scheduler's stack
  (*g_body)();
  Block* b                = g_block;
  b->fibers[b->cur].done  = true;
  b->progress++;
  switch_begin(nullptr, g_sched_stack_bottom, g_sched_stack_size);  // nullptr: this fiber's stack is not coming back
  // returning switches to uc_link = the scheduler
}

// run `body` for every thread of every block of a grid: blocks one after the other, the threads of a block as fibers
// scheduled round-robin on this OS thread (a fiber runs until it waits at a barrier or returns)
inline void launch(unsigned int grid_x, unsigned int grid_y, unsigned int block_x, const std::function<void()>& body)
{
  g_gridDim.x = grid_x; g_gridDim.y = grid_y; g_gridDim.z = 1;
  g_blockDim.x = block_x; g_blockDim.y = 1; g_blockDim.z = 1;
  const unsigned int nwarps = (block_x + 31) / 32;
  while (g_stacks.size() < block_x)
    g_stacks.push_back(static_cast<char*>(std::malloc(kStackBytes)));
  g_body = &body;
  for (unsigned int by = 0; by < grid_y; by++)
    for (unsigned int bx = 0; bx < grid_x; bx++) {
      Block blk;
      blk.warps.reset(new Warp[nwarps]);
      blk.bar.count = block_x;
      for (unsigned int w = 0; w < nwarps; w++)
        blk.warps[w].lanes = std::min(32u, block_x - 32u * w);
      blk.fibers.resize(block_x);
      g_block      = &blk;
      g_blockIdx.x = bx; g_blockIdx.y = by; g_blockIdx.z = 0;
      for (unsigned int t = 0; t < block_x; t++) {
        Fiber& f = blk.fibers[t];
        f.idx.x  = t;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp   = g_stacks[t];
        f.ctx.uc_stack.ss_size = kStackBytes;
        f.ctx.uc_link          = &blk.sched;
        makecontext(&f.ctx, fiber_entry, 0);
      }
      unsigned int remaining = block_x;
      while (remaining) {
        const unsigned long before = blk.progress;
        for (unsigned int t = 0; t < block_x; t++) {
          Fiber& f = blk.fibers[t];
          if (f.done) continue;
          blk.cur    = t;
          void* fake = nullptr;
          switch_begin(&fake, g_stacks[t], kStackBytes);
          swapcontext(&blk.sched, &f.ctx);
          switch_end(fake, nullptr, nullptr);
          if (f.done) remaining--;
        }
        if (remaining && blk.progress == before) {
          std::fprintf(stderr, "cuda_emu: deadlock in block (%u, %u): %u thread(s) wait at barriers that cannot complete\n", bx, by, remaining);
          std::abort();
        }
      }
      g_block = nullptr;
    }
  g_body = nullptr;
}

inline void launch_dim(dim3 grid, dim3 block, const std::function<void()>& body) { launch(grid.x, grid.y, block.x, body); }

}  // namespace cuda_emu

#define threadIdx (::cuda_emu::cur_idx())
#define blockIdx (::cuda_emu::g_blockIdx)
#define blockDim (::cuda_emu::g_blockDim)
#define gridDim (::cuda_emu::g_gridDim)

inline void __syncthreads() { ::cuda_emu::wait(::cuda_emu::g_block->bar); }
inline void __syncwarp(unsigned int mask = 0xffffffffu) { ::cuda_emu::warp_rendezvous(mask); }

inline unsigned int __ballot_sync(unsigned int mask, bool pred)
{
  return ::cuda_emu::exchange(mask, pred ? 1ULL : 0ULL, [mask](const unsigned long long* v) {
    unsigned int r = 0;
    for (int i = 0; i < 32; i++)
      if (((mask >> i) & 1u) && v[i]) r |= 1u << i;
    return r;
  });
}
inline bool __any_sync(unsigned int mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }

template <typename T>
inline T __shfl_sync(unsigned int mask, T val, int src, int width = 32)
{
  static_assert(sizeof(T) <= 8, "emulated shuffles move up to 64 bits");
  unsigned long long raw = 0;
  std::memcpy(&raw, &val, sizeof(T));
  const int lane = ::cuda_emu::my_lane();
  const int base = lane & ~(width - 1);
  unsigned long long got = ::cuda_emu::exchange(mask, raw, [=](const unsigned long long* v) { return v[base + (src & (width - 1))]; });
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <typename T>
inline T __shfl_up_sync(unsigned int mask, T val, unsigned int delta, int width = 32)
{
  unsigned long long raw = 0;
  std::memcpy(&raw, &val, sizeof(T));
  const int lane = ::cuda_emu::my_lane();
  const int base = lane & ~(width - 1);
  unsigned long long got = ::cuda_emu::exchange(mask, raw, [=](const unsigned long long* v) { return lane - (int)delta >= base ? v[lane - (int)delta] : v[lane]; });
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <typename T>
inline T __shfl_xor_sync(unsigned int mask, T val, int lane_mask, int width = 32)
{
  unsigned long long raw = 0;
  std::memcpy(&raw, &val, sizeof(T));
  const int lane = ::cuda_emu::my_lane();
  (void)width;
  unsigned long long got = ::cuda_emu::exchange(mask, raw, [=](const unsigned long long* v) { return v[(lane ^ lane_mask) & 31]; });
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
inline unsigned int __match_any_sync(unsigned int mask, int value)
{
  const int lane = ::cuda_emu::my_lane();
  (void)lane;
  return ::cuda_emu::exchange(mask, (unsigned long long)(unsigned int)value, [=](const unsigned long long* v) {
    unsigned int r = 0;
    for (int i = 0; i < 32; i++)
      if (((mask >> i) & 1u) && v[i] == (unsigned long long)(unsigned int)value) r |= 1u << i;
    return r;
  });
}

inline unsigned int __match_any_sync(unsigned int mask, unsigned long long value)
{
  return ::cuda_emu::exchange(mask, value, [=](const unsigned long long* v) {
    unsigned int r = 0;
    for (int i = 0; i < 32; i++)
      if (((mask >> i) & 1u) && v[i] == value) r |= 1u << i;
    return r;
  });
}

inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T>
inline T atomicMax(T* p, T v)
{
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <typename T>
inline T atomicMin(T* p, T v)
{
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}

inline float atomicAdd(float* p, float v)
{
  float old = *p;  // fibers run on one OS thread and only switch at barriers: read-modify-write is atomic by construction
  *p        = old + v;
  return old;
}
inline double atomicAdd(double* p, double v)
{
  double old = *p;
  *p         = old + v;
  return old;
}
template <typename T>
inline T atomicCAS(T* p, T expected, T desired)
{
  __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return expected;  // the value found (== the old `expected` iff the swap happened)
}
inline float __uint_as_float(unsigned int u)
{
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline float __int_as_float(int u)
{
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline int __float_as_int(float f)
{
  int u;
  std::memcpy(&u, &f, 4);
  return u;
}

template <typename T>
inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned int x) { return __builtin_popcount(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned int)x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
inline unsigned int __float_as_uint(float f)
{
  unsigned int u;
  std::memcpy(&u, &f, 4);
  return u;
}
using std::max;
using std::min;
