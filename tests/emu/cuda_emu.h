// TEST INFRASTRUCTURE ONLY (never part of the product, never linked into libwholegraph_b200.so).
//
// A tiny SIMT emulator that lets g++ compile the product's device kernels (csrc/*.cuh) and run their LOGIC on the CPU:
// one thread block at a time, every CUDA thread an OS thread, __syncthreads() / warp intrinsics as pthread barriers.
// It exists because a round can run out of GPU minutes before a newly written kernel has been on hardware: the kernel's
// control flow, indexing and barrier placement can still be checked against the oracle here.  It says nothing about
// memory ordering, occupancy or speed; the GPU parity tests stay the gate.
//
// Supported: threadIdx/blockIdx/blockDim/gridDim (.x/.y), __shared__ (function-scope static: one block runs at a time),
// __syncthreads, __syncwarp, __ballot_sync, __any_sync, __shfl_sync/_up/_xor (int, unsigned, long long, unsigned long
// long), __match_any_sync, atomicAdd/atomicMax/atomicMin, __ldg, __popc, __clz, __ffs, __clzll, __float_as_uint.
// Every warp intrinsic is a rendezvous of ALL 32 lanes of the warp: kernels whose sub-warp groups diverge around
// intrinsics with partial masks (uniform_small_kernel) cannot be emulated; masks only select which lanes are evaluated.
#pragma once

#include <cuda_runtime.h>  // types (dim3, uint3) and the host-side definitions of __host__/__device__ (empty under g++)

#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
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

struct Warp {
  pthread_barrier_t bar;
  unsigned long long v64[32];
};

struct Block {
  pthread_barrier_t bar;
  std::vector<Warp> warps;
};

inline thread_local Idx t_threadIdx;
inline Idx g_blockIdx, g_blockDim, g_gridDim;
inline Block* g_block = nullptr;

inline Warp& my_warp() { return g_block->warps[t_threadIdx.x >> 5]; }
inline int my_lane() { return (int)(t_threadIdx.x & 31u); }

// every lane publishes `v`, then reads what it needs through `read(all 32 values)`
template <typename F>
inline auto exchange(unsigned long long v, F&& read)
{
  Warp& w = my_warp();
  w.v64[my_lane()] = v;
  pthread_barrier_wait(&w.bar);
  auto r = read(w.v64);
  pthread_barrier_wait(&w.bar);
  return r;
}

// run `body` for every thread of every block of a grid (blocks sequentially)
inline void launch(unsigned int grid_x, unsigned int grid_y, unsigned int block_x, const std::function<void()>& body)
{
  g_gridDim.x = grid_x; g_gridDim.y = grid_y; g_gridDim.z = 1;
  g_blockDim.x = block_x; g_blockDim.y = 1; g_blockDim.z = 1;
  const unsigned int nwarps = (block_x + 31) / 32;
  for (unsigned int by = 0; by < grid_y; by++)
    for (unsigned int bx = 0; bx < grid_x; bx++) {
      Block blk;
      blk.warps.resize(nwarps);
      pthread_barrier_init(&blk.bar, nullptr, block_x);
      for (unsigned int w = 0; w < nwarps; w++)
        pthread_barrier_init(&blk.warps[w].bar, nullptr, std::min(32u, block_x - 32u * w));
      g_block      = &blk;
      g_blockIdx.x = bx; g_blockIdx.y = by; g_blockIdx.z = 0;
      std::vector<std::thread> ts;
      ts.reserve(block_x);
      for (unsigned int t = 0; t < block_x; t++)
        ts.emplace_back([t, &body] {
          t_threadIdx.x = t; t_threadIdx.y = 0; t_threadIdx.z = 0;
          body();
        });
      for (auto& th : ts)
        th.join();
      for (unsigned int w = 0; w < nwarps; w++)
        pthread_barrier_destroy(&blk.warps[w].bar);
      pthread_barrier_destroy(&blk.bar);
      g_block = nullptr;
    }
}

}  // namespace cuda_emu

#define threadIdx (::cuda_emu::t_threadIdx)
#define blockIdx (::cuda_emu::g_blockIdx)
#define blockDim (::cuda_emu::g_blockDim)
#define gridDim (::cuda_emu::g_gridDim)

inline void __syncthreads() { pthread_barrier_wait(&::cuda_emu::g_block->bar); }
inline void __syncwarp(unsigned int = 0xffffffffu) { pthread_barrier_wait(&::cuda_emu::my_warp().bar); }

inline unsigned int __ballot_sync(unsigned int mask, bool pred)
{
  return ::cuda_emu::exchange(pred ? 1ULL : 0ULL, [mask](const unsigned long long* v) {
    unsigned int r = 0;
    for (int i = 0; i < 32; i++)
      if (((mask >> i) & 1u) && v[i]) r |= 1u << i;
    return r;
  });
}
inline bool __any_sync(unsigned int mask, bool pred) { return (__ballot_sync(0xffffffffu, pred) & mask) != 0u; }

template <typename T>
inline T __shfl_sync(unsigned int, T val, int src, int width = 32)
{
  static_assert(sizeof(T) <= 8, "emulated shuffles move up to 64 bits");
  unsigned long long raw = 0;
  std::memcpy(&raw, &val, sizeof(T));
  const int lane = ::cuda_emu::my_lane();
  const int base = lane & ~(width - 1);
  unsigned long long got = ::cuda_emu::exchange(raw, [=](const unsigned long long* v) { return v[base + (src & (width - 1))]; });
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <typename T>
inline T __shfl_up_sync(unsigned int, T val, unsigned int delta, int width = 32)
{
  unsigned long long raw = 0;
  std::memcpy(&raw, &val, sizeof(T));
  const int lane = ::cuda_emu::my_lane();
  const int base = lane & ~(width - 1);
  unsigned long long got = ::cuda_emu::exchange(raw, [=](const unsigned long long* v) { return lane - (int)delta >= base ? v[lane - (int)delta] : v[lane]; });
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <typename T>
inline T __shfl_xor_sync(unsigned int, T val, int lane_mask, int width = 32)
{
  unsigned long long raw = 0;
  std::memcpy(&raw, &val, sizeof(T));
  const int lane = ::cuda_emu::my_lane();
  (void)width;
  unsigned long long got = ::cuda_emu::exchange(raw, [=](const unsigned long long* v) { return v[(lane ^ lane_mask) & 31]; });
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
inline unsigned int __match_any_sync(unsigned int mask, int value)
{
  const int lane = ::cuda_emu::my_lane();
  (void)lane;
  return ::cuda_emu::exchange((unsigned long long)(unsigned int)value, [=](const unsigned long long* v) {
    unsigned int r = 0;
    for (int i = 0; i < 32; i++)
      if (((mask >> i) & 1u) && v[i] == (unsigned long long)(unsigned int)value) r |= 1u << i;
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
