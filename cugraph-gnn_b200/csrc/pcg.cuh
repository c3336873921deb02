// PCG XSH-RR 64/32 with the RAFT PCGenerator stream mapping, host + device.
//
// The reference samplers draw from raft::random::detail::PCGenerator(DeviceState{seed, 0}, gidx)
// (cpp/src/wholegraph_ops/unweighted_sample_without_replacement_func.cuh:137,
//  weighted_sample_without_replacement_func.cuh:33-51, raft_random_gen.cu:32-53).  RAFT is not
// vendored in the reference tree; the generator is restated here (see oracle/wg_oracle.cpp for the
// parity note).  Stream for (seed, subsequence k):
//     state = 0; inc = 2k+1; step; state += seed; step; skip-ahead by k draws.
//
// B200-first: the skip-ahead (an affine power, O(log k) 64-bit multiplies per thread in RAFT) is
// evaluated from a 32 KB table of per-byte affine powers (8 byte positions x 256 values), so that a
// sampler lane pays a handful of multiplies instead of ~4*log2(k).
#pragma once

#include <cstdint>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace wgb {

constexpr unsigned long long kPcgMult = 6364136223846793005ULL;

// state' = state * g + inc * s   (skip-ahead by n draws: g = h^n, s = 1 + h + ... + h^(n-1))
struct Affine {
  unsigned long long g;
  unsigned long long s;
};

__host__ __device__ __forceinline__ Affine affine_then(const Affine& a, const Affine& b)
{
  // apply a, then b
  Affine r;
  r.g = a.g * b.g;
  r.s = a.s * b.g + b.s;
  return r;
}

__host__ __device__ __forceinline__ Affine affine_skip_loop(unsigned long long n)
{
  // Brown's arbitrary-stride algorithm with the increment factored out (C = inc * s)
  Affine acc{1ULL, 0ULL};
  Affine cur{kPcgMult, 1ULL};
  while (n) {
    if (n & 1ULL) acc = affine_then(acc, cur);
    cur = affine_then(cur, cur);
    n >>= 1;
  }
  return acc;
}

constexpr int kSkipTabBytes = 8;
constexpr int kSkipTabSize  = kSkipTabBytes * 256;  // entries

// tab[p*256 + v] = skip by v * 256^p draws
__host__ __device__ __forceinline__ Affine affine_skip_tab(const Affine* __restrict__ tab, unsigned long long n)
{
  Affine acc{1ULL, 0ULL};
  int p = 0;
  while (n) {
    unsigned int v = (unsigned int)(n & 255ULL);
    if (v) acc = affine_then(acc, tab[p * 256 + v]);
    n >>= 8;
    p++;
  }
  return acc;
}

struct Pcg {
  unsigned long long state;
  unsigned long long inc;

  __host__ __device__ __forceinline__ unsigned int next_u32()
  {
    unsigned long long old  = state;
    state                   = old * kPcgMult + inc;
    unsigned int xorshifted = (unsigned int)(((old >> 18u) ^ old) >> 27u);
    unsigned int rot        = (unsigned int)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
  }
  // RAFT PCGenerator(DeviceState{seed, base_subsequence = 0}, k) given the affine skip for k draws
  __host__ __device__ __forceinline__ void init_with_skip(unsigned long long seed, unsigned long long k, const Affine& skip)
  {
    inc                   = (k << 1u) | 1ULL;
    unsigned long long s0 = (inc + seed) * kPcgMult + inc;  // state=0; step; +=seed; step
    state                 = s0 * skip.g + inc * skip.s;
  }
  __host__ __device__ __forceinline__ void init_loop(unsigned long long seed, unsigned long long k)
  {
    init_with_skip(seed, k, affine_skip_loop(k));
  }
  __host__ __device__ __forceinline__ void init_tab(unsigned long long seed, unsigned long long k, const Affine* __restrict__ tab)
  {
    init_with_skip(seed, k, affine_skip_tab(tab, k));
  }
  __host__ __device__ __forceinline__ int next_i32() { return (int)(next_u32() & 0x7fffffffu); }
  __host__ __device__ __forceinline__ unsigned long long next_u64()
  {
    unsigned int a = next_u32();
    unsigned int b = next_u32();
    return (unsigned long long)a | ((unsigned long long)b << 32);
  }
  __host__ __device__ __forceinline__ float next_float() { return (float)(next_u32() >> 8) / 16777216.0f; }
};

// device-resident skip table (built once per device on first use)
const Affine* skip_table_device();

}  // namespace wgb
