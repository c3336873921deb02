// Micro-probe: chip-wide rate of random-address primitives over a table of 16-byte slots (one op per element).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomic_probe atomic_probe.cu && ./atomic_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <algorithm>

struct __align__(16) Slot { unsigned long long key, aux; };
__device__ __forceinline__ unsigned long long mix(unsigned long long x)
{
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
enum { OP_LD16, OP_LD32, OP_ST8, OP_ST16, OP_RED_MIN64, OP_ATOM_MIN64, OP_CAS64, OP_CAS128, OP_EXCH64, OP_RED_ADD32, OP_LD32_THEN_ST16, OP_LD32_THEN_CAS128 };

template <int OP>
__global__ void __launch_bounds__(256) probe(Slot* table, unsigned int nslots, int n, unsigned long long salt, unsigned long long* sink)
{
  unsigned long long acc = 0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    unsigned int s = (unsigned int)(((mix(salt + e) >> 32) * (unsigned long long)(nslots / 2)) >> 32) * 2;
    Slot* p = table + s;
    if (OP == OP_LD16) {
      unsigned long long a, b;
      asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
      acc += a + b;
    } else if (OP == OP_LD32 || OP == OP_LD32_THEN_ST16 || OP == OP_LD32_THEN_CAS128) {
      unsigned long long a, b, c, d;
      asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
      acc += a + b + c + d;
      if (OP == OP_LD32_THEN_ST16) {
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a + 1), "l"(b + 1) : "memory");
      } else if (OP == OP_LD32_THEN_CAS128) {
        unsigned long long ok, oa;
        asm volatile("{\n .reg .b128 c, s, d;\n mov.b128 c, {%2, %3};\n mov.b128 s, {%4, %5};\n atom.relaxed.gpu.global.cas.b128 d, [%6], c, s;\n mov.b128 {%0, %1}, d;\n}"
                     : "=l"(ok), "=l"(oa) : "l"(a), "l"(b), "l"(a + 1), "l"(b + 1), "l"(p) : "memory");
        acc += ok + oa;
      }
    } else if (OP == OP_ST8) {
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"((unsigned long long)e) : "memory");
    } else if (OP == OP_ST16) {
      asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"((unsigned long long)e), "l"(salt) : "memory");
    } else if (OP == OP_RED_MIN64) {
      asm volatile("red.relaxed.gpu.global.min.u64 [%0], %1;" ::"l"(&p->aux), "l"((unsigned long long)e) : "memory");
    } else if (OP == OP_ATOM_MIN64) {
      acc += atomicMin(&p->aux, (unsigned long long)e);
    } else if (OP == OP_CAS64) {
      acc += atomicCAS(&p->key, ~0ULL, (unsigned long long)e);
    } else if (OP == OP_CAS128) {
      unsigned long long ok, oa;
      asm volatile("{\n .reg .b128 c, s, d;\n mov.b128 c, {%2, %3};\n mov.b128 s, {%4, %5};\n atom.relaxed.gpu.global.cas.b128 d, [%6], c, s;\n mov.b128 {%0, %1}, d;\n}"
                   : "=l"(ok), "=l"(oa) : "l"(~0ULL), "l"(~0ULL), "l"((unsigned long long)e), "l"(salt), "l"(p) : "memory");
      acc += ok + oa;
    } else if (OP == OP_EXCH64) {
      acc += atomicExch(&p->key, (unsigned long long)e);
    } else if (OP == OP_RED_ADD32) {
      atomicAdd((unsigned int*)&p->aux, 1u);
    }
  }
  if (acc == 0x1234567ULL) *sink = acc;
}
__global__ void flush_kernel(float* p, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] += 1.f; }

int main()
{
  const int n = 2600000;
  size_t cap = 36000000;
  Slot* table; float* junk; unsigned long long* sink;
  cudaMalloc(&table, cap * sizeof(Slot)); size_t jn = 512u << 20; cudaMalloc(&junk, jn); cudaMalloc(&sink, 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  unsigned long long salt = 1;
  auto run = [&](const char* name, auto kern, unsigned int nslots, int grid, bool flush) {
    float best = 1e9;
    for (int it = 0; it < 5; it++) {
      salt += 77777;
      cudaMemset(table, 0xFF, (size_t)nslots * 16);
      if (flush) flush_kernel<<<1184, 256>>>(junk, jn / 4);
      cudaEventRecord(a);
      kern<<<grid, 256>>>(table, nslots, n, salt, sink);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (it > 0) best = std::min(best, ms);
    }
    printf("%-22s nslots=%9u (%4.0f MB) grid=%5d flush=%d  %7.1f us  %6.1f Gop/s (%s)\n", name, nslots, nslots * 16e-6, grid, (int)flush, best * 1e3, n / (best * 1e-3) * 1e-9, cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
  };
  unsigned int sizes[] = {1000000, 5900000, 17700000};
  for (unsigned int ns : sizes)
    for (int flush = 0; flush < 2; flush++) {
      run("ld 16B", probe<OP_LD16>, ns, 1184, flush);
      run("ld 32B", probe<OP_LD32>, ns, 1184, flush);
      run("st 8B", probe<OP_ST8>, ns, 1184, flush);
      run("st 16B", probe<OP_ST16>, ns, 1184, flush);
      run("red.min.u64", probe<OP_RED_MIN64>, ns, 1184, flush);
      run("atom.min.u64", probe<OP_ATOM_MIN64>, ns, 1184, flush);
      run("atom.cas.b64", probe<OP_CAS64>, ns, 1184, flush);
      run("atom.cas.b128", probe<OP_CAS128>, ns, 1184, flush);
      run("atom.exch.b64", probe<OP_EXCH64>, ns, 1184, flush);
      run("red.add.u32", probe<OP_RED_ADD32>, ns, 1184, flush);
      run("ld32 -> st16", probe<OP_LD32_THEN_ST16>, ns, 1184, flush);
      run("ld32 -> cas128", probe<OP_LD32_THEN_CAS128>, ns, 1184, flush);
    }
  return 0;
}
