// TEST INFRASTRUCTURE ONLY: the pieces of the CUDA runtime and of libwholegraph_b200's own runtime (runtime.cu,
// sample.cu) that csrc/multihop.cu's HOST code calls, restated for the CPU emulation in tests/emu: device memory is
// malloc, copies are memcpy (launches are synchronous in the emulator, so stream order is program order), streams and
// events are no-ops.  Tensors are plain pointer tensors.
#include "cuda_emu.h"

#include "wm_common.cuh"
#include "pcg.cuh"

#include <cstdarg>
#include <cstdlib>
#include <map>

extern "C" {

// "Device" allocations are adversarial on purpose: filled with garbage (cudaMalloc does not zero memory either) and fenced by
// guard words that are checked when the block is freed -- a kernel or a host-side sizing bug that writes past a scratch
// buffer shows up as a non-zero emu_guard_violations() instead of passing silently.
namespace {
#if defined(__SANITIZE_ADDRESS__)
constexpr size_t kGuard = 0;  // under AddressSanitizer its own red zones fence the block (and catch out-of-bounds READS too)
#else
constexpr size_t kGuard = 64;
#endif
std::map<void*, size_t>& live()
{
  static std::map<void*, size_t> m;
  return m;
}
long long g_violations = 0;
bool guards_ok(char* raw, size_t n)
{
  for (size_t i = 0; i < kGuard; i++)
    if ((unsigned char)raw[i] != 0xC3 || (unsigned char)raw[kGuard + n + i] != 0xC3) return false;
  return true;
}
}  // namespace

cudaError_t cudaMalloc(void** p, size_t n)
{
  char* raw = static_cast<char*>(std::malloc(n + 2 * kGuard));
  if (!raw) return cudaErrorMemoryAllocation;
  std::memset(raw, 0xC3, kGuard);
  std::memset(raw + kGuard, 0xA5, n);
  std::memset(raw + kGuard + n, 0xC3, kGuard);
  *p = raw + kGuard;
  live()[*p] = n;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p)
{
  if (!p) return cudaSuccess;
  auto it = live().find(p);
  if (it == live().end()) {
    g_violations++;  // not a live allocation (double free / foreign pointer)
    return cudaErrorInvalidValue;
  }
  char* raw = static_cast<char*>(p) - kGuard;
  if (!guards_ok(raw, it->second)) g_violations++;
  live().erase(it);
  std::free(raw);
  return cudaSuccess;
}
// guard check over everything still allocated + the violations seen at frees so far
long long emu_guard_violations()
{
  long long v = g_violations;
  for (auto& kv : live())
    if (!guards_ok(static_cast<char*>(kv.first) - kGuard, kv.second)) v++;
  return v;
}
long long emu_live_allocations() { return (long long)live().size(); }
cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t)
{
  if (n) std::memmove(dst, src, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind)
{
  if (n) std::memmove(dst, src, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t)
{
  for (size_t r = 0; r < height; r++)
    std::memmove(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t)
{
  if (n) std::memset(p, v, n);
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e)
{
  *e = nullptr;
  return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned int)
{
  *e = nullptr;
  return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t)
{
  *ms = 0.f;
  return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d)
{
  *d = 0;
  return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA runtime"; }

// ---- tensor descriptions / pointer tensors (product: csrc/runtime.cu) ----
size_t wholememory_dtype_get_element_size(wholememory_dtype_t dt)
{
  switch (dt) {
    case WHOLEMEMORY_DT_INT8: return 1;
    case WHOLEMEMORY_DT_INT16:
    case WHOLEMEMORY_DT_HALF:
    case WHOLEMEMORY_DT_BF16: return 2;
    case WHOLEMEMORY_DT_INT:
    case WHOLEMEMORY_DT_FLOAT: return 4;
    case WHOLEMEMORY_DT_INT64:
    case WHOLEMEMORY_DT_DOUBLE: return 8;
    default: return (size_t)-1;
  }
}
void wholememory_initialize_tensor_desc(wholememory_tensor_description_t* d)
{
  std::memset(d, 0, sizeof(*d));
  d->dtype = WHOLEMEMORY_DT_UNKNOWN;
}
bool wholememory_dtype_is_floating_number(wholememory_dtype_t dt)
{
  return dt == WHOLEMEMORY_DT_FLOAT || dt == WHOLEMEMORY_DT_HALF || dt == WHOLEMEMORY_DT_DOUBLE || dt == WHOLEMEMORY_DT_BF16;
}
bool wholememory_dtype_is_integer_number(wholememory_dtype_t dt)
{
  return dt == WHOLEMEMORY_DT_INT || dt == WHOLEMEMORY_DT_INT64 || dt == WHOLEMEMORY_DT_INT16 || dt == WHOLEMEMORY_DT_INT8;
}
bool wholememory_convert_tensor_desc_to_array(wholememory_array_description_t* a, wholememory_tensor_description_t* t)
{
  if (t->dim != 1 && t->dim != 0) return false;
  if (t->dim == 1 && t->strides[0] != 1) return false;
  a->size           = t->dim == 0 ? 1 : t->sizes[0];
  a->storage_offset = t->storage_offset;
  a->dtype          = t->dtype;
  return true;
}
bool wholememory_convert_tensor_desc_to_matrix(wholememory_matrix_description_t* m, wholememory_tensor_description_t* t)
{
  if (t->dim != 2 || t->strides[1] != 1) return false;
  m->sizes[0]       = t->sizes[0];
  m->sizes[1]       = t->sizes[1];
  m->stride         = t->strides[0];
  m->storage_offset = t->storage_offset;
  m->dtype          = t->dtype;
  return true;
}
bool wholememory_unsqueeze_tensor(wholememory_tensor_description_t* t, int dim)
{
  if (dim < 0 || dim > t->dim || t->dim >= WHOLEMEMORY_MAX_TENSOR_DIM) return false;
  int64_t new_stride = dim == t->dim ? 1 : t->sizes[dim] * t->strides[dim];
  for (int i = t->dim; i > dim; i--) {
    t->sizes[i]   = t->sizes[i - 1];
    t->strides[i] = t->strides[i - 1];
  }
  t->sizes[dim]   = 1;
  t->strides[dim] = new_stride;
  t->dim++;
  return true;
}
wholememory_tensor_t wholememory_tensor_get_root(wholememory_tensor_t t) { return t ? (t->root ? t->root : t) : nullptr; }
wholememory_tensor_description_t* wholememory_tensor_get_tensor_description(wholememory_tensor_t t) { return &t->desc; }
void* wholememory_tensor_get_data_pointer(wholememory_tensor_t t)
{
  return static_cast<char*>(t->storage_ptr) + t->desc.storage_offset * wholememory_dtype_get_element_size(t->desc.dtype);
}

}  // extern "C"

namespace wgb {

int g_log_level                     = 0;
unsigned long long g_kernel_launches = 0;

void log_msg(int, const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  std::vfprintf(stderr, fmt, ap);
  std::fprintf(stderr, "\n");
  va_end(ap);
}

int num_sms() { return 2; }  // small grids: every grid-stride loop and persistent ticket loop iterates

// emu_set_split_world(W > 1): every tensor is presented to the kernels as if it were CHUNKED over W ranks (row-aligned
// boundaries, chunk r based at ptr + start[r], i.e. the same memory): the CHUNKED template variants and their owner lookup run.
int g_split_world = 1;

ChunkRef make_chunk_ref(wholememory_tensor_t t)
{
  ChunkRef r;
  std::memset(&r, 0, sizeof(r));
  const unsigned long long elt = wholememory_dtype_get_element_size(t->desc.dtype);
  const long long rows   = t->desc.sizes[0];
  const long long stride = t->desc.dim == 2 ? t->desc.strides[0] : 1;
  const int W            = std::max(1, std::min(g_split_world, kMaxWorld));
  r.world                = W;
  r.base[0]              = static_cast<char*>(t->storage_ptr);
  for (int k = 1; k < W; k++) {
    r.start[k] = (unsigned long long)(t->desc.storage_offset + (rows * k / W) * stride) * elt;
    r.base[k]  = static_cast<char*>(t->storage_ptr) + r.start[k];
  }
  r.start[W] = (unsigned long long)(t->desc.storage_offset + rows * stride) * elt;
  return r;
}

#ifndef EMU_PRODUCT_SKIP_TABLE  // the library that contains csrc/sample.cu has the product's own skip_table_device()
const Affine* skip_table_device()
{
  static std::vector<Affine> host;
  if (host.empty()) {
    host.resize(kSkipTabSize);
    for (int p = 0; p < kSkipTabBytes; p++) {
      Affine unit = affine_skip_loop(1ULL << (8 * p));
      Affine acc{1ULL, 0ULL};
      for (int v = 0; v < 256; v++) {
        host[p * 256 + v] = acc;
        acc               = affine_then(acc, unit);
      }
    }
  }
  return host.data();
}
#endif

// runtime.cu drops an embedding's hot-row replica when its table is written; the emulated library has no embeddings
void hot_rows_invalidate_for_tensor(wholememory_tensor_t) {}

}  // namespace wgb

extern "C" void emu_set_split_world(int w) { wgb::g_split_world = w; }
