// Shared internals of libwholegraph_b200: error mapping, handle/tensor structs, the by-value
// chunk reference used by every kernel, env-callback RAII helpers.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <wholememory/embedding.h>
#include <wholememory/env_func_ptrs.h>
#include <wholememory/graph_op.h>
#include <wholememory/tensor_description.h>
#include <wholememory/wholegraph_op.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_op.h>
#include <wholememory/wholememory_tensor.h>

namespace wgb {

constexpr int kMaxWorld = 8;  // one NVSwitch box

// ---- errors ----------------------------------------------------------------------------------
struct cuda_error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct logic_error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct invalid_input : std::runtime_error {
  using std::runtime_error::runtime_error;
};

extern int g_log_level;
void log_msg(int level, const char* fmt, ...);

#define WGB_CUDA_TRY(call)                                                                     \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      cudaGetLastError();                                                                      \
      throw ::wgb::cuda_error(std::string(#call) + " failed: " + cudaGetErrorString(e__) +     \
                              " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")");        \
    }                                                                                          \
  } while (0)

// every kernel launch of this library goes through this check; the counter backs bench.py's
// "gpu_launches" claim (wholememory_b200_kernel_launch_count()).
extern unsigned long long g_kernel_launches;
#define WGB_CHECK_LAUNCH()                   \
  do {                                       \
    ++::wgb::g_kernel_launches;              \
    WGB_CUDA_TRY(cudaGetLastError());        \
  } while (0)

#define WGB_EXPECTS(cond, msg)                                                                 \
  do {                                                                                         \
    if (!(cond)) throw ::wgb::logic_error(std::string(msg) + " [" #cond "]");                  \
  } while (0)

#define WGB_CHECK_INPUT(cond, msg)                                                             \
  do {                                                                                         \
    if (!(cond)) throw ::wgb::invalid_input(std::string(msg) + " [" #cond "]");                \
  } while (0)

// Runs `fn`, mapping exceptions to the reference's error codes
// (reference: cpp/src/wholememory_ops/gather_op_impl_nccl.cu:160-168, cpp/src/error.hpp).
template <typename F>
wholememory_error_code_t guarded(const char* what, F&& fn)
{
  try {
    fn();
  } catch (const invalid_input& e) {
    log_msg(LEVEL_ERROR, "%s: invalid input: %s", what, e.what());
    return WHOLEMEMORY_INVALID_INPUT;
  } catch (const cuda_error& e) {
    log_msg(LEVEL_ERROR, "%s: CUDA error: %s", what, e.what());
    return WHOLEMEMORY_CUDA_ERROR;
  } catch (const logic_error& e) {
    log_msg(LEVEL_ERROR, "%s: logic error: %s", what, e.what());
    return WHOLEMEMORY_LOGIC_ERROR;
  } catch (const std::bad_alloc&) {
    return WHOLEMEMORY_OUT_OF_MEMORY;
  } catch (const std::exception& e) {
    log_msg(LEVEL_ERROR, "%s: %s", what, e.what());
    return WHOLEMEMORY_UNKNOW_ERROR;
  } catch (...) {
    return WHOLEMEMORY_UNKNOW_ERROR;
  }
  return WHOLEMEMORY_SUCCESS;
}

// ---- by-value chunk reference ----------------------------------------------------------------
// What the reference reaches through wholememory_gref_t + device_reference<T>
// (cpp/include/wholememory/device_reference.cuh:14-62: a device array of base pointers, one 64-bit
// divide per access) is passed here BY VALUE in the kernel parameter block: up to 8 base pointers
// and the byte offset at which each chunk starts.  The owning rank of a byte offset is found by a
// branch-free compare chain over <= 7 boundaries (no divide, handles uneven partitions too).
struct ChunkRef {
  char* base[kMaxWorld];
  unsigned long long start[kMaxWorld + 1];  // start[r] = first byte of rank r's chunk; start[world] = total
  int world;

  template <bool CHUNKED>
  __device__ __forceinline__ char* at(unsigned long long byte_off) const
  {
    if (!CHUNKED) return base[0] + byte_off;
    int r = 0;
#pragma unroll
    for (int i = 1; i < kMaxWorld; i++)
      r += (i < world && byte_off >= start[i]) ? 1 : 0;
    return base[r] + (byte_off - start[r]);
  }
};

}  // namespace wgb

// ---- opaque handle definitions (global namespace: they are the C ABI's opaque struct tags) ------
struct wholememory_comm_ {
  int rank       = 0;
  int size       = 1;
  int device_id  = -1;
  void* shm      = nullptr;  // shared rendezvous region (size > 1)
  size_t shm_len = 0;
  unsigned long long seq = 0;  // per-rank collective sequence number
  std::string shm_name;
  wholememory_distributed_backend_t backend = WHOLEMEMORY_DB_NCCL;
};

struct wholememory_handle_ {
  wholememory_comm_t comm                = nullptr;
  wholememory_memory_type_t type         = WHOLEMEMORY_MT_NONE;
  wholememory_memory_location_t location = WHOLEMEMORY_ML_NONE;
  size_t total_size                      = 0;
  size_t granularity                     = 1;
  int world                              = 1;
  int rank                               = 0;
  bool same_chunk                        = true;
  size_t stride                          = 0;  // bytes per chunk when same_chunk
  void* local_ptr                        = nullptr;
  size_t local_alloc                     = 0;
  void* peer_ptr[wgb::kMaxWorld]         = {};
  size_t chunk_start[wgb::kMaxWorld + 1] = {};  // byte offsets
  void** d_ptrs                          = nullptr;  // device copies for wholememory_gref_t
  size_t* d_offsets                      = nullptr;
  bool vmm                               = false;    // chunks are cuMem allocations mapped with cuMemMap
  unsigned long long vmm_handle[wgb::kMaxWorld] = {};
  size_t vmm_size[wgb::kMaxWorld]        = {};
  void* flat_ptr                         = nullptr;
};

struct wholememory_tensor_ {
  wholememory_tensor_description_t desc;
  void* storage_ptr           = nullptr;  // for pointer tensors
  wholememory_handle_t handle = nullptr;  // for WholeMemory tensors
  wholememory_tensor_t root   = nullptr;
  bool own_handle             = false;
  bool is_wholememory         = false;
};

// Replica of the hottest rows of a read-only table on THIS GPU (wholememory_embedding_set_hot_rows, b200_ops.h):
// slot[row] = index into `rows`, or -1.  Gathers consult it before going to the owning rank.
struct wgb_hot_rows {
  const int* slot                = nullptr;
  const char* rows               = nullptr;
  unsigned long long stride_bytes = 0;
};

struct wholememory_embedding_optimizer_;
struct wholememory_embedding_ {
  wholememory_tensor_t tensor = nullptr;
  int user_sms                = -1;
  wgb_hot_rows hot;            // valid iff hot.slot != nullptr
  void* hot_slot_mem = nullptr;
  void* hot_rows_mem = nullptr;
  long long hot_count = 0;
  // trainable embeddings (embedding_optimizer.cu)
  wholememory_embedding_optimizer_* optimizer = nullptr;
  std::vector<std::pair<std::string, wholememory_tensor_t>> states;  // per-row optimizer states, partitioned like the table
  std::vector<const char*> state_names;                              // nullptr-terminated view of `states`
  wholememory_tensor_t inbox_idx = nullptr, inbox_grad = nullptr;    // peer-mapped (index, gradient) mailboxes
  size_t inbox_cap                = 0;                               // entries per rank
  void* scratch[6]                = {};
  size_t scratch_bytes[6]         = {};
};

namespace wgb {

ChunkRef make_chunk_ref(wholememory_tensor_t t);
void comm_barrier(wholememory_comm_t c);
void comm_allgather(wholememory_comm_t c, const void* in, void* out, size_t bytes);  // bytes <= 256 per rank
void embedding_release_training_state(wholememory_embedding_t e);
void embedding_drop_hot_rows(wholememory_embedding_t e);
void hot_rows_invalidate_for_tensor(wholememory_tensor_t written);  // scatter / file load into a table that has a replica
// gather / scatter of rows; `hot` (may be null) is consulted by same-dtype gathers only
wholememory_error_code_t rows_op(wholememory_tensor_t wm_tensor, wholememory_tensor_t indices_tensor, wholememory_tensor_t dense_tensor,
                                 void* stream, int sms, bool scatter, const wgb_hot_rows* hot);
inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }
int num_sms();

inline size_t dtype_size(wholememory_dtype_t dt) { return wholememory_dtype_get_element_size(dt); }

// ---- env-callback RAII (reference: cpp/src/wholememory_ops/temp_memory_handle.hpp,
//      output_memory_handle.hpp) ---------------------------------------------------------------
class temp_memory {
 public:
  explicit temp_memory(wholememory_env_func_t* env) : fns_(&env->temporary_fns)
  {
    fns_->create_memory_context_fn(&ctx_, fns_->global_context);
  }
  temp_memory(const temp_memory&) = delete;
  ~temp_memory()
  {
    if (allocated_) fns_->free_fn(ctx_, fns_->global_context);
    fns_->destroy_memory_context_fn(ctx_, fns_->global_context);
  }
  void* alloc(int64_t elt_count, wholememory_dtype_t dt,
              wholememory_memory_allocation_type_t where = WHOLEMEMORY_MA_DEVICE)
  {
    WGB_EXPECTS(!allocated_, "temp_memory already allocated");
    wholememory_tensor_description_t d;
    wholememory_initialize_tensor_desc(&d);
    d.dim      = 1;
    d.sizes[0] = elt_count;
    d.dtype    = dt;
    void* p    = fns_->malloc_fn(&d, where, ctx_, fns_->global_context);
    allocated_ = true;
    if (elt_count > 0 && p == nullptr) throw std::bad_alloc();
    return p;
  }
  void* device(int64_t elt_count, wholememory_dtype_t dt) { return alloc(elt_count, dt, WHOLEMEMORY_MA_DEVICE); }
  void* bytes(int64_t n) { return alloc((n + 7) / 8, WHOLEMEMORY_DT_INT64, WHOLEMEMORY_MA_DEVICE); }

 private:
  wholememory_temp_memory_func_t* fns_;
  void* ctx_      = nullptr;
  bool allocated_ = false;
};

inline void* output_alloc(wholememory_env_func_t* env, void* ctx, int64_t elt_count, wholememory_dtype_t dt)
{
  wholememory_tensor_description_t d;
  wholememory_initialize_tensor_desc(&d);
  d.dim      = 1;
  d.sizes[0] = elt_count;
  d.dtype    = dt;
  void* p    = env->output_fns.malloc_fn(&d, WHOLEMEMORY_MA_DEVICE, ctx, env->output_fns.global_context);
  if (elt_count > 0 && p == nullptr) throw std::bad_alloc();
  return p;
}

}  // namespace wgb
