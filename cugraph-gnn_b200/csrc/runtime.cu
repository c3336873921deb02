// Host runtime of libwholegraph_b200: descriptors, shared-memory communicator, peer-mapped
// WholeMemory handles, tensors, default allocation callbacks, non-cached embedding.
//
// B200-first design: one process per GPU on one NVSwitch box.
//  * communicator = POSIX shared-memory rendezvous keyed by the 128-byte unique id: counters
//    for barriers and 256-byte slots for all-gathers.  It replaces the reference's
//    ncclCommInitRank + unix-socket FD passing (cpp/src/wholememory/communicator.cpp:398-767,
//    memory_handle.cpp:852-901) because the only thing the hot path needs from it is the
//    exchange of 64-byte cudaIpc handles at allocation time.
//  * every device memory type is one cudaMalloc per rank, exported with cudaIpcGetMemHandle and
//    opened by every peer (reference CHUNKED: cpp/src/wholememory/memory_handle.cpp:1054-1211);
//    kernels address it through wgb::ChunkRef (by-value base pointers, no divide).
//  * partition plan: reference cpp/src/wholememory/memory_handle.cpp:1597-1629, 2116-2122.

#include "wm_common.cuh"

#include <wholememory/b200_ops.h>

#include <cuda.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>

#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <random>

namespace wgb {

int g_log_level = LEVEL_WARN;
unsigned long long g_kernel_launches = 0;

void log_msg(int level, const char* fmt, ...)
{
  if (level > g_log_level) return;
  static const char* names[] = {"FATAL", "ERROR", "WARN", "INFO", "DEBUG", "TRACE"};
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "[wholegraph_b200 %s] ", names[level < 0 ? 0 : (level > 5 ? 5 : level)]);
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}

int num_sms()
{
  static int sms[64] = {};
  int dev            = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

ChunkRef make_chunk_ref(wholememory_tensor_t t)
{
  ChunkRef r;
  memset(&r, 0, sizeof(r));
  wholememory_tensor_t root = t->root ? t->root : t;
  if (root->handle != nullptr) {
    wholememory_handle_t h = root->handle;
    r.world                = h->world;
    for (int i = 0; i < h->world; i++) {
      r.base[i]  = static_cast<char*>(h->peer_ptr[i]);
      r.start[i] = h->chunk_start[i];
    }
    r.start[h->world] = h->chunk_start[h->world];
  } else {
    r.world    = 1;
    r.base[0]  = static_cast<char*>(root->storage_ptr);
    r.start[0] = 0;
    r.start[1] = ~0ULL;
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// shared-memory communicator
// ---------------------------------------------------------------------------------------------
constexpr int kSlotBytes = 256;
struct ShmRegion {
  std::atomic<unsigned long long> barrier_count;
  std::atomic<unsigned int> attached;
  unsigned int world;
  char pad[64 - 16];
  char slots[2][kMaxWorld][kSlotBytes];
};

static double now_sec()
{
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

void comm_barrier(wholememory_comm_t c)
{
  if (c->size <= 1) return;
  auto* reg = static_cast<ShmRegion*>(c->shm);
  unsigned long long target = (c->seq + 1) * (unsigned long long)c->size;
  c->seq++;
  reg->barrier_count.fetch_add(1, std::memory_order_acq_rel);
  double t0 = now_sec();
  int spins = 0;
  while (reg->barrier_count.load(std::memory_order_acquire) < target) {
    if (++spins > 1000) {
      timespec ts = {0, 50000};
      nanosleep(&ts, nullptr);
      if (now_sec() - t0 > 300.0) throw std::runtime_error("communicator barrier timed out (peer died?)");
    }
  }
}

// every rank contributes `bytes` (<= kSlotBytes); out receives size*bytes
void comm_allgather(wholememory_comm_t c, const void* in, void* out, size_t bytes)
{
  if (bytes > (size_t)kSlotBytes) throw logic_error("allgather payload too large");
  if (c->size <= 1) {
    memcpy(out, in, bytes);
    return;
  }
  auto* reg  = static_cast<ShmRegion*>(c->shm);
  int parity = (int)(c->seq & 1);
  memcpy(reg->slots[parity][c->rank], in, bytes);
  comm_barrier(c);  // release/acquire on barrier_count orders the slot writes
  for (int r = 0; r < c->size; r++)
    memcpy(static_cast<char*>(out) + r * bytes, reg->slots[parity][r], bytes);
  // double-buffered: the same parity is rewritten only after the NEXT barrier, which every rank
  // reaches after it finished reading these slots.
}


// ---------------------------------------------------------------------------------------------
// peer-mapped device memory through the CUDA virtual memory management API
// ---------------------------------------------------------------------------------------------
// cudaIpc mappings of cudaMalloc memory were MEASURED at 58 GB/s for random 512-byte rows between two
// B200s (PCIe BAR path) against 762 GB/s for the same kernel over a cuMem peer mapping
// (profiles/p2p_*_probe.py), so multi-rank chunks are cuMemCreate allocations exported as POSIX file
// descriptors, passed between the processes of the box over an abstract unix datagram socket
// (SCM_RIGHTS) and mapped with cuMemMap + cuMemSetAccess -- NVLink loads/stores from inside kernels.
// The driver entry points are resolved through the runtime, so the library has no link-time
// dependency on libcuda and still loads on a machine without a GPU.
struct DriverApi {
  CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*memRelease)(CUmemGenericAllocationHandle);
  CUresult (*memExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*memImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*addrReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*addrFree)(CUdeviceptr, size_t);
  CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*memUnmap)(CUdeviceptr, size_t);
  CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*memGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  bool ok = false;
};

static const DriverApi& driver_api()
{
  static DriverApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult st;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess && *fn != nullptr;
    };
    api.ok = get("cuMemCreate", (void**)&api.memCreate) && get("cuMemRelease", (void**)&api.memRelease) &&
             get("cuMemExportToShareableHandle", (void**)&api.memExport) &&
             get("cuMemImportFromShareableHandle", (void**)&api.memImport) &&
             get("cuMemAddressReserve", (void**)&api.addrReserve) && get("cuMemAddressFree", (void**)&api.addrFree) &&
             get("cuMemMap", (void**)&api.memMap) && get("cuMemUnmap", (void**)&api.memUnmap) &&
             get("cuMemSetAccess", (void**)&api.memSetAccess) &&
             get("cuMemGetAllocationGranularity", (void**)&api.memGranularity);
    cudaGetLastError();
  }
  return api;
}

#define WGB_CU_TRY(call)                                                                        \
  do {                                                                                          \
    CUresult r__ = (call);                                                                      \
    if (r__ != CUDA_SUCCESS)                                                                    \
      throw ::wgb::cuda_error(std::string(#call) + " failed with CUresult " + std::to_string((int)r__)); \
  } while (0)

// every rank sends `fd` to every other rank and receives theirs (out[r] for r != rank)
static void exchange_fds(wholememory_comm_t c, int fd, int* out)
{
  static std::atomic<unsigned int> local_seq{0};
  unsigned int seq = local_seq.fetch_add(1);
  auto make_addr = [&](int rank, sockaddr_un* a) {
    memset(a, 0, sizeof(*a));
    a->sun_family = AF_UNIX;
    // abstract namespace: sun_path[0] == 0, no file system entry to clean up
    int n = snprintf(a->sun_path + 1, sizeof(a->sun_path) - 1, "%s_fd%u_%d", c->shm_name.c_str() + 1, seq, rank);
    return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
  };
  int sock = socket(AF_UNIX, SOCK_DGRAM, 0);
  if (sock < 0) throw std::runtime_error("socket() failed");
  sockaddr_un me;
  socklen_t me_len = make_addr(c->rank, &me);
  if (bind(sock, (sockaddr*)&me, me_len) != 0) {
    close(sock);
    throw std::runtime_error("bind() of the fd-exchange socket failed");
  }
  try {
    comm_barrier(c);  // everybody is bound
    for (int r = 0; r < c->size; r++) {
      if (r == c->rank) continue;
      sockaddr_un to;
      socklen_t to_len = make_addr(r, &to);
      int payload      = c->rank;
      iovec iov        = {&payload, sizeof(payload)};
      char ctrl[CMSG_SPACE(sizeof(int))];
      memset(ctrl, 0, sizeof(ctrl));
      msghdr msg;
      memset(&msg, 0, sizeof(msg));
      msg.msg_name       = &to;
      msg.msg_namelen    = to_len;
      msg.msg_iov        = &iov;
      msg.msg_iovlen     = 1;
      msg.msg_control    = ctrl;
      msg.msg_controllen = sizeof(ctrl);
      cmsghdr* cm        = CMSG_FIRSTHDR(&msg);
      cm->cmsg_level     = SOL_SOCKET;
      cm->cmsg_type      = SCM_RIGHTS;
      cm->cmsg_len       = CMSG_LEN(sizeof(int));
      memcpy(CMSG_DATA(cm), &fd, sizeof(int));
      if (sendmsg(sock, &msg, 0) < 0) throw std::runtime_error("sendmsg(SCM_RIGHTS) failed");
    }
    for (int k = 0; k < c->size - 1; k++) {
      int payload = -1;
      iovec iov   = {&payload, sizeof(payload)};
      char ctrl[CMSG_SPACE(sizeof(int))];
      msghdr msg;
      memset(&msg, 0, sizeof(msg));
      msg.msg_iov        = &iov;
      msg.msg_iovlen     = 1;
      msg.msg_control    = ctrl;
      msg.msg_controllen = sizeof(ctrl);
      if (recvmsg(sock, &msg, 0) < 0) throw std::runtime_error("recvmsg(SCM_RIGHTS) failed");
      cmsghdr* cm = CMSG_FIRSTHDR(&msg);
      if (!cm || cm->cmsg_type != SCM_RIGHTS || payload < 0 || payload >= c->size) throw std::runtime_error("malformed fd message");
      memcpy(&out[payload], CMSG_DATA(cm), sizeof(int));
    }
    comm_barrier(c);
  } catch (...) {
    close(sock);
    throw;
  }
  close(sock);
}

}  // namespace wgb

using namespace wgb;

// ---------------------------------------------------------------------------------------------
// tensor_description.h
// ---------------------------------------------------------------------------------------------
extern "C" {

size_t wholememory_dtype_get_element_size(wholememory_dtype_t dtype)
{
  switch (dtype) {
    case WHOLEMEMORY_DT_INT8: return 1;
    case WHOLEMEMORY_DT_HALF:
    case WHOLEMEMORY_DT_BF16:
    case WHOLEMEMORY_DT_INT16: return 2;
    case WHOLEMEMORY_DT_FLOAT:
    case WHOLEMEMORY_DT_INT: return 4;
    case WHOLEMEMORY_DT_DOUBLE:
    case WHOLEMEMORY_DT_INT64: return 8;
    default: return static_cast<size_t>(-1);
  }
}

bool wholememory_dtype_is_floating_number(wholememory_dtype_t dtype)
{
  return dtype == WHOLEMEMORY_DT_FLOAT || dtype == WHOLEMEMORY_DT_HALF || dtype == WHOLEMEMORY_DT_DOUBLE ||
         dtype == WHOLEMEMORY_DT_BF16;
}

bool wholememory_dtype_is_integer_number(wholememory_dtype_t dtype)
{
  return dtype == WHOLEMEMORY_DT_INT || dtype == WHOLEMEMORY_DT_INT64 || dtype == WHOLEMEMORY_DT_INT16 ||
         dtype == WHOLEMEMORY_DT_INT8;
}

wholememory_array_description_t wholememory_create_array_desc(int64_t size, int64_t storage_offset,
                                                              wholememory_dtype_t dtype)
{
  wholememory_array_description_t d;
  d.size           = size;
  d.storage_offset = storage_offset;
  d.dtype          = dtype;
  return d;
}

wholememory_matrix_description_t wholememory_create_matrix_desc(int64_t sizes[2], int64_t stride,
                                                                int64_t storage_offset,
                                                                wholememory_dtype_t dtype)
{
  wholememory_matrix_description_t d;
  d.sizes[0]       = sizes[0];
  d.sizes[1]       = sizes[1];
  d.stride         = stride;
  d.storage_offset = storage_offset;
  d.dtype          = dtype;
  return d;
}

void wholememory_initialize_tensor_desc(wholememory_tensor_description_t* p)
{
  for (int i = 0; i < WHOLEMEMORY_MAX_TENSOR_DIM; i++) {
    p->sizes[i]   = 1;
    p->strides[i] = 1;
  }
  p->storage_offset = 0;
  p->dim            = 0;
  p->dtype          = WHOLEMEMORY_DT_UNKNOWN;
}

void wholememory_copy_array_desc_to_matrix(wholememory_matrix_description_t* m,
                                           wholememory_array_description_t* a)
{
  m->sizes[0]       = a->size;
  m->sizes[1]       = 1;
  m->stride         = 1;
  m->storage_offset = a->storage_offset;
  m->dtype          = a->dtype;
}

void wholememory_copy_array_desc_to_tensor(wholememory_tensor_description_t* t,
                                           wholememory_array_description_t* a)
{
  wholememory_initialize_tensor_desc(t);
  t->dim            = 1;
  t->sizes[0]       = a->size;
  t->strides[0]     = 1;
  t->storage_offset = a->storage_offset;
  t->dtype          = a->dtype;
}

void wholememory_copy_matrix_desc_to_tensor(wholememory_tensor_description_t* t,
                                            wholememory_matrix_description_t* m)
{
  wholememory_initialize_tensor_desc(t);
  t->dim            = 2;
  t->sizes[0]       = m->sizes[0];
  t->sizes[1]       = m->sizes[1];
  t->strides[0]     = m->stride;
  t->strides[1]     = 1;
  t->storage_offset = m->storage_offset;
  t->dtype          = m->dtype;
}

bool wholememory_convert_tensor_desc_to_array(wholememory_array_description_t* a,
                                              wholememory_tensor_description_t* t)
{
  if (t->dim != 1 && t->dim != 0) return false;
  if (t->dim == 1 && t->strides[0] != 1) return false;
  a->size           = t->dim == 0 ? 1 : t->sizes[0];
  a->storage_offset = t->storage_offset;
  a->dtype          = t->dtype;
  return true;
}

bool wholememory_convert_tensor_desc_to_matrix(wholememory_matrix_description_t* m,
                                               wholememory_tensor_description_t* t)
{
  if (t->dim != 2) return false;
  if (t->strides[1] != 1) return false;
  m->sizes[0]       = t->sizes[0];
  m->sizes[1]       = t->sizes[1];
  m->stride         = t->strides[0];
  m->storage_offset = t->storage_offset;
  m->dtype          = t->dtype;
  return true;
}

int64_t wholememory_get_memory_element_count_from_array(wholememory_array_description_t* a) { return a->size; }
int64_t wholememory_get_memory_size_from_array(wholememory_array_description_t* a)
{
  return a->size * (int64_t)wholememory_dtype_get_element_size(a->dtype);
}
int64_t wholememory_get_memory_element_count_from_matrix(wholememory_matrix_description_t* m)
{
  return m->sizes[0] * m->stride;
}
int64_t wholememory_get_memory_size_from_matrix(wholememory_matrix_description_t* m)
{
  return wholememory_get_memory_element_count_from_matrix(m) * (int64_t)wholememory_dtype_get_element_size(m->dtype);
}
int64_t wholememory_get_memory_element_count_from_tensor(wholememory_tensor_description_t* t)
{
  if (t->dim == 0) return 1;
  return t->sizes[0] * t->strides[0];
}
int64_t wholememory_get_memory_size_from_tensor(wholememory_tensor_description_t* t)
{
  return wholememory_get_memory_element_count_from_tensor(t) * (int64_t)wholememory_dtype_get_element_size(t->dtype);
}

bool wholememory_squeeze_tensor(wholememory_tensor_description_t* t, int dim)
{
  if (dim < 0 || dim >= t->dim) return false;
  if (t->sizes[dim] != 1) return false;
  for (int i = dim; i < t->dim - 1; i++) {
    t->sizes[i]   = t->sizes[i + 1];
    t->strides[i] = t->strides[i + 1];
  }
  t->dim--;
  return true;
}

bool wholememory_unsqueeze_tensor(wholememory_tensor_description_t* t, int dim)
{
  if (dim < 0 || dim > t->dim) return false;
  if (t->dim >= WHOLEMEMORY_MAX_TENSOR_DIM) return false;
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

wholememory_gref_t wholememory_create_continuous_global_reference(void* ptr)
{
  wholememory_gref_t g;
  g.pointer             = ptr;
  g.rank_memory_offsets = nullptr;
  g.world_size          = 1;
  g.stride              = 0;
  g.same_chunk          = true;
  return g;
}

// ---------------------------------------------------------------------------------------------
// init / communicator
// ---------------------------------------------------------------------------------------------
static std::mutex g_mu;
static bool g_inited = false;

wholememory_error_code_t wholememory_init(unsigned int flags, LogLevel log_level)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (flags != 0) return WHOLEMEMORY_INVALID_INPUT;
  g_log_level = (int)log_level;
  g_inited    = true;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_finalize()
{
  std::lock_guard<std::mutex> lk(g_mu);
  g_inited = false;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_create_unique_id(wholememory_unique_id_t* unique_id)
{
  if (unique_id == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  std::random_device rd;
  for (int i = 0; i < WHOLEMEMORY_UNIQUE_ID_BYTES; i += 4) {
    unsigned int v = rd();
    memcpy(unique_id->internal + i, &v, 4);
  }
  // mix in pid/time so that a weak random_device cannot collide between jobs
  unsigned long long salt = (unsigned long long)getpid() ^ (unsigned long long)(now_sec() * 1e6);
  for (int i = 0; i < 8; i++)
    unique_id->internal[i] ^= (char)(salt >> (8 * i));
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_create_communicator(wholememory_comm_t* comm,
                                                         wholememory_unique_id_t unique_id, int rank,
                                                         int size)
{
  return guarded("wholememory_create_communicator", [&] {
    WGB_CHECK_INPUT(comm != nullptr, "comm is null");
    WGB_CHECK_INPUT(size >= 1 && rank >= 0 && rank < size, "bad rank/size");
    if (size > kMaxWorld) throw invalid_input("world size > 8: one NVSwitch box only");
    auto* c = new wholememory_comm_();
    c->rank = rank;
    c->size = size;
    cudaGetDevice(&c->device_id);
    if (size > 1) {
      char name[64] = "/wgb200_";
      static const char* hex = "0123456789abcdef";
      for (int i = 0; i < 16; i++) {
        unsigned char b   = (unsigned char)unique_id.internal[i];
        name[8 + 2 * i]   = hex[b >> 4];
        name[8 + 2 * i + 1] = hex[b & 15];
      }
      name[8 + 32] = 0;
      c->shm_name  = name;
      c->shm_len   = sizeof(ShmRegion);
      int fd       = shm_open(name, O_CREAT | O_RDWR, 0600);
      if (fd < 0) {
        delete c;
        throw std::runtime_error("shm_open failed");
      }
      if (ftruncate(fd, (off_t)c->shm_len) != 0) {
        close(fd);
        delete c;
        throw std::runtime_error("ftruncate failed");
      }
      c->shm = mmap(nullptr, c->shm_len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      close(fd);
      if (c->shm == MAP_FAILED) {
        delete c;
        throw std::runtime_error("mmap failed");
      }
      auto* reg = static_cast<ShmRegion*>(c->shm);
      reg->attached.fetch_add(1);
      comm_barrier(c);
      if (rank == 0) shm_unlink(name);  // everybody is attached: the name can go away
    }
    *comm = c;
  });
}

wholememory_error_code_t wholememory_split_communicator(wholememory_comm_t*, wholememory_comm_t, int, int)
{
  return WHOLEMEMORY_NOT_SUPPORTED;  // multi-level (cross-node) communicators are out of scope
}

wholememory_error_code_t wholememory_destroy_communicator(wholememory_comm_t comm)
{
  if (comm == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  if (comm->shm) munmap(comm->shm, comm->shm_len);
  delete comm;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_communicator_support_type_location(
  wholememory_comm_t comm, wholememory_memory_type_t memory_type, wholememory_memory_location_t memory_location)
{
  if (comm == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  if (memory_location != WHOLEMEMORY_ML_DEVICE) return WHOLEMEMORY_NOT_SUPPORTED;
  if (memory_type == WHOLEMEMORY_MT_CONTINUOUS || memory_type == WHOLEMEMORY_MT_CHUNKED ||
      memory_type == WHOLEMEMORY_MT_DISTRIBUTED)
    return WHOLEMEMORY_SUCCESS;
  return WHOLEMEMORY_NOT_SUPPORTED;
}

wholememory_error_code_t wholememory_communicator_get_rank(int* rank, wholememory_comm_t comm)
{
  if (!rank || !comm) return WHOLEMEMORY_INVALID_INPUT;
  *rank = comm->rank;
  return WHOLEMEMORY_SUCCESS;
}
wholememory_error_code_t wholememory_communicator_get_size(int* size, wholememory_comm_t comm)
{
  if (!size || !comm) return WHOLEMEMORY_INVALID_INPUT;
  *size = comm->size;
  return WHOLEMEMORY_SUCCESS;
}
wholememory_error_code_t wholememory_communicator_get_local_size(int* local_size, wholememory_comm_t comm)
{
  if (!local_size || !comm) return WHOLEMEMORY_INVALID_INPUT;
  *local_size = comm->size;  // single box
  return WHOLEMEMORY_SUCCESS;
}
wholememory_error_code_t wholememory_communicator_set_distributed_backend(
  wholememory_comm_t comm, wholememory_distributed_backend_t backend)
{
  if (!comm) return WHOLEMEMORY_INVALID_INPUT;
  if (backend == WHOLEMEMORY_DB_NVSHMEM) return WHOLEMEMORY_NOT_SUPPORTED;
  comm->backend = backend;
  return WHOLEMEMORY_SUCCESS;
}
wholememory_distributed_backend_t wholememory_communicator_get_distributed_backend(wholememory_comm_t comm)
{
  return comm ? comm->backend : WHOLEMEMORY_DB_NONE;
}
wholememory_error_code_t wholememory_communicator_barrier(wholememory_comm_t comm)
{
  if (!comm) return WHOLEMEMORY_INVALID_INPUT;
  return guarded("wholememory_communicator_barrier", [&] { comm_barrier(comm); });
}
bool wholememory_is_intranode_communicator(wholememory_comm_t) { return true; }
bool wholememory_is_intra_mnnvl_communicator(wholememory_comm_t) { return false; }
bool wholememory_is_build_with_nvshmem() { return false; }

unsigned long long wholememory_b200_kernel_launch_count() { return wgb::g_kernel_launches; }

int fork_get_device_count()
{
  // count devices in a child so that the caller never creates a CUDA context
  // (reference: cpp/src/parallel_utils.cpp ForkGetDeviceCount)
  int fds[2];
  if (pipe(fds) != 0) return -1;
  pid_t pid = fork();
  if (pid < 0) return -1;
  if (pid == 0) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
    ssize_t w = write(fds[1], &n, sizeof(n));
    (void)w;
    _exit(0);
  }
  int n     = -1;
  ssize_t r = read(fds[0], &n, sizeof(n));
  (void)r;
  close(fds[0]);
  close(fds[1]);
  int st = 0;
  waitpid(pid, &st, 0);
  return n;
}

// ---------------------------------------------------------------------------------------------
// WholeMemory handles
// ---------------------------------------------------------------------------------------------
wholememory_error_code_t wholememory_equal_entry_partition_plan(size_t* entry_per_rank,
                                                                size_t total_entry_count, int world_size)
{
  if (!entry_per_rank || world_size <= 0) return WHOLEMEMORY_INVALID_INPUT;
  *entry_per_rank = (total_entry_count + world_size - 1) / world_size;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_malloc(wholememory_handle_t* handle_ptr, size_t total_size,
                                            wholememory_comm_t comm, wholememory_memory_type_t memory_type,
                                            wholememory_memory_location_t memory_location,
                                            size_t data_granularity, size_t* rank_entry_partition)
{
  return guarded("wholememory_malloc", [&] {
    WGB_CHECK_INPUT(handle_ptr != nullptr && comm != nullptr, "null argument");
    WGB_CHECK_INPUT(data_granularity > 0, "data_granularity must be > 0");
    WGB_CHECK_INPUT(total_size % data_granularity == 0, "total_size must be a multiple of data_granularity");
    if (memory_location != WHOLEMEMORY_ML_DEVICE)
      throw invalid_input("only WHOLEMEMORY_ML_DEVICE is supported (B200 HBM-resident tables)");
    if (memory_type != WHOLEMEMORY_MT_CONTINUOUS && memory_type != WHOLEMEMORY_MT_CHUNKED &&
        memory_type != WHOLEMEMORY_MT_DISTRIBUTED)
      throw invalid_input("unsupported memory type");
    auto* h        = new wholememory_handle_();
    h->comm        = comm;
    h->world       = comm->size;
    h->rank        = comm->rank;
    h->location    = memory_location;
    h->total_size  = total_size;
    h->granularity = data_granularity;
    // CONTINUOUS over several ranks would need a cuMem VMM flat mapping; on one NVSwitch box the
    // chunked peer mapping gives the same load/store reachability, so multi-rank CONTINUOUS is
    // served by the chunked layout and reports itself as CHUNKED.
    h->type = (memory_type == WHOLEMEMORY_MT_CONTINUOUS && comm->size > 1) ? WHOLEMEMORY_MT_CHUNKED : memory_type;
    size_t entries = total_size / data_granularity;
    int W          = h->world;
    size_t acc     = 0;
    if (rank_entry_partition != nullptr) {
      size_t sum = 0;
      for (int r = 0; r < W; r++)
        sum += rank_entry_partition[r];
      if (sum != entries) {
        delete h;
        throw invalid_input("rank_entry_partition does not sum to the entry count");
      }
      h->same_chunk = true;
      for (int r = 0; r < W; r++) {
        h->chunk_start[r] = acc * data_granularity;
        acc += rank_entry_partition[r];
        if (r + 1 < W - 0 && r > 0 && rank_entry_partition[r] != rank_entry_partition[0] && r != W - 1) h->same_chunk = false;
      }
      if (W > 1 && rank_entry_partition[W - 1] > rank_entry_partition[0]) h->same_chunk = false;
      h->chunk_start[W] = acc * data_granularity;
      h->stride         = h->same_chunk ? rank_entry_partition[0] * data_granularity : total_size / W;
    } else {
      size_t per = (entries + W - 1) / W;
      for (int r = 0; r <= W; r++) {
        size_t e          = std::min(entries, per * (size_t)r);
        h->chunk_start[r] = e * data_granularity;
      }
      h->same_chunk = true;
      h->stride     = per * data_granularity;
    }
    size_t local_bytes = h->chunk_start[h->rank + 1] - h->chunk_start[h->rank];
    for (int r = 0; r < W; r++)
      h->peer_ptr[r] = nullptr;
    const DriverApi& drv = driver_api();
    if (W > 1 && drv.ok) {
      // cuMem allocation + fd passing + peer mapping (NVLink loads/stores from kernels)
      int dev = 0;
      WGB_CUDA_TRY(cudaGetDevice(&dev));
      WGB_CUDA_TRY(cudaFree(nullptr));  // make sure the primary context exists
      CUmemAllocationProp prop;
      memset(&prop, 0, sizeof(prop));
      prop.type                 = CU_MEM_ALLOCATION_TYPE_PINNED;
      prop.location.type        = CU_MEM_LOCATION_TYPE_DEVICE;
      prop.location.id          = dev;
      prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      size_t gran               = 0;
      WGB_CU_TRY(drv.memGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
      size_t my_size = ((std::max<size_t>(local_bytes, 1) + gran - 1) / gran) * gran;
      CUmemGenericAllocationHandle mine;
      CUresult cr = drv.memCreate(&mine, my_size, &prop, 0);
      if (cr == CUDA_ERROR_OUT_OF_MEMORY) {
        delete h;
        throw std::bad_alloc();
      }
      WGB_CU_TRY(cr);
      int my_fd = -1;
      WGB_CU_TRY(drv.memExport(&my_fd, mine, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
      std::vector<size_t> sizes(W);
      comm_allgather(comm, &my_size, sizes.data(), sizeof(size_t));
      std::vector<int> fds(W, -1);
      exchange_fds(comm, my_fd, fds.data());
      CUmemAccessDesc access;
      memset(&access, 0, sizeof(access));
      access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      access.location.id   = dev;
      access.flags         = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      for (int r = 0; r < W; r++) {
        CUmemGenericAllocationHandle hr = mine;
        if (r != h->rank) WGB_CU_TRY(drv.memImport(&hr, (void*)(uintptr_t)fds[r], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
        CUdeviceptr va = 0;
        WGB_CU_TRY(drv.addrReserve(&va, sizes[r], gran, 0, 0));
        WGB_CU_TRY(drv.memMap(va, sizes[r], 0, hr, 0));
        WGB_CU_TRY(drv.memSetAccess(va, sizes[r], &access, 1));
        h->peer_ptr[r]  = reinterpret_cast<void*>(va);
        h->vmm_handle[r] = (unsigned long long)hr;
        h->vmm_size[r]   = sizes[r];
        if (r != h->rank && fds[r] >= 0) close(fds[r]);
      }
      close(my_fd);
      h->vmm         = true;
      h->local_ptr   = h->peer_ptr[h->rank];
      h->local_alloc = my_size;
      comm_barrier(comm);
    } else {
      h->local_alloc = std::max<size_t>(local_bytes, 256);
      cudaError_t e  = cudaMalloc(&h->local_ptr, h->local_alloc);
      if (e != cudaSuccess) {
        cudaGetLastError();
        delete h;
        throw std::bad_alloc();
      }
      h->peer_ptr[h->rank] = h->local_ptr;
      if (W > 1) {
        // driver without the VMM entry points: legacy cudaIpc (works, but peer rows travel the slow path)
        log_msg(LEVEL_WARN, "cuMem VMM API unavailable, falling back to cudaIpc peer mappings");
        cudaIpcMemHandle_t mine;
        WGB_CUDA_TRY(cudaIpcGetMemHandle(&mine, h->local_ptr));
        std::vector<cudaIpcMemHandle_t> all(W);
        comm_allgather(comm, &mine, all.data(), sizeof(mine));
        for (int r = 0; r < W; r++) {
          if (r == h->rank) continue;
          WGB_CUDA_TRY(cudaIpcOpenMemHandle(&h->peer_ptr[r], all[r], cudaIpcMemLazyEnablePeerAccess));
        }
        comm_barrier(comm);
      }
    }
    WGB_CUDA_TRY(cudaMalloc(&h->d_ptrs, sizeof(void*) * kMaxWorld));
    WGB_CUDA_TRY(cudaMalloc(&h->d_offsets, sizeof(size_t) * (kMaxWorld + 1)));
    WGB_CUDA_TRY(cudaMemcpy(h->d_ptrs, h->peer_ptr, sizeof(void*) * kMaxWorld, cudaMemcpyHostToDevice));
    WGB_CUDA_TRY(cudaMemcpy(h->d_offsets, h->chunk_start, sizeof(size_t) * (kMaxWorld + 1), cudaMemcpyHostToDevice));
    if (h->type == WHOLEMEMORY_MT_CONTINUOUS) h->flat_ptr = h->local_ptr;  // world == 1
    *handle_ptr = h;
  });
}

wholememory_error_code_t wholememory_free(wholememory_handle_t h)
{
  if (h == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  return guarded("wholememory_free", [&] {
    cudaDeviceSynchronize();
    if (h->vmm) {
      const DriverApi& drv = driver_api();
      comm_barrier(h->comm);  // nobody unmaps while a peer kernel may still read the chunk
      for (int r = 0; r < h->world; r++) {
        if (!h->peer_ptr[r]) continue;
        drv.memUnmap((CUdeviceptr)h->peer_ptr[r], h->vmm_size[r]);
        drv.addrFree((CUdeviceptr)h->peer_ptr[r], h->vmm_size[r]);
        drv.memRelease((CUmemGenericAllocationHandle)h->vmm_handle[r]);
      }
      comm_barrier(h->comm);
    } else {
      if (h->world > 1) {
        for (int r = 0; r < h->world; r++)
          if (r != h->rank && h->peer_ptr[r]) cudaIpcCloseMemHandle(h->peer_ptr[r]);
        comm_barrier(h->comm);  // nobody frees while a peer still has the chunk mapped
      }
      if (h->local_ptr) cudaFree(h->local_ptr);
    }
    if (h->d_ptrs) cudaFree(h->d_ptrs);
    if (h->d_offsets) cudaFree(h->d_offsets);
    cudaGetLastError();
    delete h;
  });
}

wholememory_error_code_t wholememory_get_communicator(wholememory_comm_t* comm, wholememory_handle_t h)
{
  if (!comm || !h) return WHOLEMEMORY_INVALID_INPUT;
  *comm = h->comm;
  return WHOLEMEMORY_SUCCESS;
}
wholememory_memory_type_t wholememory_get_memory_type(wholememory_handle_t h) { return h ? h->type : WHOLEMEMORY_MT_NONE; }
wholememory_memory_location_t wholememory_get_memory_location(wholememory_handle_t h)
{
  return h ? h->location : WHOLEMEMORY_ML_NONE;
}
wholememory_distributed_backend_t wholememory_get_distributed_backend(wholememory_handle_t h)
{
  return h ? h->comm->backend : WHOLEMEMORY_DB_NONE;
}
size_t wholememory_get_total_size(wholememory_handle_t h) { return h ? h->total_size : 0; }
size_t wholememory_get_data_granularity(wholememory_handle_t h) { return h ? h->granularity : 0; }

wholememory_error_code_t wholememory_get_local_memory(void** local_ptr, size_t* local_size, size_t* local_offset,
                                                      wholememory_handle_t h)
{
  if (!h) return WHOLEMEMORY_INVALID_INPUT;
  if (local_ptr) *local_ptr = h->local_ptr;
  if (local_size) *local_size = h->chunk_start[h->rank + 1] - h->chunk_start[h->rank];
  if (local_offset) *local_offset = h->chunk_start[h->rank];
  return WHOLEMEMORY_SUCCESS;
}
wholememory_error_code_t wholememory_get_local_size(size_t* local_size, wholememory_handle_t h)
{
  return wholememory_get_local_memory(nullptr, local_size, nullptr, h);
}
wholememory_error_code_t wholememory_get_local_offset(size_t* local_offset, wholememory_handle_t h)
{
  return wholememory_get_local_memory(nullptr, nullptr, local_offset, h);
}
wholememory_error_code_t wholememory_get_rank_memory(void** ptr, size_t* size, size_t* offset, int rank,
                                                     wholememory_handle_t h)
{
  if (!h || rank < 0 || rank >= h->world) return WHOLEMEMORY_INVALID_INPUT;
  if (ptr) *ptr = h->peer_ptr[rank];
  if (size) *size = h->chunk_start[rank + 1] - h->chunk_start[rank];
  if (offset) *offset = h->chunk_start[rank];
  return WHOLEMEMORY_SUCCESS;
}
wholememory_error_code_t wholememory_get_global_pointer(void** global_ptr, wholememory_handle_t h)
{
  if (!h || !global_ptr) return WHOLEMEMORY_INVALID_INPUT;
  if (h->type != WHOLEMEMORY_MT_CONTINUOUS) return WHOLEMEMORY_INVALID_INPUT;
  *global_ptr = h->flat_ptr;
  return WHOLEMEMORY_SUCCESS;
}
wholememory_error_code_t wholememory_get_global_reference(wholememory_gref_t* gref, wholememory_handle_t h)
{
  if (!h || !gref) return WHOLEMEMORY_INVALID_INPUT;
  if (h->type == WHOLEMEMORY_MT_CONTINUOUS) {
    *gref = wholememory_create_continuous_global_reference(h->flat_ptr);
    return WHOLEMEMORY_SUCCESS;
  }
  gref->pointer             = h->d_ptrs;
  gref->rank_memory_offsets = h->d_offsets;
  gref->world_size          = h->world;
  gref->stride              = h->stride;
  gref->same_chunk          = h->same_chunk;
  return WHOLEMEMORY_SUCCESS;
}
wholememory_error_code_t wholememory_get_rank_partition_sizes(size_t* sizes, wholememory_handle_t h)
{
  if (!h || !sizes) return WHOLEMEMORY_INVALID_INPUT;
  for (int r = 0; r < h->world; r++)
    sizes[r] = h->chunk_start[r + 1] - h->chunk_start[r];
  return WHOLEMEMORY_SUCCESS;
}
wholememory_error_code_t wholememory_get_rank_partition_offsets(size_t* offsets, wholememory_handle_t h)
{
  if (!h || !offsets) return WHOLEMEMORY_INVALID_INPUT;
  for (int r = 0; r <= h->world; r++)
    offsets[r] = h->chunk_start[r];
  return WHOLEMEMORY_SUCCESS;
}

// ---- binary part files (reference: cpp/src/wholememory/file_io.cpp:1849-2165) ----------
// Each rank reads exactly the byte range of its own chunk from the concatenation of the files
// (round_robin_size must be 0, as cugraph-pyg uses it), through a pinned bounce buffer.
wholememory_error_code_t wholememory_load_from_file(wholememory_handle_t h, size_t memory_offset,
                                                    size_t memory_entry_size, size_t file_entry_size,
                                                    const char** file_names, int file_count, int round_robin_size)
{
  return guarded("wholememory_load_from_file", [&] {
    WGB_CHECK_INPUT(h != nullptr && file_names != nullptr && file_count > 0, "bad arguments");
    WGB_CHECK_INPUT(round_robin_size == 0, "round_robin_size != 0 is not supported");
    WGB_CHECK_INPUT(file_entry_size > 0 && memory_entry_size >= file_entry_size + memory_offset % memory_entry_size,
                    "entry sizes inconsistent");
    WGB_CHECK_INPUT(h->granularity % memory_entry_size == 0 || memory_entry_size % h->granularity == 0,
                    "memory_entry_size must align with the handle's granularity");
    std::vector<size_t> fsize(file_count);
    size_t total_file = 0;
    for (int i = 0; i < file_count; i++) {
      struct stat st;
      if (stat(file_names[i], &st) != 0) throw invalid_input(std::string("cannot stat ") + file_names[i]);
      WGB_CHECK_INPUT((size_t)st.st_size % file_entry_size == 0, "file size is not a multiple of file_entry_size");
      fsize[i] = (size_t)st.st_size;
      total_file += fsize[i];
    }
    size_t total_entries = total_file / file_entry_size;
    size_t local_start   = h->chunk_start[h->rank];
    size_t local_end     = h->chunk_start[h->rank + 1];
    size_t e_begin       = (local_start + memory_entry_size - 1) / memory_entry_size;
    size_t e_end         = std::min(total_entries, local_end / memory_entry_size);
    // entries that straddle a chunk boundary cannot happen: granularity aligns with the entry size
    const size_t kBuf = 8u << 20;
    void* bounce      = nullptr;
    WGB_CUDA_TRY(cudaMallocHost(&bounce, kBuf));
    size_t entries_per_buf = std::max<size_t>(1, kBuf / file_entry_size);
    int fi                 = 0;
    size_t fbase           = 0;  // first entry of file fi
    FILE* fp               = nullptr;
    try {
      for (size_t e = e_begin; e < e_end;) {
        while (fi < file_count && e >= fbase + fsize[fi] / file_entry_size) {
          fbase += fsize[fi] / file_entry_size;
          fi++;
          if (fp) {
            fclose(fp);
            fp = nullptr;
          }
        }
        if (fi >= file_count) break;
        if (!fp) {
          fp = fopen(file_names[fi], "rb");
          if (!fp) throw invalid_input(std::string("cannot open ") + file_names[fi]);
        }
        size_t in_file = e - fbase;
        size_t n       = std::min({entries_per_buf, e_end - e, fsize[fi] / file_entry_size - in_file});
        if (fseeko(fp, (off_t)(in_file * file_entry_size), SEEK_SET) != 0) throw std::runtime_error("fseek failed");
        if (fread(bounce, file_entry_size, n, fp) != n) throw std::runtime_error("short read");
        char* dst = static_cast<char*>(h->local_ptr) + (e * memory_entry_size - local_start) + memory_offset;
        WGB_CUDA_TRY(cudaMemcpy2D(dst, memory_entry_size, bounce, file_entry_size, file_entry_size, n,
                                  cudaMemcpyHostToDevice));
        e += n;
      }
    } catch (...) {
      if (fp) fclose(fp);
      cudaFreeHost(bounce);
      throw;
    }
    if (fp) fclose(fp);
    cudaFreeHost(bounce);
    comm_barrier(h->comm);
  });
}

wholememory_error_code_t wholememory_store_to_file(wholememory_handle_t h, size_t memory_offset,
                                                   size_t memory_entry_stride, size_t file_entry_size,
                                                   const char* local_file_name)
{
  return guarded("wholememory_store_to_file", [&] {
    WGB_CHECK_INPUT(h != nullptr && local_file_name != nullptr, "bad arguments");
    WGB_CHECK_INPUT(file_entry_size > 0 && memory_entry_stride >= file_entry_size, "entry sizes inconsistent");
    size_t local_bytes = h->chunk_start[h->rank + 1] - h->chunk_start[h->rank];
    size_t n_entries   = local_bytes / memory_entry_stride;
    FILE* fp           = fopen(local_file_name, "wb");
    if (!fp) throw invalid_input(std::string("cannot open ") + local_file_name);
    const size_t kBuf = 8u << 20;
    void* bounce      = nullptr;
    if (cudaMallocHost(&bounce, kBuf) != cudaSuccess) {
      fclose(fp);
      throw cuda_error("cudaMallocHost failed");
    }
    size_t per = std::max<size_t>(1, kBuf / file_entry_size);
    try {
      for (size_t e = 0; e < n_entries; e += per) {
        size_t n        = std::min(per, n_entries - e);
        const char* src = static_cast<const char*>(h->local_ptr) + e * memory_entry_stride + memory_offset;
        WGB_CUDA_TRY(cudaMemcpy2D(bounce, file_entry_size, src, memory_entry_stride, file_entry_size, n,
                                  cudaMemcpyDeviceToHost));
        if (fwrite(bounce, file_entry_size, n, fp) != n) throw std::runtime_error("short write");
      }
    } catch (...) {
      fclose(fp);
      cudaFreeHost(bounce);
      throw;
    }
    fclose(fp);
    cudaFreeHost(bounce);
  });
}

// ---------------------------------------------------------------------------------------------
// tensors
// ---------------------------------------------------------------------------------------------
static std::atomic<int64_t> g_tensor_count{0};
int64_t get_wholememory_tensor_count() { return g_tensor_count.load(); }

static bool desc_ok(const wholememory_tensor_description_t* d)
{
  if (d == nullptr) return false;
  if (d->dim != 1 && d->dim != 2) return false;
  if (wholememory_dtype_get_element_size(d->dtype) == static_cast<size_t>(-1)) return false;
  if (d->strides[d->dim - 1] != 1) return false;
  return true;
}

wholememory_error_code_t wholememory_create_tensor(wholememory_tensor_t* tensor,
                                                   wholememory_tensor_description_t* desc, wholememory_comm_t comm,
                                                   wholememory_memory_type_t memory_type,
                                                   wholememory_memory_location_t memory_location,
                                                   size_t* tensor_entry_partition)
{
  if (!tensor || !desc_ok(desc) || !comm) return WHOLEMEMORY_INVALID_INPUT;
  if (desc->storage_offset != 0) return WHOLEMEMORY_INVALID_INPUT;
  if (desc->dim == 2 && desc->strides[0] < desc->sizes[1]) return WHOLEMEMORY_INVALID_INPUT;
  size_t elt  = wholememory_dtype_get_element_size(desc->dtype);
  size_t gran = (desc->dim == 2 ? (size_t)desc->strides[0] : 1) * elt;
  size_t total = (size_t)desc->sizes[0] * gran;
  wholememory_handle_t h = nullptr;
  auto err = wholememory_malloc(&h, total, comm, memory_type, memory_location, gran, tensor_entry_partition);
  if (err != WHOLEMEMORY_SUCCESS) return err;
  auto* t           = new wholememory_tensor_();
  t->desc           = *desc;
  t->handle         = h;
  t->own_handle     = true;
  t->is_wholememory = true;
  t->root           = nullptr;
  g_tensor_count++;
  *tensor = t;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_destroy_tensor(wholememory_tensor_t t)
{
  if (t == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_error_code_t err = WHOLEMEMORY_SUCCESS;
  if (t->own_handle && t->handle) err = wholememory_free(t->handle);
  delete t;
  g_tensor_count--;
  return err;
}

wholememory_error_code_t wholememory_make_tensor_from_pointer(wholememory_tensor_t* tensor, void* storage_ptr,
                                                              wholememory_tensor_description_t* desc)
{
  if (!tensor || !desc) return WHOLEMEMORY_INVALID_INPUT;
  if (desc->dim < 0 || desc->dim > 2) return WHOLEMEMORY_INVALID_INPUT;
  bool empty = false;
  for (int i = 0; i < desc->dim; i++)
    if (desc->sizes[i] == 0) empty = true;
  if (!empty && desc->dim >= 1 && desc->strides[desc->dim - 1] != 1) {
    log_msg(LEVEL_ERROR, "last stride must be 1");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (storage_ptr == nullptr && wholememory_get_memory_element_count_from_tensor(desc) > 0 && desc->dim > 0 &&
      desc->sizes[0] > 0)
    return WHOLEMEMORY_INVALID_INPUT;
  auto* t           = new wholememory_tensor_();
  t->desc           = *desc;
  t->storage_ptr    = storage_ptr;
  t->is_wholememory = false;
  g_tensor_count++;
  *tensor = t;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_make_tensor_from_handle(wholememory_tensor_t* tensor, wholememory_handle_t h,
                                                             wholememory_tensor_description_t* desc)
{
  if (!tensor || !h || !desc_ok(desc)) return WHOLEMEMORY_INVALID_INPUT;
  auto* t           = new wholememory_tensor_();
  t->desc           = *desc;
  t->handle         = h;
  t->own_handle     = false;
  t->is_wholememory = true;
  g_tensor_count++;
  *tensor = t;
  return WHOLEMEMORY_SUCCESS;
}

bool wholememory_tensor_has_handle(wholememory_tensor_t t)
{
  if (!t) return false;
  wholememory_tensor_t root = t->root ? t->root : t;
  return root->is_wholememory;
}

wholememory_handle_t wholememory_tensor_get_memory_handle(wholememory_tensor_t t)
{
  if (!t) return nullptr;
  wholememory_tensor_t root = t->root ? t->root : t;
  return root->handle;
}

wholememory_tensor_description_t* wholememory_tensor_get_tensor_description(wholememory_tensor_t t)
{
  return t ? &t->desc : nullptr;
}

wholememory_error_code_t wholememory_tensor_get_global_reference(wholememory_tensor_t t, wholememory_gref_t* gref)
{
  if (!t || !gref) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_tensor_t root = t->root ? t->root : t;
  if (root->is_wholememory) return wholememory_get_global_reference(gref, root->handle);
  *gref = wholememory_create_continuous_global_reference(root->storage_ptr);
  return WHOLEMEMORY_SUCCESS;
}

void* wholememory_tensor_get_data_pointer(wholememory_tensor_t t)
{
  if (!t) return nullptr;
  wholememory_tensor_t root = t->root ? t->root : t;
  char* base                = nullptr;
  if (root->is_wholememory) {
    if (root->handle->type != WHOLEMEMORY_MT_CONTINUOUS) return nullptr;
    base = static_cast<char*>(root->handle->flat_ptr);
  } else {
    base = static_cast<char*>(root->storage_ptr);
  }
  if (base == nullptr) return nullptr;
  return base + t->desc.storage_offset * (int64_t)wholememory_dtype_get_element_size(t->desc.dtype);
}

static size_t tensor_entry_bytes(wholememory_tensor_t root)
{
  size_t elt = wholememory_dtype_get_element_size(root->desc.dtype);
  return (root->desc.dim == 2 ? (size_t)root->desc.strides[0] : 1) * elt;
}

wholememory_error_code_t wholememory_tensor_get_entry_offsets(size_t* entry_offsets, wholememory_tensor_t t)
{
  if (!t || !entry_offsets || !wholememory_tensor_has_handle(t)) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_tensor_t root = t->root ? t->root : t;
  size_t eb                 = tensor_entry_bytes(root);
  for (int r = 0; r <= root->handle->world; r++)
    entry_offsets[r] = root->handle->chunk_start[r] / eb;
  return WHOLEMEMORY_SUCCESS;
}
wholememory_error_code_t wholememory_tensor_get_entry_partition_sizes(size_t* entry_partition, wholememory_tensor_t t)
{
  if (!t || !entry_partition || !wholememory_tensor_has_handle(t)) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_tensor_t root = t->root ? t->root : t;
  size_t eb                 = tensor_entry_bytes(root);
  for (int r = 0; r < root->handle->world; r++)
    entry_partition[r] = (root->handle->chunk_start[r + 1] - root->handle->chunk_start[r]) / eb;
  return WHOLEMEMORY_SUCCESS;
}
wholememory_error_code_t wholememory_tensor_get_local_entry_count(size_t* count, wholememory_tensor_t t)
{
  if (!t || !count || !wholememory_tensor_has_handle(t)) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_tensor_t root = t->root ? t->root : t;
  auto* h                   = root->handle;
  *count                    = (h->chunk_start[h->rank + 1] - h->chunk_start[h->rank]) / tensor_entry_bytes(root);
  return WHOLEMEMORY_SUCCESS;
}
wholememory_error_code_t wholememory_tensor_get_local_entry_start(size_t* start, wholememory_tensor_t t)
{
  if (!t || !start || !wholememory_tensor_has_handle(t)) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_tensor_t root = t->root ? t->root : t;
  auto* h                   = root->handle;
  *start                    = h->chunk_start[h->rank] / tensor_entry_bytes(root);
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_tensor_map_local_tensor(wholememory_tensor_t t, wholememory_tensor_t* local)
{
  if (!t || !local || !wholememory_tensor_has_handle(t)) return WHOLEMEMORY_INVALID_INPUT;
  if (t->desc.dim == 1 && t->desc.storage_offset != 0) return WHOLEMEMORY_INVALID_INPUT;
  if (t->desc.dim == 2 && t->desc.storage_offset + t->desc.sizes[1] > t->desc.strides[0]) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_tensor_t root = t->root ? t->root : t;
  auto* h                   = root->handle;
  size_t cnt = 0;
  wholememory_tensor_get_local_entry_count(&cnt, t);
  wholememory_tensor_description_t d = t->desc;
  d.sizes[0]                         = (int64_t)cnt;
  return wholememory_make_tensor_from_pointer(local, h->local_ptr, &d);
}

wholememory_error_code_t wholememory_tensor_get_subtensor(wholememory_tensor_t t, int64_t* starts, int64_t* ends,
                                                          wholememory_tensor_t* sub)
{
  if (!t || !starts || !ends || !sub) return WHOLEMEMORY_INVALID_INPUT;
  if (t->desc.dim != 1 && t->desc.dim != 2) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_tensor_description_t d = t->desc;
  int64_t off                        = d.storage_offset;
  for (int i = 0; i < d.dim; i++) {
    int64_t s = starts[i] < 0 ? 0 : starts[i];
    int64_t e = ends[i] < 0 ? d.sizes[i] : ends[i];
    if (s > e || e > d.sizes[i]) return WHOLEMEMORY_INVALID_VALUE;
    off += s * d.strides[i];
    d.sizes[i] = e - s;
  }
  d.storage_offset = off;
  auto* n          = new wholememory_tensor_();
  n->desc          = d;
  n->root          = t->root ? t->root : t;
  n->storage_ptr   = nullptr;
  n->handle        = nullptr;
  g_tensor_count++;
  *sub = n;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_tensor_t wholememory_tensor_get_root(wholememory_tensor_t t) { return t ? (t->root ? t->root : t) : nullptr; }

// ---------------------------------------------------------------------------------------------
// default allocation callbacks (plain cudaMalloc; Python supplies torch-backed ones instead)
// ---------------------------------------------------------------------------------------------
struct default_temp_ctx {
  void* ptr    = nullptr;
  int location = 0;
};
static void def_create_ctx(void** ctx, void*) { *ctx = new default_temp_ctx(); }
static void def_destroy_ctx(void* ctx, void*) { delete static_cast<default_temp_ctx*>(ctx); }
static void* def_alloc_bytes(size_t bytes, wholememory_memory_allocation_type_t where)
{
  void* p = nullptr;
  if (bytes == 0) bytes = 8;
  if (where == WHOLEMEMORY_MA_DEVICE) {
    if (cudaMalloc(&p, bytes) != cudaSuccess) p = nullptr;
  } else if (where == WHOLEMEMORY_MA_PINNED) {
    if (cudaMallocHost(&p, bytes) != cudaSuccess) p = nullptr;
  } else {
    p = malloc(bytes);
  }
  return p;
}
static void def_free_bytes(void* p, int where)
{
  if (!p) return;
  if (where == WHOLEMEMORY_MA_DEVICE)
    cudaFree(p);
  else if (where == WHOLEMEMORY_MA_PINNED)
    cudaFreeHost(p);
  else
    free(p);
}
static void* def_temp_malloc(wholememory_tensor_description_t* d, wholememory_memory_allocation_type_t where, void* ctx,
                             void*)
{
  auto* c     = static_cast<default_temp_ctx*>(ctx);
  c->ptr      = def_alloc_bytes((size_t)wholememory_get_memory_size_from_tensor(d), where);
  c->location = where;
  return c->ptr;
}
static void def_temp_free(void* ctx, void*)
{
  auto* c = static_cast<default_temp_ctx*>(ctx);
  def_free_bytes(c->ptr, c->location);
  c->ptr = nullptr;
}
static void* def_out_malloc(wholememory_tensor_description_t* d, wholememory_memory_allocation_type_t where, void* ctx,
                            void*)
{
  auto* o = static_cast<wholememory_default_output_t*>(ctx);
  o->ptr  = def_alloc_bytes((size_t)wholememory_get_memory_size_from_tensor(d), where);
  o->desc = *d;
  return o->ptr;
}
static void def_out_free(void* ctx, void*)
{
  auto* o = static_cast<wholememory_default_output_t*>(ctx);
  if (o->ptr) cudaFree(o->ptr);
  o->ptr = nullptr;
}

wholememory_env_func_t* wholememory_get_default_env_func()
{
  static wholememory_env_func_t env = {
    {def_create_ctx, def_destroy_ctx, def_temp_malloc, def_temp_free, nullptr},
    {def_out_malloc, def_out_free, nullptr},
  };
  return &env;
}
void wholememory_default_output_release(wholememory_default_output_t* out) { def_out_free(out, nullptr); }

// ---------------------------------------------------------------------------------------------
// non-cached embedding (reference: cpp/src/wholememory/embedding.cpp:545-554, 1045-1073)
// ---------------------------------------------------------------------------------------------
wholememory_error_code_t wholememory_create_embedding(wholememory_embedding_t* emb,
                                                      wholememory_tensor_description_t* desc, wholememory_comm_t comm,
                                                      wholememory_memory_type_t memory_type,
                                                      wholememory_memory_location_t memory_location,
                                                      wholememory_embedding_cache_policy_t cache_policy,
                                                      size_t* embedding_entry_partition, int user_defined_sms,
                                                      int round_robin_size)
{
  if (!emb || !desc || !comm) return WHOLEMEMORY_INVALID_INPUT;
  if (cache_policy != nullptr) return WHOLEMEMORY_NOT_SUPPORTED;
  if (round_robin_size != 0) return WHOLEMEMORY_NOT_SUPPORTED;
  if (desc->dim != 2) return WHOLEMEMORY_INVALID_INPUT;
  wholememory_tensor_t t = nullptr;
  auto err = wholememory_create_tensor(&t, desc, comm, memory_type, memory_location, embedding_entry_partition);
  if (err != WHOLEMEMORY_SUCCESS) return err;
  auto* e     = new wholememory_embedding_();
  e->tensor   = t;
  e->user_sms = user_defined_sms;
  *emb        = e;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_destroy_embedding(wholememory_embedding_t e)
{
  if (!e) return WHOLEMEMORY_INVALID_INPUT;
  wgb::embedding_release_training_state(e);  // optimizer states, inbox, scratch (collective, like the table itself)
  wgb::embedding_drop_hot_rows(e);
  auto err = wholememory_destroy_tensor(e->tensor);
  delete e;
  return err;
}

wholememory_tensor_t wholememory_embedding_get_embedding_tensor(wholememory_embedding_t e) { return e ? e->tensor : nullptr; }

wholememory_error_code_t wholememory_embedding_gather(wholememory_embedding_t e, wholememory_tensor_t indices,
                                                      wholememory_tensor_t output, bool /*adjust_cache*/,
                                                      wholememory_env_func_t* p_env_fns, int64_t stream_int)
{
  if (!e) return WHOLEMEMORY_INVALID_INPUT;
  return wgb::rows_op(e->tensor, indices, output, reinterpret_cast<void*>(stream_int), e->user_sms, false,
                      e->hot.slot ? &e->hot : nullptr);
}

}  // extern "C"

// ---- replicated hot rows (b200_ops.h) --------------------------------------------------------------------------------
namespace wgb {
__global__ void hot_slot_kernel(const long long* __restrict__ idx, long long n, long long rows, int* __restrict__ slot)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = idx[i];
    if (r >= 0 && r < rows) slot[r] = (int)i;  // duplicates: any of the copies is as good as another
  }
}

// embeddings that currently hold a replica: scatters into their table drop it (a replica of rows that have changed would
// serve stale rows silently)
static std::mutex g_hot_mutex;
static std::vector<wholememory_embedding_t> g_hot_embeddings;

void hot_rows_invalidate_for_tensor(wholememory_tensor_t written)
{
  if (!written) return;
  std::vector<wholememory_embedding_t> hit;
  {
    std::lock_guard<std::mutex> lock(g_hot_mutex);
    if (g_hot_embeddings.empty()) return;
    wholememory_tensor_t wroot = written->root ? written->root : written;
    for (auto* e : g_hot_embeddings) {
      wholememory_tensor_t eroot = e->tensor && e->tensor->root ? e->tensor->root : e->tensor;
      if (eroot == wroot || (eroot && wroot && eroot->handle && eroot->handle == wroot->handle)) hit.push_back(e);
    }
  }
  for (auto* e : hit) {
    log_msg(LEVEL_WARN, "scatter into a table with %lld replicated hot rows: the replica on this GPU is dropped "
            "(replicas on other GPUs are theirs to refresh: call set_hot_rows again on every rank)", e->hot_count);
    cudaDeviceSynchronize();  // gathers that still read the replica finish first
    embedding_drop_hot_rows(e);
  }
}

void embedding_drop_hot_rows(wholememory_embedding_t e)
{
  {
    std::lock_guard<std::mutex> lock(g_hot_mutex);
    for (size_t i = 0; i < g_hot_embeddings.size(); i++)
      if (g_hot_embeddings[i] == e) {
        g_hot_embeddings.erase(g_hot_embeddings.begin() + i);
        break;
      }
  }
  if (e->hot_slot_mem) cudaFree(e->hot_slot_mem);
  if (e->hot_rows_mem) cudaFree(e->hot_rows_mem);
  e->hot_slot_mem = e->hot_rows_mem = nullptr;
  e->hot                            = wgb_hot_rows();
  e->hot_count                      = 0;
  cudaGetLastError();
}
}  // namespace wgb

extern "C" {

wholememory_error_code_t wholememory_embedding_set_hot_rows(wholememory_embedding_t e, wholememory_tensor_t hot_indices, void* stream)
{
  using namespace wgb;
  if (!e) return WHOLEMEMORY_INVALID_INPUT;
  return guarded("wholememory_embedding_set_hot_rows", [&] {
    cudaStream_t st = as_stream(stream);
    WGB_CUDA_TRY(cudaStreamSynchronize(st));
    embedding_drop_hot_rows(e);
    if (!hot_indices) return;
    auto* id = wholememory_tensor_get_tensor_description(hot_indices);
    WGB_CHECK_INPUT(id->dim == 1 && id->dtype == WHOLEMEMORY_DT_INT64, "hot_indices must be a 1-D int64 device tensor");
    const long long n = id->sizes[0];
    if (n == 0) return;
    WGB_CHECK_INPUT(n < (1LL << 31), "too many hot rows");
    auto* td = wholememory_tensor_get_tensor_description(e->tensor);
    const long long rows = td->sizes[0];
    const size_t row_bytes = (size_t)td->sizes[1] * wholememory_dtype_get_element_size(td->dtype);
    const size_t stride    = (row_bytes + 15) / 16 * 16;
    WGB_CUDA_TRY(cudaMalloc(&e->hot_slot_mem, sizeof(int) * (size_t)rows));
    WGB_CUDA_TRY(cudaMalloc(&e->hot_rows_mem, stride * (size_t)n));
    WGB_CUDA_TRY(cudaMemsetAsync(e->hot_slot_mem, 0xFF, sizeof(int) * (size_t)rows, st));
    const long long* idx = static_cast<const long long*>(wholememory_tensor_get_data_pointer(hot_indices)) + id->storage_offset;
    hot_slot_kernel<<<(int)std::min<long long>((n + 255) / 256, 1184), 256, 0, st>>>(idx, n, rows, static_cast<int*>(e->hot_slot_mem));
    WGB_CHECK_LAUNCH();
    // fill the replica with a plain gather (peer rows arrive over NVLink once)
    wholememory_tensor_description_t od = *td;
    od.sizes[0]       = n;
    od.strides[0]     = (int64_t)(stride / wholememory_dtype_get_element_size(td->dtype));
    od.strides[1]     = 1;
    od.storage_offset = 0;
    wholememory_tensor_t out = nullptr;
    WGB_EXPECTS(wholememory_make_tensor_from_pointer(&out, e->hot_rows_mem, &od) == WHOLEMEMORY_SUCCESS, "wrap replica");
    auto err = rows_op(e->tensor, hot_indices, out, st, e->user_sms, false, nullptr);
    wholememory_destroy_tensor(out);
    WGB_EXPECTS(err == WHOLEMEMORY_SUCCESS, "could not fill the hot-row replica");
    WGB_CUDA_TRY(cudaStreamSynchronize(st));
    e->hot.slot         = static_cast<const int*>(e->hot_slot_mem);
    e->hot.rows         = static_cast<const char*>(e->hot_rows_mem);
    e->hot.stride_bytes = stride;
    e->hot_count        = n;
    {
      std::lock_guard<std::mutex> lock(g_hot_mutex);
      g_hot_embeddings.push_back(e);
    }
  });
}

long long wholememory_embedding_hot_row_count(wholememory_embedding_t e) { return e ? e->hot_count : 0; }

}  // extern "C"
