/*
 * Core C ABI: error codes, memory types, communicator, WholeMemory handles.
 *
 * Drop-in for the part of libwholegraph's C API the sampler / gather hot path needs.
 * Every declaration keeps the reference's name, argument order and meaning:
 *   error codes                /root/reference/cpp/include/wholememory/wholememory.h:21-33
 *   memory type / location     :50-66
 *   init / finalize            :91,97
 *   unique id + communicator   :116-148,170-245
 *   wholememory_malloc / free  :264-278
 *   handle queries             :285-409
 *
 * B200-first scope: one process per GPU on ONE NVSwitch box.  Every peer is
 * reachable by load/store, so all device memory types are backed by the same peer-mapped
 * (cudaIpc) chunk layout and are read by P2P from inside the kernels; there is no NCCL on the
 * data path.  The communicator is a shared-memory rendezvous between the processes of the box
 * (no NCCL, no sockets); multi-node features return WHOLEMEMORY_NOT_SUPPORTED.
 */
#pragma once

#include <stdio.h>
#include <unistd.h>

#include <wholememory/global_reference.h>

#ifdef __cplusplus
extern "C" {
#endif

enum wholememory_error_code_t {
  WHOLEMEMORY_SUCCESS = 0,
  WHOLEMEMORY_UNKNOW_ERROR,
  WHOLEMEMORY_NOT_IMPLEMENTED,
  WHOLEMEMORY_LOGIC_ERROR,
  WHOLEMEMORY_CUDA_ERROR,
  WHOLEMEMORY_COMMUNICATION_ERROR,
  WHOLEMEMORY_INVALID_INPUT,
  WHOLEMEMORY_INVALID_VALUE,
  WHOLEMEMORY_OUT_OF_MEMORY,
  WHOLEMEMORY_NOT_SUPPORTED,
  WHOLEMEMORY_SYSTEM_ERROR,
};

#define WHOLEMEMORY_RETURN_ON_FAIL(X)                                                    \
  do {                                                                                   \
    auto err__ = X;                                                                      \
    if (err__ != WHOLEMEMORY_SUCCESS) {                                                  \
      fprintf(stderr, "File %s line %d %s failed.\n", __FILE__, __LINE__, #X);           \
      return err__;                                                                      \
    }                                                                                    \
  } while (0)

enum wholememory_memory_type_t {
  WHOLEMEMORY_MT_NONE = 0,
  WHOLEMEMORY_MT_CONTINUOUS,  /* one flat VA range over all ranks (cuMem VMM) */
  WHOLEMEMORY_MT_CHUNKED,     /* one base pointer per rank (cudaIpc) */
  WHOLEMEMORY_MT_DISTRIBUTED, /* reference: private memory + NCCL all-to-all; here: peer-mapped like CHUNKED */
  WHOLEMEMORY_MT_HIERARCHY,   /* multi-node: not supported */
};

enum wholememory_memory_location_t {
  WHOLEMEMORY_ML_NONE = 0,
  WHOLEMEMORY_ML_DEVICE,
  WHOLEMEMORY_ML_HOST,
};

enum wholememory_distributed_backend_t {
  WHOLEMEMORY_DB_NONE = 0,
  WHOLEMEMORY_DB_NCCL,
  WHOLEMEMORY_DB_NVSHMEM,
};

enum LogLevel { LEVEL_FATAL = 0, LEVEL_ERROR, LEVEL_WARN, LEVEL_INFO, LEVEL_DEBUG, LEVEL_TRACE };

wholememory_error_code_t wholememory_init(unsigned int flags, LogLevel log_level = LEVEL_INFO);
wholememory_error_code_t wholememory_finalize();

typedef struct wholememory_comm_* wholememory_comm_t;

#define WHOLEMEMORY_UNIQUE_ID_BYTES (128)
struct wholememory_unique_id_t {
  char internal[WHOLEMEMORY_UNIQUE_ID_BYTES];
};

wholememory_error_code_t wholememory_create_unique_id(wholememory_unique_id_t* unique_id);
/* collective over the `size` processes that were handed the same unique_id */
wholememory_error_code_t wholememory_create_communicator(wholememory_comm_t* comm,
                                                         wholememory_unique_id_t unique_id,
                                                         int rank,
                                                         int size);
wholememory_error_code_t wholememory_split_communicator(wholememory_comm_t* new_comm,
                                                        wholememory_comm_t comm,
                                                        int color,
                                                        int key);
wholememory_error_code_t wholememory_destroy_communicator(wholememory_comm_t comm);
wholememory_error_code_t wholememory_communicator_support_type_location(
  wholememory_comm_t comm,
  wholememory_memory_type_t memory_type,
  wholememory_memory_location_t memory_location);
wholememory_error_code_t wholememory_communicator_get_rank(int* rank, wholememory_comm_t comm);
wholememory_error_code_t wholememory_communicator_get_size(int* size, wholememory_comm_t comm);
wholememory_error_code_t wholememory_communicator_get_local_size(int* local_size,
                                                                 wholememory_comm_t comm);
wholememory_error_code_t wholememory_communicator_set_distributed_backend(
  wholememory_comm_t comm, wholememory_distributed_backend_t distributed_backend);
wholememory_distributed_backend_t wholememory_communicator_get_distributed_backend(
  wholememory_comm_t comm);
wholememory_error_code_t wholememory_communicator_barrier(wholememory_comm_t comm);
bool wholememory_is_intranode_communicator(wholememory_comm_t comm);
bool wholememory_is_intra_mnnvl_communicator(wholememory_comm_t comm);
bool wholememory_is_build_with_nvshmem();

typedef struct wholememory_handle_* wholememory_handle_t;

/*
 * Collective allocation of `total_size` bytes striped over the ranks of `comm` in units of
 * `data_granularity` bytes.  rank_entry_partition (entries per rank, length world_size) may be
 * NULL, in which case rank r owns entries [r*ceil(N/W), (r+1)*ceil(N/W)).
 */
wholememory_error_code_t wholememory_malloc(wholememory_handle_t* wholememory_handle_ptr,
                                            size_t total_size,
                                            wholememory_comm_t comm,
                                            wholememory_memory_type_t memory_type,
                                            wholememory_memory_location_t memory_location,
                                            size_t data_granularity,
                                            size_t* rank_entry_partition = nullptr);
wholememory_error_code_t wholememory_free(wholememory_handle_t wholememory_handle);

wholememory_error_code_t wholememory_get_communicator(wholememory_comm_t* comm,
                                                      wholememory_handle_t wholememory_handle);
wholememory_memory_type_t wholememory_get_memory_type(wholememory_handle_t wholememory_handle);
wholememory_memory_location_t wholememory_get_memory_location(
  wholememory_handle_t wholememory_handle);
wholememory_distributed_backend_t wholememory_get_distributed_backend(
  wholememory_handle_t wholememory_handle);
size_t wholememory_get_total_size(wholememory_handle_t wholememory_handle);
size_t wholememory_get_data_granularity(wholememory_handle_t wholememory_handle);
wholememory_error_code_t wholememory_get_local_memory(void** local_ptr,
                                                      size_t* local_size,
                                                      size_t* local_offset,
                                                      wholememory_handle_t wholememory_handle);
wholememory_error_code_t wholememory_get_local_size(size_t* local_size,
                                                    wholememory_handle_t wholememory_handle);
wholememory_error_code_t wholememory_get_local_offset(size_t* local_offset,
                                                      wholememory_handle_t wholememory_handle);
wholememory_error_code_t wholememory_get_rank_memory(void** rank_memory_ptr,
                                                     size_t* rank_memory_size,
                                                     size_t* rank_memory_offset,
                                                     int rank,
                                                     wholememory_handle_t wholememory_handle);
wholememory_error_code_t wholememory_equal_entry_partition_plan(size_t* entry_per_rank,
                                                                size_t total_entry_count,
                                                                int world_size);
wholememory_error_code_t wholememory_get_global_pointer(void** global_ptr,
                                                        wholememory_handle_t wholememory_handle);
wholememory_error_code_t wholememory_get_global_reference(wholememory_gref_t* wholememory_gref,
                                                          wholememory_handle_t wholememory_handle);
wholememory_error_code_t wholememory_get_rank_partition_sizes(
  size_t* rank_mem_sizes, wholememory_handle_t wholememory_handle);
wholememory_error_code_t wholememory_get_rank_partition_offsets(
  size_t* rank_mem_offsets, wholememory_handle_t wholememory_handle);

int fork_get_device_count();

/* binary part-file IO (reference cpp/src/wholememory/file_io.cpp:1849,2048) */
wholememory_error_code_t wholememory_load_from_file(wholememory_handle_t wholememory_handle,
                                                    size_t memory_offset,
                                                    size_t memory_entry_size,
                                                    size_t file_entry_size,
                                                    const char** file_names,
                                                    int file_count,
                                                    int round_robin_size);
wholememory_error_code_t wholememory_store_to_file(wholememory_handle_t wholememory_handle,
                                                   size_t memory_offset,
                                                   size_t memory_entry_stride,
                                                   size_t file_entry_size,
                                                   const char* local_file_name);

#ifdef __cplusplus
}
#endif
