/*
 * B200-native extensions of the C ABI: entry points that have no counterpart in libwholegraph's
 * headers because the reference reaches the same functionality through other libraries
 * (pylibcugraph's multi-hop sampler, torch_geometric's aggregation) or not at all.
 * Plain handles, pointers and sizes only.
 */
#pragma once

#include <wholememory/embedding.h>
#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_tensor.h>

#ifdef __cplusplus
extern "C" {
#endif

/* number of CUDA kernels this library has launched in this process (bench.py "gpu_launches") */
unsigned long long wholememory_b200_kernel_launch_count();

/* ---------------------------------------------------------------------------------------------
 * S0: multi-hop, multi-label ("batch") neighbour sampling with renumbering in ONE call.
 *
 * Replaces the call cugraph-pyg makes into the un-vendored libcugraph,
 *   pylibcugraph.homogeneous_{uniform,biased}_neighbor_sample(..., renumber=True, retain_seeds=True,
 *   deduplicate_sources=True, prior_sources_behavior="exclude", return_hops=True, compression=COO|CSR)
 *   (/root/reference/python/cugraph-pyg/cugraph_pyg/sampler/distributed_sampler.py:784-819, 888-902)
 * and the per-hop python loop of the in-repo path
 *   (/root/reference/python/pylibwholegraph/pylibwholegraph/torch/graph_structure.py:136-196).
 *
 * The sampler object owns the persistent scratch (hash table keyed (label, vertex), frontier and edge
 * buffers); it is bound to the CUDA device that was current when it was created and must be used from
 * one stream at a time.
 *
 *   csr_row_ptr   int64 [V+1]           WholeMemory or local tensor (peer chunks are read by P2P)
 *   csr_col       int32|int64 [E]
 *   csr_weight    fp32|fp64 [E] or NULL  non-NULL selects weight-biased (A-Res) sampling
 *   csr_edge_id   int64 [E] or NULL      value reported as edge id (NULL: the CSR position)
 *   seeds         int32|int64 [S]        device; seeds of label l are seeds[label_offsets[l] : label_offsets[l+1]]
 *   label_offsets int64 [B+1]            device
 *   fanout[h]     > 0 sample, -1 take all, 0 skip the hop;  random_state: hop h uses random_state + h*0x9E3779B97F4A7C15
 *
 * Outputs (allocated through p_env_fns->output_fns, contexts as in wholegraph_op.h):
 *   majors / minors        int32 (int64 with WHOLEGRAPH_MULTIHOP_INT64_IDS) [E_s] label-local ids, grouped by
 *                          label then hop; local ids: seeds first, then vertices in first-occurrence order
 *   edge_id                int64 [E_s]
 *   label_hop_offsets      int64 [B*L+1]   COO: edge offsets.  CSR: offsets into major_offsets
 *   renumber_map           int64 [N_s]     global id of every local id, labels concatenated
 *   renumber_map_offsets   int64 [B+1]
 *   major_offsets          int64 [R+1]     only with WHOLEGRAPH_MULTIHOP_CSR (majors is then not produced)
 *   label_step_base        int32 [(L+1)*B] optional (ctx may be NULL): label_step_base[t*B + l] = first local id of the
 *                                          vertices label l discovered at step t (t = 0: seeds, t = h+1: hop h)
 * One host synchronisation per call (to size the outputs); take-all hops add one each.
 */
typedef struct wholegraph_multihop_sampler_* wholegraph_multihop_sampler_t;

#define WHOLEGRAPH_MULTIHOP_COO 0
#define WHOLEGRAPH_MULTIHOP_CSR 1
#define WHOLEGRAPH_MULTIHOP_INT64_IDS 2

wholememory_error_code_t wholegraph_create_multihop_sampler(wholegraph_multihop_sampler_t* sampler);
wholememory_error_code_t wholegraph_destroy_multihop_sampler(wholegraph_multihop_sampler_t sampler);

wholememory_error_code_t wholegraph_multihop_neighbor_sample(wholegraph_multihop_sampler_t sampler,
                                                             wholememory_tensor_t csr_row_ptr,
                                                             wholememory_tensor_t csr_col,
                                                             wholememory_tensor_t csr_weight,
                                                             wholememory_tensor_t csr_edge_id,
                                                             wholememory_tensor_t seeds,
                                                             wholememory_tensor_t label_offsets,
                                                             const int* fanout,
                                                             int num_hops,
                                                             unsigned long long random_state,
                                                             int flags,
                                                             void* out_majors_ctx,
                                                             void* out_minors_ctx,
                                                             void* out_edge_id_ctx,
                                                             void* out_label_hop_offsets_ctx,
                                                             void* out_renumber_map_ctx,
                                                             void* out_renumber_map_offsets_ctx,
                                                             void* out_major_offsets_ctx,
                                                             void* out_label_step_base_ctx,
                                                             wholememory_env_func_t* p_env_fns,
                                                             void* stream);

/* The same call in two halves, so that a loader can keep the GPU busy across call groups:
 *   _begin   enqueues every hop on `stream` and returns without waiting for the device;
 *   _finish  waits until the output sizes have reached the host (an event recorded by _begin, NOT the whole
 *            stream), allocates the outputs through the callbacks and enqueues the final scatter on `stream`.
 * Work enqueued between the two (e.g. _begin of the next call group on another sampler object, or the feature
 * gather of the previous one) overlaps the host wait.  One call may be in flight per sampler object; the graph
 * tensors must stay alive until _finish returns, seeds/label_offsets until the stream has run _begin's kernels
 * (stream-ordered frees, such as torch's caching allocator on the same stream, are fine). */
wholememory_error_code_t wholegraph_multihop_neighbor_sample_begin(wholegraph_multihop_sampler_t sampler,
                                                                   wholememory_tensor_t csr_row_ptr,
                                                                   wholememory_tensor_t csr_col,
                                                                   wholememory_tensor_t csr_weight,
                                                                   wholememory_tensor_t csr_edge_id,
                                                                   wholememory_tensor_t seeds,
                                                                   wholememory_tensor_t label_offsets,
                                                                   const int* fanout,
                                                                   int num_hops,
                                                                   unsigned long long random_state,
                                                                   int flags,
                                                                   void* stream);

wholememory_error_code_t wholegraph_multihop_neighbor_sample_finish(wholegraph_multihop_sampler_t sampler,
                                                                    void* out_majors_ctx,
                                                                    void* out_minors_ctx,
                                                                    void* out_edge_id_ctx,
                                                                    void* out_label_hop_offsets_ctx,
                                                                    void* out_renumber_map_ctx,
                                                                    void* out_renumber_map_offsets_ctx,
                                                                    void* out_major_offsets_ctx,
                                                                    void* out_label_step_base_ctx,
                                                                    wholememory_env_func_t* p_env_fns,
                                                                    void* stream);

/* Local id of every INPUT seed of the call that was last finished on `sampler` (duplicates included, int32 [S]; for a
 * heterogeneous call the id is local to the seed's (label, vertex type)).  The link-prediction loaders use it as
 * edge_label_index: the reference sorts and unique_consecutive's the endpoints of every batch in python to get the same
 * mapping (/root/reference/python/cugraph-pyg/cugraph_pyg/sampler/distributed_sampler.py:487-533, "a good target for a
 * C++ function").  Valid between _finish and the next _begin on the same sampler object. */
wholememory_error_code_t wholegraph_multihop_seed_local_ids(wholegraph_multihop_sampler_t sampler,
                                                            void* out_seed_local_id_ctx,
                                                            wholememory_env_func_t* p_env_fns,
                                                            void* stream);

/* Heterogeneous form: what cugraph-pyg asks of pylibcugraph.heterogeneous_{uniform,biased}_neighbor_sample
 * (/root/reference/python/cugraph-pyg/cugraph_pyg/sampler/distributed_sampler.py:53-94, 784-819) and decodes in
 * HeterogeneousSampleReader (/root/reference/python/cugraph-pyg/cugraph_pyg/sampler/sampler.py:280-490).
 *
 *   num_edge_types T (<= 16): one CSR per edge type over ONE global vertex id space (csr_row_ptr[t] int64 [V+1], csr_col[t],
 *     optional csr_weight[t] -- all types or none --, optional csr_edge_id[t]); all cols share one dtype
 *   vertex_type_offsets  host int64 [Vt+1] (Vt <= 16): vertex type vt owns the global ids [off[vt], off[vt+1])
 *   fanout  [num_hops * T], laid out [hop * T + edge type]  (neighbor_loader.py:192-201)
 *   hop h samples the whole frontier once per edge type with seed  random_state + h*0x9E3779B97F4A7C15 + t*0xD1B54A32D192ED03;
 *   edge order inside a hop is (frontier row, edge type, slot), vertices are deduplicated per label in that order.
 *
 * Outputs (COO only, like the reference):
 *   majors / minors           [E_s] ids local to (label, vertex type of the endpoint), hop-monotone, seeds first
 *   edge_id                   int64 [E_s] position inside the (label, edge type) group
 *   edge_type                 int32 [E_s]
 *   label_type_hop_offsets    int64 [B*T*L+1], groups ordered [label][edge type][hop]
 *   renumber_map              int64 [N_s] global ids; renumber_map_offsets int64 [B*Vt+1], segments [label][vertex type]
 *   edge_renumber_map         int64 [E_s] original edge id of every output edge (csr_edge_id value, else the CSR position);
 *   edge_renumber_map_offsets int64 [B*T+1]      => original id = edge_renumber_map[offsets[l*T+t] + edge_id]
 *   label_type_step_base      int32 [(L+1)*Vt*B] optional: first local id of the type-vt vertices label l discovered at step s
 */
wholememory_error_code_t wholegraph_hetero_multihop_neighbor_sample_begin(wholegraph_multihop_sampler_t sampler,
                                                                          int num_edge_types,
                                                                          const wholememory_tensor_t* csr_row_ptr,
                                                                          const wholememory_tensor_t* csr_col,
                                                                          const wholememory_tensor_t* csr_weight,
                                                                          const wholememory_tensor_t* csr_edge_id,
                                                                          const long long* vertex_type_offsets,
                                                                          int num_vertex_types,
                                                                          wholememory_tensor_t seeds,
                                                                          wholememory_tensor_t label_offsets,
                                                                          const int* fanout,
                                                                          int num_hops,
                                                                          unsigned long long random_state,
                                                                          int flags,
                                                                          void* stream);

wholememory_error_code_t wholegraph_hetero_multihop_neighbor_sample_finish(wholegraph_multihop_sampler_t sampler,
                                                                           void* out_majors_ctx,
                                                                           void* out_minors_ctx,
                                                                           void* out_edge_id_ctx,
                                                                           void* out_edge_type_ctx,
                                                                           void* out_label_type_hop_offsets_ctx,
                                                                           void* out_renumber_map_ctx,
                                                                           void* out_renumber_map_offsets_ctx,
                                                                           void* out_edge_renumber_map_ctx,
                                                                           void* out_edge_renumber_map_offsets_ctx,
                                                                           void* out_label_type_step_base_ctx,
                                                                           wholememory_env_func_t* p_env_fns,
                                                                           void* stream);

/* Temporal form: what cugraph-pyg asks of pylibcugraph.{homogeneous,heterogeneous}_{uniform,biased}_temporal_neighbor_sample
 * (/root/reference/python/cugraph-pyg/cugraph_pyg/sampler/distributed_sampler.py:56-80, 808-810, 897-900; behaviour
 * pinned by /root/reference/python/cugraph-pyg/cugraph_pyg/tests/loader/test_neighbor_loader.py:943-1058).
 *
 *   csr_edge_time[t]   int64 [E_t], the time of every edge of type t, in CSR order (required, one per edge type)
 *   seed_times         int64 [S] local device tensor, starting time of every seed (first occurrence wins for a repeated seed)
 *   time_comparison    0 strictly increasing, 1 monotonically increasing, 2 strictly decreasing, 3 monotonically
 *                      decreasing: an edge of a frontier vertex is eligible iff its time compares so with the vertex's
 *                      time; a sampled vertex takes the time of the edge that reached it first
 *   heterogeneous      != 0: arguments and outputs of the heterogeneous form, finish with
 *                      wholegraph_hetero_multihop_neighbor_sample_finish; 0: num_edge_types must be 1,
 *                      vertex_type_offsets is ignored, finish with wholegraph_multihop_neighbor_sample_finish
 *   The uniform one-hop algorithm runs over the eligible edges of a row in CSR order with the plain sampler's random
 *   streams: with every edge eligible the call returns exactly what the non-temporal call returns.
 *   csr_weight         NULL: uniform.  Otherwise fp32 | fp64 [E_t] per edge type (all types): the A-Res selection of the
 *                      biased sampler in which only eligible edges draw a key and compete; a row with at most `fanout`
 *                      eligible edges returns all of them in CSR order.  Same open-window identity.
 *   STATUS: compiled for sm_100a, not yet run on a GPU.
 */
wholememory_error_code_t wholegraph_temporal_multihop_neighbor_sample_begin(wholegraph_multihop_sampler_t sampler,
                                                                            int num_edge_types,
                                                                            const wholememory_tensor_t* csr_row_ptr,
                                                                            const wholememory_tensor_t* csr_col,
                                                                            const wholememory_tensor_t* csr_weight,
                                                                            const wholememory_tensor_t* csr_edge_time,
                                                                            const wholememory_tensor_t* csr_edge_id,
                                                                            const long long* vertex_type_offsets,
                                                                            int num_vertex_types,
                                                                            int heterogeneous,
                                                                            wholememory_tensor_t seeds,
                                                                            wholememory_tensor_t seed_times,
                                                                            wholememory_tensor_t label_offsets,
                                                                            const int* fanout,
                                                                            int num_hops,
                                                                            unsigned long long random_state,
                                                                            int time_comparison,
                                                                            int flags,
                                                                            void* stream);

/* ---------------------------------------------------------------------------------------------
 * Replicated hot rows: a read-only copy of the given rows of a WholeMemory embedding on THIS GPU.
 * wholememory_embedding_gather then serves those rows locally instead of from the owning rank (same bytes out, bit
 * for bit).  Role of the reference's device cache for remote tables (wholememory_create_embedding_cache_policy,
 * /root/reference/cpp/include/wholememory/embedding.h:82-111), made static: the caller names the rows (e.g. the
 * vertices of highest degree; on RMAT the hottest 10 % of the rows are 84 % of what a call group gathers) and pays
 * hot_rows * row_bytes + 4 bytes per table row of HBM on every GPU that calls it.
 *   hot_indices  int64 [H] local device tensor, or NULL to drop the replica.  Rank-local (not collective), synchronises
 *   `stream`.  The replica is a snapshot: call again after the table has been modified (scatter / file load);
 *   gradient apply drops it.
 */
wholememory_error_code_t wholememory_embedding_set_hot_rows(wholememory_embedding_t wholememory_embedding,
                                                            wholememory_tensor_t hot_indices,
                                                            void* stream);
long long wholememory_embedding_hot_row_count(wholememory_embedding_t wholememory_embedding);

/* ---------------------------------------------------------------------------------------------
 * A1: CSR neighbourhood aggregation of the sampled block (the SpMM that consumes the sampler output).
 *
 * The reference has no such kernel; its models use torch_geometric SAGEConv/GCNConv message passing over
 * the COO block (/root/reference/python/cugraph-pyg/cugraph_pyg/examples/gcn_dist_mnmg.py:239,
 * /root/reference/python/pylibwholegraph/pylibwholegraph/torch/gnn_model.py:119-125).
 *
 *   out[i, :] = reduce_{e in [indptr[i], indptr[i+1])} x[ map[indices[e]], : ]     (fp32 accumulate and output)
 *
 *   indptr     int32|int64 [n_dst+1]   local device tensor (the sampler's major_offsets)
 *   indices    int32|int64 [nnz]       local device tensor (the sampler's minors)
 *   gather_map int64 [*] or NULL       NULL: indices address rows of x directly.  Non-NULL (the renumber map):
 *                                      x is the full feature table and the gather is fused into the aggregation
 *   x          fp32|fp16|bf16 [n, F]   local or WholeMemory tensor (peer rows are read by P2P), rows 16-byte aligned
 *   out        fp32 [n_dst, F]         local device tensor
 * wholegraph_csr_aggregate_backward adds grad_out[i,:] (* 1/deg_i for mean) into grad_x[indices[e],:].
 */
#define WHOLEGRAPH_AGG_SUM 0
#define WHOLEGRAPH_AGG_MEAN 1

wholememory_error_code_t wholegraph_csr_aggregate(wholememory_tensor_t indptr,
                                                  wholememory_tensor_t indices,
                                                  wholememory_tensor_t gather_map,
                                                  wholememory_tensor_t x,
                                                  int reduce,
                                                  wholememory_tensor_t out,
                                                  void* stream);

wholememory_error_code_t wholegraph_csr_aggregate_backward(wholememory_tensor_t indptr,
                                                           wholememory_tensor_t indices,
                                                           wholememory_tensor_t grad_out,
                                                           int reduce,
                                                           wholememory_tensor_t grad_x,
                                                           void* stream);

/**
 * Fused GraphSAGE layer on a sampled CSR block, dense part on the tensor cores (csrc/sage_tile.cu: warp-per-row gather-mean
 * into a shared-memory bf16 operand tile, W brought by TMA, tcgen05.mma with the accumulator in tensor memory):
 *     out[i, :] = [ mean over e in [indptr[i], indptr[i+1]) of x[indices[e], :]  ||  x[i, :] ] . w_cat^T (+ bias)
 * Replaces, for one layer, the aggregation kernel + two GEMMs of the reference's consumer
 * (python/pylibwholegraph/pylibwholegraph/torch/gnn_model.py:119-125; PyG SAGEConv maths with aggr = "mean", root weight).
 *   indptr int32 | int64 [n_dst + 1], indices int32 | int64 [nnz]: the sampler's CSR block; destination rows are the first
 *   n_dst rows of x.   x bf16 [n_src, 128].   w_cat bf16 [F_out, 256] = [W_l || W_r], contiguous, F_out a multiple of 16 <= 256.
 *   bias fp32 [F_out] or NULL.   out fp32 [n_dst, F_out].   All local device tensors.
 * WHOLEMEMORY_NOT_IMPLEMENTED for other feature widths (use wholegraph_csr_aggregate + GEMMs).
 */
wholememory_error_code_t wholegraph_sage_layer_forward(wholememory_tensor_t indptr,
                                                       wholememory_tensor_t indices,
                                                       wholememory_tensor_t x,
                                                       wholememory_tensor_t w_cat,
                                                       wholememory_tensor_t bias,
                                                       wholememory_tensor_t out,
                                                       void* stream);

#ifdef __cplusplus
}
#endif
