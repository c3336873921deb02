/*
 * G1 / G2: row gather and scatter over a (possibly multi-GPU) table.
 *
 * Replaces /root/reference/cpp/include/wholememory/wholememory_op.h:25-47 (same signatures);
 * kernels replaced: cpp/src/wholememory_ops/functions/gather_scatter_func.cuh:243-365, 509-587.
 *   out[i, :]        = cast(table[idx[i], :])      idx[i] < 0  -> row i of out is left untouched
 *   table[idx[i], :] = cast(in[i, :])              idx[i] < 0  -> skipped
 * table and out/in must both be floating (fp64/fp32/fp16/bf16) or both integer (int8..int64);
 * indices int32 or int64.  1-D tensors are treated as [N, 1].
 * Peer chunks are read/written by P2P from inside the kernel; no collective is issued, so the
 * call is rank-local for every memory type.
 */
#pragma once

#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_tensor.h>

#ifdef __cplusplus
extern "C" {
#endif

/* gather_sms / scatter_sms: upper bound on the SMs the copy kernel may occupy (-1: all 148; the grid is 8 CTAs per SM).
 * p_env_fns is accepted for signature compatibility: neither op allocates. */
wholememory_error_code_t wholememory_gather(wholememory_tensor_t wholememory_tensor, wholememory_tensor_t indices_tensor,
                                            wholememory_tensor_t output_tensor, wholememory_env_func_t* p_env_fns, void* stream,
                                            int gather_sms = -1);

wholememory_error_code_t wholememory_scatter(wholememory_tensor_t input_tensor, wholememory_tensor_t indices_tensor,
                                             wholememory_tensor_t wholememory_tensor, wholememory_env_func_t* p_env_fns,
                                             void* stream, int scatter_sms = -1);

#ifdef __cplusplus
}
#endif
