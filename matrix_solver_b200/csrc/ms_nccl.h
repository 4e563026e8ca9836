// ms_nccl.h — the exchange steps of the multi-GPU path (one rank = one process = one GPU).
// NCCL is bound lazily (dlopen of libnccl.so.2) so that single-GPU use never loads it and so
// that, inside a torch process, the already-loaded torch-bundled NCCL is the one that is used.
#pragma once
#include "ms_common.cuh"

int ms_nccl_unique_id (void *id128);
int ms_nccl_init (ms_ctx *c, const void *id128);
void ms_nccl_destroy (ms_ctx *c);
// step I of the factorisation: panel buffer columns >= I*m and d_I, from `owner` to everyone
int ms_nccl_bcast_panel (ms_ctx *c, size_t I, int pb, int owner, cudaStream_t s);
int ms_nccl_bcast (ms_ctx *c, double *p, size_t cnt, int root, cudaStream_t s);
int ms_nccl_reduce_sum (ms_ctx *c, double *p, size_t cnt, int root, cudaStream_t s);
int ms_nccl_allreduce_sum (ms_ctx *c, double *p, size_t cnt, cudaStream_t s);
int ms_nccl_allgather_bytes (ms_ctx *c, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t s);
// OR of a host flag over all ranks (check_res, src/init.cpp:565-581)
int ms_nccl_agree_flag (ms_ctx *c, int *flag);
