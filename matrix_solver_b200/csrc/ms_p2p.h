// ms_p2p.h — panel transport over NVLink peer memory without SM-resident collectives.
//
// Measured problem (profiles/README.md): while NCCL's broadcast CTAs sit on the receivers waiting
// for the owner, every one of them displaces a tile CTA and the trailing update loses ~15 %.
// Here the owner pushes the factored panel into each peer's panel buffer with the copy engines
// (cudaMemcpyAsync on CUDA-IPC mapped peer pointers) and raises a flag in the peer's memory; the
// receivers wait with a stream memory operation (cuStreamWaitValue32) — no kernel, no SM.
// Flow control in the other direction is a credit counter each rank writes into the memory of the
// rank that will next overwrite the same panel buffer.
#pragma once
#include "ms_common.cuh"

int ms_p2p_init (ms_ctx *c);      // needs c->nccl (used once, to exchange the IPC handles)
void ms_p2p_destroy (ms_ctx *c);
bool ms_p2p_enabled (const ms_ctx *c);
// owner: after the panel chain of block row I (event `factored` recorded on the panel stream)
int ms_p2p_send_panel (ms_ctx *c, size_t I, int pb, cudaEvent_t factored);
// everyone else: make stream `s` wait until panel I has landed in PB[pb]
int ms_p2p_wait_panel (ms_ctx *c, size_t I, int pb, int owner, cudaStream_t s);
// every rank, after its main stream is done with panel I (event `freed`)
int ms_p2p_release_panel (ms_ctx *c, size_t I, cudaEvent_t freed);
void ms_p2p_begin_factor (ms_ctx *c);
