// ms_p2p.h — panel transport over NVLink peer memory without SM-resident collectives.
//
// Measured problem (profiles/README.md): while NCCL's broadcast CTAs sit on the receivers waiting
// for the owner, every one of them displaces a tile CTA and the trailing update loses ~15 %.
// Here the owner raises a flag in each peer's memory when the panel is factored; the receivers wait
// with a stream memory operation (cuStreamWaitValue32) and PULL the panel out of the owner's buffer
// with their own copy engines (cudaMemcpyAsync on CUDA-IPC mapped peer pointers) — no kernel, no SM.
// Flow control in the other direction is a credit each receiver writes into the owner's memory once
// its pull is complete; the owner waits for them before it writes that buffer again.
#pragma once
#include "ms_common.cuh"

int ms_p2p_init (ms_ctx *c);      // needs c->nccl (used once, to exchange the IPC handles)
void ms_p2p_destroy (ms_ctx *c);
bool ms_p2p_enabled (const ms_ctx *c);
// owner: after the panel chain of block row I (event `factored` recorded on the panel stream)
int ms_p2p_send_panel (ms_ctx *c, size_t I, int pb, cudaEvent_t factored);
// every rank, on the stream that is about to write PB[I % 3] (factor or pull)
int ms_p2p_before_overwrite (ms_ctx *c, size_t I, cudaStream_t s);
// everyone else: wait for the owner's flag, pull panel I into PB[pb], give the owner its credit
int ms_p2p_wait_panel (ms_ctx *c, size_t I, int pb, int owner, cudaStream_t s);
// every rank, after its main stream is done with panel I (event `freed`)
int ms_p2p_release_panel (ms_ctx *c, size_t I, cudaEvent_t freed);
void ms_p2p_begin_factor (ms_ctx *c);
int ms_p2p_end_factor (ms_ctx *c);
// peer views used by the single-launch backward sweep (null when the transport is off)
double *ms_p2p_peer_x (const ms_ctx *c, int r);
int *ms_p2p_peer_sweep_ready (const ms_ctx *c, int r);
