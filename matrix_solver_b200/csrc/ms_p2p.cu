// ms_p2p.cu — see ms_p2p.h.
#include <nccl.h>

#include <algorithm>

#include "ms_nccl.h"
#include "ms_p2p.h"

namespace
{
constexpr int SIG_INTS = 256;      // [I % 16] ready flag of panel I, [16 + r] credit from rank r,
                                   // [32 + (I % 4) * 16 + r] "rank r holds its slice of panel I"
constexpr int SIG_READY = 0;
constexpr int SIG_READY_SLOTS = 16;
constexpr int SIG_CREDIT = 16;
constexpr int SIG_SLICE = 32;
constexpr int SIG_SLICE_SLOTS = 4;
constexpr int SIG_MAX_RANKS = 16;
constexpr int STAGE_SLOTS = 64;

struct p2p_state
{
  bool on = false;
  int *sig = nullptr;                       // this rank's flags (peers write, we wait)
  int *stage = nullptr;                     // local source words for the 4-byte signal copies:
  std::vector<int> stage_next;              // one ring of STAGE_SLOTS per stream (index = peer, P = misc)
  std::vector<double *> peer_pb[MS_NPB];    // peers' panel buffers (IPC mapped), [rank]
  std::vector<double *> peer_d;             // peers' d vectors
  std::vector<double *> peer_x;             // peers' x vectors (backward sweep: owners store x_t into them)
  std::vector<int *> peer_sweep;            // peers' sweep flags
  std::vector<int *> peer_sig;
  std::vector<void *> opened;
  std::vector<cudaStream_t> s_peer;         // one copy stream per peer: transfers to different
                                            // peers run on different copy engines / links
  cudaStream_t s_misc = nullptr;            // credits
  cudaEvent_t ev_tmp = nullptr;
  unsigned epoch = 0;                       // grows by (nblocks + 8) per factorisation
  CUresult (*waitValue32) (CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
  CUresult (*writeValue32) (CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
};

p2p_state *
st (const ms_ctx *c)
{
  return (p2p_state *)c->p2p;
}

int
driver_fn (const char *name, void **fn)
{
  cudaDriverEntryPointQueryResult q;
  MS_CUDA (cudaGetDriverEntryPoint (name, fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !*fn)
    {
      ms_set_error ("driver entry point %s unavailable", name);
      return MS_ERR_CUDA;
    }
  return MS_OK;
}

// post `value` into the int at `remote` (a peer or local device address), ordered on stream s;
// `ring` selects the staging ring that belongs to that stream
int
post_value (p2p_state *p, int ring, int *remote, unsigned value, cudaStream_t s)
{
  int *slot = p->stage + (size_t)ring * STAGE_SLOTS + (p->stage_next[ring]++ % STAGE_SLOTS);
  CUresult r = p->writeValue32 ((CUstream)s, (CUdeviceptr)slot, value, 0);
  if (r != CUDA_SUCCESS)
    {
      ms_set_error ("cuStreamWriteValue32 failed: %d", (int)r);
      return MS_ERR_CUDA;
    }
  MS_CUDA (cudaMemcpyAsync (remote, slot, sizeof (int), cudaMemcpyDefault, s));
  return MS_OK;
}

int
wait_value (p2p_state *p, int *addr, unsigned value, cudaStream_t s)
{
  CUresult r = p->waitValue32 ((CUstream)s, (CUdeviceptr)addr, value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS)
    {
      ms_set_error ("cuStreamWaitValue32 failed: %d", (int)r);
      return MS_ERR_CUDA;
    }
  return MS_OK;
}
}  // namespace

bool
ms_p2p_enabled (const ms_ctx *c)
{
  return c->p2p && st (c)->on;
}

namespace
{
constexpr int NH = MS_NPB + 4;  // IPC handles per rank: [PB0 PB1 PB2 d sig x sweep_ready]

// everything that involves only this rank: flag words, streams, and the handles the peers will open
int
local_setup (ms_ctx *c, p2p_state *p, std::vector<char> &mine)
{
  const int P = c->nranks;
  const size_t hb = sizeof (cudaIpcMemHandle_t);
  MS_TRY (driver_fn ("cuStreamWaitValue32", (void **)&p->waitValue32));
  MS_TRY (driver_fn ("cuStreamWriteValue32", (void **)&p->writeValue32));
  MS_CUDA (cudaMalloc (&p->sig, SIG_INTS * sizeof (int)));
  MS_CUDA (cudaMemset (p->sig, 0, SIG_INTS * sizeof (int)));
  MS_CUDA (cudaMalloc (&p->stage, (size_t)(P + 1) * STAGE_SLOTS * sizeof (int)));
  p->stage_next.assign (P + 1, 0);
  p->s_peer.assign (P, nullptr);
  for (int r = 0; r < P; r++)
    if (r != c->rank)
      MS_CUDA (cudaStreamCreateWithFlags (&p->s_peer[r], cudaStreamNonBlocking));
  MS_CUDA (cudaStreamCreateWithFlags (&p->s_misc, cudaStreamNonBlocking));
  MS_CUDA (cudaEventCreateWithFlags (&p->ev_tmp, cudaEventDisableTiming));
  void *ptrs[NH];
  for (int b = 0; b < MS_NPB; b++)
    ptrs[b] = c->PB[b];
  ptrs[MS_NPB] = c->vec[MS_VEC_D];
  ptrs[MS_NPB + 1] = p->sig;
  ptrs[MS_NPB + 2] = c->vec[MS_VEC_X];
  ptrs[MS_NPB + 3] = c->sweep_ready;
  for (int i = 0; i < NH; i++)
    MS_CUDA (cudaIpcGetMemHandle ((cudaIpcMemHandle_t *)(mine.data () + i * hb), ptrs[i]));
  return MS_OK;
}

int
open_peers (ms_ctx *c, p2p_state *p, const std::vector<char> &all)
{
  const int P = c->nranks;
  const size_t hb = sizeof (cudaIpcMemHandle_t);
  for (int b = 0; b < MS_NPB; b++)
    p->peer_pb[b].assign (P, nullptr);
  p->peer_d.assign (P, nullptr);
  p->peer_sig.assign (P, nullptr);
  p->peer_x.assign (P, nullptr);
  p->peer_sweep.assign (P, nullptr);
  const char *force = getenv ("MS_P2P_FORCE_FAIL");  // test hook: pretend rank <value> cannot map its peers
  if (force && atoi (force) == c->rank)
    {
      ms_set_error ("peer mapping failure forced by MS_P2P_FORCE_FAIL");
      return MS_ERR_CUDA;
    }
  for (int r = 0; r < P; r++)
    {
      if (r == c->rank)
        continue;
      for (int i = 0; i < NH; i++)
        {
          void *q = nullptr;
          cudaError_t err = cudaIpcOpenMemHandle (&q, *(cudaIpcMemHandle_t *)(all.data () + ((size_t)r * NH + i) * hb),
                                                  cudaIpcMemLazyEnablePeerAccess);
          if (err != cudaSuccess)
            {
              cudaGetLastError ();
              ms_set_error ("cudaIpcOpenMemHandle (rank %d, buffer %d) -> %s", r, i, cudaGetErrorString (err));
              return MS_ERR_CUDA;
            }
          p->opened.push_back (q);
          if (i < MS_NPB)
            p->peer_pb[i][r] = (double *)q;
          else if (i == MS_NPB)
            p->peer_d[r] = (double *)q;
          else if (i == MS_NPB + 1)
            p->peer_sig[r] = (int *)q;
          else if (i == MS_NPB + 2)
            p->peer_x[r] = (double *)q;
          else
            p->peer_sweep[r] = (int *)q;
        }
    }
  return MS_OK;
}
}  // namespace

// Every rank walks through the same collective calls whatever happens locally, and the ranks then
// AGREE on the transport: if peer memory cannot be set up anywhere (no P2P between two of the
// devices, IPC refused in a container ...), everybody uses the ncclBroadcast transport instead.
int
ms_p2p_init (ms_ctx *c)
{
  const char *e = getenv ("MS_P2P");
  if (e && !atoi (e))
    return MS_OK;  // NCCL broadcast instead
  p2p_state *p = new p2p_state ();
  c->p2p = p;
  const int P = c->nranks;
  const size_t hb = sizeof (cudaIpcMemHandle_t);
  std::vector<char> mine (NH * hb, 0), all ((size_t)P * NH * hb);
  int bad = local_setup (c, p, mine) != MS_OK;
  if (P > SIG_MAX_RANKS)
    bad = 1;  // the flag page holds 16 ranks; larger jobs keep the NCCL transport

  // every rank walks the SAME collectives from here on, whatever failed locally: failures only set
  // `bad`, and the agreement below sends everybody to the NCCL transport together.  The exchange buffer
  // is the context's device staging area (allocated long ago), so nothing here can fail half way.
  char *dbuf = (char *)c->d_stage;
  if (all.size () > 2 * c->stage_doubles * sizeof (double))
    bad = 1;
  if (!bad && cudaMemcpy (dbuf + (size_t)c->rank * NH * hb, mine.data (), mine.size (), cudaMemcpyHostToDevice) != cudaSuccess)
    {
      ms_set_error ("ms_p2p_init: staging the IPC handles failed: %s", cudaGetErrorString (cudaGetLastError ()));
      bad = 1;
    }
  int coll = ms_nccl_allgather_bytes (c, dbuf + (size_t)c->rank * NH * hb, dbuf, NH * hb, c->s_main);
  if (coll == MS_OK && cudaStreamSynchronize (c->s_main) != cudaSuccess)
    coll = MS_ERR_CUDA;
  if (coll == MS_OK && !bad && cudaMemcpy (all.data (), dbuf, all.size (), cudaMemcpyDeviceToHost) != cudaSuccess)
    bad = 1;
  if (coll != MS_OK)
    bad = 1;  // a failed collective is fatal for NCCL too, but the agreement is still attempted by all

  if (!bad)
    bad = open_peers (c, p, all) != MS_OK;
  // the agreement doubles as the barrier after which peers may post into this rank's memory
  int any_bad = bad;
  MS_TRY (ms_nccl_agree_flag (c, &any_bad));
  if (any_bad)
    {
      if (bad)
        fprintf (stderr, "matrix_solver_b200: rank %d: %s; all ranks use the ncclBroadcast panel transport\n",
                 c->rank, ms_last_error ());
      for (void *q : p->opened)
        cudaIpcCloseMemHandle (q);
      p->opened.clear ();
      p->on = false;
      return MS_OK;
    }
  p->on = true;
  return MS_OK;
}

void
ms_p2p_destroy (ms_ctx *c)
{
  p2p_state *p = st (c);
  if (!p)
    return;
  cudaDeviceSynchronize ();
  for (void *q : p->opened)
    cudaIpcCloseMemHandle (q);
  cudaFree (p->sig);
  cudaFree (p->stage);
  for (cudaStream_t t : p->s_peer)
    if (t)
      cudaStreamDestroy (t);
  if (p->s_misc)
    cudaStreamDestroy (p->s_misc);
  if (p->ev_tmp)
    cudaEventDestroy (p->ev_tmp);
  delete p;
  c->p2p = nullptr;
}

// end of a factorisation: drain the flag posts so that none of them can land after the next
// factorisation (next epoch) has started writing the same words
int
ms_p2p_end_factor (ms_ctx *c)
{
  p2p_state *p = st (c);
  if (!p || !p->on)
    return MS_OK;
  for (cudaStream_t t : p->s_peer)
    if (t)
      MS_CUDA (cudaStreamSynchronize (t));
  MS_CUDA (cudaStreamSynchronize (p->s_misc));
  return MS_OK;
}

void
ms_p2p_begin_factor (ms_ctx *c)
{
  if (ms_p2p_enabled (c))
    st (c)->epoch += (unsigned)c->nblocks + 8;
}

// PULL transport, scatter + all-gather.  Pushing from the owner serialises on the owner's copy path
// (measured ~260 GB/s aggregate for 7 peers in round 1), so every receiver pulls with its own copy engine.
// Round 2 measured what seven full-panel pulls out of ONE GPU cost: the owner's NVLink egress carries the
// panel seven times (434 MB at r = 30k: ~0.56 ms of a 0.92 ms step cadence at 8 GPUs,
// profiles/r02_trace_factor_8gpu_pull7.log).  Now the panel is cut into P-1 slices of 16-column chunks:
//   owner      : after the panel chain, raises ready[I % 16] = epoch + I + 1 in every peer's memory
//                (4-byte copies on per-peer streams; next owner first)
//   receiver r : panel stream waits for the flag, pulls ITS slice (and d_I) out of the owner's buffer — the
//                owner's egress carries the panel once —, credits the owner, tells every other receiver
//                "slice r is here" (slice[I % 4][r] = epoch + I + 1), then pulls the other P-2 slices from
//                the receivers that hold them (rotating start, so that every link is busy) and credits each
//   any rank   : before it writes the same buffer again (panel I+3, factored or pulled) it waits for the
//                credits of everybody who read panel I out of it: all peers if it owned panel I, the other
//                receivers otherwise
// Every flag word has exactly one writer whose posts are ordered on one stream and only ever grow (ready:
// the owner of block rows I, I+16, ... for the rank counts of one box; slice and credit: rank r), so a
// flag can never move backwards.
namespace
{
struct span
{
  size_t off, cnt;  // doubles
};

// slice k (0 .. P-2) of the range of panel I that travels: 16-column chunks >= I*m / 16
span
panel_slice (const ms_ctx *c, size_t I, int k)
{
  const size_t chunk0 = (I * c->m) / MS_CHUNK, nch = c->ld / MS_CHUNK;
  const size_t per = (nch - chunk0 + (size_t)(c->nranks - 2)) / (size_t)(c->nranks - 1);
  const size_t a = std::min (nch, chunk0 + (size_t)k * per), b = std::min (nch, a + per);
  return {a * c->hpb * MS_CHUNK, (b - a) * c->hpb * MS_CHUNK};
}

// receivers of panel I are numbered 0 .. P-2 starting with the next owner
int
receiver_index (const ms_ctx *c, int owner, int rank)
{
  return (rank - owner - 1 + c->nranks) % c->nranks;
}

int
receiver_rank (const ms_ctx *c, int owner, int index)
{
  return (owner + 1 + index) % c->nranks;
}
}  // namespace

int
ms_p2p_send_panel (ms_ctx *c, size_t I, int pb, cudaEvent_t factored)
{
  (void)pb;
  p2p_state *p = st (c);
  const int P = c->nranks;
  for (int k = 1; k < P; k++)
    {
      const int r = (int)((I + (size_t)k) % (size_t)P);
      if (r == c->rank)
        continue;
      cudaStream_t s = p->s_peer[r];
      MS_CUDA (cudaStreamWaitEvent (s, factored, 0));
      MS_TRY (post_value (p, r, p->peer_sig[r] + SIG_READY + (int)(I % SIG_READY_SLOTS), p->epoch + (unsigned)I + 1, s));
    }
  return MS_OK;
}

int
ms_p2p_wait_panel (ms_ctx *c, size_t I, int pb, int owner, cudaStream_t s)
{
  p2p_state *p = st (c);
  const int P = c->nranks;
  const unsigned stamp = p->epoch + (unsigned)I + 1;
  MS_TRY (wait_value (p, p->sig + SIG_READY + (int)(I % SIG_READY_SLOTS), stamp, s));
  const bool has_trailing = (I + 1) * c->m < c->n;
  const size_t rows = c->block_rows (I);
  const int me = receiver_index (c, owner, c->rank);
  if (has_trailing)
    {
      const span sl = panel_slice (c, I, me);
      if (sl.cnt)
        MS_CUDA (cudaMemcpyAsync (c->PB[pb] + sl.off, p->peer_pb[pb][owner] + sl.off, sl.cnt * sizeof (double),
                                  cudaMemcpyDefault, s));
    }
  MS_CUDA (cudaMemcpyAsync (c->vec[MS_VEC_D] + I * c->m, p->peer_d[owner] + I * c->m, rows * sizeof (double),
                            cudaMemcpyDefault, s));
  // the owner may reuse its buffer once everybody has pulled
  MS_TRY (post_value (p, P, p->peer_sig[owner] + SIG_CREDIT + c->rank, stamp, s));
  if (!has_trailing || P <= 2)
    return MS_OK;
  // all-gather among the receivers: announce my slice first (nobody waits before everybody has posted)
  int *slice_word = nullptr;
  const int slot = SIG_SLICE + (int)(I % SIG_SLICE_SLOTS) * SIG_MAX_RANKS;
  for (int j = 1; j < P - 1; j++)
    {
      const int r = receiver_rank (c, owner, (me + j) % (P - 1));
      MS_TRY (post_value (p, P, p->peer_sig[r] + slot + c->rank, stamp, s));
    }
  for (int j = 1; j < P - 1; j++)
    {
      const int idx = (me + j) % (P - 1);
      const int r = receiver_rank (c, owner, idx);
      slice_word = p->sig + slot + r;
      MS_TRY (wait_value (p, slice_word, stamp, s));
      const span sl = panel_slice (c, I, idx);
      if (sl.cnt)
        MS_CUDA (cudaMemcpyAsync (c->PB[pb] + sl.off, p->peer_pb[pb][r] + sl.off, sl.cnt * sizeof (double),
                                  cudaMemcpyDefault, s));
      MS_TRY (post_value (p, P, p->peer_sig[r] + SIG_CREDIT + c->rank, stamp, s));
    }
  return MS_OK;
}

// before this rank writes PB[pb] at step I (factor or pull): everybody who read panel I-3 out of this
// buffer must be done — all peers if this rank owned that panel, the other receivers (who fetched this
// rank's slice here) otherwise
int
ms_p2p_before_overwrite (ms_ctx *c, size_t I, cudaStream_t s)
{
  p2p_state *p = st (c);
  if (I < MS_NPB)
    return MS_OK;
  const size_t J = I - MS_NPB;
  const int owner = c->owner_of_block (J);
  const bool sliced = (J + 1) * c->m < c->n && c->nranks > 2;
  if (owner != c->rank && !sliced)
    return MS_OK;
  for (int r = 0; r < c->nranks; r++)
    if (r != c->rank && (owner == c->rank || r != owner))
      MS_TRY (wait_value (p, p->sig + SIG_CREDIT + r, p->epoch + (unsigned)J + 1, s));
  return MS_OK;
}

int
ms_p2p_release_panel (ms_ctx *c, size_t I, cudaEvent_t freed)
{
  (void)c;
  (void)I;
  (void)freed;
  return MS_OK;  // credits are posted by the pulls themselves
}

// backward sweep over peer memory (ms_kernels_solve.cu): rank r's x vector and sweep flags as seen from here
double *
ms_p2p_peer_x (const ms_ctx *c, int r)
{
  return ms_p2p_enabled (c) ? st (c)->peer_x[r] : nullptr;
}

int *
ms_p2p_peer_sweep_ready (const ms_ctx *c, int r)
{
  return ms_p2p_enabled (c) ? st (c)->peer_sweep[r] : nullptr;
}
