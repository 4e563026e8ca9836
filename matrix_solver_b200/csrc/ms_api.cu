// ms_api.cu — the C ABI (include/ms_b200.h) and the host-side orchestration of the blocked
// right-looking R^T D R factorisation and the two triangular sweeps.
//
// Factorisation (replaces the left-looking loop of src/solver.cpp:8-104).  For block row I:
//   panel stream : for every 64-row piece of the block row: diag kernel (K1, with the block row's
//                  pre-update pivot test of src/solver.cpp:117 in its first launch) -> TRSM panel
//                  kernel (K3: in place + panel buffer) -> update of the later pieces of the same
//                  block row (K = 64)
//                  -> [nranks > 1] the panel travels to the peers (ms_p2p.cu, or ncclBroadcast)
//   row stream   : rank-m update (K4+K5) of block row I+1, its first piece first
//   main stream  : rank-m update of block row I+2, then of all later block rows.
// Look-ahead depth 2: the panel stream starts block row I+1 as soon as ITS rows have received update
// I, while the main stream is still applying updates I-1 and I to the rest (three panel buffers).
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "ms_common.cuh"
#include "ms_nccl.h"
#include "ms_p2p.h"

static thread_local char g_err[1024] = "";

void
ms_set_error (const char *fmt, ...)
{
  va_list ap;
  va_start (ap, fmt);
  vsnprintf (g_err, sizeof (g_err), fmt, ap);
  va_end (ap);
}

extern "C" const char *
ms_last_error (void)
{
  return g_err;
}

extern "C" int
ms_device_count (int *count)
{
  MS_CHECK_ARG (count != nullptr, "count is null");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount (&n);
  if (e != cudaSuccess)
    {
      cudaGetLastError ();
      n = 0;
    }
  *count = n;
  return MS_OK;
}

static int
require_device (int device)
{
  int n = 0;
  if (cudaGetDeviceCount (&n) != cudaSuccess || n == 0)
    {
      cudaGetLastError ();
      ms_set_error ("no CUDA device visible: this library has no CPU fallback");
      return MS_ERR_NO_DEVICE;
    }
  if (device < 0 || device >= n)
    {
      ms_set_error ("device %d out of range (%d visible)", device, n);
      return MS_ERR_NO_DEVICE;
    }
  cudaDeviceProp prop;
  MS_CUDA (cudaGetDeviceProperties (&prop, device));
  if (prop.major != 10)
    {
      ms_set_error ("device %d is sm_%d%d; this build contains sm_100a code only", device, prop.major,
                    prop.minor);
      return MS_ERR_NO_DEVICE;
    }
  MS_CUDA (cudaSetDevice (device));
  return MS_OK;
}

// ----------------------------------------------------------------------------- lifetime
constexpr size_t MS_STAGE_DOUBLES = (size_t)4 << 20;  // 32 MB staging halves

static void
build_pieces (ms_ctx *c)
{
  c->pieces.clear ();
  c->block_first_piece.assign (c->nblocks + 1, 0);
  for (size_t I = 0; I < c->nblocks; I++)
    {
      c->block_first_piece[I] = c->pieces.size ();
      size_t rows = c->block_rows (I);
      for (size_t off = 0; off < rows; off += MS_SUB)
        {
          ms_piece p;
          p.row0 = I * c->m + off;
          p.h = (int)std::min ((size_t)MS_SUB, rows - off);
          p.block = (int)I;
          p.owner = c->owner_of_block (I);
          p.lrow0 = (I / (size_t)c->nranks) * c->m + off;
          c->pieces.push_back (p);
        }
    }
  c->block_first_piece[c->nblocks] = c->pieces.size ();
}

// Switch the blocking the kernels walk (m_api <-> m_super, one GPU only: local row == global row under either).
// The matrix storage does not depend on it — rows x ld with everything right of the diagonal meaningful — and
// the panel buffers were sized for the taller panel at creation, so only the block / piece tables change.
static void
set_blocking (ms_ctx *c, size_t m_new)
{
  if (m_new == c->m)
    return;
  c->m = m_new;
  c->nblocks = c->n / m_new + (c->n % m_new > 0);
  build_pieces (c);
}

static size_t
wanted_blocking (const ms_ctx *c)
{
  return (c->m_super && c->fast && !c->force_generic && !c->exact_pivots) ? c->m_super : c->m_api;
}

static int
ctx_free (ms_ctx *c)
{
  if (!c)
    return MS_OK;
  cudaSetDevice (c->device);
  cudaDeviceSynchronize ();
  if (c->p2p)
    ms_p2p_destroy (c);
  if (c->nccl)
    ms_nccl_destroy (c);
  cudaFree (c->A);
  for (int b = 0; b < MS_NPB; b++)
    cudaFree (c->PB[b]);
  cudaFree (c->Wt);
  for (int i = 0; i < 5; i++)
    cudaFree (c->vec[i]);
  cudaFree (c->acc);
  cudaFree (c->scratch);
  cudaFree (c->sweep_ready);
  cudaFree (c->sweep_inv);
  cudaFree (c->flags);
  cudaFree (c->d_stage);
  if (c->h_flags)
    cudaFreeHost (c->h_flags);
  if (c->h_pinned)
    cudaFreeHost (c->h_pinned);
  for (auto e : c->ev_pool)
    cudaEventDestroy (e);
  for (auto e : c->prof_pool)
    cudaEventDestroy (e);
  if (c->ev_t0)
    cudaEventDestroy (c->ev_t0);
  if (c->ev_t1)
    cudaEventDestroy (c->ev_t1);
  if (c->s_main)
    cudaStreamDestroy (c->s_main);
  if (c->s_panel)
    cudaStreamDestroy (c->s_panel);
  if (c->s_row)
    cudaStreamDestroy (c->s_row);
  cudaGetLastError ();
  delete c;
  return MS_OK;
}

extern "C" int
ms_get_unique_id (void *id128)
{
  MS_CHECK_ARG (id128 != nullptr, "id buffer is null");
  return ms_nccl_unique_id (id128);
}

static int g_last_device = -1;  // device of the most recent ms_create in this process (blk_ctx follows it)

static int
create_ctx (ms_ctx **out, size_t n, size_t m, int device, int rank, int nranks, const void *unique_id,
            size_t stage_doubles)
{
  MS_CHECK_ARG (out != nullptr, "out is null");
  *out = nullptr;
  MS_CHECK_ARG (n >= 1 && m >= 1, "n and m must be positive");
  MS_CHECK_ARG (nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  MS_CHECK_ARG (nranks == 1 || unique_id != nullptr, "unique_id required when nranks > 1");
  MS_CHECK_ARG (n < ((size_t)1 << 31), "n too large");
  MS_TRY (require_device (device));

  ms_ctx *c = new ms_ctx ();
  c->n = n;
  c->m = m;
  c->m_api = m;
  c->device = device;
  c->rank = rank;
  c->nranks = nranks;
  c->nblocks = n / m + (n % m > 0);
  // + the augmented column (b during the factorisation); MS_FUSE_FORWARD=0 at creation drops it from the
  // layout altogether (A/B of the leading dimension: n = 65536 then has power-of-two rows again)
  const char *ff = getenv ("MS_FUSE_FORWARD");
  c->aug_cols = (ff && !atoi (ff)) ? 0 : 1;
  c->fuse_forward = c->aug_cols > 0;
  c->ld = round_up (n + (size_t)c->aug_cols, MS_CHUNK);
  c->ncols = n + (size_t)c->aug_cols;
  c->hpb = round_up (m, MS_SUB);
  c->fast = (m % MS_SUB == 0 && m <= 1024);  // the tile kernel keeps K = m sign words in shared memory
  // Internal blocking (one GPU, tensor-core path): the K = 64 / 128 trailing update runs at 19 % / 74 % of the
  // DMMA peak against 89 % at K = 256 (profiles/README.md), and which rows form a block row is invisible in R
  // and d.  A caller's m = 64 or 128 is therefore factored in block rows of 256: the same chain of 64-row pieces,
  // the reference's per-block-row pivot pre-test still taken per m rows (factor_block_row), one update of depth
  // 256 instead of four of depth 64.  Ownership would change with several ranks, so P = 1 only; the
  // exact-arithmetic modes ("force_generic", "exact_pivots", the block API) keep the caller's blocking.
  // MS_SUPER_BLOCK=0 switches it off, another multiple of m (<= 1024) selects that.
  if (nranks == 1 && c->fast)
    {
      size_t want = 256;
      const char *sb = getenv ("MS_SUPER_BLOCK");
      if (sb)
        want = (size_t)atol (sb);
      if (want > m && want % m == 0 && want <= 1024)
        {
          c->m_super = want;
          c->hpb = round_up (want, MS_SUB);  // the panel buffers are laid out for the taller panel from the start
        }
    }
  c->rows_local = 0;
  for (size_t I = rank; I < c->nblocks; I += (size_t)nranks)
    c->rows_local += c->block_rows (I);
  c->rows_alloc = round_up (std::max (c->rows_local, (size_t)1), MS_BM);
  cudaDeviceProp prop;
  cudaGetDeviceProperties (&prop, device);
  c->sm_count = prop.multiProcessorCount;
  build_pieces (c);

  auto fail = [&] (int code) {
    ctx_free (c);
    return code;
  };
#define CK(expr)                                                                       \
  do                                                                                   \
    {                                                                                  \
      cudaError_t e__ = (expr);                                                        \
      if (e__ != cudaSuccess)                                                          \
        {                                                                              \
          ms_set_error ("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                        cudaGetErrorString (e__));                                     \
          return fail (MS_ERR_CUDA);                                                   \
        }                                                                              \
    }                                                                                  \
  while (0)

  size_t a_bytes = c->rows_alloc * c->ld * sizeof (double);
  size_t pb_bytes = c->ld * c->hpb * sizeof (double);
  CK (cudaMalloc (&c->A, a_bytes));
  CK (cudaMemset (c->A, 0, a_bytes));
  for (int b = 0; b < MS_NPB; b++)
    {
      CK (cudaMalloc (&c->PB[b], pb_bytes));
      CK (cudaMemset (c->PB[b], 0, pb_bytes));
    }
  CK (cudaMalloc (&c->Wt, MS_SUB * MS_SUB * sizeof (double)));
  CK (cudaMemset (c->Wt, 0, MS_SUB * MS_SUB * sizeof (double)));
  for (int i = 0; i < 5; i++)
    {
      CK (cudaMalloc (&c->vec[i], (c->ld + MS_SUB) * sizeof (double)));
      CK (cudaMemset (c->vec[i], 0, (c->ld + MS_SUB) * sizeof (double)));
    }
  CK (cudaMalloc (&c->acc, (c->ld + MS_SUB) * sizeof (double)));
  // epoch flags of the single-launch sweeps, one per 64-row piece (peers write into them when nranks > 1,
  // so they exist before the IPC handles are exchanged)
  CK (cudaMalloc (&c->sweep_ready, (c->pieces.size () + 8) * sizeof (int)));
  CK (cudaMemset (c->sweep_ready, 0, (c->pieces.size () + 8) * sizeof (int)));
  CK (cudaMalloc (&c->flags, 8 * sizeof (int)));
  CK (cudaMemset (c->flags, 0, 8 * sizeof (int)));
  CK (cudaHostAlloc (&c->h_flags, 8 * sizeof (int), cudaHostAllocDefault));
  c->stage_doubles = stage_doubles;
  CK (cudaMalloc (&c->d_stage, 2 * c->stage_doubles * sizeof (double)));
  CK (cudaHostAlloc (&c->h_pinned, 2 * c->stage_doubles * sizeof (double), cudaHostAllocDefault));

  int lo = 0, hi = 0;
  CK (cudaDeviceGetStreamPriorityRange (&lo, &hi));
  CK (cudaStreamCreateWithPriority (&c->s_main, cudaStreamNonBlocking, lo));
  CK (cudaStreamCreateWithPriority (&c->s_panel, cudaStreamNonBlocking, hi));
  CK (cudaStreamCreateWithPriority (&c->s_row, cudaStreamNonBlocking, hi));
  CK (cudaEventCreate (&c->ev_t0));
  CK (cudaEventCreate (&c->ev_t1));
  c->ev_pool.resize (5 * c->nblocks + 10);
  c->trace = getenv ("MS_TRACE") && atoi (getenv ("MS_TRACE"));
  for (auto &e : c->ev_pool)
    CK (cudaEventCreateWithFlags (&e, c->trace ? cudaEventDefault : cudaEventDisableTiming));
#undef CK
  int r = ms_tile_setup (c);
  if (r < 0)
    return fail (r);
  if (nranks > 1)
    {
      r = ms_nccl_init (c, unique_id);
      if (r < 0)
        return fail (r);
      r = ms_p2p_init (c);
      if (r < 0)
        {
          // ms_p2p_init only fails when a collective inside it failed; a rank that merely cannot map
          // its peers makes ALL ranks agree on the NCCL broadcast there.  Keep going without the
          // peer-memory transport unless it was requested explicitly
          const char *force = getenv ("MS_P2P");
          if (force && atoi (force))
            return fail (r);
          fprintf (stderr, "matrix_solver_b200: peer-memory panel transport unavailable (%s); using ncclBroadcast\n",
                   ms_last_error ());
          ms_p2p_destroy (c);
        }
    }
  const char *e = getenv ("MS_FORCE_GENERIC");
  if (e && atoi (e))
    c->force_generic = true;
  e = getenv ("MS_NO_SWEEP");
  if (e && atoi (e))
    c->no_sweep = true;
  e = getenv ("MS_PANEL_GEMM");
  if (e && atoi (e))
    c->panel_trsm = false;
  e = getenv ("MS_LEFT_LOOKING");
  if (e)
    c->left_looking = std::min (2, std::max (0, atoi (e)));
  e = getenv ("MS_NO_LOOKAHEAD");
  if (e && atoi (e))
    c->lookahead = false;
  *out = c;
  return MS_OK;
}

extern "C" int
ms_create (ms_ctx **out, size_t n, size_t m, int device, int rank, int nranks, const void *unique_id)
{
  int r = create_ctx (out, n, m, device, rank, nranks, unique_id, MS_STAGE_DOUBLES);
  if (r == MS_OK)
    g_last_device = device;
  return r;
}

extern "C" int
ms_destroy (ms_ctx *ctx)
{
  return ctx_free (ctx);
}

extern "C" int
ms_set_option (ms_ctx *c, const char *key, long value)
{
  MS_CHECK_ARG (c && key, "null argument");
  if (!strcmp (key, "force_generic"))
    c->force_generic = value != 0;
  else if (!strcmp (key, "lookahead"))
    c->lookahead = value != 0;
  else if (!strcmp (key, "profile_update"))
    {
      c->profile_update = value != 0;
      MS_CUDA (cudaSetDevice (c->device));
      if (c->profile_update && c->prof_pool.empty ())
        {
          c->prof_pool.resize (2 * (4 * (c->n / c->m_api + 1) + 8));
          for (auto &e : c->prof_pool)
            MS_CUDA (cudaEventCreate (&e));
        }
    }
  else if (!strcmp (key, "panel_trsm"))
    c->panel_trsm = value != 0;
  else if (!strcmp (key, "left_looking"))
    c->left_looking = (int)std::min (2l, std::max (0l, value));
  else if (!strcmp (key, "single_launch_sweeps"))
    c->no_sweep = value == 0;
  else if (!strcmp (key, "fuse_forward"))
    c->fuse_forward = value != 0 && c->aug_cols > 0;
  else if (!strcmp (key, "exact_pivots"))
    c->exact_pivots = value != 0;
  else if (!strcmp (key, "super_block"))
    {
      // 0 = factor in block rows of the caller's m; otherwise a multiple of m the panel buffers were sized for
      MS_CHECK_ARG (value == 0
                      || (c->nranks == 1 && c->fast && value > 0 && (size_t)value % c->m_api == 0
                          && round_up ((size_t)value, MS_SUB) <= c->hpb),
                    "super_block: 0, or a multiple of m within the panel buffers of this context (one GPU)");
      c->m_super = (size_t)value > c->m_api ? (size_t)value : 0;
    }
  else if (!strcmp (key, "tile_band_w"))
    c->tile_band_w = (int)std::max (0l, value);
  else if (!strcmp (key, "tile_tma3"))
    c->tm3_state = value ? 0 : -1;  // 0 = encode the rank-3 maps on the next update launch
  else
    {
      ms_set_error ("unknown option %s", key);
      return MS_ERR_ARG;
    }
  return MS_OK;
}

extern "C" int
ms_get_stats (ms_ctx *c, ms_stats *out)
{
  MS_CHECK_ARG (c && out, "null argument");
  *out = c->stats;
  return MS_OK;
}

extern "C" int
ms_sync (ms_ctx *c)
{
  MS_CHECK_ARG (c, "null ctx");
  MS_CUDA (cudaSetDevice (c->device));
  MS_CUDA (cudaStreamSynchronize (c->s_panel));
  MS_CUDA (cudaStreamSynchronize (c->s_row));
  MS_CUDA (cudaStreamSynchronize (c->s_main));
  return MS_OK;
}

struct phase_timer
{
  ms_ctx *c;
  double *slot;
  phase_timer (ms_ctx *c_, double *slot_) : c (c_), slot (slot_) { cudaEventRecord (c->ev_t0, c->s_main); }
  int finish ()
  {
    MS_CUDA (cudaEventRecord (c->ev_t1, c->s_main));
    MS_CUDA (cudaEventSynchronize (c->ev_t1));
    float ms = 0;
    MS_CUDA (cudaEventElapsedTime (&ms, c->ev_t0, c->ev_t1));
    *slot = ms;
    return MS_OK;
  }
};

// ----------------------------------------------------------------------------- init
extern "C" int
ms_init_formula (ms_ctx *c, int kind)
{
  MS_CHECK_ARG (c, "null ctx");
  MS_CHECK_ARG (kind == MS_FORMULA_F1 || kind == MS_FORMULA_HILBERT, "unknown formula");
  MS_CUDA (cudaSetDevice (c->device));
  phase_timer t (c, &c->stats.init_ms);
  MS_TRY (ms_k_fill_formula (c, kind, c->s_main));
  return t.finish ();
}

static bool
is_pinned_host (const void *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes (&a, p) != cudaSuccess)
    {
      cudaGetLastError ();
      return false;
    }
  return a.type == cudaMemoryTypeHost;
}

// host -> device copy of `cnt` doubles into d_stage half `half`; pageable sources go through the
// pinned staging buffer ("staged once through pinned memory").
static int
stage_h2d (ms_ctx *c, int half, const double *src, size_t cnt)
{
  double *dst = c->d_stage + (size_t)half * c->stage_doubles;
  const double *from = src;
  if (!is_pinned_host (src))
    {
      double *pin = c->h_pinned + (size_t)half * c->stage_doubles;
      memcpy (pin, src, cnt * sizeof (double));
      from = pin;
    }
  MS_CUDA (cudaMemcpyAsync (dst, from, cnt * sizeof (double), cudaMemcpyHostToDevice, c->s_main));
  c->stats.h2d_bytes += (long long)(cnt * sizeof (double));
  return MS_OK;
}

// One strided DMA per block row between a page-locked host array of full rows and the device
// layout: only columns >= the start of the block row's diagonal block move (n(n+m)/2 doubles
// in total instead of n^2).
static int
rows_dma (ms_ctx *c, double *rows, size_t row0, size_t nrows, size_t ld_host, bool to_device, bool wait = true)
{
  // block rows of the CALLER's m (with an internal blocking the rows lie where they always do — one GPU —, and
  // "from the start of each row's diagonal block" keeps the meaning the header gives it)
  const size_t ma = c->m_api, nba = c->n / ma + (c->n % ma > 0);
  for (size_t I = row0 / ma; I < nba && I * ma < row0 + nrows; I++)
    {
      if (c->owner_of_block (I) != c->rank)
        continue;
      size_t g0 = std::max (I * ma, row0);
      size_t g1 = std::min (std::min ((I + 1) * ma, c->n), row0 + nrows);
      size_t col0 = I * ma;
      double *dev = c->row_ptr (g0) + col0;
      double *host = rows + (g0 - row0) * ld_host + col0;
      size_t width = (c->n - col0) * sizeof (double);
      if (to_device)
        {
          MS_CUDA (cudaMemcpy2DAsync (dev, c->ld * 8, host, ld_host * 8, width, g1 - g0, cudaMemcpyHostToDevice, c->s_main));
          c->stats.h2d_bytes += (long long)(width * (g1 - g0));
        }
      else
        {
          MS_CUDA (cudaMemcpy2DAsync (host, ld_host * 8, dev, c->ld * 8, width, g1 - g0, cudaMemcpyDeviceToHost, c->s_main));
          c->stats.d2h_bytes += (long long)(width * (g1 - g0));
        }
    }
  if (wait)
    MS_CUDA (cudaStreamSynchronize (c->s_main));
  return MS_OK;
}

// The same transfer without waiting for it: the copies are queued on the library's stream and the call
// returns.  `rows` must be page-locked and must stay untouched until ms_sync (or any later synchronous
// call on this context) has returned — the file reader parses its next chunk into a second buffer meanwhile
// (parse of chunk k+1 overlapped with the DMA of chunk k).
extern "C" int
ms_upload_rows_async (ms_ctx *c, const double *rows, size_t row0, size_t nrows, size_t ld_host)
{
  MS_CHECK_ARG (c && rows, "null argument");
  MS_CHECK_ARG (row0 + nrows <= c->n && ld_host >= c->n, "row range / ld out of bounds");
  MS_CHECK_ARG (is_pinned_host (rows), "ms_upload_rows_async needs page-locked rows (ms_host_alloc)");
  MS_CUDA (cudaSetDevice (c->device));
  return rows_dma (c, const_cast<double *> (rows), row0, nrows, ld_host, true, false);
}

extern "C" int
ms_upload_rows (ms_ctx *c, const double *rows, size_t row0, size_t nrows, size_t ld_host)
{
  MS_CHECK_ARG (c && rows, "null argument");
  MS_CHECK_ARG (row0 + nrows <= c->n && ld_host >= c->n, "row range / ld out of bounds");
  MS_CUDA (cudaSetDevice (c->device));
  phase_timer t (c, &c->stats.init_ms);
  if (is_pinned_host (rows))
    {
      MS_TRY (rows_dma (c, const_cast<double *> (rows), row0, nrows, ld_host, true));
      return t.finish ();
    }
  MS_CHECK_ARG (ld_host <= c->stage_doubles, "row longer than the staging buffer");
  size_t per = std::max ((size_t)1, c->stage_doubles / ld_host);
  int half = 0;
  for (size_t r = 0; r < nrows; r += per)
    {
      size_t cnt = std::min (per, nrows - r);
      // the pinned half is rewritten by the host: wait until the copy that read it has finished
      MS_CUDA (cudaStreamSynchronize (c->s_main));
      MS_TRY (stage_h2d (c, half, rows + r * ld_host, cnt * ld_host));
      MS_TRY (ms_k_unpack_rows (c, c->d_stage + (size_t)half * c->stage_doubles, row0 + r, cnt, ld_host,
                                c->s_main));
      half ^= 1;
    }
  return t.finish ();
}

extern "C" int
ms_download_rows (ms_ctx *c, double *rows, size_t row0, size_t nrows, size_t ld_host)
{
  MS_CHECK_ARG (c && rows, "null argument");
  MS_CHECK_ARG (row0 + nrows <= c->n && ld_host >= c->n, "row range / ld out of bounds");
  MS_CUDA (cudaSetDevice (c->device));
  return rows_dma (c, rows, row0, nrows, ld_host, false);
}

extern "C" int
ms_host_alloc (void **out, size_t bytes)
{
  MS_CHECK_ARG (out != nullptr, "out is null");
  int n = 0;
  if (cudaGetDeviceCount (&n) != cudaSuccess || n == 0)
    {
      cudaGetLastError ();
      ms_set_error ("no CUDA device visible: this library has no CPU fallback");
      return MS_ERR_NO_DEVICE;
    }
  MS_CUDA (cudaHostAlloc (out, bytes, cudaHostAllocPortable));
  return MS_OK;
}

extern "C" int
ms_host_free (void *p)
{
  if (p)
    MS_CUDA (cudaFreeHost (p));
  return MS_OK;
}

// Walk the reference's packed body block column by block column, in row ranges that fit the
// staging buffer.  Row i of block column J (width w) starts at packed offset:
//   i < J*m : base_J + i*w ;  else base_J + J*m*w + ii*w - ii(ii-1)/2  (ii = i - J*m; row ii has
//   w - ii entries), base_J = sum_{K<J} (K*m*m + m(m+1)/2)        (src/init.cpp:29-58)
static int
packed_transfer (ms_ctx *c, double *body, int to_packed)
{
  MS_CUDA (cudaSetDevice (c->device));
  const size_t n = c->n, m = c->m_api;  // the reference's layout: block columns of the caller's m
  const size_t nb_api = n / m + (n % m > 0);
  size_t base = 0;
  bool pinned = is_pinned_host (body);
  int half = 0;
  for (size_t J = 0; J < nb_api; J++)
    {
      size_t w = std::min (m, n - J * m);  // column width == height of block row J
      size_t rows_total = J * m + w;
      auto row_off = [&] (size_t i) -> size_t {
        if (i <= J * m)
          return i * w;
        size_t ii = i - J * m;
        return J * m * w + ii * w - (ii * (ii - 1)) / 2;
      };
      size_t i0 = 0;
      while (i0 < rows_total)
        {
          size_t i1 = i0;
          // (one grid row per matrix row in the scatter/gather kernel: at most 65535 of them per segment — narrow
          // block columns of a large matrix would otherwise fit more rows into the staging buffer than a grid has)
          while (i1 < rows_total && i1 - i0 < 65535 && row_off (i1 + 1) - row_off (i0) <= c->stage_doubles)
            i1++;
          MS_CHECK_ARG (i1 > i0, "one row of a block column exceeds the staging buffer");
          size_t seg_base = row_off (i0);
          size_t cnt = row_off (i1) - seg_base;
          double *dseg = c->d_stage + (size_t)half * c->stage_doubles;
          double *hseg = body + base + seg_base;
          if (!to_packed)
            {
              if (!pinned)
                MS_CUDA (cudaStreamSynchronize (c->s_main));
              MS_TRY (stage_h2d (c, half, hseg, cnt));
              MS_TRY (ms_k_packed_rows (c, dseg, J, w, i0, i1, seg_base, 0, c->s_main));
            }
          else
            {
              MS_TRY (ms_k_packed_rows (c, dseg, J, w, i0, i1, seg_base, 1, c->s_main));
              if (pinned)
                MS_CUDA (cudaMemcpyAsync (hseg, dseg, cnt * sizeof (double), cudaMemcpyDeviceToHost, c->s_main));
              else
                {
                  double *pin = c->h_pinned + (size_t)half * c->stage_doubles;
                  MS_CUDA (cudaMemcpyAsync (pin, dseg, cnt * sizeof (double), cudaMemcpyDeviceToHost, c->s_main));
                  MS_CUDA (cudaStreamSynchronize (c->s_main));
                  memcpy (hseg, pin, cnt * sizeof (double));
                }
              c->stats.d2h_bytes += (long long)(cnt * sizeof (double));
            }
          half ^= 1;
          i0 = i1;
        }
      base += row_off (rows_total);
    }
  (void)n;
  MS_CUDA (cudaStreamSynchronize (c->s_main));
  return MS_OK;
}

extern "C" int
ms_upload_packed (ms_ctx *c, const double *body)
{
  MS_CHECK_ARG (c && body, "null argument");
  MS_CHECK_ARG (c->nranks == 1, "packed transfers use the single-rank reference layout");
  phase_timer t (c, &c->stats.init_ms);
  MS_TRY (packed_transfer (c, const_cast<double *> (body), 0));
  return t.finish ();
}

extern "C" int
ms_download_packed (ms_ctx *c, double *body)
{
  MS_CHECK_ARG (c && body, "null argument");
  MS_CHECK_ARG (c->nranks == 1, "packed transfers use the single-rank reference layout");
  return packed_transfer (c, body, 1);
}

extern "C" int
ms_get_row (ms_ctx *c, size_t i, double *buf)
{
  MS_CHECK_ARG (c && buf && i < c->n, "bad argument");
  MS_CUDA (cudaSetDevice (c->device));
  size_t I = i / c->m;
  int owner = c->owner_of_block (I);
  size_t cnt = c->n - i;
  const double *src = nullptr;
  if (c->nranks == 1)
    src = c->row_ptr (i) + i;
  else
    {
      // collective: the owner's row travels to every rank (the reference gathers it on the root)
      double *tmp = c->d_stage;
      if (owner == c->rank)
        MS_CUDA (cudaMemcpyAsync (tmp, c->row_ptr (i) + i, cnt * sizeof (double), cudaMemcpyDeviceToDevice,
                                  c->s_main));
      MS_TRY (ms_nccl_bcast (c, tmp, cnt, owner, c->s_main));
      src = tmp;
    }
  MS_CUDA (cudaMemcpyAsync (buf + i, src, cnt * sizeof (double), cudaMemcpyDeviceToHost, c->s_main));
  MS_CUDA (cudaStreamSynchronize (c->s_main));
  c->stats.d2h_bytes += (long long)(cnt * sizeof (double));
  return MS_OK;
}

// A rectangular window of the stored matrix for the printing path (print_mat_beauty,
// src/matrix_op.cpp:443-505, shows two corners of at most 7 x 7 entries): rows [row0, row0 + nrows),
// columns [col0, col0 + ncols) -> buf (row-major, nrows x ncols).  One strided copy per block row the
// window touches instead of one whole-row gather + transfer per printed row; with several ranks the
// owners' pieces are summed into every rank's copy (the others contribute zeros) before the one D2H.
extern "C" int
ms_get_block (ms_ctx *c, size_t row0, size_t nrows, size_t col0, size_t ncols, double *buf)
{
  MS_CHECK_ARG (c && buf && nrows >= 1 && ncols >= 1 && row0 + nrows <= c->n && col0 + ncols <= c->n, "bad window");
  MS_CHECK_ARG (nrows * ncols <= c->stage_doubles, "window larger than the staging buffer");
  MS_CUDA (cudaSetDevice (c->device));
  cudaStream_t s = c->s_main;
  double *stage = c->d_stage;
  const size_t bytes = nrows * ncols * sizeof (double);
  if (c->nranks > 1)
    MS_CUDA (cudaMemsetAsync (stage, 0, bytes, s));
  for (size_t I = row0 / c->m; I < c->nblocks && I * c->m < row0 + nrows; I++)
    {
      if (c->owner_of_block (I) != c->rank)
        continue;
      const size_t g0 = std::max (I * c->m, row0);
      const size_t g1 = std::min (I * c->m + c->block_rows (I), row0 + nrows);
      const double *src = c->row_ptr (g0) + col0;
      if (c->nranks == 1)
        MS_CUDA (cudaMemcpy2DAsync (buf + (g0 - row0) * ncols, ncols * 8, src, c->ld * 8, ncols * 8, g1 - g0,
                                    cudaMemcpyDeviceToHost, s));
      else
        MS_CUDA (cudaMemcpy2DAsync (stage + (g0 - row0) * ncols, ncols * 8, src, c->ld * 8, ncols * 8, g1 - g0,
                                    cudaMemcpyDeviceToDevice, s));
    }
  if (c->nranks > 1)
    {
      MS_TRY (ms_nccl_allreduce_sum (c, stage, nrows * ncols, s));
      MS_CUDA (cudaMemcpyAsync (buf, stage, bytes, cudaMemcpyDeviceToHost, s));
    }
  MS_CUDA (cudaStreamSynchronize (s));
  c->stats.d2h_bytes += (long long)bytes;
  return MS_OK;
}

extern "C" int
ms_get_col (ms_ctx *c, size_t j, double *buf)
{
  MS_CHECK_ARG (c && buf && j < c->n, "bad argument");
  MS_CUDA (cudaSetDevice (c->device));
  double *tmp = c->d_stage;
  MS_TRY (ms_k_get_col (c, j, tmp, c->s_main));
  if (c->nranks > 1)
    MS_TRY (ms_nccl_allreduce_sum (c, tmp, j + 1, c->s_main));
  MS_CUDA (cudaMemcpyAsync (buf, tmp, (j + 1) * sizeof (double), cudaMemcpyDeviceToHost, c->s_main));
  MS_CUDA (cudaStreamSynchronize (c->s_main));
  c->stats.d2h_bytes += (long long)((j + 1) * sizeof (double));
  return MS_OK;
}

extern "C" int
ms_set_vec (ms_ctx *c, int which, const double *host, size_t count)
{
  MS_CHECK_ARG (c && host && which >= 0 && which < 5 && count <= c->n, "bad argument");
  MS_CUDA (cudaSetDevice (c->device));
  if (which == MS_VEC_B && c->b_version_factored == c->b_version && count == c->n)
    {
      // the mirrored compute_y hands back the b that init_b_get_norm produced (test/test_mat.cpp:114,132):
      // if these are the very bits the factorisation already carried through, the fused forward solve
      // stays valid and nothing has to be written
      std::vector<double> cur (count);
      MS_CUDA (cudaMemcpyAsync (cur.data (), c->vec[which], count * sizeof (double), cudaMemcpyDeviceToHost, c->s_main));
      MS_CUDA (cudaStreamSynchronize (c->s_main));
      c->stats.d2h_bytes += (long long)(count * sizeof (double));
      if (memcmp (cur.data (), host, count * sizeof (double)) == 0)
        return MS_OK;
    }
  MS_CUDA (cudaMemcpyAsync (c->vec[which], host, count * sizeof (double), cudaMemcpyHostToDevice, c->s_main));
  MS_CUDA (cudaStreamSynchronize (c->s_main));
  if (which == MS_VEC_B)
    c->b_version++;
  c->stats.h2d_bytes += (long long)(count * sizeof (double));
  return MS_OK;
}

extern "C" int
ms_get_vec (ms_ctx *c, int which, double *host, size_t count)
{
  MS_CHECK_ARG (c && host && which >= 0 && which < 5 && count <= c->n, "bad argument");
  MS_CUDA (cudaSetDevice (c->device));
  MS_CUDA (cudaMemcpyAsync (host, c->vec[which], count * sizeof (double), cudaMemcpyDeviceToHost, c->s_main));
  MS_CUDA (cudaStreamSynchronize (c->s_main));
  c->stats.d2h_bytes += (long long)(count * sizeof (double));
  return MS_OK;
}

// ----------------------------------------------------------------------------- b and norm
extern "C" int
ms_b_and_norm (ms_ctx *c, double *norm)
{
  MS_CHECK_ARG (c && norm, "null argument");
  MS_CUDA (cudaSetDevice (c->device));
  phase_timer t (c, &c->stats.bnorm_ms);
  c->b_version++;
  MS_TRY (ms_k_sym_rows (c, MS_SYM_BNORM, nullptr, c->vec[MS_VEC_B], nullptr, c->s_main));
  double *absp = ms_sym_abs_partial (c);
  if (c->nranks > 1)
    {
      MS_TRY (ms_nccl_allreduce_sum (c, c->vec[MS_VEC_B], c->n, c->s_main));
      MS_TRY (ms_nccl_allreduce_sum (c, absp, c->n, c->s_main));
    }
  double *dmax = absp + c->n;
  MS_TRY (ms_k_max (c, absp, c->n, dmax, c->s_main));
  MS_CUDA (cudaMemcpyAsync (norm, dmax, sizeof (double), cudaMemcpyDeviceToHost, c->s_main));
  MS_TRY (t.finish ());
  c->stats.d2h_bytes += 8;
  return MS_OK;
}

// ----------------------------------------------------------------------------- factor
static double
trapezoid_elems (const ms_ctx *c, size_t lrow_begin, size_t lrow_end)
{
  // stored elements (j >= i) of local rows [lrow_begin, lrow_end)
  double total = 0;
  size_t lr = lrow_begin;
  while (lr < lrow_end)
    {
      size_t lb = lr / c->m;
      size_t seg_end = std::min ((lb + 1) * c->m, lrow_end);
      size_t g0 = (lb * (size_t)c->nranks + (size_t)c->rank) * c->m + (lr - lb * c->m);
      if (g0 < c->n)
        {
          double rows = (double)std::min (seg_end - lr, c->n - g0);
          total += rows * (double)(c->n - g0) - 0.5 * rows * (rows - 1.0);
        }
      lr = seg_end;
    }
  return total;
}

static int
do_update (ms_ctx *c, int pb, size_t I, size_t lrow_begin, size_t lrow_end, int K, size_t k0,
           cudaStream_t s, bool count_flops)
{
  if (lrow_end <= lrow_begin)
    return MS_OK;
  if (count_flops)
    c->stats.update_flops += 2.0 * K * trapezoid_elems (c, lrow_begin, lrow_end);
  if (c->fast && !c->force_generic)
    return ms_k_tile_update (c, pb, I, lrow_begin, lrow_end, K, k0, s);
  ms_operand P;
  P.p = c->PB[pb] + k0 * MS_CHUNK;
  P.ld = 0;
  P.H = c->hpb;
  return ms_k_generic_update (c, P, c->vec[MS_VEC_D] + I * c->m + k0, lrow_begin, lrow_end, K, s);
}

static int
factor_block_row (ms_ctx *c, size_t I, int pb, double thresh, cudaStream_t s, cudaEvent_t rest_ready)
{
  const size_t rowsI = c->block_rows (I);
  const size_t lrowI = (I / (size_t)c->nranks) * c->m;
  double *diag0 = c->A + lrowI * c->ld + I * c->m;
  for (size_t pi = c->block_first_piece[I]; pi < c->block_first_piece[I + 1]; pi++)
    {
      const ms_piece &pc = c->pieces[pi];
      // the block row's pre-update pivot test (src/solver.cpp:117) rides along with its first piece's
      // diag kernel — except that with look-ahead the later pieces receive their last update (panel
      // I-1) only at `rest_ready`: their share of the test runs right after that wait, below.
      // "Block row" here is the CALLER's: with an internal blocking (c->m a multiple of m_api) the test
      // of rows [k0, k0 + m_api) is taken when the piece that opens them comes up, i.e. after the pieces
      // above it have been applied to them — the state the reference tests them in.
      const size_t k0 = pc.row0 - I * c->m;
      const bool first = pi == c->block_first_piece[I];
      const bool sub_first = k0 % c->m_api == 0;
      const size_t sub_rows = std::min (c->m_api, rowsI - k0);
      const bool split_pre = first && rest_ready != nullptr && (size_t)pc.h < rowsI;
      const double *pre = sub_first ? diag0 + k0 * (c->ld + 1) : nullptr;
      const int pre_count = sub_first ? (split_pre ? pc.h : (int)sub_rows) : 0;
      double *blk = c->A + pc.lrow0 * c->ld + pc.row0;
      double *rows = c->A + pc.lrow0 * c->ld;
      size_t col_lo = pc.row0 + (size_t)pc.h;
      const bool tiles = c->fast && !c->force_generic;
      // Inside the block row the tensor-core path works left-looking: a piece receives the pieces above it in ONE
      // update of depth K = 64, 128, 192 right before its diag kernel, instead of every factored piece being pushed
      // into all rows below it with K = 64 (where the tile kernel runs at a fifth of its K = 256 rate) — the same
      // flops in deeper, smaller launches, and a shorter dependent chain.  The piece that opens one of the CALLER's
      // block rows brings all of that block row's rows up to date (they are pre-tested with it, see above); the
      // pieces after it then only lack the pieces of their own caller block row.  The exact-arithmetic kernels keep
      // the right-looking order.
      const bool left = tiles && (c->left_looking == 2 || (c->left_looking == 1 && c->ncols - I * c->m >= MS_LEFT_MIN_COLS));
      if (left && k0 > 0)
        {
          const size_t ksub = k0 / c->m_api * c->m_api;
          if (sub_first)
            MS_TRY (do_update (c, pb, I, pc.lrow0, pc.lrow0 + sub_rows, (int)k0, 0, s, false));
          else if (k0 > ksub)
            MS_TRY (do_update (c, pb, I, pc.lrow0, pc.lrow0 + (size_t)pc.h, (int)(k0 - ksub), ksub, s, false));
        }
      if (c->panel_trsm)
        {
          // K1 then K3 as a triangular solve: no explicit inverse on the critical path
          MS_TRY (ms_k_diag (c, blk, c->ld, pc.h, thresh, c->vec[MS_VEC_D] + pc.row0, nullptr, nullptr, pre, c->ld,
                             pre_count, s));
          if (col_lo >= c->ncols)
            continue;
          MS_TRY (ms_k_trsm_panel (c, blk, c->ld, rows, c->ld, pc.h, col_lo, c->ncols, c->vec[MS_VEC_D] + pc.row0,
                                   c->PB[pb], c->hpb, k0, s));
        }
      else
        {
          // the reference's formulation: K1 + K2 (inverse) then K3 = (V D)^T * rows
          MS_TRY (ms_k_diag (c, blk, c->ld, pc.h, thresh, c->vec[MS_VEC_D] + pc.row0, c->Wt, nullptr, pre, c->ld,
                             pre_count, s));
          if (col_lo >= c->ncols)
            continue;
          if (tiles && pc.h == MS_SUB)
            MS_TRY (ms_k_tile_panel (c, pb, pc, I, s));
          else
            {
              ms_operand W;
              W.p = c->Wt;
              W.ld = 0;
              W.H = MS_SUB;
              MS_TRY (ms_k_generic_panel (c, W, rows, c->ld, rows, c->ld, pc.h, col_lo, c->PB[pb], c->hpb, k0,
                                          c->ncols, s));
            }
        }
      // later pieces of the same block row (their rows must have received the previous panel: the
      // first piece was updated ahead of them so that its diag/TRSM could start early)
      size_t lr_next = pc.lrow0 + (size_t)pc.h;
      size_t lr_end = lrowI + rowsI;
      if (rest_ready && first && lr_next < lr_end)
        {
          MS_CUDA (cudaStreamWaitEvent (s, rest_ready, 0));
          MS_TRY (ms_k_pretest (c, diag0, c->ld, pc.h, (int)std::min (c->m_api, rowsI), thresh, s));
        }
      if (!left)
        MS_TRY (do_update (c, pb, I, lr_next, lr_end, pc.h, k0, s, false));
    }
  return MS_OK;
}

// Streams of the look-ahead (depth 2):
//   s_panel (high priority): panel chain of block row I (owner) and the NCCL broadcast of panel I
//   s_row   (high priority): "U1" — panel I applied to block row I+1 only, so that its owner can
//                            start the next panel chain at once
//   s_main  (low priority) : "U2a" — panel I applied to block row I+2 (what U1 of the NEXT step
//                            waits for), then "U2b" — panel I applied to everything after it
// Panel I+1 can therefore be factored and broadcast while the bulk updates of steps I-1 and I are
// still running; three panel buffers rotate.
extern "C" int
ms_factor (ms_ctx *c, double norma, double *elapsed_s)
{
  MS_CHECK_ARG (c, "null ctx");
  MS_CUDA (cudaSetDevice (c->device));
  const double thresh = 1.e-14 * norma;  // EPS * norma, src/defines.h:9
  set_blocking (c, wanted_blocking (c));
  c->stats.block_internal = (int)c->m;
  const size_t nb = c->nblocks;
  const int P = c->nranks;
  const bool la = c->lookahead;
  const bool p2p = ms_p2p_enabled (c);
  ms_p2p_begin_factor (c);
  cudaStream_t sm = c->s_main;
  cudaStream_t sp = la ? c->s_panel : c->s_main;
  cudaStream_t su = la ? c->s_row : c->s_main;
  c->stats.update_flops = 0;
  c->stats.update_launches = 0;
  c->stats.update_ms = 0;
  c->stats.used_dmma = 0;
  MS_CUDA (cudaMemsetAsync (c->flags, 0, 8 * sizeof (int), sm));
  MS_CUDA (cudaEventRecord (c->ev_t0, sm));
  // forward solve fused into the factorisation: the current b becomes column n of the matrix and comes
  // out as z = D y; ms_solve_forward then only has to read it back
  c->ncols = c->n + (c->fuse_forward ? 1 : 0);
  c->b_version_factored = 0;
  if (c->fuse_forward)
    MS_TRY (ms_k_aug_scatter (c, c->vec[MS_VEC_B], sm));
  cudaEvent_t *ev_panel = c->ev_pool.data ();            // panel I factored (and received)
  cudaEvent_t *ev_strip = ev_panel + nb + 2;              // block row I updated through panel I-1
  cudaEvent_t *ev_u2a = ev_strip + nb + 2;                // block row I updated through panel I-2
  cudaEvent_t *ev_free = ev_u2a + nb + 2;                 // main stream is done with panel I
  cudaEvent_t *ev_rest = ev_free + nb + 2;                // pieces 1.. of block row I updated through panel I-1
  if (la)
    {
      MS_CUDA (cudaEventRecord (ev_strip[0], sm));
      MS_CUDA (cudaStreamWaitEvent (sp, ev_strip[0], 0));
      MS_CUDA (cudaStreamWaitEvent (su, ev_strip[0], 0));
    }
  // profile_update: every update launch sits between two timing events taken from a pool that
  // ms_set_option created (nothing is created or destroyed inside the factorisation)
  size_t prof_used = 0;
  auto timed = [&] (cudaStream_t st, size_t I, int pb, size_t lr0, size_t lr1) -> int {
    const int K = (int)c->m;  // block row I is full whenever later block rows exist
    if (!c->profile_update || prof_used + 2 > c->prof_pool.size () || lr1 <= lr0)
      return do_update (c, pb, I, lr0, lr1, K, 0, st, true);
    MS_CUDA (cudaEventRecord (c->prof_pool[prof_used], st));
    int r = do_update (c, pb, I, lr0, lr1, K, 0, st, true);
    MS_CUDA (cudaEventRecord (c->prof_pool[prof_used + 1], st));
    prof_used += 2;
    return r;
  };
  auto mine = [&] (size_t J) { return J < nb && c->owner_of_block (J) == c->rank; };
  auto lrow_of_block = [&] (size_t J) { return (J / (size_t)P) * c->m; };
  for (size_t I = 0; I < nb; I++)
    {
      const int pb = (int)(I % MS_NPB);
      const int owner = c->owner_of_block (I);
      // panel buffer pb last held panel I-3: the main stream must be through with it before the
      // panel stream overwrites it (by factoring or by receiving the broadcast)
      if (la && I >= MS_NPB)
        MS_CUDA (cudaStreamWaitEvent (sp, ev_free[I - MS_NPB], 0));
      if (p2p)
        MS_TRY (ms_p2p_before_overwrite (c, I, sp));
      if (owner == c->rank)
        {
          if (la && I > 0)
            MS_CUDA (cudaStreamWaitEvent (sp, ev_strip[I], 0));
          MS_TRY (factor_block_row (c, I, pb, thresh, sp, (la && I > 0) ? ev_rest[I] : nullptr));
        }
      if (P > 1 && !p2p)
        MS_TRY (ms_nccl_bcast_panel (c, I, pb, owner, sp));
      if (P > 1 && p2p && owner != c->rank)
        MS_TRY (ms_p2p_wait_panel (c, I, pb, owner, sp));
      if (la || p2p)
        MS_CUDA (cudaEventRecord (ev_panel[I], sp));
      if (la)
        MS_CUDA (cudaStreamWaitEvent (sm, ev_panel[I], 0));
      if (P > 1 && p2p && owner == c->rank)
        MS_TRY (ms_p2p_send_panel (c, I, pb, ev_panel[I]));
      if (I + 1 >= nb)
        break;
      // U1: block row I+1
      if (mine (I + 1))
        {
          if (la)
            {
              MS_CUDA (cudaStreamWaitEvent (su, ev_panel[I], 0));
              if (I >= 1)
                MS_CUDA (cudaStreamWaitEvent (su, ev_u2a[I + 1], 0));
            }
          // first piece of block row I+1 first: its diag kernel is the head of the next panel chain
          size_t lr = lrow_of_block (I + 1);
          size_t lr_end = lr + c->block_rows (I + 1);
          size_t lr_p0 = std::min (lr + (size_t)MS_SUB, lr_end);
          MS_TRY (timed (su, I, pb, lr, lr_p0));
          if (la)
            MS_CUDA (cudaEventRecord (ev_strip[I + 1], su));
          if (lr_p0 < lr_end)
            MS_TRY (timed (su, I, pb, lr_p0, lr_end));
          if (la)
            MS_CUDA (cudaEventRecord (ev_rest[I + 1], su));
        }
      // U2a: block row I+2 (the next U1 depends on it), then U2b: the rest
      size_t lr_rest;
      if (mine (I + 2))
        {
          size_t lr = lrow_of_block (I + 2);
          size_t lr_end = lr + c->block_rows (I + 2);
          MS_TRY (timed (sm, I, pb, lr, lr_end));
          if (la)
            MS_CUDA (cudaEventRecord (ev_u2a[I + 2], sm));
          lr_rest = lr_end;
        }
      else
        {
          size_t Jn = I + 3;
          while (Jn < nb && !mine (Jn))
            Jn++;
          lr_rest = Jn < nb ? lrow_of_block (Jn) : c->rows_local;
        }
      if (lr_rest < c->rows_local)
        MS_TRY (timed (sm, I, pb, lr_rest, c->rows_local));
      if (la || p2p)
        MS_CUDA (cudaEventRecord (ev_free[I], sm));
      if (p2p)
        MS_TRY (ms_p2p_release_panel (c, I, ev_free[I]));
    }
  MS_CUDA (cudaMemcpyAsync (c->h_flags, c->flags, 8 * sizeof (int), cudaMemcpyDeviceToHost, sm));
  MS_CUDA (cudaEventRecord (c->ev_t1, sm));
  MS_CUDA (cudaEventSynchronize (c->ev_t1));
  MS_CUDA (cudaStreamSynchronize (c->s_panel));
  MS_CUDA (cudaStreamSynchronize (c->s_row));
  MS_TRY (ms_p2p_end_factor (c));
  float ms = 0;
  MS_CUDA (cudaEventElapsedTime (&ms, c->ev_t0, c->ev_t1));
  c->stats.factor_ms = ms;
  if (elapsed_s)
    *elapsed_s = ms * 1e-3;
  // device time in which at least one update launch was running: the union of the launch intervals
  // (two launches overlap under look-ahead, so their plain sum over-counts wall time)
  c->stats.update_sum_ms = 0;
  if (prof_used)
    {
      std::vector<std::pair<float, float>> iv (prof_used / 2);
      for (size_t i = 0; i < prof_used; i += 2)
        {
          float a = 0, b = 0;
          MS_CUDA (cudaEventElapsedTime (&a, c->ev_t0, c->prof_pool[i]));
          MS_CUDA (cudaEventElapsedTime (&b, c->ev_t0, c->prof_pool[i + 1]));
          iv[i / 2] = {a, b};
          c->stats.update_sum_ms += b - a;
        }
      std::sort (iv.begin (), iv.end ());
      double total = 0;
      float lo = iv[0].first, hi = iv[0].second;
      for (size_t i = 1; i < iv.size (); i++)
        {
          if (iv[i].first > hi)
            {
              total += hi - lo;
              lo = iv[i].first;
              hi = iv[i].second;
            }
          else
            hi = std::max (hi, iv[i].second);
        }
      c->stats.update_ms = total + (hi - lo);
    }
  if (c->trace && la)
    {
      c->trace_panel.assign (nb, -1.f);
      c->trace_free.assign (nb, -1.f);
      c->trace_strip.assign (nb, -1.f);
      for (size_t I = 0; I < nb; I++)
        {
          cudaEventElapsedTime (&c->trace_panel[I], c->ev_t0, ev_panel[I]);
          if (I + 1 < nb)
            cudaEventElapsedTime (&c->trace_free[I], c->ev_t0, ev_free[I]);
          if (I > 0 && mine (I))
            cudaEventElapsedTime (&c->trace_strip[I], c->ev_t0, ev_strip[I]);
        }
      cudaGetLastError ();
    }
  int failed = c->h_flags[0] != 0;
  if (P > 1)
    MS_TRY (ms_nccl_agree_flag (c, &failed));
  if (c->fuse_forward && !failed)
    c->b_version_factored = c->b_version;
  return failed ? MS_NUMERICAL_FAILURE : MS_OK;
}

// debug hook (not in the public header): per-step timeline of the last ms_factor in ms since its
// start — panel I available, main stream done with panel I, block row I ready to factor (owner)
extern "C" int
ms_debug_trace (ms_ctx *c, float *panel_ms, float *free_ms, float *strip_ms, size_t cap)
{
  MS_CHECK_ARG (c && c->trace, "create the context with MS_TRACE=1");
  for (size_t i = 0; i < c->trace_panel.size () && i < cap; i++)
    {
      panel_ms[i] = c->trace_panel[i];
      free_ms[i] = c->trace_free[i];
      strip_ms[i] = c->trace_strip[i];
    }
  return (int)c->trace_panel.size ();
}

// ----------------------------------------------------------------------------- solves
extern "C" int
ms_solve_forward (ms_ctx *c, double norma)
{
  MS_CHECK_ARG (c, "null ctx");
  MS_CUDA (cudaSetDevice (c->device));
  const double thresh = 1.e-14 * norma;
  cudaStream_t s = c->s_main;
  phase_timer t (c, &c->stats.forward_ms);
  c->stats.forward_fused = 0;
  if (c->b_version_factored != 0 && c->b_version_factored == c->b_version)
    {
      // the factorisation already carried this b through (augmented column): y = D z.  No pivot can
      // fail here that ms_factor has not reported already (same |R_jj| < EPS*norm test, src/solver.cpp:195-199
      // vs :122).  With several ranks every rank contributes its own rows and y is replicated like the
      // reference's (MPI_Bcast per block, src/solver.cpp:204-205).
      if (c->nranks > 1)
        MS_CUDA (cudaMemsetAsync (c->vec[MS_VEC_Y], 0, c->ld * sizeof (double), s));
      MS_TRY (ms_k_aug_gather (c, c->vec[MS_VEC_Y], s));
      if (c->nranks > 1)
        MS_TRY (ms_nccl_allreduce_sum (c, c->vec[MS_VEC_Y], c->n, s));
      c->stats.forward_fused = 1;
      return t.finish ();
    }
  MS_CUDA (cudaMemsetAsync (c->flags, 0, 8 * sizeof (int), s));
  MS_CUDA (cudaMemsetAsync (c->acc, 0, (c->ld + MS_SUB) * sizeof (double), s));
  if (c->nranks > 1)
    MS_CUDA (cudaMemsetAsync (c->vec[MS_VEC_Y], 0, c->ld * sizeof (double), s));
  int swept = 0;
  MS_TRY (ms_k_sweep (c, true, thresh, s, &swept));
  for (size_t I = 0; I < c->nblocks && !swept; I++)
    {
      int owner = c->owner_of_block (I);
      if (c->nranks > 1)
        MS_TRY (ms_nccl_reduce_sum (c, c->acc + I * c->m, c->block_rows (I), owner, s));
      if (owner != c->rank)
        continue;
      for (size_t pi = c->block_first_piece[I]; pi < c->block_first_piece[I + 1]; pi++)
        MS_TRY (ms_k_forward_step (c, c->pieces[pi], thresh, s));
    }
  if (c->nranks > 1)
    MS_TRY (ms_nccl_allreduce_sum (c, c->vec[MS_VEC_Y], c->n, s));
  MS_CUDA (cudaMemcpyAsync (c->h_flags, c->flags, 8 * sizeof (int), cudaMemcpyDeviceToHost, s));
  MS_TRY (t.finish ());
  int failed = c->h_flags[0] != 0;
  if (c->nranks > 1)
    MS_TRY (ms_nccl_agree_flag (c, &failed));
  return failed ? MS_NUMERICAL_FAILURE : MS_OK;
}

extern "C" int
ms_solve_backward (ms_ctx *c, double norma)
{
  MS_CHECK_ARG (c, "null ctx");
  MS_CUDA (cudaSetDevice (c->device));
  const double thresh = 1.e-14 * norma;
  cudaStream_t s = c->s_main;
  phase_timer t (c, &c->stats.backward_ms);
  MS_CUDA (cudaMemsetAsync (c->flags, 0, 8 * sizeof (int), s));
  MS_CUDA (cudaMemsetAsync (c->acc, 0, (c->ld + MS_SUB) * sizeof (double), s));
  int swept = 0;
  MS_TRY (ms_k_sweep (c, false, thresh, s, &swept));
  for (size_t I = c->nblocks; !swept && I-- > 0;)
    {
      int owner = c->owner_of_block (I);
      // local rows strictly above block row I on this rank
      size_t blocks_before = I / (size_t)c->nranks + ((I % (size_t)c->nranks) > (size_t)c->rank ? 1 : 0);
      size_t lrows_before_block = blocks_before * c->m;
      if (owner == c->rank)
        {
          for (size_t pi = c->block_first_piece[I + 1]; pi-- > c->block_first_piece[I];)
            {
              const ms_piece &pc = c->pieces[pi];
              MS_TRY (ms_k_backward_step (c, pc, pc.lrow0, thresh, s));
            }
        }
      if (c->nranks > 1)
        {
          MS_TRY (ms_nccl_bcast (c, c->vec[MS_VEC_X] + I * c->m, c->block_rows (I), owner, s));
          if (owner != c->rank)
            for (size_t pi = c->block_first_piece[I + 1]; pi-- > c->block_first_piece[I];)
              MS_TRY (ms_k_backward_step (c, c->pieces[pi], lrows_before_block, thresh, s));
        }
    }
  MS_CUDA (cudaMemcpyAsync (c->h_flags, c->flags, 8 * sizeof (int), cudaMemcpyDeviceToHost, s));
  MS_TRY (t.finish ());
  int failed = c->h_flags[0] != 0;
  const int timed_out = c->h_flags[3] != 0;
  if (c->nranks > 1)
    MS_TRY (ms_nccl_agree_flag (c, &failed));
  if (timed_out)
    {
      ms_set_error ("backward sweep: a peer rank never published its part of x (bounded wait expired)");
      return MS_ERR_CUDA;
    }
  return failed ? MS_NUMERICAL_FAILURE : MS_OK;
}

// ----------------------------------------------------------------------------- residual
extern "C" int
ms_residual (ms_ctx *c, double *resid2, double *err2)
{
  MS_CHECK_ARG (c && resid2 && err2, "null argument");
  MS_CUDA (cudaSetDevice (c->device));
  cudaStream_t s = c->s_main;
  phase_timer t (c, &c->stats.residual_ms);
  // r <- A x (symmetric product from the stored triangle), then r -= b inside norms? keep r = Ax - b
  MS_TRY (ms_k_sym_rows (c, MS_SYM_SYMV, c->vec[MS_VEC_X], c->vec[MS_VEC_R], nullptr, s));
  if (c->nranks > 1)
    MS_TRY (ms_nccl_allreduce_sum (c, c->vec[MS_VEC_R], c->n, s));
  MS_TRY (ms_k_axpy_minus (c, c->vec[MS_VEC_R], c->vec[MS_VEC_B], c->n, s));
  double *out2 = c->scratch;  // reuse: sym_rows scratch is no longer needed
  MS_TRY (ms_k_norms (c, c->vec[MS_VEC_R], c->vec[MS_VEC_X], c->n, out2, s));
  double h[2];
  MS_CUDA (cudaMemcpyAsync (h, out2, 2 * sizeof (double), cudaMemcpyDeviceToHost, s));
  MS_TRY (t.finish ());
  *resid2 = sqrt (h[0]);
  *err2 = sqrt (h[1]);
  c->stats.d2h_bytes += 16;
  return MS_OK;
}

// ----------------------------------------------------------------------------- block API
// Stateless entry points working on host buffers; a small private context supplies the stream,
// the flag word and scratch memory.
static ms_ctx *g_blk = nullptr;
static std::mutex g_blk_mu;

static int
blk_ctx (ms_ctx **out)
{
  std::lock_guard<std::mutex> lk (g_blk_mu);
  if (!g_blk)
    {
      // on the device of the context this process created last (a rank of an MS_NGPU run works on its own GPU;
      // round 1 opened an extra CUDA context on GPU 0 from every rank), else MS_DEVICE, else 0.  The block
      // routines stage through buffers of their own, so this context carries the smallest staging area.
      int dev = g_last_device;
      if (dev < 0)
        {
          const char *e = getenv ("MS_DEVICE");
          dev = e ? atoi (e) : 0;
        }
      int r = create_ctx (&g_blk, 64, 64, dev, 0, 1, nullptr, (size_t)1 << 16);
      if (r < 0)
        return r;
      // the block API mirrors the reference's routines one to one: its diagonal-block factor keeps the
      // reference's sqrt / divide roundings (bit-identical to cholesky_decomp_U), whatever the pipeline does
      g_blk->exact_pivots = true;
    }
  MS_CUDA (cudaSetDevice (g_blk->device));
  *out = g_blk;
  return MS_OK;
}

struct dev_buf
{
  double *p = nullptr;
  ~dev_buf () { cudaFree (p); }
  int alloc (size_t cnt) { MS_CUDA (cudaMalloc (&p, std::max (cnt, (size_t)1) * sizeof (double))); return MS_OK; }
};

static void
unpack_upper (const double *packed, double *full, int w)
{
  size_t pos = 0;
  for (int i = 0; i < w; i++)
    for (int j = 0; j < w; j++)
      full[(size_t)i * w + j] = j >= i ? packed[pos++] : 0.0;
}

static void
pack_upper (const double *full, double *packed, int w)
{
  size_t pos = 0;
  for (int i = 0; i < w; i++)
    for (int j = i; j < w; j++)
      packed[pos++] = full[(size_t)i * w + j];
}

static int
blk_flag (ms_ctx *c, int *flag)
{
  MS_CUDA (cudaMemcpyAsync (c->h_flags, c->flags, sizeof (int), cudaMemcpyDeviceToHost, c->s_main));
  MS_CUDA (cudaStreamSynchronize (c->s_main));
  *flag = c->h_flags[0] != 0;
  return MS_OK;
}

// cholesky_decomp_U for blocks that do not fit one CTA's shared memory (w > 112; the BASELINE block sizes
// 128 and 256 arrive here through the mirrored cholesky_decomp_U): one CTA works on the block in global
// memory (it stays in L2), pivot by pivot, with the reference's roundings in the reference's order —
// entry (k, j) loses (r_ik * r_ij) * d_i for i ascending, each product and difference rounded on its own,
// then row k is divided by r_kk * d_k (src/solver.cpp:106-160).  Bit-identical to the oracle at every w;
// speed is irrelevant here (the factorisation itself never factors more than 64 pivots per launch).
__global__ void __launch_bounds__ (1024)
diag_global_kernel (double *__restrict__ a, int w, double thresh, double *__restrict__ d_out, int *flags)
{
  const int tid = threadIdx.x;
  // the reference tests the not-yet-updated a_ii of every pivot (src/solver.cpp:117)
  for (int i = tid; i < w; i += 1024)
    if (fabs (a[(size_t)i * w + i]) < thresh)
      flags[0] = 1;
  __syncthreads ();
  for (int i = 0; i < w; i++)
    {
      double *ri = a + (size_t)i * w;
      const double s = ri[i];
      const double di = s > 0 ? 1.0 : -1.0;
      const double rii = __dsqrt_rn (fabs (s));
      const double den = __dmul_rn (rii, di);
      __syncthreads ();  // everybody has read the pivot
      if (tid == 0)
        {
          ri[i] = rii;
          d_out[i] = di;
          if (fabs (rii) < thresh)
            flags[0] = 1;
        }
      for (int j = i + 1 + tid; j < w; j += 1024)
        ri[j] = __ddiv_rn (ri[j], den);
      __syncthreads ();
      // rows k > i, columns j >= k
      const int rem = w - i - 1;
      for (int e = tid; e < rem * rem; e += 1024)
        {
          const int k = i + 1 + e / rem, j = i + 1 + e % rem;
          if (j >= k)
            a[(size_t)k * w + j] = __dsub_rn (a[(size_t)k * w + j], __dmul_rn (__dmul_rn (ri[k], ri[j]), di));
        }
      __syncthreads ();
    }
}

extern "C" int
ms_block_diag_factor (double *a_packed, double *d, int w, double norma)
{
  MS_CHECK_ARG (a_packed && d && w >= 1 && w <= 8192, "bad argument (1 <= w <= 8192)");
  ms_ctx *c;
  MS_TRY (blk_ctx (&c));
  std::vector<double> full ((size_t)w * w);
  unpack_upper (a_packed, full.data (), w);
  dev_buf da, dd;
  MS_TRY (da.alloc ((size_t)w * w));
  MS_TRY (dd.alloc (w));
  cudaStream_t s = c->s_main;
  MS_CUDA (cudaMemsetAsync (c->flags, 0, 8 * sizeof (int), s));
  MS_CUDA (cudaMemcpyAsync (da.p, full.data (), full.size () * 8, cudaMemcpyHostToDevice, s));
  if (w <= 112)
    MS_TRY (ms_k_diag (c, da.p, w, w, 1.e-14 * norma, dd.p, nullptr, nullptr, da.p, w, w, s));
  else
    {
      diag_global_kernel<<<1, 1024, 0, s>>> (da.p, w, 1.e-14 * norma, dd.p, c->flags);
      c->stats.kernel_launches++;
      MS_CUDA (cudaGetLastError ());
    }
  MS_CUDA (cudaMemcpyAsync (full.data (), da.p, full.size () * 8, cudaMemcpyDeviceToHost, s));
  MS_CUDA (cudaMemcpyAsync (d, dd.p, w * 8, cudaMemcpyDeviceToHost, s));
  int flag;
  MS_TRY (blk_flag (c, &flag));
  pack_upper (full.data (), a_packed, w);
  return flag ? MS_NUMERICAL_FAILURE : MS_OK;
}

// The inverse is produced by the same kernel that factors; to invert a GIVEN factor R we hand it
// R^T R's ... no: we run only the K2 half by factoring nothing.  The diag kernel's inverse stage
// reads the factored block from shared memory, so we feed it A := R^T R?  That would change the
// rounding.  Instead a dedicated thin kernel below reuses the same column sweep.
__global__ void __launch_bounds__ (256)
upper_inverse_kernel (const double *__restrict__ a, int w, double thresh, double *__restrict__ inv, int *flags)
{
  extern __shared__ double sm[];
  const int S = w + 1;
  double *r = sm, *v = sm + (size_t)w * S;
  for (int e = threadIdx.x; e < w * w; e += 256)
    {
      int i = e / w, j = e - i * w;
      r[i * S + j] = a[(size_t)i * w + j];
      v[i * S + j] = 0.0;
    }
  __syncthreads ();
  for (int i = threadIdx.x; i < w; i += 256)
    if (fabs (r[i * S + i]) < thresh)
      flags[0] = 1;
  for (int j = threadIdx.x; j < w; j += 256)
    for (int i = j; i >= 0; i--)
      {
        double val;
        if (i == j)
          val = __ddiv_rn (1.0, r[i * S + i]);
        else
          {
            double sum = 0;
            for (int k = i + 1; k <= j; k++)
              sum = __dadd_rn (sum, __dmul_rn (r[i * S + k], v[k * S + j]));
            val = __ddiv_rn (-sum, r[i * S + i]);
          }
        v[i * S + j] = val;
      }
  __syncthreads ();
  for (int e = threadIdx.x; e < w * w; e += 256)
    {
      int i = e / w, j = e - i * w;
      inv[(size_t)i * w + j] = j >= i ? v[i * S + j] : 0.0;
    }
}

// reverse_upper for w > 112: same recurrence and order as upper_inverse_kernel, operands in global memory
// (one thread per column of the inverse; column j only ever reads its own earlier results)
__global__ void __launch_bounds__ (256)
upper_inverse_global_kernel (const double *__restrict__ a, int w, double thresh, double *__restrict__ inv, int *flags)
{
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= w)
    return;
  if (fabs (a[(size_t)j * w + j]) < thresh)
    flags[0] = 1;
  for (int i = 0; i < j; i++)
    inv[(size_t)j * w + i] = 0.0;  // strictly lower part of row j
  for (int i = j; i >= 0; i--)
    {
      double val;
      if (i == j)
        val = __ddiv_rn (1.0, a[(size_t)i * w + i]);
      else
        {
          double sum = 0;
          for (int k = i + 1; k <= j; k++)
            sum = __dadd_rn (sum, __dmul_rn (a[(size_t)i * w + k], inv[(size_t)k * w + j]));
          val = __ddiv_rn (-sum, a[(size_t)i * w + i]);
        }
      inv[(size_t)i * w + j] = val;
    }
}

extern "C" int
ms_block_upper_inverse (const double *a_packed, double *inv_packed, int w, double norma)
{
  MS_CHECK_ARG (a_packed && inv_packed && w >= 1 && w <= 8192, "bad argument (1 <= w <= 8192)");
  ms_ctx *c;
  MS_TRY (blk_ctx (&c));
  std::vector<double> full ((size_t)w * w);
  unpack_upper (a_packed, full.data (), w);
  dev_buf da, di;
  MS_TRY (da.alloc ((size_t)w * w));
  MS_TRY (di.alloc ((size_t)w * w));
  cudaStream_t s = c->s_main;
  MS_CUDA (cudaMemsetAsync (c->flags, 0, 8 * sizeof (int), s));
  MS_CUDA (cudaMemcpyAsync (da.p, full.data (), full.size () * 8, cudaMemcpyHostToDevice, s));
  if (w <= 112)
    {
      size_t smem = 2 * (size_t)w * (w + 1) * sizeof (double);
      MS_CUDA (cudaFuncSetAttribute (upper_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      upper_inverse_kernel<<<1, 256, smem, s>>> (da.p, w, 1.e-14 * norma, di.p, c->flags);
    }
  else
    upper_inverse_global_kernel<<<(w + 255) / 256, 256, 0, s>>> (da.p, w, 1.e-14 * norma, di.p, c->flags);
  c->stats.kernel_launches++;
  MS_CUDA (cudaGetLastError ());
  MS_CUDA (cudaMemcpyAsync (full.data (), di.p, full.size () * 8, cudaMemcpyDeviceToHost, s));
  int flag;
  MS_TRY (blk_flag (c, &flag));
  pack_upper (full.data (), inv_packed, w);
  return flag ? MS_NUMERICAL_FAILURE : MS_OK;
}

__global__ void
scale_cols_to_wt_kernel (const double *__restrict__ inv, const double *__restrict__ D, int h, double *__restrict__ wt)
{
  // wt(k, i) = inv(k, i) * D_i, row-major h x h
  int e = blockIdx.x * 256 + threadIdx.x;
  if (e < h * h)
    {
      int i = e % h;
      wt[e] = inv[e] * D[i];
    }
}

extern "C" int
ms_block_panel (const double *inv_packed, double *A, const double *D, int h, int w)
{
  MS_CHECK_ARG (inv_packed && A && D && h >= 1 && w >= 1, "bad argument");
  ms_ctx *c;
  MS_TRY (blk_ctx (&c));
  std::vector<double> inv ((size_t)h * h);
  unpack_upper (inv_packed, inv.data (), h);
  dev_buf dinv, dwt, dsrc, dout, dD;
  MS_TRY (dinv.alloc ((size_t)h * h));
  MS_TRY (dwt.alloc ((size_t)h * h));
  MS_TRY (dsrc.alloc ((size_t)h * w));
  MS_TRY (dout.alloc ((size_t)h * w));
  MS_TRY (dD.alloc (h));
  cudaStream_t s = c->s_main;
  MS_CUDA (cudaMemcpyAsync (dinv.p, inv.data (), inv.size () * 8, cudaMemcpyHostToDevice, s));
  MS_CUDA (cudaMemcpyAsync (dsrc.p, A, (size_t)h * w * 8, cudaMemcpyHostToDevice, s));
  MS_CUDA (cudaMemcpyAsync (dD.p, D, h * 8, cudaMemcpyHostToDevice, s));
  scale_cols_to_wt_kernel<<<(h * h + 255) / 256, 256, 0, s>>> (dinv.p, dD.p, h, dwt.p);
  c->stats.kernel_launches++;
  ms_operand W;
  W.p = dwt.p;
  W.ld = (size_t)h;
  W.H = 0;
  MS_TRY (ms_k_generic_panel (c, W, dsrc.p, (size_t)w, dout.p, (size_t)w, h, 0, nullptr, 0, 0, (size_t)w, s));
  MS_CUDA (cudaMemcpyAsync (A, dout.p, (size_t)h * w * 8, cudaMemcpyDeviceToHost, s));
  MS_CUDA (cudaStreamSynchronize (s));
  return MS_OK;
}

extern "C" int
ms_block_trsm_panel (const double *r_packed, double *A, const double *D, int h, int w)
{
  MS_CHECK_ARG (r_packed && A && D && h >= 1 && w >= 1, "bad argument");
  ms_ctx *c;
  MS_TRY (blk_ctx (&c));
  std::vector<double> R ((size_t)h * h);
  unpack_upper (r_packed, R.data (), h);
  dev_buf dR, dA, dD, dPB;
  size_t ldw = round_up ((size_t)w, MS_CHUNK);
  MS_TRY (dR.alloc ((size_t)h * h));
  MS_TRY (dA.alloc ((size_t)h * w));
  MS_TRY (dD.alloc (h));
  MS_TRY (dPB.alloc (ldw * MS_SUB));
  cudaStream_t s = c->s_main;
  MS_CUDA (cudaMemcpyAsync (dR.p, R.data (), R.size () * 8, cudaMemcpyHostToDevice, s));
  MS_CUDA (cudaMemcpyAsync (dA.p, A, (size_t)h * w * 8, cudaMemcpyHostToDevice, s));
  MS_CUDA (cudaMemcpyAsync (dD.p, D, h * 8, cudaMemcpyHostToDevice, s));
  // what the factorisation does inside a block row: 64-row piece t is solved against its own triangle,
  // then the pieces below it lose R_tu^T D_t X_t (X_t = D_t R_tt^{-T}(...), so D_t X_t is the unscaled solve)
  for (int t0 = 0; t0 < h; t0 += MS_SUB)
    {
      const int ht = std::min (MS_SUB, h - t0);
      MS_TRY (ms_k_trsm_panel (c, dR.p + (size_t)t0 * h + t0, (size_t)h, dA.p + (size_t)t0 * w, (size_t)w, ht, 0,
                               (size_t)w, dD.p + t0, dPB.p, MS_SUB, 0, s));
      const int below = h - t0 - ht;
      if (below > 0)
        MS_TRY (ms_k_generic_block_update (c, dA.p + (size_t)(t0 + ht) * w, (size_t)w, dR.p + (size_t)t0 * h + t0 + ht,
                                           (size_t)h, dD.p + t0, dA.p + (size_t)t0 * w, (size_t)w, ht, below, w, 0, s));
    }
  MS_CUDA (cudaMemcpyAsync (A, dA.p, (size_t)h * w * 8, cudaMemcpyDeviceToHost, s));
  MS_CUDA (cudaStreamSynchronize (s));
  return MS_OK;
}

// The same contraction on the DMMA / TMA tile kernel of the factorisation (k and h multiples of 64):
// a throw-away context whose panel buffer holds P = [R1 | R2] (k x (h + w), chunk-major) and whose matrix
// rows 0..h-1 hold A in columns [h, h + w) — one update launch then computes rows 0..h-1 of
// C -= P^T D P, i.e. A -= R1^T D R2 in the columns we read back (the columns left of h receive
// R1^T D R1, which nobody looks at).  For the packed-upper twin R2 == R1 and P = R1, C = A.
static int
block_update_dmma (double *A_full, const double *R1, const double *D, const double *R2, int k, int h, int w,
                   int syrk)
{
  ms_ctx *b;
  MS_TRY (blk_ctx (&b));
  const size_t ncols = syrk ? (size_t)w : (size_t)h + (size_t)w;
  const size_t col0 = syrk ? 0 : (size_t)h;
  ms_ctx *t = nullptr;
  MS_TRY (ms_create (&t, std::max (ncols, (size_t)k), (size_t)k, b->device, 0, 1, nullptr));
  struct closer
  {
    ms_ctx *t;
    ~closer () { ms_destroy (t); }
  } guard{t};
  t->ncols = ncols;
  std::vector<double> pb (t->ld * t->hpb, 0.0);
  for (int kk = 0; kk < k; kk++)
    for (size_t j = 0; j < ncols; j++)
      pb[chunk_major_index ((size_t)kk, j, t->hpb)]
        = (syrk || j < (size_t)h) ? R1[(size_t)kk * h + j] : R2[(size_t)kk * w + (j - h)];
  cudaStream_t s = t->s_main;
  MS_CUDA (cudaMemcpyAsync (t->PB[0], pb.data (), pb.size () * 8, cudaMemcpyHostToDevice, s));
  MS_CUDA (cudaMemcpyAsync (t->vec[MS_VEC_D], D, (size_t)k * 8, cudaMemcpyHostToDevice, s));
  MS_CUDA (cudaMemcpy2DAsync (t->A + col0, t->ld * 8, A_full, (size_t)w * 8, (size_t)w * 8, (size_t)h,
                              cudaMemcpyHostToDevice, s));
  MS_TRY (ms_k_tile_update (t, 0, 0, 0, (size_t)h, k, 0, s));
  MS_CUDA (cudaMemcpy2DAsync (A_full, (size_t)w * 8, t->A + col0, t->ld * 8, (size_t)w * 8, (size_t)h,
                              cudaMemcpyDeviceToHost, s));
  MS_CUDA (cudaStreamSynchronize (s));
  b->stats.kernel_launches += t->stats.kernel_launches;
  b->stats.used_dmma = 1;
  return MS_OK;
}

static int
block_update_common (double *A_full, size_t a_cnt, const double *R1, const double *D, const double *R2, int k,
                     int h, int w, int upper_only)
{
  ms_ctx *c;
  MS_TRY (blk_ctx (&c));
  c->stats.used_dmma = 0;
  // multiples of the tile kernel's piece height go to the DMMA kernel the factorisation runs
  // (MS_BLOCK_GENERIC=1 keeps them on the DFMA kernel for comparison)
  const char *g = getenv ("MS_BLOCK_GENERIC");
  if (k % MS_SUB == 0 && h % MS_SUB == 0 && k <= 1024 && !(g && atoi (g)))
    return block_update_dmma (A_full, R1, D, R2, k, h, w, upper_only);
  dev_buf dA, d1, d2, dD;
  MS_TRY (dA.alloc (a_cnt));
  MS_TRY (d1.alloc ((size_t)k * h));
  MS_TRY (d2.alloc ((size_t)k * w));
  MS_TRY (dD.alloc (k));
  cudaStream_t s = c->s_main;
  MS_CUDA (cudaMemcpyAsync (dA.p, A_full, a_cnt * 8, cudaMemcpyHostToDevice, s));
  MS_CUDA (cudaMemcpyAsync (d1.p, R1, (size_t)k * h * 8, cudaMemcpyHostToDevice, s));
  MS_CUDA (cudaMemcpyAsync (d2.p, R2, (size_t)k * w * 8, cudaMemcpyHostToDevice, s));
  MS_CUDA (cudaMemcpyAsync (dD.p, D, k * 8, cudaMemcpyHostToDevice, s));
  MS_TRY (ms_k_generic_block_update (c, dA.p, (size_t)w, d1.p, (size_t)h, dD.p, d2.p, (size_t)w, k, h, w,
                                     upper_only, s));
  MS_CUDA (cudaMemcpyAsync (A_full, dA.p, a_cnt * 8, cudaMemcpyDeviceToHost, s));
  MS_CUDA (cudaStreamSynchronize (s));
  return MS_OK;
}

// debug hook (not in the public header): did the last ms_block_update(_diag) run on the DMMA tile kernel?
extern "C" int
ms_debug_block_used_dmma (void)
{
  return g_blk ? g_blk->stats.used_dmma : 0;
}

extern "C" int
ms_block_update (double *A, const double *R1, const double *D, const double *R2, int k, int h, int w)
{
  MS_CHECK_ARG (A && R1 && D && R2 && k >= 1 && h >= 1 && w >= 1, "bad argument");
  return block_update_common (A, (size_t)h * w, R1, D, R2, k, h, w, 0);
}

extern "C" int
ms_block_update_diag (double *a_packed, const double *R1, const double *D, int k, int w)
{
  MS_CHECK_ARG (a_packed && R1 && D && k >= 1 && w >= 1, "bad argument");
  std::vector<double> full ((size_t)w * w);
  unpack_upper (a_packed, full.data (), w);
  MS_TRY (block_update_common (full.data (), full.size (), R1, D, R1, k, w, w, 1));
  pack_upper (full.data (), a_packed, w);
  return MS_OK;
}

static int
vec_norms (const double *r, const double *x, size_t n, double *out2)
{
  ms_ctx *c;
  MS_TRY (blk_ctx (&c));
  dev_buf dv, dout;
  MS_TRY (dv.alloc (n));
  MS_TRY (dout.alloc (2));
  cudaStream_t s = c->s_main;
  MS_CUDA (cudaMemcpyAsync (dv.p, r ? r : x, n * 8, cudaMemcpyHostToDevice, s));
  MS_TRY (ms_k_norms (c, r ? dv.p : nullptr, x ? dv.p : nullptr, n, dout.p, s));
  MS_CUDA (cudaMemcpyAsync (out2, dout.p, 16, cudaMemcpyDeviceToHost, s));
  MS_CUDA (cudaStreamSynchronize (s));
  return MS_OK;
}

extern "C" int
ms_norm2 (const double *v, size_t n, double *out)
{
  MS_CHECK_ARG (v && out, "null argument");
  double o[2];
  MS_TRY (vec_norms (v, nullptr, n, o));
  *out = sqrt (o[0]);
  return MS_OK;
}

extern "C" int
ms_error_vs_planted (const double *x, size_t n, double *out)
{
  MS_CHECK_ARG (x && out, "null argument");
  double o[2];
  MS_TRY (vec_norms (nullptr, x, n, o));
  *out = sqrt (o[1]);
  return MS_OK;
}

// debug hook (not part of the public header): raw flag words of the block-API context
extern "C" int
ms_debug_blk_flags (int *out8)
{
  ms_ctx *c;
  MS_TRY (blk_ctx (&c));
  MS_CUDA (cudaMemcpy (out8, c->flags, 8 * sizeof (int), cudaMemcpyDeviceToHost));
  return MS_OK;
}
