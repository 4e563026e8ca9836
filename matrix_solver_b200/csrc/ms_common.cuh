// ms_common.cuh — shared declarations of the sm_100a library: context, layout helpers,
// error plumbing and the PTX wrappers (mbarrier, TMA, DMMA) used by the tile kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/ms_b200.h"

// ----------------------------------------------------------------------------- errors
void ms_set_error (const char *fmt, ...);

#define MS_CUDA(expr)                                                                  \
  do                                                                                   \
    {                                                                                  \
      cudaError_t e__ = (expr);                                                        \
      if (e__ != cudaSuccess)                                                          \
        {                                                                              \
          ms_set_error ("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                        cudaGetErrorString (e__));                                     \
          return MS_ERR_CUDA;                                                          \
        }                                                                              \
    }                                                                                  \
  while (0)

#define MS_CHECK_ARG(cond, msg)                                                        \
  do                                                                                   \
    {                                                                                  \
      if (!(cond))                                                                     \
        {                                                                              \
          ms_set_error ("%s:%d: bad argument: %s", __FILE__, __LINE__, msg);           \
          return MS_ERR_ARG;                                                           \
        }                                                                              \
    }                                                                                  \
  while (0)

#define MS_TRY(expr)                                                                   \
  do                                                                                   \
    {                                                                                  \
      int r__ = (expr);                                                                \
      if (r__ < 0)                                                                     \
        return r__;                                                                    \
    }                                                                                  \
  while (0)

// ----------------------------------------------------------------------------- layout
// Constants of the device layout (DESIGN.md "Data layout in HBM").
constexpr int MS_SUB = 64;    // sub-block (piece) height: diag kernel / solve granularity
constexpr int MS_CHUNK = 16;  // doubles per 128-byte column chunk (TMA box inner dim)
constexpr int MS_BM = 64;     // tile rows of the DMMA kernel
constexpr int MS_BN = 128;    // tile cols of the DMMA kernel
constexpr int MS_KC = 16;     // k rows per pipeline stage
constexpr int MS_STAGES = 4;  // TMA ring depth
constexpr int MS_NPB = 3;     // panel buffers in rotation (look-ahead depth 2)
constexpr int MS_TILE_CONSUMER_WARPS = 8;

static inline size_t
round_up (size_t v, size_t a)
{
  return (v + a - 1) / a * a;
}

// Left-looking in-block updates from this many remaining columns on: below it the right-looking strips (<= 3 row
// tiles x r/128 column tiles at K = 64) fit one wave of CTAs and are the shorter chain (n = 4096: 4.23 vs 4.45 ms);
// above it the deeper, smaller launches win (n = 16384: 52.4 vs 52.7 ms, n = 65536: 2817.5 vs 2822.1 ms).
constexpr size_t MS_LEFT_MIN_COLS = 8192;

// One piece = up to MS_SUB consecutive rows inside one block row.
struct ms_piece
{
  size_t row0;   // global first row (== first column of its diagonal block)
  int h;         // rows in the piece (<= MS_SUB)
  int block;     // block row I it belongs to
  int owner;     // rank owning that block row
  size_t lrow0;  // local row of row0 on the owner
};

struct ms_ctx
{
  size_t n = 0, m = 0;     // m = the blocking the kernels currently walk (see m_api / m_super)
  size_t m_api = 0;        // the caller's block size: the reference's semantics (packed layout, which rows a
                           // block row's pre-update pivot test covers, row transfers) follow this one
  size_t m_super = 0;      // 0, or the internal blocking of the tensor-core path when the caller's m is a
                           // small multiple of 64 on one GPU: block rows of m_super rows are factored in 64-row
                           // pieces exactly as before, but the trailing update then runs with K = m_super
  int device = 0, rank = 0, nranks = 1;
  size_t nblocks = 0;      // block rows: n/m + (n%m > 0)
  size_t ld = 0;           // leading dimension (doubles), multiple of 16, >= n + 1
  size_t ncols = 0;        // columns the factorisation kernels walk: n, or n + 1 when b rides along as
                           // an augmented column (forward solve fused into the factorisation)
  bool fuse_forward = true;
  bool exact_pivots = false;  // diag kernel with the reference's sqrt / divide roundings ("exact_pivots" option)
  int tile_band_w = 0;     // column tiles per band of the tile walk (0 = default 64; "tile_band_w" option, tests)
  int aug_cols = 1;        // columns reserved right of column n-1 (0 only with MS_FUSE_FORWARD=0 at creation)
  unsigned long long b_version = 1;           // bumped whenever vec[B] is rewritten
  unsigned long long b_version_factored = 0;  // b_version the last ms_factor carried as its augmented column
  size_t rows_local = 0;   // rows stored on this rank
  size_t rows_alloc = 0;   // rows_local rounded up to MS_BM
  size_t hpb = 0;          // panel buffer chunk height: m rounded up to MS_SUB
  bool fast = false;       // m % 64 == 0 -> DMMA/TMA tile kernel
  bool force_generic = false;
  bool lookahead = true;
  bool profile_update = false;
  bool trace = false;       // MS_TRACE=1: timing-enabled step events, read back by ms_debug_trace
  std::vector<float> trace_panel, trace_free, trace_strip;
  int left_looking = 1;     // inside a block row the tensor-core path can update each 64-row piece just before it is
                            // factored, with all pieces above it at once (K = 64, 128, 192) — instead of pushing every
                            // factored piece into all rows below it (K = 64 each time).  0 = never, 2 = always,
                            // 1 = where it pays: block rows with at least MS_LEFT_MIN_COLS columns ("left_looking" option)
  bool panel_trsm = true;   // K3 as a triangular solve (default) or as a DMMA product with the inverse

  double *A = nullptr;       // rows_alloc x ld, row-major; only j >= i meaningful
  double *PB[MS_NPB] = {nullptr, nullptr, nullptr};  // panel buffers, chunk-major [ld/16][hpb][16]
  double *Wt = nullptr;      // D-scaled inverse of the current piece, chunk-major [4][64][16]
  double *vec[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // d b y x r (n each, padded)
  double *acc = nullptr;     // solve accumulators (n)
  double *scratch = nullptr; // reductions / partial sums
  size_t scratch_doubles = 0;
  int *flags = nullptr;      // device int[8]: [0] numerical failure
  int *h_flags = nullptr;    // pinned mirror
  double *h_pinned = nullptr;  // pinned staging, MS_PINNED_DOUBLES each half
  double *d_stage = nullptr;   // device staging for uploads/downloads
  size_t stage_doubles = 0;

  cudaStream_t s_main = nullptr, s_panel = nullptr, s_row = nullptr;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<cudaEvent_t> prof_pool;  // timing events of "profile_update", created by ms_set_option
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;

  CUtensorMap tm_A, tm_PB[MS_NPB], tm_Wt;
  bool have_tm = false;
  CUtensorMap tm_PB3[MS_NPB];  // rank-3 view {16, hpb, ld/16} of the panel buffers (update launches)
  int tm3_state = 0;           // 0 = not tried yet, 1 = encoded, -1 = unavailable or MS_TILE_TMA3=0 (2-D loads instead)

  std::vector<ms_piece> pieces;
  std::vector<size_t> block_first_piece;  // index into pieces, size nblocks+1

  double *sweep_inv = nullptr; // inverses of the local 64 x 64 diagonal pieces (backward sweep), allocated on first use
  int *sweep_ready = nullptr;  // epoch flags of the single-launch sweeps
  int sweep_epoch = 0;
  bool no_sweep = false;       // MS_NO_SWEEP=1: per-piece solve launches
  void *nccl = nullptr;  // ncclComm_t when nranks > 1
  void *p2p = nullptr;   // copy-engine panel transport (ms_p2p.cu) when nranks > 1

  ms_stats stats = {};
  int sm_count = 148;

  // ownership helpers
  int owner_of_block (size_t I) const { return (int)(I % (size_t)nranks); }
  size_t block_rows (size_t I) const { return (I + 1) * m <= n ? m : n - I * m; }
  // local row index of global row g (must be owned by this rank)
  size_t local_row (size_t g) const
  {
    size_t I = g / m;
    return (I / (size_t)nranks) * m + (g - I * m);
  }
  double *row_ptr (size_t g) const { return A + local_row (g) * ld; }
};

// ----------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// chunk-major addressing used by PB and Wt: element (k, j) of a k-major operand
__host__ __device__ __forceinline__ size_t
chunk_major_index (size_t k, size_t j, size_t H)
{
  return ((j / MS_CHUNK) * H + k) * MS_CHUNK + (j % MS_CHUNK);
}

__device__ __forceinline__ uint32_t
smem_u32 (const void *p)
{
  return (uint32_t)__cvta_generic_to_shared (p);
}

__device__ __forceinline__ void
mbar_init (uint64_t *bar, uint32_t count)
{
  asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32 (bar)), "r"(count));
}

__device__ __forceinline__ void
mbar_fence_init ()
{
  asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void
mbar_arrive_expect_tx (uint64_t *bar, uint32_t bytes)
{
  asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32 (bar)),
                "r"(bytes)
                : "memory");
}

__device__ __forceinline__ void
mbar_arrive (uint64_t *bar)
{
  asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32 (bar)) : "memory");
}

__device__ __forceinline__ void
mbar_wait (uint64_t *bar, uint32_t parity)
{
  asm volatile ("{\n"
                ".reg .pred p;\n"
                "WAIT_LOOP:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra WAIT_DONE;\n"
                "bra WAIT_LOOP;\n"
                "WAIT_DONE:\n"
                "}\n" ::"r"(smem_u32 (bar)),
                "r"(parity)
                : "memory");
}

// 2-D TMA tile load: box lands at `dst`, completion counted in bytes on `bar`.
__device__ __forceinline__ void
tma_load_2d (void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
  asm volatile ("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32 (dst)),
                "l"(map), "r"(x), "r"(y), "r"(smem_u32 (bar))
                : "memory");
}

// 3-D TMA box load (x fastest): used with the rank-3 view of a chunk-major panel buffer, where one box
// {16 columns, 16 k rows, 4 chunks} lands as four consecutive 2 KB chunk images.
__device__ __forceinline__ void
tma_load_3d (void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
  asm volatile ("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32 (dst)),
                "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32 (bar))
                : "memory");
}

__device__ __forceinline__ void
tma_prefetch_desc (const CUtensorMap *map)
{
  asm volatile ("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// D(8x8) = A(8x4, row) * B(4x8, col) + C, fp64 — lowers to one SASS DMMA.8x8x4.
__device__ __forceinline__ void
dmma_8x8x4 (double &c0, double &c1, double a, double b)
{
  asm volatile ("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(c0), "+d"(c1)
                : "d"(a), "d"(b));
}

__device__ __forceinline__ double
flip_sign (double v, uint32_t mask_hi)
{
  return __hiloint2double (__double2hiint (v) ^ (int)mask_hi, __double2loint (v));
}

#endif  // __CUDACC__

// ----------------------------------------------------------------------------- launchers
// (implemented across the .cu files; all enqueue on the given stream and count launches)

struct ms_operand  // k-major operand: element (k, j)
{
  const double *p;
  size_t ld;  // row-major leading dimension, or 0 when chunk-major
  size_t H;   // chunk height when chunk-major
};

// ms_kernels_util.cu
int ms_k_fill_formula (ms_ctx *c, int kind, cudaStream_t s);
int ms_k_sym_rows (ms_ctx *c, int mode, const double *xvec, double *out, double *norm_out,
                   cudaStream_t s);
double *ms_sym_abs_partial (ms_ctx *c);
int ms_k_max (ms_ctx *c, const double *v, size_t n, double *d_out, cudaStream_t s);
int ms_k_unpack_rows (ms_ctx *c, const double *d_rows, size_t row0, size_t nrows, size_t ld_src,
                      cudaStream_t s);
int ms_k_packed_rows (ms_ctx *c, double *d_seg, size_t J, size_t w, size_t i0, size_t i1,
                      size_t seg_base, int to_packed, cudaStream_t s);
int ms_k_get_col (ms_ctx *c, size_t j, double *d_out, cudaStream_t s);
int ms_k_norms (ms_ctx *c, const double *r, const double *x, size_t n, double *d_out2,
                cudaStream_t s);
int ms_k_axpy_minus (ms_ctx *c, double *r, const double *b, size_t n, cudaStream_t s);
int ms_k_aug_scatter (ms_ctx *c, const double *b, cudaStream_t s);
int ms_k_aug_gather (ms_ctx *c, double *y, cudaStream_t s);

// ms_kernels_diag.cu
int ms_k_diag (ms_ctx *c, double *blk, size_t ld, int h, double thresh, double *d_out,
               double *wt_out, double *inv_out, const double *pre_diag, size_t pre_ld, int pre_count,
               cudaStream_t s);
int ms_k_pretest (ms_ctx *c, const double *diag0, size_t ld, int first, int count, double thresh, cudaStream_t s);
int ms_k_trsm_panel (ms_ctx *c, const double *Rblk, size_t ldr, double *rows, size_t ld, int h, size_t col_lo,
                     size_t n, const double *dpiece, double *PB, size_t pb_H, size_t pb_k0,
                     cudaStream_t s);

// ms_kernels_tile.cu
int ms_tile_setup (ms_ctx *c);
int ms_k_tile_update (ms_ctx *c, int pb, size_t I, size_t lrow_begin, size_t lrow_end, int K,
                      size_t k0, cudaStream_t s);
int ms_k_tile_panel (ms_ctx *c, int pb, const ms_piece &pc, size_t I, cudaStream_t s);
int ms_k_generic_update (ms_ctx *c, ms_operand P, const double *dsign, size_t lrow_begin,
                         size_t lrow_end, int K, cudaStream_t s);
int ms_k_generic_panel (ms_ctx *c, ms_operand Wt, const double *src_rows, size_t ld_src, double *out_rows,
                        size_t ld_out, int h, size_t col_lo, double *C2, size_t c2_H, size_t c2_k0,
                        size_t n, cudaStream_t s);
int ms_k_generic_block_update (ms_ctx *c, double *A, size_t lda, const double *R1, size_t ld1,
                               const double *D, const double *R2, size_t ld2, int K, int h, int w,
                               int upper_only, cudaStream_t s);

// ms_kernels_solve.cu
int ms_k_forward_step (ms_ctx *c, const ms_piece &pc, double thresh, cudaStream_t s);
int ms_k_backward_step (ms_ctx *c, const ms_piece &pc, size_t lrows_above, double thresh,
                        cudaStream_t s);
int ms_k_sweep (ms_ctx *c, bool forward, double thresh, cudaStream_t s, int *done);

enum { MS_MODE_UPDATE = 0, MS_MODE_PANEL = 1 };
enum { MS_SYM_BNORM = 0, MS_SYM_SYMV = 1 };
