/* ms_b200.h — C ABI of the B200-native R^T D R (Cholesky) factor + solve library.
 *
 * This is the drop-in boundary: plain C, pointers and sizes only, no
 * C++/torch types.  The reference has no FFI of its own; the functions below are what
 * its C++ entry points bind to in our host mirror (matrix_solver_b200/host/), and the
 * reference file:line each one stands in for is cited next to it.  INTEGRATION.md shows
 * the binding.
 *
 * Model: one context = one rank = one GPU.  The matrix lives on the device for the
 * whole run in block-row cyclic layout (block row I, rows [I*m, I*m+m), belongs to rank
 * I % nranks — the reference's rule src/init.cpp:20 applied to rows of R instead of
 * columns); the vectors d, b, y, x, r live on the device too.  Host buffers cross the
 * boundary only in the upload/download/get/set calls.
 *
 * Return codes: 0 = ok; 1 = NUMERICAL failure (tiny pivot — the reference's `return 1`
 * in src/solver.cpp:51,119,124,198,235 and src/matrix_op.cpp:252); < 0 = CUDA / NCCL /
 * argument error, text available from ms_last_error().  There is no CPU fallback: every
 * compute entry point fails with MS_ERR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef MS_B200_H
#define MS_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MS_OK 0
#define MS_NUMERICAL_FAILURE 1
#define MS_ERR_CUDA (-1)
#define MS_ERR_ARG (-2)
#define MS_ERR_NO_DEVICE (-3)
#define MS_ERR_NCCL (-4)
#define MS_ERR_IO (-5)

#define MS_UNIQUE_ID_BYTES 128

typedef struct ms_ctx ms_ctx;

/* which device-resident vector (test/test_mat.cpp:102-107: d | b | y | x | r) */
enum { MS_VEC_D = 0, MS_VEC_B = 1, MS_VEC_Y = 2, MS_VEC_X = 3, MS_VEC_R = 4 };

/* generators (test/test_mat.cpp:13-22): 0 = f1 = n - max(i,j); 1 = hilb = 1/(i+j+1) */
enum { MS_FORMULA_F1 = 0, MS_FORMULA_HILBERT = 1 };

typedef struct ms_stats
{
  double factor_ms;       /* device time of the last ms_factor (CUDA events) */
  double update_ms;       /* device time in which at least one rank-m trailing-update launch was running:
                             the UNION of the launch intervals (0 unless ms_set_option(ctx,
                             "profile_update", 1)); never exceeds factor_ms */
  double forward_ms;      /* last ms_solve_forward */
  double backward_ms;     /* last ms_solve_backward */
  double init_ms;         /* last ms_init_formula / upload */
  double bnorm_ms;        /* last ms_b_and_norm */
  double residual_ms;     /* last ms_residual */
  double update_flops;    /* algorithmic flops issued by the trailing-update launches */
  long long kernel_launches; /* kernels of this library launched since ms_create */
  long long update_launches; /* trailing-update launches in the last ms_factor */
  long long h2d_bytes, d2h_bytes; /* bytes moved across PCIe since ms_create */
  int used_dmma;          /* 1 if the DMMA/TMA tile kernel ran in the last factor */
  int forward_fused;      /* 1 if the last ms_solve_forward only read back the y that ms_factor carried
                             along as an augmented column (no sweep over R) */
  double update_sum_ms;   /* plain sum of the same intervals (look-ahead lets two of them overlap) */
  int block_internal;     /* rows per block row the last ms_factor walked: m, or the internal blocking (256) a
                             single-GPU context uses for m = 64 / 128 on the tensor-core path (same R and d, the
                             reference's pivot pre-tests still per m rows; MS_SUPER_BLOCK=0 switches it off) */
} ms_stats;

/* ---- library / error state -------------------------------------------------- */
const char *ms_last_error (void);
int ms_device_count (int *count);

/* ---- lifetime (replaces compute_alloc_size + alloc_rows, src/init.cpp:4-58) -- */
/* rank 0 calls ms_get_unique_id and hands the 128 bytes to every rank by any means
 * (torch.distributed broadcast, a pipe, a file); ignored when nranks == 1.  The id bootstraps an
 * NCCL communicator (reductions of the solves, b/norm, flags); the factored panels themselves move
 * over CUDA-IPC peer memory with copy engines unless MS_P2P=0 (then ncclBroadcast).
 * Options for ms_set_option (0/1 unless noted): "lookahead", "force_generic", "panel_trsm",
 * "single_launch_sweeps", "profile_update", "fuse_forward" (b rides through ms_factor as an augmented
 * column; default 1), "exact_pivots" (the reference's sqrt / divide roundings in every diagonal piece;
 * default 0: full 64-row pieces use rsqrt and a multiply), "tile_tma3", "tile_band_w" (column tiles per
 * band of the update's tile walk; 0 = default 64). */
int ms_get_unique_id (void *id128);
int ms_create (ms_ctx **out, size_t n, size_t m, int device, int rank, int nranks,
               const void *unique_id);
int ms_destroy (ms_ctx *ctx);
int ms_set_option (ms_ctx *ctx, const char *key, long value);
int ms_get_stats (ms_ctx *ctx, ms_stats *out);
int ms_sync (ms_ctx *ctx);

/* ---- init (src/init.cpp:60-98 init_mat; :100-294 init_mat_file) ------------- */
int ms_init_formula (ms_ctx *ctx, int kind);
/* `rows` holds nrows full rows (ld_host doubles apart) starting at global row row0;
 * only j >= i is used, like the reference's file reader.  Rows owned by other ranks are
 * skipped.  Page-locked sources are DMA-ed straight into place (one strided copy per block row);
 * pageable sources are staged through the library's pinned buffer first. */
int ms_upload_rows (ms_ctx *ctx, const double *rows, size_t row0, size_t nrows,
                    size_t ld_host);
/* the same for page-locked rows without waiting: the DMA is queued and the call returns; `rows` must not be
 * touched until ms_sync (or a later synchronous call) returns.  Lets a reader parse the next chunk of a
 * file while the previous one is on its way (host/init.cpp init_mat_file). */
int ms_upload_rows_async (ms_ctx *ctx, const double *rows, size_t row0, size_t nrows, size_t ld_host);
/* the reverse: the upper part (j >= i, from the start of each row's diagonal block) of the rows
 * this rank owns inside [row0, row0+nrows) is written into `rows`; other entries are untouched */
int ms_download_rows (ms_ctx *ctx, double *rows, size_t row0, size_t nrows, size_t ld_host);
/* page-locked host memory for callers that want the direct-DMA path of the transfers above */
int ms_host_alloc (void **out, size_t bytes);
int ms_host_free (void *p);
/* whole matrix in the reference's single-rank packed block-column layout
 * (src/init.cpp:29-98): n(n+1)/2 doubles, nranks == 1 only. */
int ms_upload_packed (ms_ctx *ctx, const double *body);
int ms_download_packed (ms_ctx *ctx, double *body);
/* gather_row / gather_col (src/init.cpp:309-490): buf[i..n) resp. buf[0..j]; collective when
 * nranks > 1 (every rank calls, every rank receives) */
int ms_get_row (ms_ctx *ctx, size_t i, double *buf);
int ms_get_col (ms_ctx *ctx, size_t j, double *buf);
/* a window of the stored matrix, rows [row0, row0+nrows) x columns [col0, col0+ncols) -> buf (row-major):
 * what print_mat_beauty (src/matrix_op.cpp:443-505) needs for its two corners, as strided copies;
 * entries left of a row's diagonal block are whatever the factorisation left there (never printed).
 * Collective when nranks > 1. */
int ms_get_block (ms_ctx *ctx, size_t row0, size_t nrows, size_t col0, size_t ncols, double *buf);

/* ---- right-hand side and norm (src/init.cpp:492-531 init_b_get_norm) -------- */
int ms_b_and_norm (ms_ctx *ctx, double *norm);

/* ---- factor and solves (src/solver.cpp:8-104, :162-208, :210-245) ----------- */
int ms_factor (ms_ctx *ctx, double norma, double *elapsed_s);
int ms_solve_forward (ms_ctx *ctx, double norma);
int ms_solve_backward (ms_ctx *ctx, double norma);

/* ---- residual and norms (src/init.cpp:533-563; src/matrix_op.cpp:524-548) ---- */
/* r = A x - b with the matrix currently stored (the caller re-initialises it first,
 * test/test_mat.cpp:149-160); returns ||r||_2 and ||x - (1,0,1,0,...)||_2 */
int ms_residual (ms_ctx *ctx, double *resid2, double *err2);
int ms_set_vec (ms_ctx *ctx, int which, const double *host, size_t count);
int ms_get_vec (ms_ctx *ctx, int which, double *host, size_t count);

/* ---- stateless block kernels on HOST buffers (device 0 unless MS_DEVICE is set) --
 * One-to-one with the reference's block routines so each kernel can be checked in
 * isolation; layouts are the reference's (packed upper = src/defines.h:12-18). */
/* cholesky_decomp_U, src/solver.cpp:106-160 */
int ms_block_diag_factor (double *a_packed, double *d, int w, double norma);
/* reverse_upper, src/matrix_op.cpp:244-285 */
int ms_block_upper_inverse (const double *a_packed, double *inv_packed, int w,
                            double norma);
/* DRtA / DRtA_l, src/matrix_op.cpp:287-368: A(h x w) <- D inv^T A */
int ms_block_panel (const double *inv_packed, double *A, const double *D, int h,
                    int w);
/* the same quantity from the FACTOR itself (what the factorisation's panel step runs):
 * A(h x w) <- D R^{-T} A by a triangular solve, R packed upper, h <= 64 */
int ms_block_trsm_panel (const double *r_packed, double *A, const double *D, int h, int w);
/* minus_RTDR_o1 / minus_RTDR_l_o1, src/matrix_op.cpp:4-39,156-222:
 * A(h x w) -= R1(k x h)^T diag(D) R2(k x w) */
int ms_block_update (double *A, const double *R1, const double *D, const double *R2,
                     int k, int h, int w);
/* minus_RTDRu_ofast / minus_RTDRu_l, src/matrix_op.cpp:41-99,224-242:
 * packed-upper A(w x w) -= upper(R1(k x w)^T diag(D) R1) */
int ms_block_update_diag (double *a_packed, const double *R1, const double *D, int k,
                          int w);
/* norma_vec2 / find_error, src/matrix_op.cpp:524-548 */
int ms_norm2 (const double *v, size_t n, double *out);
int ms_error_vs_planted (const double *x, size_t n, double *out);

/* ---- microbenchmarks used by bench.py for the fp64 roofline denominator ------ */
/* register-only DMMA.8x8x4 issue loop on every SM: TFLOP/s */
int ms_probe_dmma_peak (int device, double *tflops);
/* streaming copy: GB/s (read + write) */
int ms_probe_hbm_copy (int device, size_t bytes, double *gbs);

#ifdef __cplusplus
}
#endif
#endif
