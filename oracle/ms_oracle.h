/* ms_oracle.h — CPU restatement of the reference's factor/solve path.
 *
 * TEST INFRASTRUCTURE ONLY: this is the checker the CUDA path is compared against.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  Nothing under matrix_solver_b200/ includes or links it.
 *
 * Every routine follows the arithmetic ORDER of the reference routine it cites, so
 * that, compiled with plain `-O2` (no fast-math, no FMA contraction), it reproduces
 * the reference's `-O2` build bit for bit (pinned by tests/test_oracle_vs_ref.py and
 * by tests/golden/).  Storage differs on purpose: the oracle keeps A / R as a dense
 * row-major n x n array (leading dimension n) of which only j >= i is meaningful;
 * the reference's packed block-column layout is restated separately in
 * ora_alloc_size / ora_packed_offset because the host mirror has to reproduce it.
 */
#ifndef MS_ORACLE_H
#define MS_ORACLE_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_EPS 1.e-14 /* src/defines.h:9 */

/* ---- layout (src/init.cpp:4-58, src/solver.h:4-9, src/defines.h:12-18) ---- */
size_t ora_alloc_size (size_t n, size_t m, int rank, int nranks);
size_t ora_packed_index (size_t i, size_t j, size_t w);
/* offset of logical element (i, j), j >= i, inside the single-rank allocation */
size_t ora_packed_offset (size_t i, size_t j, size_t n, size_t m);

/* ---- init (src/init.cpp:60-98,100-294; test/test_mat.cpp:13-17) ---- */
void ora_init_formula (double *a, size_t n);
/* returns 1 on success, 0 on open failure, -1 on short file (src/init.cpp:128,135) */
int ora_init_file (double *a, size_t n, const char *path);
void ora_b_norm (const double *a, size_t n, double *b, double *norm);

/* ---- block kernels (src/solver.cpp:106-160, src/matrix_op.cpp:4-368) ----
 * Blocks are addressed inside the dense array: pointer to the block's (0,0) element
 * plus the leading dimension.  h x w shapes replace the reference's n / l pairs. */
int ora_diag_factor (double *a, size_t ld, double *d, int w, double norma);
int ora_upper_inverse (const double *a, size_t ld, double *inv, int w,
                       double norma);
void ora_update (double *A, size_t lda, const double *R1, size_t ld1,
                 const double *D, const double *R2, size_t ld2, int k, int h,
                 int w, int upper_only);
void ora_panel (const double *inv, double *A, size_t lda, const double *D, int h,
                int w);

/* ---- driver and solves (src/solver.cpp:8-104,162-245) ---- */
int ora_factor (double *a, double *d, size_t n, size_t m, double norma);
int ora_forward (const double *r, const double *b, double *y, size_t n,
                 double norma);
int ora_backward (const double *r, double *x, const double *y, const double *d,
                  size_t n, double norma);

/* ---- checks (src/init.cpp:533-563, src/matrix_op.cpp:524-548) ---- */
void ora_residual (const double *a, const double *b, const double *x, double *r,
                   size_t n);
double ora_norm2 (const double *v, size_t n);
double ora_error (const double *x, size_t n);

#ifdef __cplusplus
}
#endif
#endif
