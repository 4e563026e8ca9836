// solver.cpp — host mirror of the reference's src/solver.cpp: thin calls into the CUDA library.
//   cholesky_decomp   :8-104    -> ms_factor   (right-looking, DMMA trailing update, look-ahead)
//   cholesky_decomp_U :106-160  -> ms_block_diag_factor (the diag kernel on one host block)
//   compute_y         :162-208  -> ms_solve_forward
//   compute_x         :210-245  -> ms_solve_backward
// Polarity as in the reference: true / 1 means "pivot too small".
#include "solver.h"

#include <cstdio>
#include <cstdlib>

#include "ms_rt.h"

static ms_ctx *
ctx_of (double **rows_p)
{
  ms_ctx *c = ms_rt_ctx (rows_p);
  if (!c)
    {
      fprintf (stderr, "matrix_solver_b200: rows_p was not produced by alloc_rows\n");
      abort ();
    }
  return c;
}

static bool
verdict (int rc, const char *what)
{
  if (rc < 0)
    ms_rt_die (what, rc);
  return rc == MS_NUMERICAL_FAILURE;
}

double g_last_factor_seconds = 0;  // device time of the last cholesky_decomp (CUDA events)

bool
cholesky_decomp (double **&rows_p, vec d, size_t n, size_t m, double norma)
{
  (void)m;
  ms_ctx *c = ctx_of (rows_p);
  bool failed = verdict (ms_factor (c, norma, &g_last_factor_seconds), "ms_factor");
  if (d && ms_get_vec (c, MS_VEC_D, d, n) < 0)
    ms_rt_die ("ms_get_vec", -1);
  return failed;
}

bool
cholesky_decomp_U (mat a, vec d, int n, double norma)
{
  return verdict (ms_block_diag_factor (a, d, n, norma), "ms_block_diag_factor");
}

bool
compute_y (double **&rows_p, vec b, vec y, size_t n, size_t m, double norma)
{
  (void)m;
  ms_ctx *c = ctx_of (rows_p);
  if (b && ms_set_vec (c, MS_VEC_B, b, n) < 0)
    ms_rt_die ("ms_set_vec", -1);
  bool failed = verdict (ms_solve_forward (c, norma), "ms_solve_forward");
  if (y && ms_get_vec (c, MS_VEC_Y, y, n) < 0)
    ms_rt_die ("ms_get_vec", -1);
  return failed;
}

bool
compute_x (double **&rows_p, vec x, vec y, vec d, size_t n, size_t m, double norma, int root)
{
  (void)m;
  (void)root;
  ms_ctx *c = ctx_of (rows_p);
  if (y && ms_set_vec (c, MS_VEC_Y, y, n) < 0)
    ms_rt_die ("ms_set_vec", -1);
  if (d && ms_set_vec (c, MS_VEC_D, d, n) < 0)
    ms_rt_die ("ms_set_vec", -1);
  bool failed = verdict (ms_solve_backward (c, norma), "ms_solve_backward");
  if (x && ms_get_vec (c, MS_VEC_X, x, n) < 0)
    ms_rt_die ("ms_get_vec", -1);
  return failed;
}
