// matrix_op.cpp — host mirror of the reference's src/matrix_op.cpp.
//   minus_RTDR_o1 / _l_o1 (:4-39, :156-222)     -> ms_block_update       (tile contraction)
//   minus_RTDRu_ofast / _l (:41-99, :224-242)   -> ms_block_update_diag
//   reverse_upper (:244-285)                    -> ms_block_upper_inverse
//   DRtA / DRtA_l (:287-368)                    -> ms_block_panel
//   print_mat_beauty / print_vec (:443-522)     -> same text, the two printed corners fetched as two strided
//                                                  window copies (ms_get_block)
//   norma_vec2 / find_error (:524-548)          -> ms_norm2 / ms_error_vs_planted
#include "matrix_op.h"

#include <vector>

#include "init.h"
#include "ms_rt.h"

static void
must (int rc, const char *what)
{
  if (rc < 0)
    ms_rt_die (what, rc);
}

void
minus_RTDR_o1 (mat A, mat R1, vec D, mat R2, int n)
{
  must (ms_block_update (A, R1, D, R2, n, n, n), "ms_block_update");
}

void
minus_RTDR_l_o1 (mat A, mat R1, vec D, mat R2, int n, int l)
{
  must (ms_block_update (A, R1, D, R2, n, n, l), "ms_block_update");
}

void
minus_RTDRu_ofast (mat A, mat R1, vec D, mat R2, int n)
{
  (void)R2;  // the reference's tail loop already reads R1 for both factors (:93)
  must (ms_block_update_diag (A, R1, D, n, n), "ms_block_update_diag");
}

void
minus_RTDRu_l (mat A, mat R1, vec D, int n, int l)
{
  must (ms_block_update_diag (A, R1, D, n, l), "ms_block_update_diag");
}

bool
reverse_upper (mat A, mat B, int n, double norma)
{
  int rc = ms_block_upper_inverse (A, B, n, norma);
  must (rc, "ms_block_upper_inverse");
  return rc == MS_NUMERICAL_FAILURE;
}

void
DRtA (mat R, mat A, vec D, int n)
{
  must (ms_block_panel (R, A, D, n, n), "ms_block_panel");
}

void
DRtA_l (mat R, mat A, vec D, int n, int l)
{
  must (ms_block_panel (R, A, D, n, l), "ms_block_panel");
}

// Corner view of the upper triangle: the leading len x len part, then the trailing rows from
// max(n - len, len) on, each prefixed by a run of dots.  Skipped cells are blanks of the cell width.
void
print_mat_beauty (int root, size_t n, size_t m, double **&rows_p, size_t len, bool log, int width,
                  int precision)
{
  (void)m;
  if (len == -1UL || len > n)
    len = n;
  const bool me = ms_rt_rank () == root;
  const int cell = width + (log ? 1 : 0);
  auto put = [&] (double v) { printf (log ? "% *.*e " : "%*.*lf ", width, precision, v); };
  auto blanks = [&] (size_t cells) {
    for (size_t c = 0; c < cells; c++)
      printf ("%*s ", cell, "");
  };
  // the two corners arrive as two strided window copies (ms_get_block), not one gathered row per line
  ms_ctx *c = ms_rt_ctx (rows_p);
  if (!c)
    ms_rt_die ("print_mat_beauty: rows_p was not produced by alloc_rows", -1);
  std::vector<double> win (len * len);
  if (len > 0)
    must (ms_get_block (c, 0, len, 0, len, win.data ()), "ms_get_block");
  if (me)
    for (size_t i = 0; i < len; i++)
      {
        blanks (i);
        for (size_t j = i; j < len; j++)
          put (win[i * len + j]);
        printf ("\n");
      }
  const size_t first = n - len > len ? n - len : len;
  const size_t tail = n - first;
  if (tail > 0)
    {
      win.assign (tail * tail, 0.0);
      must (ms_get_block (c, first, tail, first, tail, win.data ()), "ms_get_block");
    }
  if (me)
    for (size_t i = first; i < n; i++)
      {
        for (int k = 0; k < cell; k++)
          putchar ('.');
        putchar (' ');
        blanks (i - first);
        for (size_t j = i; j < n; j++)
          put (win[(i - first) * tail + (j - first)]);
        printf ("\n");
      }
}

void
print_vec (vec a, size_t n, size_t len)
{
  if (len == -1UL || len > n)
    len = n;
  for (size_t i = 0; i < len; i++)
    printf ("%lf ", a[i]);
  if (len > 0)
    printf ("\n");
}

double
norma_vec2 (double *r, int n)
{
  double out = 0;
  must (ms_norm2 (r, (size_t)n, &out), "ms_norm2");
  return out;
}

double
find_error (double *x, int n)
{
  double out = 0;
  must (ms_error_vs_planted (x, (size_t)n, &out), "ms_error_vs_planted");
  return out;
}
