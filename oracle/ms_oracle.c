/* ms_oracle.c — CPU restatement of the reference's blocked R^T D R factorisation,
 * triangular solves, right-hand-side builder and residual.
 *
 * TEST INFRASTRUCTURE ONLY (see ms_oracle.h).  Build with plain -O2: the summation
 * order below is the reference's, and fast-math / FMA contraction would break the
 * bit-for-bit agreement with the reference's -O2 build that the tests pin.
 *
 * Dense row-major storage, leading dimension n unless a routine takes an explicit ld.
 * U(i,j) addresses element (i,j) of such an array.
 */
#include "ms_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define U(p, ld, i, j) ((p)[(size_t)(i) * (ld) + (size_t)(j)])

/* src/defines.h:20-24 */
static int
too_small (double v, double eps)
{
  return fabs (v) < eps;
}

/* ------------------------------------------------------------------ layout */

/* src/init.cpp:4-27: elements rank `rank` of `nranks` allocates (block column J is
 * owned by J % nranks; J*m full rows of width w plus a packed w x w triangle). */
size_t
ora_alloc_size (size_t n, size_t m, int rank, int nranks)
{
  size_t full = n / m, tail = n - full * m;
  size_t cols = full + (tail > 0);
  size_t total = 0;
  for (size_t J = 0; J < cols; J++)
    {
      size_t w = (J == cols - 1 && tail > 0) ? tail : m;
      if ((int)(J % (size_t)nranks) == rank)
        total += J * m * w + (w * w + w) / 2;
    }
  return total;
}

/* src/defines.h:12-18: row-major packed upper triangle of width w. */
size_t
ora_packed_index (size_t i, size_t j, size_t w)
{
  return i * w + j - (i * i + i) / 2;
}

/* src/init.cpp:29-58 + src/solver.h:4-9 at one rank: where (i,j), j>=i, lives. */
size_t
ora_packed_offset (size_t i, size_t j, size_t n, size_t m)
{
  size_t full = n / m, tail = n - full * m;
  size_t J = j / m;
  if (J > full)
    J = full;
  size_t w = (J == full && tail > 0) ? tail : m;
  size_t base = 0;
  for (size_t K = 0; K < J; K++)
    base += K * m * m + (m * m + m) / 2;
  if (i < J * m)
    return base + i * w + (j - J * m);
  return base + J * m * w + ora_packed_index (i - J * m, j - J * m, w);
}

/* -------------------------------------------------------------------- init */

/* src/init.cpp:60-98 with f1 of test/test_mat.cpp:13-17: integer n - max(i,j). */
void
ora_init_formula (double *a, size_t n)
{
  for (size_t i = 0; i < n; i++)
    for (size_t j = i; j < n; j++)
      U (a, n, i, j) = (double)(n - j);
}

/* src/init.cpp:100-210: n rows of n whitespace separated decimals; only j >= i kept. */
int
ora_init_file (double *a, size_t n, const char *path)
{
  FILE *f = fopen (path, "r");
  if (!f)
    return 0;
  int status = 1;
  for (size_t i = 0; i < n && status == 1; i++)
    for (size_t j = 0; j < n; j++)
      {
        double v;
        if (fscanf (f, "%lf", &v) != 1)
          {
            status = -1;
            break;
          }
        if (j >= i)
          U (a, n, i, j) = v;
      }
  fclose (f);
  return status;
}

static double
sym_at (const double *a, size_t n, size_t i, size_t k)
{
  return k >= i ? U (a, n, i, k) : U (a, n, k, i);
}

/* src/init.cpp:492-531: b_i = sum over even k of the symmetric row, norm = max abs
 * row sum; both accumulated left to right over k = 0..n-1. */
void
ora_b_norm (const double *a, size_t n, double *b, double *norm)
{
  double best = 0;
  for (size_t i = 0; i < n; i++)
    {
      double s = 0, sa = 0;
      for (size_t k = 0; k < n; k++)
        {
          double v = sym_at (a, n, i, k);
          if (k % 2 == 0)
            s += v;
          sa += fabs (v);
        }
      b[i] = s;
      if (sa > best)
        best = sa;
    }
  *norm = best;
}

/* ----------------------------------------------------------- block kernels */

/* src/solver.cpp:106-160.  Row i: pivot s = a_ii - sum_k a_ki^2 d_k; d_i = sign(s)
 * with sign(0) = -1 (src/defines.h:26-30); the OLD a_ii is tested first (:117), then
 * sqrt|s| (:122); the rest of the row is (a_ij - sum_k a_ki a_kj d_k)/(a_ii d_i). */
int
ora_diag_factor (double *a, size_t ld, double *d, int w, double norma)
{
  for (int i = 0; i < w; i++)
    {
      double s = U (a, ld, i, i);
      for (int k = 0; k < i; k++)
        s -= U (a, ld, k, i) * U (a, ld, k, i) * d[k];
      d[i] = s > 0 ? 1 : -1;
      if (too_small (U (a, ld, i, i), ORA_EPS * norma))
        return 1;
      U (a, ld, i, i) = sqrt (fabs (s));
      if (too_small (U (a, ld, i, i), ORA_EPS * norma))
        return 1;
      for (int j = i + 1; j < w; j++)
        {
          double t = U (a, ld, i, j);
          for (int k = 0; k < i; k++)
            t -= U (a, ld, k, i) * U (a, ld, k, j) * d[k];
          U (a, ld, i, j) = t / (U (a, ld, i, i) * d[i]);
        }
    }
  return 0;
}

/* src/matrix_op.cpp:244-285.  Rows bottom-up; inv_ii = 1/a_ii and
 * inv_ij = -(sum_{k=i+1..j} a_ik inv_kj)/a_ii.  `inv` is dense w x w, ld = w. */
int
ora_upper_inverse (const double *a, size_t ld, double *inv, int w, double norma)
{
  for (int i = w - 1; i >= 0; i--)
    {
      if (too_small (U (a, ld, i, i), ORA_EPS * norma))
        return 1;
      U (inv, w, i, i) = 1 / U (a, ld, i, i);
      for (int j = i + 1; j < w; j++)
        {
          double s = 0;
          for (int k = i + 1; k <= j; k++)
            s += U (a, ld, i, k) * U (inv, w, k, j);
          U (inv, w, i, j) = -s / U (a, ld, i, i);
        }
    }
  return 0;
}

/* src/matrix_op.cpp:4-39 (square), :156-222 (h x w tail), :41-99 and :224-242
 * (upper_only: diagonal block).  A(h x w) -= R1^T(k x h) diag(D) R2(k x w): each entry
 * accumulates (R1*R2)*D over the k rows in ascending order, then ONE subtraction. */
void
ora_update (double *A, size_t lda, const double *R1, size_t ld1, const double *D,
            const double *R2, size_t ld2, int k, int h, int w, int upper_only)
{
  for (int i = 0; i < h; i++)
    for (int j = upper_only ? i : 0; j < w; j++)
      {
        double s = 0;
        for (int q = 0; q < k; q++)
          s += U (R1, ld1, q, i) * U (R2, ld2, q, j) * D[q];
        U (A, lda, i, j) -= s;
      }
}

/* src/matrix_op.cpp:287-368.  In place A <- D * inv^T * A, rows bottom-up so that the
 * rows k <= i still hold the input when row i is formed. */
void
ora_panel (const double *inv, double *A, size_t lda, const double *D, int h, int w)
{
  for (int i = h - 1; i >= 0; i--)
    for (int j = 0; j < w; j++)
      {
        double s = 0;
        for (int k = 0; k <= i; k++)
          s += U (inv, h, k, i) * U (A, lda, k, j);
        U (A, lda, i, j) = s * D[i];
      }
}

/* ------------------------------------------------------------------ driver */

/* src/solver.cpp:8-104 at one rank (rank count changes who computes a block, not how:
 * SURVEY.md §8c).  Left-looking by block row I: diagonal block first, then every
 * block to its right; each block receives one update per earlier block row k. */
int
ora_factor (double *a, double *d, size_t n, size_t m, double norma)
{
  size_t full = n / m, tail = n - full * m;
  double *inv = (double *)calloc (m > 0 ? m * m : 1, sizeof (double));
  int fail = 0;
  for (size_t I = 0; I < full; I++)
    {
      double *T = &U (a, n, I * m, I * m);
      for (size_t k = 0; k < I; k++)
        {
          const double *Rk = &U (a, n, k * m, I * m);
          ora_update (T, n, Rk, n, d + k * m, Rk, n, (int)m, (int)m, (int)m, 1);
        }
      if (ora_diag_factor (T, n, d + I * m, (int)m, norma))
        fail = 1;
      if (ora_upper_inverse (T, n, inv, (int)m, norma))
        fail = 1;
      if (fail)
        break;
      for (size_t J = I + 1; J < full + (tail > 0); J++)
        {
          int w = J == full ? (int)tail : (int)m;
          double *B = &U (a, n, I * m, J * m);
          for (size_t k = 0; k < I; k++)
            ora_update (B, n, &U (a, n, k * m, I * m), n, d + k * m,
                        &U (a, n, k * m, J * m), n, (int)m, (int)m, w, 0);
          ora_panel (inv, B, n, d + I * m, (int)m, w);
        }
    }
  if (!fail && tail > 0)
    {
      double *T = &U (a, n, full * m, full * m);
      for (size_t k = 0; k < full; k++)
        {
          const double *Rk = &U (a, n, k * m, full * m);
          ora_update (T, n, Rk, n, d + k * m, Rk, n, (int)m, (int)tail,
                      (int)tail, 1);
        }
      if (ora_diag_factor (T, n, d + full * m, (int)tail, norma))
        fail = 1;
    }
  free (inv);
  return fail;
}

/* src/solver.cpp:162-208: R^T y = b, column j at a time, sum over i < j ascending. */
int
ora_forward (const double *r, const double *b, double *y, size_t n, double norma)
{
  for (size_t j = 0; j < n; j++)
    {
      double s = 0;
      for (size_t i = 0; i < j; i++)
        s += U (r, n, i, j) * y[i];
      if (too_small (U (r, n, j, j), ORA_EPS * norma))
        return 1;
      y[j] = (b[j] - s) / U (r, n, j, j);
    }
  return 0;
}

/* src/solver.cpp:210-245: x_i = d_i (y_i - d_i sum_{j>i} R_ij x_j) / R_ii, i down. */
int
ora_backward (const double *r, double *x, const double *y, const double *d,
              size_t n, double norma)
{
  for (size_t i = n; i-- > 0;)
    {
      double s = 0;
      for (size_t j = i + 1; j < n; j++)
        s += U (r, n, i, j) * x[j];
      if (too_small (U (r, n, i, i), ORA_EPS * norma))
        return 1;
      x[i] = d[i] * (y[i] - d[i] * s) / U (r, n, i, i);
    }
  return 0;
}

/* ------------------------------------------------------------------ checks */

/* src/init.cpp:533-563: r = A x - b over the full symmetric row, k ascending. */
void
ora_residual (const double *a, const double *b, const double *x, double *r,
              size_t n)
{
  for (size_t i = 0; i < n; i++)
    {
      double s = 0;
      for (size_t k = 0; k < n; k++)
        s += sym_at (a, n, i, k) * x[k];
      r[i] = s - b[i];
    }
}

/* src/matrix_op.cpp:524-533 */
double
ora_norm2 (const double *v, size_t n)
{
  double s = 0;
  for (size_t i = 0; i < n; i++)
    s += v[i] * v[i];
  return sqrt (s);
}

/* src/matrix_op.cpp:534-548: distance to the planted solution (1,0,1,0,...). */
double
ora_error (const double *x, size_t n)
{
  double s = 0;
  for (size_t i = 0; i < n; i++)
    {
      double t = x[i];
      if (i % 2 == 0)
        t -= 1.;
      s += t * t;
    }
  return sqrt (s);
}
