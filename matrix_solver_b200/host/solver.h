// solver.h — factorisation and triangular-solve entry points of the drop-in API
// (same names and signatures as the reference's src/solver.h:4-24).
#ifndef SOLVER_H
#define SOLVER_H
#include "defines.h"

// address of block (I, J) inside block column J: I full m x w blocks precede it
inline double *
get_bl (double **&rows_p, size_t I, size_t J, size_t m, size_t l = 0)
{
  const size_t w = l > 0 ? l : m;
  return rows_p[J] + I * m * w;
}

// A = R^T D R in place; d receives the +-1 diagonal on every rank; true = tiny pivot
bool cholesky_decomp (double **&rows_p, vec d, size_t n, size_t m, double norma);
// the same for one packed-upper n x n block held in HOST memory
bool cholesky_decomp_U (mat a, vec d, int n, double norma);
// R^T y = b
bool compute_y (double **&rows_p, vec b, vec y, size_t n, size_t m, double norma);
// D R x = y
bool compute_x (double **&rows_p, vec x, vec y, vec d, size_t n, size_t m, double norma, int root);
#endif
