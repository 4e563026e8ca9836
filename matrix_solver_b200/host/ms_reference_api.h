// ms_reference_api.h — the complete drop-in surface in one place.
//
// Code written against the reference includes "defines.h", "init.h", "solver.h" and "matrix_op.h";
// in this directory those four names are one-line forwarders to this header.  Everything the
// reference exposes there is declared here with the same name, argument list and polarity
// (numeric routines: true = pivot too small; I/O routines: 1 = success) — what differs is where
// the work happens (the CUDA library behind include/ms_b200.h) and that `rows_p` entries are opaque.
//
//   group                       reference header : lines          implemented in
//   types, constants, helpers   src/defines.h    : 7-40           here (inline)
//   storage / init / gathers    src/init.h       : 10-56          init.cpp
//   scope guard                 src/init.h       : 58-76          here (inline)
//   factor and solves           src/solver.h     : 4-24           solver.cpp
//   block kernels, print, norms src/matrix_op.h  : 6-53           matrix_op.cpp
#ifndef MS_REFERENCE_API_H
#define MS_REFERENCE_API_H

#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <utility>

// ------------------------------------------------------------------ types and constants
using mat = double *;
using vec = double *;
using initializer = double (*) (size_t i, size_t j, size_t);

constexpr double EPS = 1.e-14;  // pivots are compared with EPS * ||A||_inf
constexpr int MAX = 8;

// ------------------------------------------------------------------ small inline helpers
// Offset of (i, j), j >= i, in a row-major packed upper triangle of order n: rows 0..i-1 hold
// n, n-1, ... entries (i*n - i(i-1)/2 in total), then j - i.
inline size_t
get_elU (size_t i, size_t j, size_t n)
{
  return i * n - (i * i - i) / 2 + (j - i);
}

inline bool
is_double_too_small (double a, double eps)
{
  return std::fabs (a) < eps;
}

// sign with sgn(0) == -1, as the factorisation's D needs it
inline int
sgn (double a)
{
  return a > 0 ? 1 : -1;
}

inline void
assert_and_do (bool ok, std::function<void ()> on_failure)
{
  if (ok)
    return;
  on_failure ();
  assert (ok);
}

// address of block (I, J) inside block column J: I full m x w blocks precede it
inline double *
get_bl (double **&rows_p, size_t I, size_t J, size_t m, size_t l = 0)
{
  const size_t w = l > 0 ? l : m;
  return rows_p[J] + I * m * w;
}

// runs a callable when the scope is left (the reference's driver finalises MPI this way)
class ScopeGuard
{
public:
  template <class Callable> ScopeGuard (Callable &&fn) : fn_ (std::forward<Callable> (fn)) {}
  ScopeGuard (ScopeGuard &&other) : fn_ (std::move (other.fn_)) { other.fn_ = nullptr; }
  ~ScopeGuard ()
  {
    if (fn_)
      fn_ ();
  }
  ScopeGuard (const ScopeGuard &) = delete;
  void operator= (const ScopeGuard &) = delete;

private:
  std::function<void ()> fn_;
};

// ------------------------------------------------------------------ generators
// the two generators of the reference's driver (test/test_mat.cpp:13-22); init_mat runs them on the
// device when it is handed one of these (or any function that agrees with them)
double f1 (size_t i, size_t j, size_t n);
double hilb (size_t i, size_t j, size_t n);

// ------------------------------------------------------------------ storage, init, gathers
size_t compute_alloc_size (size_t n, size_t m);
double *alloc_rows (size_t n, size_t m, size_t n_to_alloc, double **rows_p);
void init_mat (double **rows_p, size_t n, size_t m, initializer f);
bool init_mat_file (double **rows_p, size_t n, size_t m, const char *filename);
bool check_args (const int argc, char **argv);
void gather_row (size_t i, size_t n, size_t m, int root, vec buf, double **rows_p, int action = 0);
void gather_col (size_t j, size_t n, size_t m, int root, vec buf, double **&rows_p);
void init_b_get_norm (vec b, size_t n, size_t m, int root, double **&rows_p, double &norm);
void find_disrep_vec (vec b, vec x, vec r, size_t n, size_t m, int root, double **&rows_p);
bool check_res (int root, int loc_res);
void printf_root (int root, const char *format, ...);

// ------------------------------------------------------------------ factor and solves
// A = R^T D R in place; d receives the +-1 diagonal on every rank
bool cholesky_decomp (double **&rows_p, vec d, size_t n, size_t m, double norma);
// the same for one packed-upper n x n block held in HOST memory
bool cholesky_decomp_U (mat a, vec d, int n, double norma);
// R^T y = b
bool compute_y (double **&rows_p, vec b, vec y, size_t n, size_t m, double norma);
// D R x = y
bool compute_x (double **&rows_p, vec x, vec y, vec d, size_t n, size_t m, double norma, int root);

// ------------------------------------------------------------------ block kernels (HOST blocks)
// A(n x n) -= R1^T diag(D) R2
void minus_RTDR_o1 (mat A, mat R1, vec D, mat R2, int n);
// packed-upper A -= upper(R1^T diag(D) R2)   (the callers pass R2 == R1)
void minus_RTDRu_ofast (mat A, mat R1, vec D, mat R2, int n);
// A(n x l) -= R1(n x n)^T diag(D) R2(n x l)
void minus_RTDR_l_o1 (mat A, mat R1, vec D, mat R2, int n, int l);
// packed-upper A(l x l) -= upper(R1(n x l)^T diag(D) R1)
void minus_RTDRu_l (mat A, mat R1, vec D, int n, int l);
// B = A^{-1} for packed-upper n x n
bool reverse_upper (mat A, mat B, int n, double norma);
// A <- D R^T A with R the packed-upper INVERSE factor (n x n resp. n x l block)
void DRtA (mat R, mat A, vec D, int n);
void DRtA_l (mat R, mat A, vec D, int n, int l);

// ------------------------------------------------------------------ printing and norms
void print_mat_beauty (int root, size_t n, size_t m, double **&rows_p, size_t len = -1, bool log = 0,
                       int width = 8, int precision = 2);
void print_vec (vec a, size_t n, size_t len = -1UL);
double norma_vec2 (double *r, int n);
double find_error (double *x, int n);

#endif
