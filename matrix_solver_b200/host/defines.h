// defines.h — host-side constants and index helpers of the drop-in API.
// Mirrors the public names of the reference's src/defines.h:7-40 (mat, vec, EPS, MAX, get_elU,
// is_double_too_small, sgn, assert_and_do) so that code written against the reference compiles
// unchanged against this directory.
#ifndef DEFINES_H
#define DEFINES_H
#include <cassert>
#include <cmath>
#include <cstddef>
#include <functional>

using mat = double *;
using vec = double *;

constexpr double EPS = 1.e-14;  // pivots are compared with EPS * ||A||_inf
constexpr int MAX = 8;

// Offset of (i, j), j >= i, in a row-major packed upper triangle of order n:
// rows 0..i-1 hold n, n-1, ... entries, i.e. i*n - i(i-1)/2 in total, then j - i.
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
#endif
