// init.h — storage, initialisation and distribution helpers of the drop-in API.
// Same public names and argument meaning as the reference's src/init.h:10-76; the bodies drive the
// GPU library through its C ABI (include/ms_b200.h).  `rows_p` keeps its shape (one entry per block
// column, nullptr for columns of other ranks) but the entries are opaque: the matrix itself lives on
// the device from alloc_rows() until the process ends.
#ifndef INIT_H
#define INIT_H

#include <cassert>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#include "defines.h"

using initializer = double (*) (size_t i, size_t j, size_t);

// the two generators the reference's driver defines (test/test_mat.cpp:13-22); init_mat runs them
// on the device when it is handed one of these (or any function that agrees with them)
double f1 (size_t i, size_t j, size_t n);
double hilb (size_t i, size_t j, size_t n);

size_t compute_alloc_size (size_t n, size_t m);
double *alloc_rows (size_t n, size_t m, size_t n_to_alloc, double **rows_p);
void init_mat (double **rows_p, size_t n, size_t m, initializer f);
bool init_mat_file (double **rows_p, size_t n, size_t m, const char *filename);  // 1 = success
bool check_args (const int argc, char **argv);

void gather_row (size_t i, size_t n, size_t m, int root, vec buf, double **rows_p, int action = 0);
void gather_col (size_t j, size_t n, size_t m, int root, vec buf, double **&rows_p);
void init_b_get_norm (vec b, size_t n, size_t m, int root, double **&rows_p, double &norm);
void find_disrep_vec (vec b, vec x, vec r, size_t n, size_t m, int root, double **&rows_p);

bool check_res (int root, int loc_res);
void printf_root (int root, const char *format, ...);

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
#endif
