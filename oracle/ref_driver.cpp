// ref_driver.cpp — thin harness around the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with
// /root/reference/src/{init,matrix_op,solver}.cpp (never copied into this repo) and
// the minimpi shim, into oracle/_ref/.  Two artefacts come out of this file:
//
//   * `ref_dump n m [file] out.bin`  (MS_REF_DUMP_MAIN defined): runs the reference's
//     pipeline in the order of test/test_mat.cpp:69-165 by calling the reference's own
//     functions, and dumps every intermediate in full precision (raw float64) so that
//     oracle/ms_oracle.c and the CUDA path can be compared against it.  One problem
//     size per process because the reference's gather_row keeps function-static
//     buffers sized by its first call (src/init.cpp:313-329).
//   * `ref_time n m`  (MS_REF_TIME_MAIN defined, the reference's Release flags): the TIMING harness of
//     bench.py's CPU baseline.  Same call order, any number of minimpi ranks; wall-clock around
//     cholesky_decomp (what the reference's own `elapsed` covers, test/test_mat.cpp:118-128) and
//     around compute_y + compute_x (which the reference never times); prints one JSON line.
//   * `libref_blocks.so`: extern "C" trampolines onto the reference's stateless block
//     kernels (src/matrix_op.h:6-37, src/solver.h:14-15) for ctypes.
//
// The only arithmetic defined here is the generator f1, restated from
// test/test_mat.cpp:13-17 (it lives in the file that owns main(), so it cannot be
// linked): A_ij = n - max(i, j) evaluated in size_t, then converted to double.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "init.h"
#include "matrix_op.h"
#include "solver.h"

static double
formula_n_minus_max (size_t i, size_t j, size_t n)
{
  return n - (i > j ? i : j);
}

extern "C" {

int
ref_diag_factor (double *a, double *d, int n, double norma)
{
  return cholesky_decomp_U (a, d, n, norma);
}

int
ref_upper_inverse (double *A, double *B, int n, double norma)
{
  return reverse_upper (A, B, n, norma);
}

// l == n selects the square routine, otherwise the rectangular "_l" twin, exactly as
// cholesky_decomp chooses between them (src/solver.cpp:60-85).
void
ref_panel (double *R, double *A, double *D, int n, int l)
{
  if (l == n)
    DRtA (R, A, D, n);
  else
    DRtA_l (R, A, D, n, l);
}

void
ref_update (double *A, double *R1, double *D, double *R2, int n, int l)
{
  if (l == n)
    minus_RTDR_o1 (A, R1, D, R2, n);
  else
    minus_RTDR_l_o1 (A, R1, D, R2, n, l);
}

void
ref_update_diag (double *A, double *R1, double *D, int n, int l)
{
  if (l == n)
    minus_RTDRu_ofast (A, R1, D, R1, n);
  else
    minus_RTDRu_l (A, R1, D, n, l);
}

double
ref_norm2 (double *r, int n)
{
  return norma_vec2 (r, n);
}

double
ref_error (double *x, int n)
{
  return find_error (x, n);
}

size_t
ref_packed_index (size_t i, size_t j, size_t n)
{
  return get_elU (i, j, n);
}
}

#ifdef MS_REF_DUMP_MAIN
static void
put (FILE *f, const double *p, size_t cnt)
{
  if (fwrite (p, sizeof (double), cnt, f) != cnt)
    {
      perror ("ref_dump: fwrite");
      exit (2);
    }
}

// Layout of out.bin (all float64):
//   [0] n  [1] m  [2] alloc_size  [3] init_ok  [4] norm  [5] factor_fail
//   [6] forward_fail  [7] backward_fail  [8] residual_2norm  [9] error_2norm
//   then b[n], d[n], y[n], x[n], r[n], then A (upper triangle by rows, n(n+1)/2),
//   then R (upper triangle by rows, n(n+1)/2).  Sections after a failure are zero.
int
main (int argc, char **argv)
{
  MPI_Init (&argc, &argv);
  if (argc < 4)
    {
      fprintf (stderr, "usage: %s n m [file] out.bin\n", argv[0]);
      return 2;
    }
  size_t n = strtoul (argv[1], nullptr, 10);
  size_t m = strtoul (argv[2], nullptr, 10);
  const char *file = argc >= 5 ? argv[3] : "";
  const char *out = argc >= 5 ? argv[4] : argv[3];

  size_t cols = n / m + (n % m > 0);
  size_t alloc = compute_alloc_size (n, m);
  std::vector<double *> rows_v (cols);
  double **rows_p = rows_v.data ();
  std::unique_ptr<double[]> body (alloc_rows (n, m, alloc, rows_p));

  std::vector<double> head (10, 0.0), b (n, 0.0), d (n, 0.0), y (n, 0.0),
      x (n, 0.0), r (n, 0.0), row (n, 0.0);
  std::vector<double> Atri (n * (n + 1) / 2, 0.0), Rtri (n * (n + 1) / 2, 0.0);
  head[0] = (double)n;
  head[1] = (double)m;
  head[2] = (double)alloc;

  auto fill = [&] () -> bool {
    if (file[0])
      return init_mat_file (rows_p, n, m, file);
    init_mat (rows_p, n, m, formula_n_minus_max);
    return true;
  };
  auto snapshot = [&] (std::vector<double> &tri) {
    size_t pos = 0;
    for (size_t i = 0; i < n; i++)
      {
        gather_row (i, n, m, 0, row.data (), rows_p);
        for (size_t j = i; j < n; j++)
          tri[pos++] = row[j];
      }
  };

  bool ok = fill ();
  head[3] = ok;
  if (ok)
    {
      snapshot (Atri);
      double norm = 0;
      init_b_get_norm (b.data (), n, m, 0, rows_p, norm);
      head[4] = norm;
      MPI_Bcast (b.data (), (int)n, MPI_DOUBLE, 0, MPI_COMM_WORLD); // test_mat.cpp:116
      bool fail = cholesky_decomp (rows_p, d.data (), n, m, norm);
      head[5] = fail;
      if (!fail)
        {
          snapshot (Rtri);
          bool f2 = compute_y (rows_p, b.data (), y.data (), n, m, norm);
          head[6] = f2;
          bool f3 = f2
                    || compute_x (rows_p, x.data (), y.data (), d.data (), n, m,
                                  norm, 0);
          head[7] = f3;
          if (!f3 && fill ())
            {
              find_disrep_vec (b.data (), x.data (), r.data (), n, m, 0, rows_p);
              head[8] = norma_vec2 (r.data (), (int)n);
              head[9] = find_error (x.data (), (int)n);
            }
        }
    }
  int rank = 0;
  MPI_Comm_rank (MPI_COMM_WORLD, &rank);
  if (rank != 0)
    {
      MPI_Finalize ();
      return 0;
    }
  FILE *f = fopen (out, "wb");
  if (!f)
    {
      perror ("ref_dump: fopen");
      return 2;
    }
  put (f, head.data (), head.size ());
  put (f, b.data (), n);
  put (f, d.data (), n);
  put (f, y.data (), n);
  put (f, x.data (), n);
  put (f, r.data (), n);
  put (f, Atri.data (), Atri.size ());
  put (f, Rtri.data (), Rtri.size ());
  fclose (f);
  MPI_Finalize ();
  return 0;
}
#endif

#ifdef MS_REF_TIME_MAIN
#include <sys/time.h>
static double
now_s ()
{
  struct timeval tv;
  gettimeofday (&tv, nullptr);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}

int
main (int argc, char **argv)
{
  MPI_Init (&argc, &argv);
  if (argc < 3)
    {
      fprintf (stderr, "usage: %s n m\n", argv[0]);
      return 2;
    }
  int rank = 0, size = 1;
  MPI_Comm_rank (MPI_COMM_WORLD, &rank);
  MPI_Comm_size (MPI_COMM_WORLD, &size);
  size_t n = strtoul (argv[1], nullptr, 10);
  size_t m = strtoul (argv[2], nullptr, 10);
  size_t cols = n / m + (n % m > 0);
  size_t alloc = compute_alloc_size (n, m);
  std::vector<double *> rows_v (cols);
  double **rows_p = rows_v.data ();
  std::unique_ptr<double[]> body (alloc_rows (n, m, alloc, rows_p));
  std::vector<double> b (n, 0.0), d (n, 0.0), y (n, 0.0), x (n, 0.0), r (n, 0.0);
  init_mat (rows_p, n, m, formula_n_minus_max);
  double norm = 0;
  init_b_get_norm (b.data (), n, m, 0, rows_p, norm);
  MPI_Bcast (b.data (), (int)n, MPI_DOUBLE, 0, MPI_COMM_WORLD);
  MPI_Barrier (MPI_COMM_WORLD);
  double t0 = now_s ();
  bool fail = cholesky_decomp (rows_p, d.data (), n, m, norm);
  MPI_Barrier (MPI_COMM_WORLD);
  double t1 = now_s ();
  bool f2 = fail || compute_y (rows_p, b.data (), y.data (), n, m, norm);
  bool f3 = f2 || compute_x (rows_p, x.data (), y.data (), d.data (), n, m, norm, 0);
  MPI_Barrier (MPI_COMM_WORLD);
  double t2 = now_s ();
  double resid = -1, err = -1;
  if (!f3)
    {
      init_mat (rows_p, n, m, formula_n_minus_max);
      find_disrep_vec (b.data (), x.data (), r.data (), n, m, 0, rows_p);
      resid = norma_vec2 (r.data (), (int)n);
      err = find_error (x.data (), (int)n);
    }
  if (rank == 0)
    printf ("{\"n\": %zu, \"m\": %zu, \"ranks\": %d, \"factor_s\": %.6f, \"solve_s\": %.6f, \"failed\": %d, "
            "\"residual\": %.6e, \"error\": %.6e}\n",
            n, m, size, t1 - t0, t2 - t1, (int)f3, resid, err);
  gather_row (0, 0, 0, 0, nullptr, nullptr, 1);
  fflush (stdout);
  MPI_Finalize ();
  return 0;
}
#endif
