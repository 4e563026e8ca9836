// test_mat.cpp — the CLI of the drop-in:  test_mat n m [file]
//
// Same positional arguments, same stdout lines and formats, same `log.txt` side effect and the
// same "always exit 0" behaviour as the reference's only executable (test/test_mat.cpp:24-174,
// SURVEY.md Appendix C).  The rank count that `mpirun -np P` used to supply comes from the
// environment instead: MS_NGPU=P runs P ranks, one per GPU (default 1).  MS_VERBOSE=1 adds one
// line on stderr with the relative residual, GFLOP/s and the solve times.
#include <sys/time.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "init.h"
#include "matrix_op.h"
#include "ms_rt.h"
#include "solver.h"

namespace
{
const int root = 0;

double
now_seconds ()
{
  timeval t;
  gettimeofday (&t, nullptr);
  return (double)t.tv_sec + 1e-6 * (double)t.tv_usec;
}

struct problem_args
{
  size_t n, m;
  const char *file;
  bool from_file () const { return file[0] != '\0'; }
};

bool
fill_matrix (double **rows_p, const problem_args &a)
{
  if (a.from_file ())
    return init_mat_file (rows_p, a.n, a.m, a.file);
  // the reference chooses between its two generators at compile time (test/test_mat.cpp:13-22, :98);
  // here MS_GENERATOR=hilb selects the second one at run time
  const char *gen = getenv ("MS_GENERATOR");
  init_mat (rows_p, a.n, a.m, gen && strcmp (gen, "hilb") == 0 ? hilb : f1);
  return true;
}

int
run (int argc, char **argv)
{
  if (!check_args (argc, argv))
    {
      printf_root (root, "Usage: %s n m <filename>\n", argv[0]);
      return 0;
    }
  problem_args a{(size_t)atoi (argv[1]), (size_t)atoi (argv[2]), argc == 4 ? argv[3] : ""};
  const size_t n = a.n, m = a.m;
  printf_root (root, "n = %lu, m = %lu, file = %s\n", n, m, a.file);
  if (ms_rt_rank () == root)
    if (FILE *fp = fopen ("log.txt", "a"))
      {
        fprintf (fp, "n=%lu, m=%lu, file=%s\n", n, m, a.file);
        fclose (fp);
      }

  const size_t columns = n / m + (n % m > 0);
  auto rows_holder = std::make_unique<double *[]> (columns);
  double **rows_p = rows_holder.get ();
  const size_t my_elems = compute_alloc_size (n, m);
  std::unique_ptr<double[]> body (alloc_rows (n, m, my_elems, rows_p));
  // the reference's allocation self-check (test/test_mat.cpp:78-88): the ranks' shares of the packed
  // upper triangle must add up to n(n+1)/2
  const size_t all_elems = (size_t)ms_rt_sum (my_elems);
  if (ms_rt_rank () == root)
    assert_and_do (all_elems == n * (n + 1) / 2,
                   [=] { printf_root (root, "Allocated: %lu, wanted: %lu\n", all_elems, n * (n + 1) / 2); });

  if (!fill_matrix (rows_p, a))
    {
      printf_root (root, "error reading file\n");
      return 0;
    }
  printf_root (root, "print given matrix\n");
  print_mat_beauty (root, n, m, rows_p, 7);

  auto work = std::make_unique<double[]> (5 * n);  // d | b | y | x | r
  double *d = work.get (), *b = d + n, *y = b + n, *x = y + n, *r = x + n;

  double norm = 0;
  init_b_get_norm (b, n, m, root, rows_p, norm);

  const double t1 = now_seconds ();
  bool failed = cholesky_decomp (rows_p, d, n, m, norm);
  const double elapsed = now_seconds () - t1;
  if (check_res (root, failed))
    return 0;

  printf_root (root, "R: \n");
  print_mat_beauty (root, n, m, rows_p, 5, 1);
  printf_root (root, "compute_y\n");
  const double t2 = now_seconds ();
  if (check_res (root, compute_y (rows_p, b, y, n, m, norm)))
    return 0;
  printf_root (root, "compute_x\n");
  if (check_res (root, compute_x (rows_p, x, y, d, n, m, norm, root)))
    return 0;
  const double solve_s = now_seconds () - t2;
  if (ms_rt_rank () == root)
    {
      printf ("answer vec x: \n");
      print_vec (x, n, 7);
    }

  printf_root (root, "reinit mat\n");
  if (!fill_matrix (rows_p, a))
    {
      printf_root (root, "error reading file\n");
      return 0;
    }
  find_disrep_vec (b, x, r, n, m, root, rows_p);
  const double resid = norma_vec2 (r, (int)n), error = find_error (x, (int)n);
  if (ms_rt_rank () == root)
    {
      printf ("n = %lu, m = %lu, file = %s, residual = %e, er = %e, elapsed = %lf sec\n", n, m, a.file,
              resid, error, elapsed);
      if (getenv ("MS_VERBOSE"))
        fprintf (stderr,
                 "b200: gpus = %d, rel_residual = %e, factor = %.3f GFLOP/s (n^3/3), solve = %lf sec\n",
                 ms_rt_size (), resid / norma_vec2 (b, (int)n), (double)n * n * n / 3. / elapsed / 1e9,
                 solve_s);
    }
  return 0;
}
}  // namespace

int
main (int argc, char **argv)
{
  ms_rt_init ();
  int rc = run (argc, argv);
  fflush (stdout);
  ms_rt_finalize ();
  return rc;
}
