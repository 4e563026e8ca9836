// Drives the C++ mirror of the reference API (matrix_solver_b200/host/) the way a program written against
// src/*.h would, on a GPU box:
//   host_api_test chol_u <w> <norma> <in.bin> <out.bin>   cholesky_decomp_U (src/solver.cpp:106-160): packed a in,
//                                                          [fail, d[w], a packed] out
//   host_api_test rev_u  <w> <norma> <in.bin> <out.bin>   reverse_upper (src/matrix_op.cpp:244-285): [fail, inverse packed]
//   host_api_test init_custom <n> <m> <out.bin>           init_mat with a caller-defined initializer that equals f1
//                                                          everywhere but in ONE entry: rows 0..n-1 gathered -> out
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "init.h"
#include "matrix_op.h"
#include "ms_rt.h"
#include "solver.h"

static std::vector<double>
slurp (const char *path, size_t count)
{
  std::vector<double> v (count);
  FILE *f = fopen (path, "rb");
  if (!f || fread (v.data (), sizeof (double), count, f) != count)
    {
      fprintf (stderr, "cannot read %zu doubles from %s\n", count, path);
      exit (2);
    }
  fclose (f);
  return v;
}

static void
spill (const char *path, const std::vector<double> &v)
{
  FILE *f = fopen (path, "wb");
  if (!f || fwrite (v.data (), sizeof (double), v.size (), f) != v.size ())
    {
      fprintf (stderr, "cannot write %s\n", path);
      exit (2);
    }
  fclose (f);
}

// f1 with entry (3, 5) moved by 0.5: agrees with f1 on every probe a recogniser would reasonably try
static double
almost_f1 (size_t i, size_t j, size_t n)
{
  double v = (double)(n - (i > j ? i : j));
  return (i == 3 && j == 5) ? v + 0.5 : v;
}

int
main (int argc, char **argv)
{
  if (argc >= 6 && (!strcmp (argv[1], "chol_u") || !strcmp (argv[1], "rev_u")))
    {
      const int w = atoi (argv[2]);
      const double norma = atof (argv[3]);
      const size_t cnt = (size_t)w * (w + 1) / 2;
      std::vector<double> a = slurp (argv[4], cnt), out;
      if (!strcmp (argv[1], "chol_u"))
        {
          std::vector<double> d (w);
          bool fail = cholesky_decomp_U (a.data (), d.data (), w, norma);
          out.push_back (fail);
          out.insert (out.end (), d.begin (), d.end ());
          out.insert (out.end (), a.begin (), a.end ());
        }
      else
        {
          std::vector<double> b (cnt);
          bool fail = reverse_upper (a.data (), b.data (), w, norma);
          out.push_back (fail);
          out.insert (out.end (), b.begin (), b.end ());
        }
      spill (argv[5], out);
      return 0;
    }
  if (argc >= 5 && !strcmp (argv[1], "init_custom"))
    {
      const size_t n = strtoul (argv[2], nullptr, 10), m = strtoul (argv[3], nullptr, 10);
      ms_rt_init ();
      auto rows_holder = std::make_unique<double *[]> (n / m + (n % m > 0));
      double **rows_p = rows_holder.get ();
      std::unique_ptr<double[]> body (alloc_rows (n, m, compute_alloc_size (n, m), rows_p));
      init_mat (rows_p, n, m, almost_f1);
      std::vector<double> out (n * n, 0.0), row (n);
      for (size_t i = 0; i < n; i++)
        {
          gather_row (i, n, m, 0, row.data (), rows_p);
          for (size_t j = i; j < n; j++)
            out[i * n + j] = row[j];
        }
      spill (argv[4], out);
      ms_rt_finalize ();
      return 0;
    }
  fprintf (stderr, "usage: see the head of tests/host_api_test.cpp\n");
  return 2;
}
