// Test driver for host/ms_matrix_file.cpp (no GPU, no library): reads `path` as an n x n matrix,
// either with MatrixFileReader ("fast", chunks of `chunk` rows) or with the reference's way, one
// fscanf("%lf") per entry (src/init.cpp:100-112; "scanf"), and prints the number of complete rows
// followed by the bit patterns of their values.
#include <cinttypes>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ms_matrix_file.h"

int
main (int argc, char **argv)
{
  if (argc < 5)
    return 2;
  const char *path = argv[1];
  size_t n = strtoull (argv[2], nullptr, 10);
  bool fast = strcmp (argv[3], "fast") == 0;
  if (strcmp (argv[3], "alternate") == 0)
    {
      // read a chunk, skip a chunk, ... : prints "<first row> <rows>" per delivered chunk, then values
      size_t chunk = strtoull (argv[4], nullptr, 10);
      MatrixFileReader r;
      if (!r.open (path, n))
        {
          printf ("cant open\n");
          return 0;
        }
      std::vector<double> buf (chunk * n);
      bool take = true;
      for (size_t row = 0; row < n; take = !take)
        {
          size_t want = chunk < n - row ? chunk : n - row;
          size_t got = take ? r.read_rows (buf.data (), want) : r.skip_rows (want);
          printf ("%s %zu %zu\n", take ? "read" : "skip", row, got);
          if (take)
            for (size_t i = 0; i < got * n; i++)
              {
                uint64_t u;
                memcpy (&u, &buf[i], 8);
                printf ("%016" PRIx64 "\n", u);
              }
          row += got;
          if (got < want)
            break;
        }
      return 0;
    }
  size_t chunk = strtoull (argv[4], nullptr, 10);
  std::vector<double> a (n * n, 0.0);
  size_t rows = 0;
  if (fast)
    {
      MatrixFileReader r;
      if (!r.open (path, n))
        {
          printf ("cant open\n");
          return 0;
        }
      while (rows < n)
        {
          size_t want = chunk < n - rows ? chunk : n - rows;
          size_t got = r.read_rows (a.data () + rows * n, want);
          rows += got;
          if (got < want)
            break;
        }
    }
  else
    {
      FILE *fp = fopen (path, "r");
      if (!fp)
        {
          printf ("cant open\n");
          return 0;
        }
      bool ok = true;
      for (; rows < n && ok;)
        {
          for (size_t j = 0; j < n && ok; j++)
            ok = fscanf (fp, "%lf", &a[rows * n + j]) == 1;
          if (ok)
            rows++;
        }
      fclose (fp);
    }
  printf ("%zu\n", rows);
  for (size_t i = 0; i < rows * n; i++)
    {
      uint64_t u;
      memcpy (&u, &a[i], 8);
      printf ("%016" PRIx64 "\n", u);
    }
  return 0;
}
