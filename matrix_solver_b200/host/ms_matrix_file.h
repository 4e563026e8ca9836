// ms_matrix_file.h — streaming reader for the matrix files of `test_mat n m <file>`.
//
// The reference parses its input with one fscanf("%lf") call per entry (src/init.cpp:100-112),
// which is the bottleneck for anything but toy files.  This reader keeps the
// text format — whitespace separated decimals, n rows of n values, same acceptance rules as "%lf"
// (strtod) — but parses out of a large refillable buffer, and adds a raw format for large inputs:
// a file whose name ends in ".f64" holds n*n little-endian float64 values, row-major.
// Rows are delivered in chunks so that the caller can stage them through pinned memory.
#ifndef MS_MATRIX_FILE_H
#define MS_MATRIX_FILE_H
#include <cstddef>
#include <cstdio>
#include <string>
#include <vector>

class MatrixFileReader
{
public:
  ~MatrixFileReader () { close (); }
  // false when the file cannot be opened
  bool open (const char *path, size_t n);
  // reads up to max_rows complete rows of n doubles into `rows`; returns how many were read.
  // Fewer than asked means the file ended (or held something that is not a number).
  size_t read_rows (double *rows, size_t max_rows);
  // moves past up to `count` rows without delivering them (a seek for the raw format, a parse for
  // text); returns how many complete rows were skipped, with the same meaning as read_rows
  size_t skip_rows (size_t count);
  void close ();
  bool binary () const { return binary_; }

private:
  bool next_value (double *out);
  bool refill ();

  FILE *fp_ = nullptr;
  size_t n_ = 0;
  bool binary_ = false;
  std::vector<char> buf_;
  size_t pos_ = 0, end_ = 0;
  bool eof_ = false;
};
#endif
