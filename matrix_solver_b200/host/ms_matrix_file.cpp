// ms_matrix_file.cpp — see ms_matrix_file.h.
#include "ms_matrix_file.h"

#include <cctype>
#include <charconv>
#include <cstdlib>
#include <cstring>

#include <sys/types.h>

namespace
{
constexpr size_t BUF_BYTES = (size_t)8 << 20;
constexpr size_t MAX_TOKEN = 512;  // longest run of characters one number may span

bool
ends_with (const char *s, const char *suffix)
{
  size_t a = strlen (s), b = strlen (suffix);
  return a >= b && strcmp (s + a - b, suffix) == 0;
}
}  // namespace

bool
MatrixFileReader::open (const char *path, size_t n)
{
  close ();
  n_ = n;
  binary_ = ends_with (path, ".f64");
  fp_ = fopen (path, binary_ ? "rb" : "r");
  if (!fp_)
    return false;
  if (!binary_)
    {
      buf_.resize (BUF_BYTES + 1);
      pos_ = end_ = 0;
      eof_ = false;
    }
  return true;
}

void
MatrixFileReader::close ()
{
  if (fp_)
    fclose (fp_);
  fp_ = nullptr;
}

// move the unread tail to the front and top the buffer up; false when nothing new could be read
bool
MatrixFileReader::refill ()
{
  if (eof_)
    return false;
  size_t tail = end_ - pos_;
  memmove (buf_.data (), buf_.data () + pos_, tail);
  pos_ = 0;
  end_ = tail;
  size_t got = fread (buf_.data () + end_, 1, BUF_BYTES - end_, fp_);
  end_ += got;
  buf_[end_] = '\0';
  if (got == 0)
    eof_ = true;
  return got > 0;
}

bool
MatrixFileReader::next_value (double *out)
{
  for (;;)
    {
      while (pos_ < end_ && isspace ((unsigned char)buf_[pos_]))
        pos_++;
      // make sure a whole token is in the buffer: refill when we are close to its end
      if (end_ - pos_ < MAX_TOKEN && !eof_)
        {
          refill ();
          continue;
        }
      if (pos_ >= end_)
        return false;
      // common case: plain decimal / scientific notation, parsed by the (correctly rounded)
      // from_chars; it does not take what "%lf" also accepts — a leading '+', hex floats, values
      // that overflow to inf or underflow to a denormal/zero — so anything it refuses goes to strtod
      const char *first = buf_.data () + pos_, *last = buf_.data () + end_;
      double v;
      auto res = std::from_chars (first, last, v);
      const char *stop = res.ptr;
      bool hex = first + 2 < last && (first[0] == '-' ? first[1] == '0' && (first[2] | 32) == 'x'
                                                      : first[0] == '0' && (first[1] | 32) == 'x');
      if (res.ec != std::errc () || hex)
        {
          char *e = nullptr;
          v = strtod (first, &e);
          stop = e;
        }
      if (stop == first)
        return false;  // not a number: "%lf" would fail here too
      pos_ = (size_t)(stop - buf_.data ());
      *out = v;
      return true;
    }
}

size_t
MatrixFileReader::read_rows (double *rows, size_t max_rows)
{
  if (!fp_)
    return 0;
  if (binary_)
    {
      size_t got = fread (rows, sizeof (double), max_rows * n_, fp_);
      return got / n_;
    }
  for (size_t r = 0; r < max_rows; r++)
    for (size_t j = 0; j < n_; j++)
      if (!next_value (&rows[r * n_ + j]))
        return r;
  return max_rows;
}

size_t
MatrixFileReader::skip_rows (size_t count)
{
  if (!fp_ || count == 0)
    return 0;
  if (binary_)
    {
      // how many complete rows are left decides the answer; then one seek
      const off_t here = ftello (fp_);
      if (here < 0 || fseeko (fp_, 0, SEEK_END) != 0)
        return 0;
      const off_t end = ftello (fp_);
      const size_t row_bytes = n_ * sizeof (double);
      const size_t left = (size_t)(end - here) / row_bytes;
      const size_t skipped = count < left ? count : left;
      fseeko (fp_, here + (off_t)(skipped * row_bytes), SEEK_SET);
      return skipped;
    }
  double ignored;
  for (size_t r = 0; r < count; r++)
    for (size_t j = 0; j < n_; j++)
      if (!next_value (&ignored))
        return r;
  return count;
}
