// ms_kernels_util.cu — HBM-bound helper kernels: formula generator (K8), symmetric
// row reductions for b / ||A||_inf / residual (K9, K10), layout conversion for host
// uploads (K11) and the two vector norms.
//
// Reference semantics restated by these kernels:
//   fill_formula   : init_mat + f1 / hilb        src/init.cpp:60-98, test/test_mat.cpp:13-22
//   sym_rows BNORM : init_b_get_norm             src/init.cpp:492-531
//   sym_rows SYMV  : find_disrep_vec             src/init.cpp:533-563
//   norms          : norma_vec2 / find_error     src/matrix_op.cpp:524-548
//   packed_copy    : the packed block-column layout, src/init.cpp:29-98, defines.h:12-18
#include "ms_common.cuh"

#define LAUNCH_COUNT(c) ((c)->stats.kernel_launches++)

// global row of local row lr (block-row cyclic: local block b is global block b*P+rank)
__device__ __forceinline__ size_t
global_row_of (size_t lr, size_t m, int rank, int nranks)
{
  size_t b = lr / m;
  return (b * (size_t)nranks + (size_t)rank) * m + (lr - b * m);
}

// ----------------------------------------------------------------------------- K8 generator
// One CTA row-slab: blockIdx.y walks local rows, blockIdx.x walks 2048-column segments that
// start at the row's diagonal block.  128-bit stores, 8 bytes written per stored element.
template <int KIND>
__global__ void __launch_bounds__ (256)
fill_formula_kernel (double *__restrict__ A, size_t ld, size_t n, size_t m, size_t rows_local,
                     int rank, int nranks)
{
  for (size_t lr = blockIdx.y; lr < rows_local; lr += gridDim.y)
    {
      size_t g = global_row_of (lr, m, rank, nranks);
      size_t bstart = (g / m) * m;        // first column of the diagonal block
      size_t j0 = (bstart & ~(size_t)1) + 2 * ((size_t)blockIdx.x * 1024 + threadIdx.x);
      double *row = A + lr * ld;
#pragma unroll
      for (int u = 0; u < 4; u++)
        {
          size_t j = j0 + (size_t)u * 512;
          if (j >= ld)
            break;
          double2 v;
          double t[2];
#pragma unroll
          for (int e = 0; e < 2; e++)
            {
              size_t jj = j + e;
              double val = 0.0;
              if (jj < n && jj >= bstart)
                {
                  size_t hi = jj > g ? jj : g;
                  if (KIND == MS_FORMULA_F1)
                    val = (double)(n - hi);                     // n - max(i,j) in integers
                  else
                    val = 1. / (double)(g + jj + 1ul);          // hilb
                }
              t[e] = val;
            }
          v.x = t[0];
          v.y = t[1];
          *reinterpret_cast<double2 *> (row + j) = v;
        }
    }
}

int
ms_k_fill_formula (ms_ctx *c, int kind, cudaStream_t s)
{
  if (c->rows_local == 0)
    return MS_OK;
  dim3 grid ((unsigned)((c->ld / 2 + 1023) / 1024), (unsigned)(c->rows_local < 16384 ? c->rows_local : 16384));
  if (kind == MS_FORMULA_F1)
    fill_formula_kernel<MS_FORMULA_F1><<<grid, 256, 0, s>>> (c->A, c->ld, c->n, c->m, c->rows_local, c->rank, c->nranks);
  else
    fill_formula_kernel<MS_FORMULA_HILBERT><<<grid, 256, 0, s>>> (c->A, c->ld, c->n, c->m, c->rows_local, c->rank, c->nranks);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// ----------------------------------------------------------------------------- K9 / K10
// out_i = sum_k w(k) * Asym(i,k) over the full symmetric row, from the stored upper triangle, in ONE pass:
// every stored element A(i,j), j > i, feeds row i (with w_j) and row j (with w_i); the diagonal feeds row i.
//   BNORM: w(k) = [k even], plus the absolute row sum (for ||A||_inf)    init_b_get_norm, src/init.cpp:492-531
//   SYMV : w(k) = x_k                                                     find_disrep_vec,  src/init.cpp:533-563
// (Round 1 read the triangle twice — a row pass and a column pass, 34.4 GB at n = 65536; the reference
// itself gathers every row and every column, :509-522.)
//
// One CTA owns 128 local rows (8 warps x 16 rows) and sweeps their columns from the diagonal to n in chunks
// of 256: a lane reads double2 at columns 64 k + 2 lane (k = 0..3), so a warp instruction covers 512
// contiguous bytes of one row.  Row sums stay in registers for the whole sweep (one shuffle reduction per
// row at the very end); column sums are lane-local over the warp's 16 rows, then the 8 warps are added in
// a fixed order through shared memory and the CTA writes ONE partial per column and row group.  A second
// small kernel adds the partials in group order.  No atomics: bit-reproducible.
constexpr int SYM_ROWS = 128;
constexpr int SYM_COLS = 256;

template <int MODE, int RB>
__global__ void __launch_bounds__ (256, 2)
sym_onepass_kernel (const double *__restrict__ A, size_t ld, size_t n, size_t m, size_t rows_local, int rank,
                    int nranks, const double *__restrict__ x, double *__restrict__ rowS, double *__restrict__ rowA,
                    double *__restrict__ colS, double *__restrict__ colA)
{
  constexpr bool NORM = MODE == MS_SYM_BNORM;
  __shared__ size_t grow[SYM_ROWS];
  __shared__ double wrow[SYM_ROWS];
  // dynamic shared memory: cpart[8][256] | rs[128][32] | (BNORM) cpartA[8][256] | rsA[128][32].
  // The per-lane row accumulators rs live in shared memory (lane-private slots: bank = lane, no conflicts,
  // no synchronisation): in registers they cost 64 of them and pushed the kernel to one CTA per SM.
  extern __shared__ double sym_sm[];
  double (*cpart)[SYM_COLS] = reinterpret_cast<double (*)[SYM_COLS]> (sym_sm);
  double (*rs)[32] = reinterpret_cast<double (*)[32]> (sym_sm + 8 * SYM_COLS);
  double (*cpartA)[SYM_COLS] = reinterpret_cast<double (*)[SYM_COLS]> (sym_sm + 8 * SYM_COLS + SYM_ROWS * 32);
  double (*rsA)[32] = reinterpret_cast<double (*)[32]> (sym_sm + 16 * SYM_COLS + SYM_ROWS * 32);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t lr0 = (size_t)blockIdx.x * SYM_ROWS;
  if (tid < SYM_ROWS)
    {
      const size_t lr = lr0 + tid;
      const size_t g = lr < rows_local ? global_row_of (lr, m, rank, nranks) : n;
      grow[tid] = g < n ? g : n;
      wrow[tid] = g < n ? (NORM ? ((g & 1) == 0 ? 1.0 : 0.0) : x[g]) : 0.0;
    }
  __syncthreads ();
  const size_t c_first = (grow[0] / SYM_COLS) * SYM_COLS;  // global rows ascend inside a group
#pragma unroll
  for (int r = 0; r < 16; r++)
    {
      rs[warp * 16 + r][lane] = 0.0;
      if (NORM)
        rsA[warp * 16 + r][lane] = 0.0;
    }

  for (size_t col0 = c_first; col0 < n; col0 += SYM_COLS)
    {
      // weights of this lane's 8 columns for the ROW sums
      double wc[4][2];
#pragma unroll
      for (int k = 0; k < 4; k++)
        {
          const size_t cj = col0 + 64 * k + 2 * lane;
          if (NORM)
            {
              wc[k][0] = 1.0;  // col0, 64 k and 2 lane are even
              wc[k][1] = 0.0;
            }
          else
            {
              wc[k][0] = cj < n ? x[cj] : 0.0;
              wc[k][1] = cj + 1 < n ? x[cj + 1] : 0.0;
            }
        }
      double cacc[4][2], caccA[NORM ? 4 : 1][2];
#pragma unroll
      for (int k = 0; k < 4; k++)
        {
          cacc[k][0] = cacc[k][1] = 0.0;
          if (NORM)
            caccA[k][0] = caccA[k][1] = 0.0;
        }
      // rows go in batches of RB: all loads of a batch are issued before its arithmetic (RB x 4 16-byte
      // loads in flight per lane; a warp-uniform `continue` per row would serialise the rows' latencies)
#pragma unroll
      for (int rb = 0; rb < 16; rb += RB)
        {
          double2 v[RB][4];
#pragma unroll
          for (int rr = 0; rr < RB; rr++)
            {
              const int row = warp * 16 + rb + rr;
              const size_t gi = grow[row];
              // no such row, or the whole chunk lies left of its diagonal: nothing to read
              const bool live = gi < n && col0 + SYM_COLS > gi;
              const double *rp = A + (lr0 + row) * ld + col0 + 2 * lane;
#pragma unroll
              for (int k = 0; k < 4; k++)
                v[rr][k] = (live && col0 + 64 * k + 2 * lane < ld) ? __ldcs (reinterpret_cast<const double2 *> (rp + 64 * k))
                                                                   : make_double2 (0.0, 0.0);
            }
#pragma unroll
          for (int rr = 0; rr < RB; rr++)
            {
              const int r = rb + rr;
              const int row = warp * 16 + r;
              const size_t gi = grow[row];
              const double wr = wrow[row];
              double ra = 0.0, raA = 0.0;
#pragma unroll
              for (int k = 0; k < 4; k++)
                {
                  const size_t cj = col0 + 64 * k + 2 * lane;
                  // stored and counted: gi <= column < n  (gi == n for rows that do not exist)
                  const double a0 = (cj >= gi && cj < n) ? v[rr][k].x : 0.0;
                  const double a1 = (cj + 1 >= gi && cj + 1 < n) ? v[rr][k].y : 0.0;
                  ra = fma (a0, wc[k][0], ra);
                  ra = fma (a1, wc[k][1], ra);
                  // the mirrored element A(column, gi) exists only strictly above the diagonal
                  const double b0 = cj > gi ? a0 : 0.0;
                  const double b1 = cj + 1 > gi ? a1 : 0.0;
                  cacc[k][0] = fma (b0, wr, cacc[k][0]);
                  cacc[k][1] = fma (b1, wr, cacc[k][1]);
                  if (NORM)
                    {
                      raA += fabs (a0) + fabs (a1);
                      caccA[k][0] += fabs (b0);
                      caccA[k][1] += fabs (b1);
                    }
                }
              rs[row][lane] += ra;
              if (NORM)
                rsA[row][lane] += raA;
              (void)r;
            }
        }
#pragma unroll
      for (int k = 0; k < 4; k++)
        {
          cpart[warp][64 * k + 2 * lane] = cacc[k][0];
          cpart[warp][64 * k + 2 * lane + 1] = cacc[k][1];
          if (NORM)
            {
              cpartA[warp][64 * k + 2 * lane] = caccA[k][0];
              cpartA[warp][64 * k + 2 * lane + 1] = caccA[k][1];
            }
        }
      __syncthreads ();
      if (col0 + tid < n)
        {
          double s = 0.0, sa = 0.0;
#pragma unroll
          for (int w = 0; w < 8; w++)
            {
              s += cpart[w][tid];
              if (NORM)
                sa += cpartA[w][tid];
            }
          colS[(size_t)blockIdx.x * n + col0 + tid] = s;
          if (NORM)
            colA[(size_t)blockIdx.x * n + col0 + tid] = sa;
        }
      __syncthreads ();
    }
#pragma unroll
  for (int r = 0; r < 16; r++)
    {
      double s = rs[warp * 16 + r][lane], sa = NORM ? rsA[warp * 16 + r][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        {
          s += __shfl_xor_sync (0xffffffffu, s, o);
          if (NORM)
            sa += __shfl_xor_sync (0xffffffffu, sa, o);
        }
      const size_t gi = grow[warp * 16 + r];
      if (lane == 0 && gi < n)
        {
          rowS[gi] = s;
          if (NORM)
            rowA[gi] = sa;
        }
    }
}

// total_j = rowS_j + sum over the row groups that reach column j of colS[group][j], in group order;
// BNORM additionally folds the abs sums.  A group's sweep starts at the 256-column chunk of its first row.
template <int MODE>
__global__ void __launch_bounds__ (256)
sym_finalize_kernel (size_t n, size_t m, size_t rows_local, int rank, int nranks, size_t ngroups,
                     const double *__restrict__ rowS, const double *__restrict__ rowA,
                     const double *__restrict__ colS, const double *__restrict__ colA, double *__restrict__ out,
                     double *__restrict__ outA)
{
  size_t j = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= n)
    return;
  double s = rowS[j], a = MODE == MS_SYM_BNORM ? rowA[j] : 0.0;
  for (size_t g = 0; g < ngroups; g++)
    {
      const size_t first = global_row_of (g * SYM_ROWS, m, rank, nranks);
      if ((first / SYM_COLS) * SYM_COLS > j)
        break;  // first rows ascend with the group index
      s += colS[g * n + j];
      if (MODE == MS_SYM_BNORM)
        a += colA[g * n + j];
    }
  out[j] = s;
  if (MODE == MS_SYM_BNORM)
    outA[j] = a;
}

__global__ void __launch_bounds__ (1024)
max_reduce_kernel (const double *__restrict__ v, size_t n, double *__restrict__ out)
{
  __shared__ double sm[32];
  double best = 0;
  for (size_t i = threadIdx.x; i < n; i += 1024)
    best = fmax (best, v[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    best = fmax (best, __shfl_xor_sync (0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0)
    sm[threadIdx.x >> 5] = best;
  __syncthreads ();
  if (threadIdx.x < 32)
    {
      best = sm[threadIdx.x];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        best = fmax (best, __shfl_xor_sync (0xffffffffu, best, o));
      if (threadIdx.x == 0)
        *out = best;
    }
}

// scratch layout (doubles): rowS[n] rowA[n] outA[n] max[16] colS[ngroups*n] colA[ngroups*n]
int
ms_k_sym_rows (ms_ctx *c, int mode, const double *xvec, double *out, double *norm_out,
               cudaStream_t s)
{
  (void)norm_out;
  size_t n = c->n;
  size_t ngroups = (c->rows_local + SYM_ROWS - 1) / SYM_ROWS;
  size_t need = 3 * n + 16 + 2 * ngroups * n;
  if (need > c->scratch_doubles)
    {
      if (c->scratch)
        MS_CUDA (cudaFree (c->scratch));
      c->scratch = nullptr;
      MS_CUDA (cudaMalloc (&c->scratch, need * sizeof (double)));
      c->scratch_doubles = need;
    }
  double *rowS = c->scratch, *rowA = rowS + n, *outA = rowA + n, *colS = outA + n + 16, *colA = colS + ngroups * n;
  MS_CUDA (cudaMemsetAsync (rowS, 0, 2 * n * sizeof (double), s));
  unsigned gx = (unsigned)((n + 255) / 256);
  if (ngroups)
    {
      const size_t sm1 = (8 * SYM_COLS + SYM_ROWS * 32) * sizeof (double);
      static bool configured_dev[64] = {};
      if (!configured_dev[c->device & 63])
        {
          MS_CUDA (cudaFuncSetAttribute (sym_onepass_kernel<MS_SYM_BNORM, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sm1)));
          MS_CUDA (cudaFuncSetAttribute (sym_onepass_kernel<MS_SYM_SYMV, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
          configured_dev[c->device & 63] = true;
        }
      // four rows per load batch in both modes (BNORM: 5.7 vs 6.1 ms with two, measured at n = 65536)
      if (mode == MS_SYM_BNORM)
        sym_onepass_kernel<MS_SYM_BNORM, 4><<<(unsigned)ngroups, 256, 2 * sm1, s>>> (c->A, c->ld, n, c->m, c->rows_local, c->rank,
                                                                            c->nranks, xvec, rowS, rowA, colS, colA);
      else
        sym_onepass_kernel<MS_SYM_SYMV, 4><<<(unsigned)ngroups, 256, sm1, s>>> (c->A, c->ld, n, c->m, c->rows_local, c->rank,
                                                                        c->nranks, xvec, rowS, rowA, colS, colA);
      LAUNCH_COUNT (c);
    }
  if (mode == MS_SYM_BNORM)
    sym_finalize_kernel<MS_SYM_BNORM><<<gx, 256, 0, s>>> (n, c->m, c->rows_local, c->rank, c->nranks, ngroups, rowS, rowA,
                                                         colS, colA, out, outA);
  else
    sym_finalize_kernel<MS_SYM_SYMV><<<gx, 256, 0, s>>> (n, c->m, c->rows_local, c->rank, c->nranks, ngroups, rowS, rowA,
                                                        colS, colA, out, outA);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

int
ms_k_max (ms_ctx *c, const double *v, size_t n, double *d_out, cudaStream_t s)
{
  max_reduce_kernel<<<1, 1024, 0, s>>> (v, n, d_out);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

double *
ms_sym_abs_partial (ms_ctx *c)
{
  // location of outA (this rank's absolute row sums; n doubles followed by 16 spare ones for the max)
  // inside the scratch area of the last ms_k_sym_rows call
  return c->scratch + 2 * c->n;
}

// ----------------------------------------------------------------------------- norms
// out[0] = sum r_i^2, out[1] = sum (x_i - [i even])^2   (square roots taken by the caller)
__global__ void __launch_bounds__ (1024)
norms_kernel (const double *__restrict__ r, const double *__restrict__ x, size_t n,
              double *__restrict__ out)
{
  __shared__ double sm[2][32];
  double a = 0, b = 0;
  for (size_t i = threadIdx.x; i < n; i += 1024)
    {
      if (r)
        a += r[i] * r[i];
      if (x)
        {
          double t = x[i];
          if ((i & 1) == 0)
            t -= 1.;
          b += t * t;
        }
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    {
      a += __shfl_xor_sync (0xffffffffu, a, o);
      b += __shfl_xor_sync (0xffffffffu, b, o);
    }
  if ((threadIdx.x & 31) == 0)
    {
      sm[0][threadIdx.x >> 5] = a;
      sm[1][threadIdx.x >> 5] = b;
    }
  __syncthreads ();
  if (threadIdx.x < 32)
    {
      a = sm[0][threadIdx.x];
      b = sm[1][threadIdx.x];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        {
          a += __shfl_xor_sync (0xffffffffu, a, o);
          b += __shfl_xor_sync (0xffffffffu, b, o);
        }
      if (threadIdx.x == 0)
        {
          out[0] = a;
          out[1] = b;
        }
    }
}

int
ms_k_norms (ms_ctx *c, const double *r, const double *x, size_t n, double *d_out2, cudaStream_t s)
{
  norms_kernel<<<1, 1024, 0, s>>> (r, x, n, d_out2);
  if (c)
    LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// ----------------------------------------------------------------------------- K11 uploads
// d_rows holds nrows full host rows (ld_src apart) starting at global row row0; keep j >= start
// of the diagonal block for the rows this rank owns.
__global__ void __launch_bounds__ (256)
unpack_rows_kernel (double *__restrict__ A, size_t ld, size_t n, size_t m, int rank, int nranks,
                    const double *__restrict__ src, size_t row0, size_t nrows, size_t ld_src)
{
  size_t g = row0 + blockIdx.y;
  size_t I = g / m;
  if ((int)(I % (size_t)nranks) != rank)
    return;
  size_t lr = (I / (size_t)nranks) * m + (g - I * m);
  size_t j = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (j < n && j >= I * m)
    A[lr * ld + j] = src[(size_t)blockIdx.y * ld_src + j];
}

int
ms_k_unpack_rows (ms_ctx *c, const double *d_rows, size_t row0, size_t nrows, size_t ld_src,
                  cudaStream_t s)
{
  if (nrows == 0)
    return MS_OK;
  dim3 grid ((unsigned)((c->n + 255) / 256), (unsigned)nrows);
  unpack_rows_kernel<<<grid, 256, 0, s>>> (c->A, c->ld, c->n, c->m, c->rank, c->nranks, d_rows, row0,
                                          nrows, ld_src);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// Rows [i0, i1) of reference block column J (width w, first column J*m) staged at `seg`
// exactly as they lie in the reference's packed body:
//   i <  J*m : seg[(i - i0) * w + jj]                       (full rows of the column)
//   i >= J*m : packed upper triangle, row ii = i - J*m at ii*w - ii(ii-1)/2 ... (defines.h:12-18)
// `seg_base` is the packed offset of row i0 inside the block column.
__global__ void __launch_bounds__ (256)
packed_copy_kernel (double *__restrict__ A, size_t ld, size_t m, int rank, int nranks,
                    double *__restrict__ seg, size_t J, size_t w, size_t i0, size_t i1,
                    size_t seg_base, int to_packed)
{
  size_t i = i0 + blockIdx.y;
  if (i >= i1)
    return;
  size_t jj = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (jj >= w)
    return;
  size_t I = i / m;
  if ((int)(I % (size_t)nranks) != rank)
    return;
  size_t lr = (I / (size_t)nranks) * m + (i - I * m);
  size_t off;
  if (i < J * m)
    off = i * w + jj;
  else
    {
      size_t ii = i - J * m;
      if (jj < ii)
        return;
      off = J * m * w + ii * w + jj - (ii * ii + ii) / 2;
    }
  off -= seg_base;
  double *dev = A + lr * ld + J * m + jj;
  if (to_packed)
    seg[off] = *dev;
  else
    *dev = seg[off];
}

int
ms_k_packed_rows (ms_ctx *c, double *d_seg, size_t J, size_t w, size_t i0, size_t i1,
                  size_t seg_base, int to_packed, cudaStream_t s)
{
  if (i1 <= i0)
    return MS_OK;
  dim3 grid ((unsigned)((w + 255) / 256), (unsigned)(i1 - i0));
  packed_copy_kernel<<<grid, 256, 0, s>>> (c->A, c->ld, c->m_api, c->rank, c->nranks, d_seg, J, w, i0, i1,
                                          seg_base, to_packed);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// strided column gather for ms_get_col (gather_col, src/init.cpp:446-490)
__global__ void __launch_bounds__ (256)
get_col_kernel (const double *__restrict__ A, size_t ld, size_t m, int rank, int nranks, size_t j,
                double *__restrict__ out)
{
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i > j)
    return;
  size_t I = i / m;
  double v = 0;
  if ((int)(I % (size_t)nranks) == rank)
    v = A[((I / (size_t)nranks) * m + (i - I * m)) * ld + j];
  out[i] = v;
}

int
ms_k_get_col (ms_ctx *c, size_t j, double *d_out, cudaStream_t s)
{
  get_col_kernel<<<(unsigned)(j / 256 + 1), 256, 0, s>>> (c->A, c->ld, c->m, c->rank, c->nranks, j, d_out);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// ----------------------------------------------------------------------------- augmented column
// Forward solve fused into the factorisation (replaces the separate pass of compute_y,
// src/solver.cpp:162-208): b rides along as column n of the matrix.  The panel step turns it into
// z_I = D_I R_II^{-T} (b_I - sum_{K<I} R_KI^T D_K z_K), the trailing update subtracts R_IJ^T D_I z_I
// from b_J, so after the factorisation y = D z solves R^T y = b.
__global__ void __launch_bounds__ (256)
aug_scatter_kernel (double *__restrict__ A, size_t ld, size_t n, size_t m, size_t rows_local, int rank, int nranks,
                    const double *__restrict__ b)
{
  size_t lr = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (lr >= rows_local)
    return;
  A[lr * ld + n] = b[global_row_of (lr, m, rank, nranks)];
}

// y_g = d_g * z_g for the rows this rank owns (the other entries of y are left alone: the caller
// zeroes them and sums over the ranks)
__global__ void __launch_bounds__ (256)
aug_gather_kernel (const double *__restrict__ A, size_t ld, size_t n, size_t m, size_t rows_local, int rank,
                   int nranks, const double *__restrict__ d, double *__restrict__ y)
{
  size_t lr = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (lr >= rows_local)
    return;
  size_t g = global_row_of (lr, m, rank, nranks);
  y[g] = d[g] * A[lr * ld + n];
}

int
ms_k_aug_scatter (ms_ctx *c, const double *b, cudaStream_t s)
{
  if (c->rows_local == 0)
    return MS_OK;
  aug_scatter_kernel<<<(unsigned)((c->rows_local + 255) / 256), 256, 0, s>>> (c->A, c->ld, c->n, c->m, c->rows_local,
                                                                            c->rank, c->nranks, b);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

int
ms_k_aug_gather (ms_ctx *c, double *y, cudaStream_t s)
{
  if (c->rows_local == 0)
    return MS_OK;
  aug_gather_kernel<<<(unsigned)((c->rows_local + 255) / 256), 256, 0, s>>> (c->A, c->ld, c->n, c->m, c->rows_local,
                                                                           c->rank, c->nranks, c->vec[MS_VEC_D], y);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// ----------------------------------------------------------------------------- probes
// Register-only DMMA issue loop: the fp64 tensor-pipe ceiling bench.py quotes fractions of.
__global__ void __launch_bounds__ (256)
dmma_probe_kernel (double *out, int iters)
{
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++)
    c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++)
    {
#pragma unroll
      for (int i = 0; i < 16; i++)
        dmma_8x8x4 (c[i][0], c[i][1], a, b);
    }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++)
    s += c[i][0] + c[i][1];
  if (s == 123.456)
    out[0] = s;
}

extern "C" int
ms_probe_dmma_peak (int device, double *tflops)
{
  MS_CHECK_ARG (tflops != nullptr, "tflops is null");
  MS_CUDA (cudaSetDevice (device));
  cudaDeviceProp prop;
  MS_CUDA (cudaGetDeviceProperties (&prop, device));
  double *d;
  MS_CUDA (cudaMalloc (&d, 64));
  // Two CTAs of 8 warps per SM, all co-resident in one wave (74 registers): with four CTAs per SM only three fit
  // at a time and the hardware's placement of the fourth made the result bimodal (37.1 or 29.7 TFLOP/s from
  // process to process on the same box — round 2, profiles/README.md)
  int iters = 8192;
  int blocks = prop.multiProcessorCount * 2;
  cudaEvent_t e0, e1;
  MS_CUDA (cudaEventCreate (&e0));
  MS_CUDA (cudaEventCreate (&e1));
  // The clocks ramp up over the first tens of milliseconds of load (a probe of five 4 ms launches was seen
  // to return 29.7 TFLOP/s = 0.8 x 37.1 on a GPU that had just been idle): warm up for ~150 ms, then take
  // the best of 20 launches.
  dmma_probe_kernel<<<blocks, 256>>> (d, 64);
  MS_CUDA (cudaDeviceSynchronize ());
  for (int rep = 0; rep < 36; rep++)
    dmma_probe_kernel<<<blocks, 256>>> (d, iters);
  MS_CUDA (cudaDeviceSynchronize ());
  double best = 0;
  for (int rep = 0; rep < 20; rep++)
    {
      MS_CUDA (cudaEventRecord (e0));
      dmma_probe_kernel<<<blocks, 256>>> (d, iters);
      MS_CUDA (cudaEventRecord (e1));
      MS_CUDA (cudaEventSynchronize (e1));
      float ms;
      MS_CUDA (cudaEventElapsedTime (&ms, e0, e1));
      double flops = (double)blocks * 8 /*warps*/ * (double)iters * 16 * 512.0;
      double tf = flops / (ms * 1e-3) / 1e12;
      if (tf > best)
        best = tf;
    }
  *tflops = best;
  cudaEventDestroy (e0);
  cudaEventDestroy (e1);
  cudaFree (d);
  return MS_OK;
}

__global__ void __launch_bounds__ (256)
copy_probe_kernel (const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n2)
{
  size_t stride = (size_t)gridDim.x * 256;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n2; i += stride)
    dst[i] = src[i];
}

extern "C" int
ms_probe_hbm_copy (int device, size_t bytes, double *gbs)
{
  MS_CHECK_ARG (gbs != nullptr, "gbs is null");
  MS_CUDA (cudaSetDevice (device));
  cudaDeviceProp prop;
  MS_CUDA (cudaGetDeviceProperties (&prop, device));
  double2 *a, *b;
  size_t n2 = bytes / 16;
  MS_CUDA (cudaMalloc (&a, n2 * 16));
  MS_CUDA (cudaMalloc (&b, n2 * 16));
  MS_CUDA (cudaMemset (a, 0, n2 * 16));
  cudaEvent_t e0, e1;
  MS_CUDA (cudaEventCreate (&e0));
  MS_CUDA (cudaEventCreate (&e1));
  double best = 0;
  for (int rep = 0; rep < 6; rep++)
    {
      MS_CUDA (cudaEventRecord (e0));
      copy_probe_kernel<<<prop.multiProcessorCount * 16, 256>>> (a, b, n2);
      MS_CUDA (cudaEventRecord (e1));
      MS_CUDA (cudaEventSynchronize (e1));
      float ms;
      MS_CUDA (cudaEventElapsedTime (&ms, e0, e1));
      double g = 2.0 * (double)n2 * 16 / (ms * 1e-3) / 1e9;
      if (rep > 0 && g > best)
        best = g;
    }
  *gbs = best;
  cudaEventDestroy (e0);
  cudaEventDestroy (e1);
  cudaFree (a);
  cudaFree (b);
  return MS_OK;
}

__global__ void __launch_bounds__ (256)
axpy_minus_kernel (double *__restrict__ r, const double *__restrict__ b, size_t n)
{
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n)
    r[i] -= b[i];
}

int
ms_k_axpy_minus (ms_ctx *c, double *r, const double *b, size_t n, cudaStream_t s)
{
  axpy_minus_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>> (r, b, n);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}
