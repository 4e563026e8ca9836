// ms_kernels_solve.cu — K6 / K7: the triangular solves, one launch per 64-row piece.
//
//   forward  (compute_y, src/solver.cpp:162-208):  R^T y = b
//   backward (compute_x, src/solver.cpp:210-245):  x_i = d_i (y_i - d_i sum_{j>i} R_ij x_j) / R_ii
//
// R is stored by rows (block-row cyclic across ranks), so both sweeps are organised around what
// one rank can do with its own rows, streaming every stored element of R exactly once
// (4 n^2 bytes per sweep, HBM bound):
//   forward step (piece t, rows/cols [r0, r0+h)):   every CTA first solves the h x h triangular
//       system R_tt^T y_t = b_t - acc_t redundantly in shared memory (one warp, shuffle based),
//       then applies acc_j += sum_{i in t} R_ij y_i to its own slice of columns j >= r0+h —
//       a coalesced walk along the piece's rows.  acc holds the partial dot products of the
//       reference's column sum `sum += a[col_width*i+j]*y[i]` (:185-194).
//   backward step (piece t, descending):   every CTA solves R_tt x_t = D y_t - acc_t, then adds
//       R_kt x_t to acc_k for its slice of rows k < r0 (the column block above the piece).
// With several ranks the owner's acc is completed by a reduction before the step (ms_api.cu).
// The pivot tests |R_jj| < EPS*norm of :195-199 and :232-236 raise flags[0].
#include "ms_common.cuh"

#define LAUNCH_COUNT(c) ((c)->stats.kernel_launches++)

constexpr int SOLVE_THREADS = 256;

// Solve the h x h system with warp 0; T(i,j) = Rs[i*65+j] upper triangular.
// FORWARD: R^T y = rhs (column sweep);  BACKWARD: R x = rhs (row sweep from the bottom).
template <bool FORWARD>
__device__ __forceinline__ void
warp_tri_solve (const double *Rs, double *rhs_sol, int h, double thresh, int *flags)
{
  const int lane = threadIdx.x & 31;
  double t0 = lane < h ? rhs_sol[lane] : 0.0;
  double t1 = lane + 32 < h ? rhs_sol[lane + 32] : 0.0;
  // reciprocals of the pivots up front (two divisions per lane, in parallel) so that the 64
  // dependent steps below are shuffle + multiply + FMA instead of shuffle + IEEE divide
  const double p0 = lane < h ? Rs[lane * 65 + lane] : 1.0;
  const double p1 = lane + 32 < h ? Rs[(lane + 32) * 65 + lane + 32] : 1.0;
  if ((lane < h && fabs (p0) < thresh) || (lane + 32 < h && fabs (p1) < thresh))
    flags[0] = 1;
  const double r0 = 1.0 / p0, r1 = 1.0 / p1;
  if (FORWARD)
    {
      for (int i = 0; i < h; i++)
        {
          const double yi = __shfl_sync (0xffffffffu, i < 32 ? t0 * r0 : t1 * r1, i & 31);
          if (lane == (i & 31))
            {
              if (i < 32)
                t0 = yi;
              else
                t1 = yi;
            }
          // columns j > i lose R_ij * y_i
          if (lane > i && lane < h)
            t0 -= Rs[i * 65 + lane] * yi;
          if (lane + 32 > i && lane + 32 < h)
            t1 -= Rs[i * 65 + lane + 32] * yi;
        }
    }
  else
    {
      for (int i = h - 1; i >= 0; i--)
        {
          const double xi = __shfl_sync (0xffffffffu, i < 32 ? t0 * r0 : t1 * r1, i & 31);
          if (lane == (i & 31))
            {
              if (i < 32)
                t0 = xi;
              else
                t1 = xi;
            }
          // rows k < i lose R_ki * x_i
          if (lane < i)
            t0 -= Rs[lane * 65 + i] * xi;
          if (lane + 32 < i)
            t1 -= Rs[(lane + 32) * 65 + i] * xi;
        }
    }
  if (lane < h)
    rhs_sol[lane] = t0;
  if (lane + 32 < h)
    rhs_sol[lane + 32] = t1;
}

__global__ void __launch_bounds__ (SOLVE_THREADS)
forward_step_kernel (const double *__restrict__ A, size_t ld, size_t lrow0, size_t r0, int h, size_t n,
                     const double *__restrict__ b, double *__restrict__ y, double *__restrict__ acc,
                     double thresh, int *flags)
{
  __shared__ double Rs[64 * 65];
  __shared__ double ys[64];
  for (int e = threadIdx.x; e < h * h; e += SOLVE_THREADS)
    {
      int i = e / h, j = e - i * h;
      Rs[i * 65 + j] = j >= i ? A[(lrow0 + i) * ld + r0 + j] : 0.0;
    }
  if (threadIdx.x < 64)
    ys[threadIdx.x] = threadIdx.x < h ? b[r0 + threadIdx.x] - acc[r0 + threadIdx.x] : 0.0;
  __syncthreads ();
  if (threadIdx.x < 32)
    warp_tri_solve<true> (Rs, ys, h, thresh, blockIdx.x == 0 ? flags : flags + 1);
  __syncthreads ();
  if (blockIdx.x == 0 && threadIdx.x < h)
    y[r0 + threadIdx.x] = ys[threadIdx.x];
  // trailing columns: acc_j += sum_i R(i, j) y_i.  The CTA owns 64 columns; its four 64-thread
  // groups take 16 rows of the piece each, so every thread has its 16 loads in flight at once
  // (the walk is HBM-latency bound otherwise) and a row segment is a 512-byte coalesced read.
  __shared__ double part[4][64];
  const int cx = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const size_t j = r0 + h + (size_t)blockIdx.x * 64 + cx;
  double s = 0.0;
  if (j < n)
    {
      const double *col = A + (lrow0 + (size_t)rg * 16) * ld + j;
      double v[16];
#pragma unroll
      for (int i = 0; i < 16; i++)
        v[i] = (rg * 16 + i < h) ? col[(size_t)i * ld] : 0.0;
#pragma unroll
      for (int i = 0; i < 16; i++)
        s += v[i] * ys[rg * 16 + i];
    }
  part[rg][cx] = s;
  __syncthreads ();
  if (rg == 0 && j < n)
    acc[j] += (part[0][cx] + part[1][cx]) + (part[2][cx] + part[3][cx]);
}

__global__ void __launch_bounds__ (SOLVE_THREADS)
backward_step_kernel (const double *__restrict__ A, size_t ld, size_t lrow0, size_t r0, int h,
                      const double *__restrict__ yv, const double *__restrict__ d, double *__restrict__ x,
                      double *__restrict__ acc, size_t lrows_above, size_t m, int rank, int nranks,
                      int owner_has_piece, double thresh, int *flags)
{
  __shared__ double Rs[64 * 65];
  __shared__ double xs[64];
  if (owner_has_piece)
    {
      for (int e = threadIdx.x; e < h * h; e += SOLVE_THREADS)
        {
          int i = e / h, j = e - i * h;
          Rs[i * 65 + j] = j >= i ? A[(lrow0 + i) * ld + r0 + j] : 0.0;
        }
      if (threadIdx.x < h)
        xs[threadIdx.x] = d[r0 + threadIdx.x] * yv[r0 + threadIdx.x] - acc[r0 + threadIdx.x];
      __syncthreads ();
      if (threadIdx.x < 32)
        warp_tri_solve<false> (Rs, xs, h, thresh, blockIdx.x == 0 ? flags : flags + 1);
      __syncthreads ();
      if (blockIdx.x == 0 && threadIdx.x < h)
        x[r0 + threadIdx.x] = xs[threadIdx.x];
    }
  else
    {
      // another rank solved the piece; x_t was broadcast into x[]
      if (threadIdx.x < h)
        xs[threadIdx.x] = x[r0 + threadIdx.x];
      __syncthreads ();
    }
  // rows above: acc_k += sum_j R(k, r0 + j) x_j ; 16 lanes per row, 16 rows per CTA pass
  const int sub = threadIdx.x & 15;
  const int rslot = threadIdx.x >> 4;
  for (size_t lk = (size_t)blockIdx.x * 16 + rslot; lk < lrows_above; lk += (size_t)gridDim.x * 16)
    {
      const double *row = A + lk * ld + r0;
      double s = 0.0;
      for (int j = sub; j < h; j += 16)
        s += row[j] * xs[j];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1)
        s += __shfl_xor_sync (0xffffffffu, s, o, 16);
      if (sub == 0)
        {
          size_t lb = lk / m;
          size_t g = (lb * (size_t)nranks + (size_t)rank) * m + (lk - lb * m);
          acc[g] += s;
        }
    }
}

int
ms_k_forward_step (ms_ctx *c, const ms_piece &pc, double thresh, cudaStream_t s)
{
  size_t rest = c->n - (pc.row0 + (size_t)pc.h);
  unsigned grid = (unsigned)((rest + 63) / 64);
  if (grid == 0)
    grid = 1;
  forward_step_kernel<<<grid, SOLVE_THREADS, 0, s>>> (c->A, c->ld, pc.lrow0, pc.row0, pc.h, c->n,
                                                    c->vec[MS_VEC_B], c->vec[MS_VEC_Y], c->acc, thresh,
                                                    c->flags);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// lrows_above: local rows of this rank whose global row is < pc.row0
int
ms_k_backward_step (ms_ctx *c, const ms_piece &pc, size_t lrows_above, double thresh, cudaStream_t s)
{
  int own = pc.owner == c->rank;
  if (!own && lrows_above == 0)
    return MS_OK;
  size_t want = (lrows_above + 15) / 16;
  unsigned grid = (unsigned)(want < 1 ? 1 : (want > 2368 ? 2368 : want));
  backward_step_kernel<<<grid, SOLVE_THREADS, 0, s>>> (c->A, c->ld, pc.lrow0, pc.row0, pc.h,
                                                     c->vec[MS_VEC_Y], c->vec[MS_VEC_D], c->vec[MS_VEC_X],
                                                     c->acc, lrows_above, c->m, c->rank, c->nranks, own,
                                                     thresh, c->flags);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}
