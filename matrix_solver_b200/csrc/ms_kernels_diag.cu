// ms_kernels_diag.cu — K1 + K2: the diagonal-block kernel.
//
// One CTA factors one h x h diagonal piece (h <= 64 on the product path; up to 112 through
// the stateless block API) held in shared memory:
//     A = R^T D R,  R upper with positive diagonal, D = diag(+-1)
// then inverts R in the same CTA and emits Wt = inverse scaled by D, k-major and
// chunk-major, which is the A operand of the panel kernel (panel = Wt^T * A_row).
//
// Reference routines restated (semantics AND operation order, so the block-level results
// are bit-identical to the reference's -O2 build; see tests/test_gpu_blocks.py):
//   cholesky_decomp_U  src/solver.cpp:106-160   pivot s = a_ii - sum_k a_ki^2 d_k (ascending k),
//                                              d_i = sgn(s) with sgn(0) = -1 (defines.h:26-30),
//                                              a_ii = sqrt|s|, a_ij = (a_ij - sum..)/(a_ii*d_i)
//   reverse_upper      src/matrix_op.cpp:244-285 inv_ii = 1/a_ii,
//                                              inv_ij = -(sum_{k=i+1..j} a_ik inv_kj)/a_ii
// The in-block elimination is right-looking: entry (k,j) receives its subtractions in
// ascending pivot order, exactly the order of the reference's inner k loop.  All arithmetic
// uses the _rn intrinsics so that nvcc cannot contract a*b-c into an FMA.
//
// The reference tests the not-yet-updated a_ii first (src/solver.cpp:117); that "pre" test
// runs for a whole block row in pretest_kernel before any piece of it is touched.
#include "ms_common.cuh"

#define LAUNCH_COUNT(c) ((c)->stats.kernel_launches++)

__global__ void __launch_bounds__ (256)
pretest_kernel (const double *__restrict__ blk, size_t ld, int h, double thresh, int *flags)
{
  for (int i = threadIdx.x; i < h; i += 256)
    if (fabs (blk[(size_t)i * ld + i]) < thresh)
      flags[0] = 1;
}

int
ms_k_pretest (ms_ctx *c, const double *blk, size_t ld, int h, double thresh, cudaStream_t s)
{
  pretest_kernel<<<1, 256, 0, s>>> (blk, ld, h, thresh, c->flags);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// dynamic smem: a[h][h+1] | v[h][h+1] | rdiag[h] | dd[h]
__global__ void __launch_bounds__ (256)
diag_kernel (double *__restrict__ blk, size_t ld, int h, double thresh, double *__restrict__ d_out,
             double *__restrict__ wt_out, int wt_h, double *__restrict__ inv_out, int *flags)
{
  extern __shared__ double smd[];
  const int S = h + 1;
  double *a = smd;
  double *v = a + (size_t)h * S;
  double *rdiag = v + (size_t)h * S;
  double *dd = rdiag + h;
  const int tid = threadIdx.x;

  for (int e = tid; e < h * h; e += 256)
    {
      int i = e / h, j = e - i * h;
      a[i * S + j] = j >= i ? blk[(size_t)i * ld + j] : 0.0;
      v[i * S + j] = 0.0;
    }
  __syncthreads ();

  // ---- K1: signed Cholesky, right-looking inside the piece (16 x 16 thread grid over the
  // trailing square; no integer divisions on the critical path)
  const int tx = tid & 15, ty = tid >> 4;
  for (int i = 0; i < h; i++)
    {
      double s = a[i * S + i];
      double di = s > 0 ? 1.0 : -1.0;
      double rii = __dsqrt_rn (fabs (s));
      double den = __dmul_rn (rii, di);
      if (tid == 0)
        {
          rdiag[i] = rii;
          dd[i] = di;
          if (fabs (rii) < thresh)
            flags[0] = 1;
        }
      for (int j = i + 1 + tid; j < h; j += 256)
        a[i * S + j] = __ddiv_rn (a[i * S + j], den);
      __syncthreads ();
      for (int k = i + 1 + ty; k < h; k += 16)
        {
          const double aik = a[i * S + k];
          for (int j = i + 1 + tx; j < h; j += 16)
            if (j >= k)
              a[k * S + j] = __dsub_rn (a[k * S + j], __dmul_rn (__dmul_rn (aik, a[i * S + j]), di));
        }
      __syncthreads ();
    }

  // ---- K2: inverse, one thread per column, rows bottom-up, k ascending
  for (int j = tid; j < h; j += 256)
    {
      for (int i = j; i >= 0; i--)
        {
          double val;
          if (i == j)
            val = __ddiv_rn (1.0, rdiag[i]);
          else
            {
              double sum = 0;
              for (int k = i + 1; k <= j; k++)
                sum = __dadd_rn (sum, __dmul_rn (a[i * S + k], v[k * S + j]));
              val = __ddiv_rn (-sum, rdiag[i]);
            }
          v[i * S + j] = val;
        }
    }
  __syncthreads ();

  // ---- write back: R in place (upper), d, the inverse, and Wt = inverse * D (chunk-major)
  for (int e = tid; e < h * h; e += 256)
    {
      int i = e / h, j = e - i * h;
      if (j > i)
        blk[(size_t)i * ld + j] = a[i * S + j];
      else if (j == i)
        blk[(size_t)i * ld + j] = rdiag[i];
    }
  if (d_out)
    for (int i = tid; i < h; i += 256)
      d_out[i] = dd[i];
  if (inv_out)
    for (int e = tid; e < h * h; e += 256)
      {
        int i = e / h, j = e - i * h;
        inv_out[(size_t)i * h + j] = j >= i ? v[i * S + j] : 0.0;
      }
  if (wt_out)
    for (int e = tid; e < wt_h * wt_h; e += 256)
      {
        int k = e / wt_h, i = e - k * wt_h;
        double val = (k < h && i < h && i >= k) ? v[k * S + i] * dd[i] : 0.0;
        wt_out[chunk_major_index (k, i, wt_h)] = val;
      }
}

// d_out: h doubles (device), wt_out: chunk-major [MS_SUB/16][MS_SUB][16] or null
int
ms_k_diag (ms_ctx *c, double *blk, size_t ld, int h, double thresh, double *d_out, double *wt_out,
           double *inv_out, cudaStream_t s)
{
  MS_CHECK_ARG (h >= 1 && h <= 112, "diag piece must have 1..112 rows");
  size_t smem = (2 * (size_t)h * (h + 1) + 2 * h) * sizeof (double);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured)
    {
      MS_CUDA (cudaFuncSetAttribute (diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      configured = 227 * 1024;
    }
  diag_kernel<<<1, 256, smem, s>>> (blk, ld, h, thresh, d_out, wt_out, MS_SUB, inv_out, c->flags);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}
