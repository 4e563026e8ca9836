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
// runs for a whole block row at the top of the diag kernel of its first piece, before any piece
// of the block row is touched.
#include "ms_common.cuh"

#define LAUNCH_COUNT(c) ((c)->stats.kernel_launches++)

// dynamic smem: a[h][h+1] | v[h][h+1] | rdiag[h] | dd[h]
//
// Schedule of K1 (one __syncthreads per pivot instead of two, and the expensive IEEE sqrt/divide of
// pivot i+1 overlapped with the bulk of update i):
//   state before step i : row i is final (updated by pivots < i and scaled), rdiag[i], dd[i] known
//   warp 0  ("pivot warp"): applies pivot i to row i+1 only, then immediately finishes row i+1
//                           (sign, sqrt, scale) — the critical path of the whole factorisation
//   warps 1-7             : apply pivot i to rows k >= i+2 (14 x 16 thread grid, operands gathered
//                           into registers first so that the loads and the fp64 chains pipeline)
// Every entry still receives its subtractions in ascending pivot order with the same roundings.
template <int NTR, int NTC>
__global__ void __launch_bounds__ (256)
diag_kernel (double *__restrict__ blk, size_t ld, int h, double thresh, double *__restrict__ d_out,
             double *__restrict__ wt_out, int wt_h, double *__restrict__ inv_out, int want_inverse,
             const double *__restrict__ pre_diag, size_t pre_ld, int pre_count, int *flags)
{
  extern __shared__ double smd[];
  const int S = h + 1;
  double *a = smd;
  double *v = a + (size_t)h * S;
  double *rdiag = v + (size_t)h * S;
  double *dd = rdiag + h;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lane = tid & 31;
#ifdef MS_DIAG_TIMING
  long long tc0 = clock64 ();
#endif

  // the reference's pre-update pivot test (src/solver.cpp:117) for the whole block row
  for (int i = tid; i < pre_count; i += 256)
    if (fabs (pre_diag[(size_t)i * pre_ld + i]) < thresh)
      flags[0] = 1;

  for (int i = ty; i < h; i += 16)
    for (int j = tx; j < h; j += 16)
      {
        a[i * S + j] = j >= i ? blk[(size_t)i * ld + j] : 0.0;
        if (want_inverse)
          v[i * S + j] = 0.0;
      }
  __syncthreads ();
#ifdef MS_DIAG_TIMING
  long long tc1 = clock64 ();
#endif

  // finish row `row`: its entries are fully updated; compute sign / sqrt / scale (one warp)
  auto finish_row = [&] (int row, double c0, double c1) {
    // c0 / c1: updated values of columns row+lane / row+lane+32 ... held by this lane
    const double s = __shfl_sync (0xffffffffu, c0, 0);  // column `row` itself lives in lane 0
    const double di = s > 0 ? 1.0 : -1.0;
    const double rii = __dsqrt_rn (fabs (s));
    const double den = __dmul_rn (rii, di);
    if (lane == 0)
      {
        rdiag[row] = rii;
        dd[row] = di;
        if (fabs (rii) < thresh)
          flags[0] = 1;
      }
    const int j0 = row + lane, j1 = row + lane + 32;
    if (lane > 0 && j0 < h)
      a[row * S + j0] = __ddiv_rn (c0, den);
    if (j1 < h)
      a[row * S + j1] = __ddiv_rn (c1, den);
    // columns beyond row + 63 (only for h > 64): same treatment, strided
    for (int j = row + lane + 64; j < h; j += 32)
      a[row * S + j] = __ddiv_rn (a[row * S + j], den);
  };

  if (tid < 32)
    {
      const int j0 = lane, j1 = lane + 32;
      finish_row (0, j0 < h ? a[j0] : 0.0, j1 < h ? a[j1] : 0.0);
    }
  __syncthreads ();

#ifdef MS_DIAG_TIMING
  long long acc_p = 0, acc_u = 0;
#endif
  for (int i = 0; i + 1 < h; i++)
    {
      const double di = dd[i];
#ifdef MS_DIAG_TIMING
      long long tl0 = clock64 ();
#endif
      if (tid < 32)
        {
          // pivot warp: row i+1 <- row i+1 - (r_{i,i+1} r_{i,j}) d_i, then finish it
          const int row = i + 1;
          const double rk = a[i * S + row];
          const int j0 = row + lane, j1 = row + lane + 32;
          double c0 = 0.0, c1 = 0.0;
          if (j0 < h)
            c0 = __dsub_rn (a[row * S + j0], __dmul_rn (__dmul_rn (rk, a[i * S + j0]), di));
          if (j1 < h)
            c1 = __dsub_rn (a[row * S + j1], __dmul_rn (__dmul_rn (rk, a[i * S + j1]), di));
          for (int j = row + lane + 64; j < h; j += 32)
            a[row * S + j] = __dsub_rn (a[row * S + j], __dmul_rn (__dmul_rn (rk, a[i * S + j]), di));
          finish_row (row, c0, c1);
#ifdef MS_DIAG_TIMING
          acc_p += clock64 () - tl0;
#endif
        }
      else
        {
          // rows k >= i+2, cols j >= k: gather, compute, scatter (14 x 16 threads)
          const int t = tid - 32, ux = t & 15, uy = t >> 4;
          double rk[NTR], rj[NTC], c[NTR][NTC];
#pragma unroll
          for (int u = 0; u < NTR; u++)
            {
              const int k = i + 2 + uy + 14 * u;
              rk[u] = k < h ? a[i * S + k] : 0.0;
            }
#pragma unroll
          for (int w = 0; w < NTC; w++)
            {
              const int j = i + 2 + ux + 16 * w;
              rj[w] = j < h ? a[i * S + j] : 0.0;
            }
#pragma unroll
          for (int u = 0; u < NTR; u++)
#pragma unroll
            for (int w = 0; w < NTC; w++)
              {
                const int k = i + 2 + uy + 14 * u, j = i + 2 + ux + 16 * w;
                c[u][w] = (k < h && j < h && j >= k) ? a[k * S + j] : 0.0;
              }
#pragma unroll
          for (int u = 0; u < NTR; u++)
#pragma unroll
            for (int w = 0; w < NTC; w++)
              c[u][w] = __dsub_rn (c[u][w], __dmul_rn (__dmul_rn (rk[u], rj[w]), di));
#pragma unroll
          for (int u = 0; u < NTR; u++)
#pragma unroll
            for (int w = 0; w < NTC; w++)
              {
                const int k = i + 2 + uy + 14 * u, j = i + 2 + ux + 16 * w;
                if (k < h && j < h && j >= k)
                  a[k * S + j] = c[u][w];
              }
#ifdef MS_DIAG_TIMING
          acc_u += clock64 () - tl0;
#endif
        }
      __syncthreads ();
    }
#ifdef MS_DIAG_TIMING
  long long tc2 = clock64 ();
  if (tid == 0)
    flags[2] = (int)acc_p;
  if (tid == 255)
    flags[3] = (int)acc_u;
#endif

  // ---- K2: inverse, one thread per column, rows bottom-up, k ascending
  if (want_inverse)
    {
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
    }
#ifdef MS_DIAG_TIMING
  long long tc3 = clock64 ();
#endif

  // ---- write back: R in place (upper), d, the inverse, and Wt = inverse * D (chunk-major)
  for (int i = ty; i < h; i += 16)
    for (int j = tx; j < h; j += 16)
      {
        if (j > i)
          blk[(size_t)i * ld + j] = a[i * S + j];
        else if (j == i)
          blk[(size_t)i * ld + j] = rdiag[i];
        if (inv_out)
          inv_out[(size_t)i * h + j] = j >= i ? v[i * S + j] : 0.0;
      }
  if (d_out)
    for (int i = tid; i < h; i += 256)
      d_out[i] = dd[i];
  if (wt_out)
    for (int k = ty; k < wt_h; k += 16)
      for (int i = tx; i < wt_h; i += 16)
        {
          double val = (k < h && i < h && i >= k) ? v[k * S + i] * dd[i] : 0.0;
          wt_out[chunk_major_index (k, i, wt_h)] = val;
        }
#ifdef MS_DIAG_TIMING
  __syncthreads ();
  if (tid == 0)
    {
      long long tc4 = clock64 ();
      flags[4] = (int)(tc1 - tc0);
      flags[5] = (int)(tc2 - tc1);
      flags[6] = (int)(tc3 - tc2);
      flags[7] = (int)(tc4 - tc3);
    }
#endif
}

// ---------------------------------------------------------------------------- K1, strip schedule (h <= 64)
// diag_kernel above pays one 256-thread barrier and a shared-memory round trip per pivot: ~870 cycles per
// pivot, 39 us per 64 x 64 piece — the longest link of the panel chain (profiles/README.md).  The chain of
// truly dependent operations per pivot is only sqrt -> multiply -> divide -> shuffle -> multiply, multiply,
// subtract.  Here the pivots go in strips of 8 rows:
//   pivot warp (warp 0) : keeps the strip in registers (lane = columns lane and lane + 32) and factors it
//                         with shuffles only — no barrier and no shared memory on the dependent chain;
//   update warps (1-7)  : S1  apply the previous strip's 8 pivots to the rows of the CURRENT strip (the only
//                             thing the pivot warp waits for: ~2 elements per thread), then `bar.arrive 1`;
//                         S2  apply them to all rows below the current strip while the pivot warp factors;
//                             `bar.sync 2` waits for the pivot warp's strip.
// Two named barriers per strip, both of the arrive / sync kind, so neither side waits for work it does not
// need.  Every entry still receives its subtractions in ascending pivot order, each product and difference
// rounded on its own (the _rn intrinsics keep nvcc from contracting them): results are bit-identical to
// diag_kernel and to the reference (tests/test_gpu_parity.py::test_diag_factor_and_inverse_bit_exact runs
// through this kernel for w <= 64).
__device__ __forceinline__ void
named_bar_sync (int id, int count)
{
  asm volatile ("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ void
named_bar_arrive (int id, int count)
{
  asm volatile ("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// factor the 8-row strip held in c[u][half] (rows i0 + u, columns lane + 32 * half); LO = the strip still
// has columns < 32 (i0 < 32), otherwise only the upper half of the columns exists.
//
// What bounds this routine is the pivot warp's INSTRUCTION COUNT: a single warp issues about one instruction
// every four cycles here (tools/lat_probe.cu and the per-phase stamps of tools/diag_timing.py: 32 register
// moves cost 130 cycles), so the pivot loop is fully unrolled (static register indices, no rotation, row
// updates only for the rows that still need them) and the per-pivot arithmetic is kept minimal.
// FAST: the pivot is taken through y = rsqrt(|s|) (CUDA's rsqrt, 1 ulp): r_ii = |s| y and the row is
// MULTIPLIED by y d_i — ~15 instructions instead of the ~45 of an IEEE sqrt and two IEEE divides (whose
// dependent chains, 92 + 125 + 125 cycles, also serialise).  The entries then differ from the reference's
// sqrt / divide sequence by at most a few ulp — far inside the c n eps cond(A) budget of the factorisation,
// but not bit-identical any more; !FAST keeps the reference's roundings ("exact_pivots" option, block API).
constexpr int DIAG_NB = 4;  // pivots per strip: the pivot warp updates (NB - 1 - t) rows per pivot itself

template <bool LO, bool FAST>
__device__ __forceinline__ void
strip_factor (double (&c)[DIAG_NB][2], int i0, int h, int lane, double thresh, double *rdiag, double *dd, int *flags)
{
  const int j0 = lane, j1 = lane + 32;
  double keep_r = 1.0, keep_d = 1.0;  // lane t keeps pivot t's r_ii and d_i and stores them after the strip
#pragma unroll
  for (int t = 0; t < DIAG_NB; t++)
    {
      const int row = i0 + t;
      // no `row < h` test here: a GPU does not speculate, and a branch on a kernel argument costs the
      // dependent chain ~40 cycles per pivot.  Rows >= h are zero padding; their "pivots" produce NaN/inf
      // that only ever reach other padding rows, and nothing of the padding is written back.
        {
          const double s = __shfl_sync (0xffffffffu, LO ? c[t][0] : c[t][1], row & 31);
          const double di = s > 0 ? 1.0 : -1.0;
          double rii, den;
          if (FAST)
            {
              const double y = rsqrt (fabs (s));
              rii = s != 0.0 ? fabs (s) * y : 0.0;  // a zero pivot must fail the threshold test, not turn into NaN
              den = y * di;                          // the multiplier 1 / (r_ii d_i)
            }
          else
            {
              rii = __dsqrt_rn (fabs (s));
              den = __dmul_rn (rii, di);
            }
          if (lane == t)
            {
              keep_r = rii;
              keep_d = di;
            }
          if (FAST)
            {
              if (LO && j0 > row)
                c[t][0] *= den;
              if (j1 > row)
                c[t][1] *= den;
            }
          else
            {
              // lanes left of the pivot hold 0.0; they divide 1.0 instead so that the quotient the compiler
              // evaluates for every lane stays on the fast path of the IEEE divide sequence
              if (LO)
                {
                  const double q0 = __ddiv_rn (j0 > row ? c[t][0] : 1.0, den);
                  if (j0 > row)
                    c[t][0] = q0;
                }
              const double q1 = __ddiv_rn (j1 > row ? c[t][1] : 1.0, den);
              if (j1 > row)
                c[t][1] = q1;
            }
#pragma unroll
          for (int u = t + 1; u < DIAG_NB; u++)
            {
              // r(row, k), k = i0 + u, lives in the lane that owns column k
              const double rk = __shfl_sync (0xffffffffu, LO ? c[t][0] : c[t][1], (i0 + u) & 31);
              if (FAST)
                {
                  // one FMA per element on the pivot chain (d_i = +-1: the product with it is exact)
                  const double mk = -(rk * di);
                  if (LO)
                    c[u][0] = fma (mk, c[t][0], c[u][0]);
                  c[u][1] = fma (mk, c[t][1], c[u][1]);
                }
              else
                {
                  if (LO)
                    c[u][0] = __dsub_rn (c[u][0], __dmul_rn (__dmul_rn (rk, c[t][0]), di));
                  c[u][1] = __dsub_rn (c[u][1], __dmul_rn (__dmul_rn (rk, c[t][1]), di));
                }
            }
        }
    }
  if (lane < DIAG_NB && i0 + lane < h)
    {
      rdiag[i0 + lane] = keep_r;
      dd[i0 + lane] = keep_d;
      if (fabs (keep_r) < thresh)
        flags[0] = 1;
    }
}

template <bool FAST>
__global__ void __launch_bounds__ (256)
diag_strip_kernel (double *__restrict__ blk, size_t ld, int h, double thresh, double *__restrict__ d_out,
                   double *__restrict__ wt_out, int wt_h, double *__restrict__ inv_out, int want_inverse,
                   const double *__restrict__ pre_diag, size_t pre_ld, int pre_count, int *flags,
                   long long *cyc = nullptr, int solo = 0)
{
  constexpr int NB = DIAG_NB, S = MS_SUB + 1;
  // cyc != nullptr (tools/diag_timing.py only): clock64 stamps of the phases, written by thread 0
  long long tc0 = 0, tc1 = 0, tc2 = 0, acc_wait = 0, acc_fact = 0;
  if (cyc)
    tc0 = clock64 ();
  extern __shared__ double smd[];
  double *a = smd;                       // [64][65], rows >= h and columns >= h are zero
  double *v = a + (size_t)MS_SUB * S;    // inverse (only when wanted)
  double *rdiag = v + (size_t)MS_SUB * S;
  double *dd = rdiag + MS_SUB;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lane = tid & 31, warp = tid >> 5;

  // The block first, all 16 loads of a thread in flight at once (the rolled loop of round 1 paid several global
  // round trips here: 4900 cycles of a 44 000-cycle kernel, profiles/r02_diag_timing.txt), then the reference's
  // pre-update pivot test (src/solver.cpp:117) for the block row, whose loads overlap them.
  double stage[16];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int q = 0; q < 4; q++)
      {
        const int i = ty + 16 * r, j = tx + 16 * q;
        stage[r * 4 + q] = (i < h && j < h && j >= i) ? blk[(size_t)i * ld + j] : 0.0;
      }
  bool pre_bad = false;
  for (int i = tid; i < pre_count; i += 256)
    pre_bad |= fabs (pre_diag[(size_t)i * pre_ld + i]) < thresh;
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int q = 0; q < 4; q++)
      {
        const int i = ty + 16 * r, j = tx + 16 * q;
        a[i * S + j] = stage[r * 4 + q];
        if (want_inverse)
          v[i * S + j] = 0.0;
      }
  if (pre_bad)
    flags[0] = 1;
  __syncthreads ();
  if (cyc)
    tc1 = clock64 ();

  const int nstrips = (h + NB - 1) / NB;
  for (int b = 0; b < nstrips; b++)
    {
      const int i0 = NB * b;
      if (warp == 0)
        {
          // ---- pivot warp: wait for S1 (strip rows carry every pivot < i0), factor out of registers
          long long tw0 = cyc ? clock64 () : 0;
          if (b > 0 && !solo)
            named_bar_sync (1, 256);
          long long tw1 = cyc ? clock64 () : 0;
          double c[NB][2];
#pragma unroll
          for (int u = 0; u < NB; u++)
            {
              c[u][0] = a[(i0 + u) * S + lane];
              c[u][1] = a[(i0 + u) * S + lane + 32];
            }
          if (i0 < 32)
            strip_factor<true, FAST> (c, i0, h, lane, thresh, rdiag, dd, flags);
          else
            strip_factor<false, FAST> (c, i0, h, lane, thresh, rdiag, dd, flags);
          if (cyc)
            {
              acc_wait += tw1 - tw0;
              acc_fact += clock64 () - tw1;
            }
#pragma unroll
          for (int u = 0; u < NB; u++)
            {
              const int row = i0 + u;
              if (lane > row)
                a[row * S + lane] = c[u][0];
              if (lane + 32 > row)
                a[row * S + lane + 32] = c[u][1];
            }
          if (!solo)
            named_bar_arrive (2, 256);
        }
      else if (!solo)
        {
          const int t2 = tid - 32;  // 0 .. 223
          if (b > 0)
            {
              // ---- S1: previous strip's pivots (rows i0-NB .. i0-1, final) into the rows of this strip
              for (int e = t2; e < NB * MS_SUB; e += 224)
                {
                  const int k = i0 + (e >> 6), j = e & 63;
                  if (j >= k && j < h && k < h)
                    {
                      double cv = a[k * S + j];
#pragma unroll
                      for (int t = 0; t < NB; t++)
                        {
                          const int i = i0 - NB + t;
                          cv = __dsub_rn (cv, __dmul_rn (__dmul_rn (a[i * S + k], a[i * S + j]), dd[i]));
                        }
                      a[k * S + j] = cv;
                    }
                }
              named_bar_arrive (1, 256);
              // ---- S2: the same pivots into every row below this strip (14 x 16 thread grid; at most
              // 64 - 2 NB rows and columns remain)
              const int base = i0 + NB;
              if (base < h)
                {
                  constexpr int NTR = (MS_SUB - 2 * NB + 13) / 14, NTC = (MS_SUB - 2 * NB + 15) / 16;
                  const int ux = t2 & 15, uy = t2 >> 4;
                  double c[NTR][NTC];
#pragma unroll
                  for (int u = 0; u < NTR; u++)
#pragma unroll
                    for (int w = 0; w < NTC; w++)
                      {
                        const int k = base + uy + 14 * u, j = base + ux + 16 * w;
                        c[u][w] = (k < h && j < h && j >= k) ? a[k * S + j] : 0.0;
                      }
#pragma unroll
                  for (int t = 0; t < NB; t++)
                    {
                      const int i = i0 - NB + t;
                      const double di = dd[i];
                      double rk[NTR], rj[NTC];
#pragma unroll
                      for (int u = 0; u < NTR; u++)
                        {
                          const int k = base + uy + 14 * u;
                          rk[u] = k < h ? a[i * S + k] : 0.0;
                        }
#pragma unroll
                      for (int w = 0; w < NTC; w++)
                        {
                          const int j = base + ux + 16 * w;
                          rj[w] = j < h ? a[i * S + j] : 0.0;
                        }
#pragma unroll
                      for (int u = 0; u < NTR; u++)
#pragma unroll
                        for (int w = 0; w < NTC; w++)
                          c[u][w] = __dsub_rn (c[u][w], __dmul_rn (__dmul_rn (rk[u], rj[w]), di));
                    }
#pragma unroll
                  for (int u = 0; u < NTR; u++)
#pragma unroll
                    for (int w = 0; w < NTC; w++)
                      {
                        const int k = base + uy + 14 * u, j = base + ux + 16 * w;
                        if (k < h && j < h && j >= k)
                          a[k * S + j] = c[u][w];
                      }
                }
            }
          named_bar_sync (2, 256);  // strip b is final in shared memory
        }
    }
  __syncthreads ();
  if (cyc)
    tc2 = clock64 ();

  // ---- K2: inverse, one thread per column, rows bottom-up, k ascending (as in diag_kernel)
  if (want_inverse)
    {
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
    }

  // ---- write back: R in place (upper), d, the inverse, and Wt = inverse * D (chunk-major)
  for (int i = ty; i < h; i += 16)
    for (int j = tx; j < h; j += 16)
      {
        if (j > i)
          blk[(size_t)i * ld + j] = a[i * S + j];
        else if (j == i)
          blk[(size_t)i * ld + j] = rdiag[i];
        if (inv_out)
          inv_out[(size_t)i * h + j] = j >= i ? v[i * S + j] : 0.0;
      }
  if (d_out)
    for (int i = tid; i < h; i += 256)
      d_out[i] = dd[i];
  if (wt_out)
    for (int k = ty; k < wt_h; k += 16)
      for (int i = tx; i < wt_h; i += 16)
        {
          double val = (k < h && i < h && i >= k) ? v[k * S + i] * dd[i] : 0.0;
          wt_out[chunk_major_index (k, i, wt_h)] = val;
        }
  if (cyc && tid == 0)
    {
      cyc[0] = tc1 - tc0;          // pre-test + load
      cyc[1] = tc2 - tc1;          // all strips
      cyc[2] = clock64 () - tc2;   // inverse (if any) + write back
      cyc[3] = acc_wait;           // pivot warp: waiting for S1 (bar.sync 1)
      cyc[4] = acc_fact;           // pivot warp: strip load + factor
    }
}

// d_out: h doubles (device), wt_out: chunk-major [MS_SUB/16][MS_SUB][16] or null.  The inverse is
// computed only when someone consumes it (wt_out / inv_out).  pre_diag/pre_ld/pre_count: diagonal to
// run the pre-update pivot test on (the whole block row, handed to its first piece).
int
ms_k_diag (ms_ctx *c, double *blk, size_t ld, int h, double thresh, double *d_out, double *wt_out,
           double *inv_out, const double *pre_diag, size_t pre_ld, int pre_count, cudaStream_t s)
{
  MS_CHECK_ARG (h >= 1 && h <= 112, "diag piece must have 1..112 rows");
  int want_inverse = (wt_out != nullptr || inv_out != nullptr);
  size_t smem = (((size_t)(want_inverse ? 2 : 1)) * h * (h + 1) + 2 * h) * sizeof (double);
  if (!want_inverse)
    smem = (2 * (size_t)h * (h + 1) + 2 * h) * sizeof (double);  // layout keeps v's slot reserved
  // function attributes are per device: one flag per device index
  static bool configured_dev[64] = {};
  bool &configured = configured_dev[c->device & 63];
  if (!configured)
    {
      MS_CUDA (cudaFuncSetAttribute (diag_kernel<5, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      MS_CUDA (cudaFuncSetAttribute (diag_kernel<8, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      MS_CUDA (cudaFuncSetAttribute (diag_strip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      MS_CUDA (cudaFuncSetAttribute (diag_strip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      configured = true;
    }
  // Full 64-row pieces of the DMMA path take the strip-scheduled kernel with fast pivots; everything else
  // (short pieces, block sizes that are not multiples of 64 — where the reference's fixtures with their
  // exactly singular minors live —, the block API, "exact_pivots" = 1) keeps the reference's roundings.
  static int strip = -1;
  if (strip < 0)
    {
      const char *e = getenv ("MS_DIAG_STRIP");
      strip = e ? atoi (e) : 1;
    }
  if (h == MS_SUB && strip && c->fast && !c->force_generic && !c->exact_pivots)
    {
      const size_t smem64 = (2 * (size_t)MS_SUB * (MS_SUB + 1) + 2 * MS_SUB) * sizeof (double);
      diag_strip_kernel<true><<<1, 256, smem64, s>>> (blk, ld, h, thresh, d_out, wt_out, MS_SUB, inv_out, want_inverse,
                                                     pre_diag, pre_ld, pre_count, c->flags);
    }
  else if (h <= 64)
    diag_kernel<5, 4><<<1, 256, smem, s>>> (blk, ld, h, thresh, d_out, wt_out, MS_SUB, inv_out, want_inverse,
                                           pre_diag, pre_ld, pre_count, c->flags);
  else
    diag_kernel<8, 7><<<1, 256, smem, s>>> (blk, ld, h, thresh, d_out, wt_out, MS_SUB, inv_out, want_inverse,
                                           pre_diag, pre_ld, pre_count, c->flags);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// The reference's pre-update pivot test (src/solver.cpp:117) for rows [first, count) of a block row
// whose diagonal starts at diag0: |a_ii| < thresh raises the failure flag.  Used for the pieces after
// the first one when the look-ahead delivers their last update later than the first piece's (the first
// piece's share rides along with its diag kernel).
__global__ void __launch_bounds__ (256)
pretest_kernel (const double *__restrict__ diag0, size_t ld, int first, int count, double thresh, int *flags)
{
  for (int i = first + (int)threadIdx.x; i < count; i += 256)
    if (fabs (diag0[(size_t)i * ld + i]) < thresh)
      flags[0] = 1;
}

int
ms_k_pretest (ms_ctx *c, const double *diag0, size_t ld, int first, int count, double thresh, cudaStream_t s)
{
  if (first >= count)
    return MS_OK;
  pretest_kernel<<<1, 256, 0, s>>> (diag0, ld, first, count, thresh, c->flags);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// ============================================================================ TRSM panel (K3)
// X = D * R^{-T} * A for the h rows of one piece and the columns [col_lo, n), in place, plus a copy
// into the panel buffer.  This is what DRtA / DRtA_l (src/matrix_op.cpp:287-368) compute with the
// explicit inverse; here it is a true triangular solve, one thread per column:
//     x_i = (a_i - sum_{k<i} R_ki x_k) / R_ii,   row i then scaled by d_i.
// R (64 x 64, upper) is staged in shared memory and read with 16-byte broadcast loads; each thread
// keeps its column in registers (right-looking: after x_i, rows k > i lose R_ik x_i — 63 independent
// FMAs, so the kernel runs at DFMA issue rate).  Loads/stores are coalesced across the warp (adjacent
// threads = adjacent columns): 16 B of HBM traffic per element + 8 B for the panel-buffer copy.
// (Round 2 tried a version blocked by 8 rows with rolled loops and the slab's partial sums in shared memory —
// 700 instead of 5000 instructions, 60 instead of 200 registers: 26.6 us against this kernel's 22 us at 16k
// columns.  With one warp per scheduler the kernel is bound by dependent-operation latency, and keeping the
// whole column in registers with four independent FMA chains beats the shared-memory round trips.)
__global__ void __launch_bounds__ (128, 2)
trsm_panel_kernel (const double *__restrict__ Rblk, size_t ldr, double *__restrict__ rows, size_t ld, int h, size_t col_lo,
                   size_t n, const double *__restrict__ dpiece, double *__restrict__ PB, size_t pb_H, size_t pb_k0)
{
  // RT[i][k] = R(k, i) for k < i (the column of R that row i of the solve needs), 0 elsewhere
  constexpr int RS = MS_SUB + 2;  // row stride: even (16-byte loads) and not a multiple of 32 banks
  __shared__ __align__ (16) double RT[MS_SUB * RS];
  __shared__ double rinv[MS_SUB], ds[MS_SUB];
  extern __shared__ double As[];  // [MS_SUB][128]: the CTA's slab of the block row
  const size_t j0 = col_lo + (size_t)blockIdx.x * 128;
  // stage the slab first: every thread has its 64 loads in flight at once (the serial solve below
  // would otherwise wait one HBM round trip per row)
#pragma unroll 16
  for (int k = 0; k < MS_SUB; k++)
    {
      const size_t j = j0 + threadIdx.x;
      As[k * 128 + threadIdx.x] = (k < h && j < n) ? rows[(size_t)k * ld + j] : 0.0;
    }
  for (int e = threadIdx.x; e < MS_SUB * MS_SUB; e += 128)
    {
      const int k = e >> 6, i = e & 63;  // coalesced read of row k of R
      RT[i * RS + k] = (i < h && k < i) ? Rblk[(size_t)k * ldr + i] : 0.0;
    }
  if (threadIdx.x < MS_SUB)
    {
      const int i = threadIdx.x;
      rinv[i] = i < h ? 1.0 / Rblk[(size_t)i * ldr + i] : 1.0;
      ds[i] = i < h ? dpiece[i] : 1.0;
    }
  __syncthreads ();
  const size_t j = j0 + threadIdx.x;
  if (j >= n)
    return;
  double x[MS_SUB];
  double *pcol = rows + j;
  double *ppb = PB + chunk_major_index (pb_k0, j, pb_H);
#pragma unroll
  for (int i = 0; i < MS_SUB; i++)
    {
      // four interleaved partial sums keep four independent FMA chains in flight
      double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
      for (int k = 0; k < (i & ~3); k += 4)
        {
          const double2 ra = *reinterpret_cast<const double2 *> (&RT[i * RS + k]);
          const double2 rb = *reinterpret_cast<const double2 *> (&RT[i * RS + k + 2]);
          p0 = fma (ra.x, x[k], p0);
          p1 = fma (ra.y, x[k + 1], p1);
          p2 = fma (rb.x, x[k + 2], p2);
          p3 = fma (rb.y, x[k + 3], p3);
        }
#pragma unroll
      for (int k = (i & ~3); k < i; k++)
        p0 = fma (RT[i * RS + k], x[k], p0);
      const double xi = (As[i * 128 + threadIdx.x] - ((p0 + p1) + (p2 + p3))) * rinv[i];
      x[i] = xi;
      if (i < h)
        {
          const double val = xi * ds[i];
          pcol[(size_t)i * ld] = val;
          ppb[(size_t)i * MS_CHUNK] = val;   // next k row of the same 16-column chunk
        }
    }
}

// The same solve with FOUR threads per column.  The kernel above is a single stream of ~3500 instructions per
// column and takes 20 us however few columns there are — it sits on the panel chain four times per block row.
// Here the quad of lanes 4c .. 4c+3 shares column c: lane t keeps rows k = t (mod 4) of it (16 values) in
// registers, and the solve runs right-looking,
//     step i:  x_i = a_i / R_ii  (lane i mod 4)  ->  quad broadcast  ->  a_k -= R_ik x_i  for the lane's own k > i,
// so the dependent chain per step is multiply + shuffle + FMA, and every a_k receives its products in ascending
// i with single FMAs — plain forward substitution, one rounding per product.  256 threads = 64 columns per CTA;
// a warp touches 4 rows x 8 consecutive columns per load/store (whole sectors).
// The 16 groups of four steps are ONE rolled loop of ~170 instructions: the lane's register window a[0..15] always
// starts at the current group (after a group its finished value is written out and the window shifts by one), and
// the shared-memory image of R is stored pre-shifted the same way, so every iteration indexes registers and R
// identically; each step runs all 16 FMAs (the image is 0 beyond row 63).  (A fully unrolled variant, 544 FMAs in
// 1500 instructions, measured the same as the loop.)  All global loads of a thread — its 16 values, its 16 entries
// of R, the diagonal — are issued before the first of them is used: with the staging loop rolled the isolated
// launch took 19.7 us, with the loads batched 14.9 us.
constexpr size_t MS_TRSM_QUAD_MAX_COLS = 8192;
constexpr int TQ_STRIDE = 18;  // doubles per (step, lane-of-quad) row of the R image: 16 + 2 keeps the four lanes'
                               // 16-byte loads in different banks
__global__ void __launch_bounds__ (256, 2)
trsm_quad_loop_kernel (const double *__restrict__ Rblk, size_t ldr, double *__restrict__ rows, size_t ld, int h,
                       size_t col_lo, size_t n, const double *__restrict__ dpiece, double *__restrict__ PB, size_t pb_H,
                       size_t pb_k0)
{
  // Rq[(i * 4 + t) * TQ_STRIDE + kk] = R(i, k), k = 4 (i / 4 + kk) + t, where i < k < h (k <= 63), 0 elsewhere
  __shared__ __align__ (16) double Rq[MS_SUB * 4 * TQ_STRIDE];
  __shared__ double rinv[MS_SUB], ds[MS_SUB];
  const int t = threadIdx.x & 3;
  const size_t j = col_lo + (size_t)blockIdx.x * 64 + (threadIdx.x >> 2);
  const bool live = j < n;
  double a[16];
#pragma unroll
  for (int kk = 0; kk < 16; kk++)
    {
      const int k = 4 * kk + t;
      a[kk] = (live && k < h) ? rows[(size_t)k * ld + j] : 0.0;
    }
  double rr[16];
#pragma unroll
  for (int u = 0; u < 16; u++)
    {
      const int e = threadIdx.x + 256 * u;
      const int i = e >> 6, rem = e & 63;
      const int k = 4 * (i >> 2) + rem;  // consecutive threads read consecutive entries of row i of R
      rr[u] = (k < MS_SUB && k > i && k < h) ? Rblk[(size_t)i * ldr + k] : 0.0;
    }
  double rdiag = 1.0, dval = 1.0;
  if (threadIdx.x < MS_SUB && (int)threadIdx.x < h)
    {
      rdiag = Rblk[(size_t)threadIdx.x * ldr + threadIdx.x];
      dval = dpiece[threadIdx.x];
    }
#pragma unroll
  for (int u = 0; u < 16; u++)
    {
      const int e = threadIdx.x + 256 * u;
      const int i = e >> 6, rem = e & 63;
      Rq[(i * 4 + (rem & 3)) * TQ_STRIDE + (rem >> 2)] = rr[u];
    }
  if (threadIdx.x < MS_SUB)
    {
      rinv[threadIdx.x] = 1.0 / rdiag;
      ds[threadIdx.x] = dval;
    }
  __syncthreads ();
  double *pcol = rows + j;
  double *ppb = PB + chunk_major_index (pb_k0, live ? j : col_lo, pb_H);
#pragma unroll 1
  for (int q = 0; q < 16; q++)
    {
#pragma unroll
      for (int o = 0; o < 4; o++)
        {
          const int i = 4 * q + o;
          double xi = a[0] * rinv[i];                    // meaningful on lane o of the quad only
          xi = __shfl_sync (0xffffffffu, xi, o, 4);
          if (t == o)
            a[0] = xi;                                   // a_i is final: it now holds x_i
          const double *rrow = &Rq[(i * 4 + t) * TQ_STRIDE];
#pragma unroll
          for (int kk = 0; kk < 16; kk += 2)
            {
              const double2 r2 = *reinterpret_cast<const double2 *> (rrow + kk);
              a[kk] = fma (-r2.x, xi, a[kk]);            // the image is 0 for k <= i: finished values stay
              a[kk + 1] = fma (-r2.y, xi, a[kk + 1]);
            }
        }
      const int k = 4 * q + t;
      if (live && k < h)
        {
          const double val = a[0] * ds[k];
          pcol[(size_t)k * ld] = val;
          ppb[(size_t)k * MS_CHUNK] = val;               // k-th row of the same 16-column chunk
        }
#pragma unroll
      for (int kk = 0; kk < 15; kk++)
        a[kk] = a[kk + 1];
      a[15] = 0.0;
    }
}

int
ms_k_trsm_panel (ms_ctx *c, const double *Rblk, size_t ldr, double *rows, size_t ld, int h, size_t col_lo, size_t n,
                 const double *dpiece, double *PB, size_t pb_H, size_t pb_k0, cudaStream_t s)
{
  if (col_lo >= n)
    return MS_OK;
  MS_CHECK_ARG (h >= 1 && h <= MS_SUB, "trsm panel piece must have 1..64 rows");
  // Two kernels, chosen by the number of columns: up to MS_TRSM_QUAD_MAX_COLS the launch is a handful of CTAs on an
  // otherwise waiting chain and the four-threads-per-column kernel is the shorter one (isolated 14.9 against 19.9 us;
  // n = 4096: 3.95 against 4.07 ms for the whole factorisation); on wide panels its 1024 FMAs per lane and four times
  // as many CTAs cost the concurrently running trailing update more than the shorter chain returns (n = 16384:
  // 53.1 against 52.1 ms, n = 65536: 2837 against 2815 ms; profiles/r02_trsm_ab.txt).  MS_TRSM_QUAD=0 / 1 forces one.
  static int quad = -1;
  if (quad < 0)
    {
      const char *e = getenv ("MS_TRSM_QUAD");
      quad = e ? (atoi (e) ? 1 : 0) : 2;
    }
  if (quad == 1 || (quad == 2 && n - col_lo <= MS_TRSM_QUAD_MAX_COLS))
    {
      const unsigned g = (unsigned)((n - col_lo + 63) / 64);
      trsm_quad_loop_kernel<<<g, 256, 0, s>>> (Rblk, ldr, rows, ld, h, col_lo, n, dpiece, PB, pb_H, pb_k0);
      LAUNCH_COUNT (c);
      MS_CUDA (cudaGetLastError ());
      return MS_OK;
    }
  unsigned grid = (unsigned)((n - col_lo + 127) / 128);
  const size_t smem = (size_t)MS_SUB * 128 * sizeof (double);
  static bool configured_dev[64] = {};
  bool &configured = configured_dev[c->device & 63];
  if (!configured)
    {
      MS_CUDA (cudaFuncSetAttribute (trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
  trsm_panel_kernel<<<grid, 128, smem, s>>> (Rblk, ldr, rows, ld, h, col_lo, n, dpiece, PB, pb_H, pb_k0);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// debug hook (not in the public header): average device time of one diag-kernel launch on a 64 x 64 SPD block
// (variant 0 = diag_kernel<5,4>, 1 = diag_strip_kernel), in microseconds
extern "C" int
ms_debug_time_diag (int variant, int h, int reps, double *us_out, long long *cycles5)
{
  // variant 0 = diag_kernel<5,4>, 1 = diag_strip_kernel, 2 = diag_strip_kernel with the update warps and the
  // barriers switched off (timing of the pivot warp alone; its results are wrong by construction)
  MS_CHECK_ARG (h >= 1 && h <= 64 && reps >= 1 && us_out, "bad argument");
  double *blk, *d, *src;
  int *flags;
  MS_CUDA (cudaMalloc (&blk, 64 * 64 * 8));
  MS_CUDA (cudaMalloc (&src, 64 * 64 * 8));
  MS_CUDA (cudaMalloc (&d, 64 * 8));
  MS_CUDA (cudaMalloc (&flags, 8 * sizeof (int)));
  MS_CUDA (cudaMemset (flags, 0, 8 * sizeof (int)));
  long long *cyc;
  MS_CUDA (cudaMalloc (&cyc, 64 * sizeof (long long)));
  MS_CUDA (cudaMemset (cyc, 0, 64 * sizeof (long long)));
  std::vector<double> a (64 * 64, 0.0);
  for (int i = 0; i < h; i++)
    for (int j = i; j < h; j++)
      a[i * 64 + j] = (double)(h - j) + (i == j ? h : 0);  // min-type SPD matrix plus a diagonal shift
  MS_CUDA (cudaMemcpy (src, a.data (), 64 * 64 * 8, cudaMemcpyHostToDevice));
  const size_t smem = (2 * (size_t)MS_SUB * (MS_SUB + 1) + 2 * MS_SUB) * sizeof (double);
  MS_CUDA (cudaFuncSetAttribute (diag_kernel<5, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  MS_CUDA (cudaFuncSetAttribute (diag_strip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  MS_CUDA (cudaFuncSetAttribute (diag_strip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t e0, e1;
  MS_CUDA (cudaEventCreate (&e0));
  MS_CUDA (cudaEventCreate (&e1));
  float total = 0;
  for (int r = 0; r < reps + 2; r++)
    {
      MS_CUDA (cudaMemcpy (blk, src, 64 * 64 * 8, cudaMemcpyDeviceToDevice));
      MS_CUDA (cudaEventRecord (e0));
      if (variant == 0)
        diag_kernel<5, 4><<<1, 256, smem>>> (blk, 64, h, 1e-300, d, nullptr, MS_SUB, nullptr, 0, blk, 64, h, flags);
      else if (variant == 3)
        diag_strip_kernel<true><<<1, 256, smem>>> (blk, 64, h, 1e-300, d, nullptr, MS_SUB, nullptr, 0, blk, 64, h, flags, cyc, 0);
      else
        diag_strip_kernel<false><<<1, 256, smem>>> (blk, 64, h, 1e-300, d, nullptr, MS_SUB, nullptr, 0, blk, 64, h, flags, cyc,
                                                    variant == 2);
      MS_CUDA (cudaEventRecord (e1));
      MS_CUDA (cudaEventSynchronize (e1));
      float ms;
      MS_CUDA (cudaEventElapsedTime (&ms, e0, e1));
      if (r >= 2)
        total += ms;
    }
  *us_out = total / reps * 1e3;
  if (cycles5)
    MS_CUDA (cudaMemcpy (cycles5, cyc, 40 * sizeof (long long), cudaMemcpyDeviceToHost));
  cudaFree (cyc);
  cudaEventDestroy (e0);
  cudaEventDestroy (e1);
  cudaFree (blk);
  cudaFree (src);
  cudaFree (d);
  cudaFree (flags);
  return MS_OK;
}
