// lat_probe.cu — dependent-issue latencies (cycles per operation, one warp on one SM) of the operations on
// the critical path of the diagonal-block factorisation: DFMA / DMUL / DADD, the IEEE divide and square root
// (__ddiv_rn, __dsqrt_rn), the approximate rsqrt() / 1.0 / x, a 64-bit warp shuffle, a shared-memory round trip,
// a CTA barrier.  Measurement tool only (profiles/r02_latency_probe.txt); not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o matrix_solver_b200/bin/lat_probe tools/lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 512;

template <int OP>
__global__ void
probe (double *io, long long *cycles, int threads_active)
{
  __shared__ double sm[1024];
  double x = io[0] + threadIdx.x * 1e-9, a = io[1], b = io[2];
  sm[threadIdx.x] = x;
  __syncthreads ();
  long long t0 = clock64 ();
#pragma unroll 16
  for (int i = 0; i < N; i++)
    {
      if (OP == 0)
        x = fma (x, a, b);
      else if (OP == 1)
        x = __dmul_rn (x, a);
      else if (OP == 2)
        x = __dadd_rn (x, b);
      else if (OP == 3)
        x = __ddiv_rn (b, x);
      else if (OP == 4)
        x = __dsqrt_rn (x) + 1.0;   // one dependent DADD included
      else if (OP == 5)
        x = rsqrt (x) + 1.0;        // one dependent DADD included
      else if (OP == 6)
        x = __shfl_sync (0xffffffffu, x, (threadIdx.x + 1) & 31);
      else if (OP == 7)
        {
          sm[threadIdx.x] = x;
          __syncwarp ();
          x = sm[threadIdx.x ^ 1] + 1.0;
          __syncwarp ();
        }
      else if (OP == 8)
        {
          __syncthreads ();
          x += 1.0;
        }
      else if (OP == 9)
        x = 1.0 / x + 1.0;          // default-precision divide: same IEEE sequence unless -use_fast_math
      else if (OP == 10)
        {
          // Newton reciprocal square root from the hardware seed: the candidate fast pivot
          double y = rsqrt (x);
          x = x * y + 1.0;          // sqrt(x) ~ x * rsqrt(x); one dependent DADD included
        }
    }
  long long t1 = clock64 ();
  if (threadIdx.x == 0)
    cycles[OP] = t1 - t0;
  io[3 + OP] = x;
  (void)threads_active;
}

// ---- the pivot warp's strip factorisation (copy of strip_factor<true> in csrc/ms_kernels_diag.cu) taken
// apart: which piece of the per-pivot chain costs what.  PARTS bit 0: sqrt, bit 1: the two divides,
// bit 2: update of the later strip rows (shuffle + mul, mul, sub), bit 3: FAST pivot (rsqrt, multiplies)
template <int PARTS>
__global__ void
strip_probe (double *io, long long *cycles, int slot)
{
  const int lane = threadIdx.x;
  double c[8][2];
  for (int u = 0; u < 8; u++)
    {
      c[u][0] = io[0] + 0.01 * lane + u + (lane == u ? 40.0 : 0.0);
      c[u][1] = io[1] + 0.02 * lane + u;
    }
  const int i0 = 0;
  double sink = 0;
  long long t0 = clock64 ();
  for (int rep = 0; rep < 16; rep++)
    {
#pragma unroll
      for (int t = 0; t < 8; t++)
        {
          const int row = i0 + t;
          const int j0 = lane, j1 = lane + 32;
          const double s = __shfl_sync (0xffffffffu, c[t][0], row & 31);
          const double di = s > 0 ? 1.0 : -1.0;
          double rii = fabs (s), den;
          if (PARTS & 8)
            {
              const double y = rsqrt (fabs (s));
              rii = fabs (s) * y;
              den = y * di;  // multiplier
              if (j0 > row)
                c[t][0] *= den;
              if (j1 > row)
                c[t][1] *= den;
            }
          else
            {
              if (PARTS & 1)
                rii = __dsqrt_rn (fabs (s));
              den = __dmul_rn (rii, di);
              if (PARTS & 2)
                {
                  const double q0 = __ddiv_rn (j0 > row ? c[t][0] : 1.0, den);
                  if (j0 > row)
                    c[t][0] = q0;
                  const double q1 = __ddiv_rn (j1 > row ? c[t][1] : 1.0, den);
                  if (j1 > row)
                    c[t][1] = q1;
                }
            }
          sink += rii;
          if (PARTS & 4)
            {
#pragma unroll
              for (int u = t + 1; u < 8; u++)
                {
                  const double rk = __shfl_sync (0xffffffffu, c[t][0], (i0 + u) & 31);
                  c[u][0] = __dsub_rn (c[u][0], __dmul_rn (__dmul_rn (rk, c[t][0]), di));
                  c[u][1] = __dsub_rn (c[u][1], __dmul_rn (__dmul_rn (rk, c[t][1]), di));
                }
            }
        }
      // keep the values in a sane range for the next repetition
#pragma unroll
      for (int u = 0; u < 8; u++)
        {
          c[u][0] = fabs (c[u][0]) + 1.0 + (lane == u ? 40.0 : 0.0);
          c[u][1] = fabs (c[u][1]) + 1.0;
        }
    }
  long long t1 = clock64 ();
  if (threadIdx.x == 0)
    cycles[slot] = t1 - t0;
  double acc = sink;
  for (int u = 0; u < 8; u++)
    acc += c[u][0] + c[u][1];
  io[20 + slot] = acc;
}

// ---- the library's strip_factor verbatim (csrc/ms_kernels_diag.cu), once alone and once with seven more
// warps parked at a named barrier, as in diag_strip_kernel's first strip
template <bool LO>
__device__ __forceinline__ void
strip_factor (double (&c)[8][2], int i0, int h, int lane, double thresh, double *rdiag, double *dd, int *flags)
{
  const int j0 = lane, j1 = lane + 32;
#pragma unroll
  for (int t = 0; t < 8; t++)
    {
      const int row = i0 + t;
      if (row < h)
        {
          const double s = __shfl_sync (0xffffffffu, LO ? c[t][0] : c[t][1], row & 31);
          const double di = s > 0 ? 1.0 : -1.0;
          const double rii = __dsqrt_rn (fabs (s));
          const double den = __dmul_rn (rii, di);
          if (lane == (row & 31))
            {
              rdiag[row] = rii;
              dd[row] = di;
              if (fabs (rii) < thresh)
                flags[0] = 1;
            }
          // The compiler evaluates the quotient for every lane and selects afterwards; a lane left of the
          // pivot holds 0.0, and a zero numerator fails the fast-path test of the IEEE divide sequence
          // (|num| >= 2^-120), which sends the WHOLE warp through the ~100-instruction slow path twice per
          // pivot (measured: 1400 instead of ~450 cycles per pivot).  Masked lanes divide 1.0 instead.
          if (LO)
            {
              const double q0 = __ddiv_rn (j0 > row ? c[t][0] : 1.0, den);
              if (j0 > row)
                c[t][0] = q0;
            }
          {
            const double q1 = __ddiv_rn (j1 > row ? c[t][1] : 1.0, den);
            if (j1 > row)
              c[t][1] = q1;
          }
#pragma unroll
          for (int u = t + 1; u < 8; u++)
            {
              // r(row, k), k = i0 + u, lives in the lane that owns column k
              const double rk = __shfl_sync (0xffffffffu, LO ? c[t][0] : c[t][1], (i0 + u) & 31);
              if (LO)
                c[u][0] = __dsub_rn (c[u][0], __dmul_rn (__dmul_rn (rk, c[t][0]), di));
              c[u][1] = __dsub_rn (c[u][1], __dmul_rn (__dmul_rn (rk, c[t][1]), di));
            }
        }
    }
}


__global__ void
strip_probe_lib (double *io, long long *cycles, int slot, int h, double thresh, int *flags)
{
  __shared__ double rdiag[64], dd[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0)
    {
      double c[8][2];
      for (int u = 0; u < 8; u++)
        {
          c[u][0] = lane >= u ? io[0] + 0.01 * lane + u + (lane == u ? 40.0 : 0.0) : 0.0;
          c[u][1] = io[1] + 0.02 * lane + u;
        }
      long long t0 = clock64 ();
      strip_factor<true> (c, 0, h, lane, thresh, rdiag, dd, flags);
      long long t1 = clock64 ();
      if (lane == 0)
        cycles[slot] = t1 - t0;
      double acc = 0;
      for (int u = 0; u < 8; u++)
        acc += c[u][0] + c[u][1];
      io[30 + slot] = acc + rdiag[lane] + dd[lane];
      if (blockDim.x > 32)
        asm volatile ("bar.arrive 2, 256;" ::: "memory");
    }
  else
    asm volatile ("bar.sync 2, 256;" ::: "memory");
}

int
main ()
{
  double *io;
  long long *cyc;
  cudaMalloc (&io, 64 * sizeof (double));
  cudaMalloc (&cyc, 16 * sizeof (long long));
  double h[3] = {1.25, 0.999999, 1.5};
  cudaMemcpy (io, h, sizeof (h), cudaMemcpyHostToDevice);
  const char *names[] = {"DFMA", "DMUL", "DADD", "__ddiv_rn", "__dsqrt_rn + DADD", "rsqrt() + DADD", "SHFL (64-bit)",
                         "shared st + ld + DADD (warp)", "__syncthreads (256 thr) + DADD", "1.0/x + DADD",
                         "x*rsqrt(x) + DADD"};
  for (int rep = 0; rep < 2; rep++)
    {
      probe<0><<<1, 32>>> (io, cyc, 32);
      probe<1><<<1, 32>>> (io, cyc, 32);
      probe<2><<<1, 32>>> (io, cyc, 32);
      probe<3><<<1, 32>>> (io, cyc, 32);
      probe<4><<<1, 32>>> (io, cyc, 32);
      probe<5><<<1, 32>>> (io, cyc, 32);
      probe<6><<<1, 32>>> (io, cyc, 32);
      probe<7><<<1, 32>>> (io, cyc, 32);
      probe<8><<<1, 256>>> (io, cyc, 256);
      probe<9><<<1, 32>>> (io, cyc, 32);
      probe<10><<<1, 32>>> (io, cyc, 32);
      cudaDeviceSynchronize ();
    }
  long long hc[16];
  cudaMemcpy (hc, cyc, sizeof (hc), cudaMemcpyDeviceToHost);
  printf ("# dependent chains of %d operations, one warp, B200 (cycles per operation)\n", N);
  for (int i = 0; i < 11; i++)
    printf ("%-34s %8.1f\n", names[i], (double)hc[i] / N);
  strip_probe<7><<<1, 32>>> (io, cyc, 0);
  strip_probe<3><<<1, 32>>> (io, cyc, 1);
  strip_probe<1><<<1, 32>>> (io, cyc, 2);
  strip_probe<2><<<1, 32>>> (io, cyc, 3);
  strip_probe<4><<<1, 32>>> (io, cyc, 4);
  strip_probe<12><<<1, 32>>> (io, cyc, 5);
  strip_probe<8><<<1, 32>>> (io, cyc, 6);
  cudaDeviceSynchronize ();
  cudaMemcpy (hc, cyc, sizeof (hc), cudaMemcpyDeviceToHost);
  const char *sn[] = {"exact: sqrt + 2 div + row updates", "exact: sqrt + 2 div", "sqrt only", "2 div only", "row updates only",
                      "fast: rsqrt, multiplies + row updates", "fast: rsqrt, multiplies"};
  printf ("# strip factorisation of the pivot warp (8 rows in registers), cycles per PIVOT, one warp alone on an SM\n");
  for (int i = 0; i < 7; i++)
    printf ("%-42s %8.1f\n", sn[i], (double)hc[i] / (16 * 8));
  int *flags;
  cudaMalloc (&flags, 32);
  cudaMemset (flags, 0, 32);
  for (int rep = 0; rep < 2; rep++)
    {
      strip_probe_lib<<<1, 32>>> (io, cyc, 8, 64, 1e-300, flags);
      strip_probe_lib<<<1, 256>>> (io, cyc, 9, 64, 1e-300, flags);
      cudaDeviceSynchronize ();
    }
  cudaMemcpy (hc, cyc, sizeof (hc), cudaMemcpyDeviceToHost);
  printf ("library strip_factor<true>, one strip, cycles per pivot: alone %.1f, with 7 warps parked at a barrier %.1f\n",
          (double)hc[8] / 8, (double)hc[9] / 8);
  cudaError_t e = cudaGetLastError ();
  if (e != cudaSuccess)
    printf ("CUDA error: %s\n", cudaGetErrorString (e));
  return 0;
}
