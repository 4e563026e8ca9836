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
  cudaError_t e = cudaGetLastError ();
  if (e != cudaSuccess)
    printf ("CUDA error: %s\n", cudaGetErrorString (e));
  return 0;
}
