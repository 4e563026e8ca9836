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
#include <cooperative_groups.h>

#include <algorithm>

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

// ============================================================================ single-launch sweeps
// One cooperative launch per sweep (single GPU, m % 64 == 0): G co-resident CTAs, 64-wide chunk c of
// the unknowns belongs to CTA c % G, which keeps that chunk's accumulator in shared memory.
//   forward, piece t (rows 64t..): the owner of chunk t solves R_tt^T y_t = b_t - acc_t, publishes
//   y_t (global memory + release flag); every CTA then applies rows t to the chunks it owns (> t).
//   The chunk on the critical path (t+1) is updated and solved FIRST, before the CTA's other chunks.
//   backward is the mirror image over row chunks (x_t published, rows above updated).
// Each element of R is read once, 512-byte coalesced segments, 16 loads in flight per thread.
// Flags are epochs, never reset.  All CTAs spin on flags written by other CTAs of the SAME launch,
// which is why the launch is cooperative (co-residency guaranteed).
namespace cg = cooperative_groups;

constexpr int SWEEP_MAX_SLOTS = 24;  // chunks one CTA may own: ceil(T / G)

struct sweep_args
{
  const double *A;
  size_t ld, n;
  const double *b;     // forward: b ; backward: y
  const double *d;     // backward only
  double *out;         // forward: y ; backward: x
  int *ready;          // T flags
  int epoch;
  int T;               // number of 64-row pieces
  double thresh;
  int *flags;
};

__device__ __forceinline__ void
sweep_wait (const int *flag, int epoch)
{
  if (threadIdx.x == 0)
    {
      int v;
      do
        {
          asm volatile ("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
          if (v != epoch)
            __nanosleep (40);
        }
      while (v != epoch);
    }
  __syncthreads ();
}

__device__ __forceinline__ void
sweep_publish (int *flag, int epoch)
{
  __threadfence ();
  __syncthreads ();
  if (threadIdx.x == 0)
    asm volatile ("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}

// 64 x 64 block update shared by both sweeps, split in two so that the loads can be issued before
// the vector they are multiplied with is known (critical path: the neighbour chunk).
// FORWARD : accs[j] += sum_i A(r0 + i, c0 + j) * v[i]   (column sums; thread = column x 16-row group)
// BACKWARD: accs[i] += sum_j A(r0 + i, c0 + j) * v[j]   (row sums;    thread = row x 16-col group)
template <bool FORWARD>
__device__ __forceinline__ void
sweep_block_load (const sweep_args &sa, size_t r0, int hr, size_t c0, int hc, double (&t)[16])
{
  if (FORWARD)
    {
      const int cx = threadIdx.x & 63, rg = threadIdx.x >> 6;
      const double *col = sa.A + (r0 + (size_t)rg * 16) * sa.ld + c0 + cx;
#pragma unroll
      for (int i = 0; i < 16; i++)
        t[i] = (cx < hc && rg * 16 + i < hr) ? __ldcs (col + (size_t)i * sa.ld) : 0.0;
    }
  else
    {
      const int row = threadIdx.x >> 2, cgp = threadIdx.x & 3;
      const double *rp = sa.A + (r0 + row) * sa.ld + c0 + cgp * 16;
#pragma unroll
      for (int j = 0; j < 16; j++)
        t[j] = (row < hr && cgp * 16 + j < hc) ? __ldcs (rp + j) : 0.0;
    }
}

template <bool FORWARD>
__device__ __forceinline__ void
sweep_block_apply (const double (&t)[16], const double *v, double *accs, double (*part)[64])
{
  double s = 0.0;
  if (FORWARD)
    {
      const int cx = threadIdx.x & 63, rg = threadIdx.x >> 6;
#pragma unroll
      for (int i = 0; i < 16; i++)
        s += t[i] * v[rg * 16 + i];
      part[rg][cx] = s;
      __syncthreads ();
      if (threadIdx.x < 64)
        accs[threadIdx.x] += (part[0][threadIdx.x] + part[1][threadIdx.x]) + (part[2][threadIdx.x] + part[3][threadIdx.x]);
      __syncthreads ();
    }
  else
    {
      const int row = threadIdx.x >> 2, cgp = threadIdx.x & 3;
#pragma unroll
      for (int j = 0; j < 16; j++)
        s += t[j] * v[cgp * 16 + j];
      s += __shfl_xor_sync (0xffffffffu, s, 1);
      s += __shfl_xor_sync (0xffffffffu, s, 2);
      if (cgp == 0)
        accs[row] += s;
      __syncthreads ();
    }
}

template <bool FORWARD>
__device__ __forceinline__ void
sweep_block_update (const sweep_args &sa, size_t r0, int hr, size_t c0, int hc, const double *v, double *accs,
                    double (*part)[64])
{
  double t[16];
  sweep_block_load<FORWARD> (sa, r0, hr, c0, hc, t);
  sweep_block_apply<FORWARD> (t, v, accs, part);
}

template <bool FORWARD>
__global__ void __launch_bounds__ (256)
sweep_kernel (const sweep_args sa)
{
  __shared__ double Rs[64 * 65];
  __shared__ double vs[64];               // y_t resp. x_t of the current piece
  __shared__ double accs[SWEEP_MAX_SLOTS][64];
  __shared__ double part[4][64];
  const int G = gridDim.x, me = blockIdx.x, T = sa.T;
  for (int e = threadIdx.x; e < SWEEP_MAX_SLOTS * 64; e += 256)
    (&accs[0][0])[e] = 0.0;
  __syncthreads ();

  auto piece_h = [&] (int t) { return (int)min ((size_t)64, sa.n - (size_t)t * 64); };
  auto slot_of = [&] (int c) { return c / G; };

  // diagonal block of piece t -> Rs: 16 independent loads per thread, then the stores
  auto load_diag = [&] (int t) {
    const int h = piece_h (t);
    const size_t r0 = (size_t)t * 64;
    double r[16];
#pragma unroll
    for (int u = 0; u < 16; u++)
      {
        const int e = threadIdx.x + 256 * u, i = e >> 6, j = e & 63;
        r[u] = (i < h && j < h && j >= i) ? sa.A[(r0 + i) * sa.ld + r0 + j] : 0.0;
      }
#pragma unroll
    for (int u = 0; u < 16; u++)
      {
        const int e = threadIdx.x + 256 * u, i = e >> 6, j = e & 63;
        Rs[i * 65 + j] = r[u];
      }
  };

  // solve piece t (this CTA owns chunk t and its accumulator is complete) and publish the result
  auto solve_and_publish = [&] (int t, bool diag_loaded) {
    const int h = piece_h (t);
    const size_t r0 = (size_t)t * 64;
    if (!diag_loaded)
      load_diag (t);
    if (threadIdx.x < 64)
      {
        double rhs = 0.0;
        if (threadIdx.x < h)
          rhs = FORWARD ? sa.b[r0 + threadIdx.x] - accs[slot_of (t)][threadIdx.x]
                        : sa.d[r0 + threadIdx.x] * sa.b[r0 + threadIdx.x] - accs[slot_of (t)][threadIdx.x];
        vs[threadIdx.x] = rhs;
      }
    __syncthreads ();
    if (threadIdx.x < 32)
      warp_tri_solve<FORWARD> (Rs, vs, h, sa.thresh, sa.flags);
    __syncthreads ();
    if (threadIdx.x < h)
      sa.out[r0 + threadIdx.x] = vs[threadIdx.x];
    sweep_publish (sa.ready + t, sa.epoch);
  };

  bool solved_ahead = false;  // the piece of this iteration was already solved in the previous one
  for (int step = 0; step < T; step++)
    {
      const int t = FORWARD ? step : T - 1 - step;
      const int h = piece_h (t);
      const size_t r0 = (size_t)t * 64;
      const int tn = FORWARD ? t + 1 : t - 1;
      const bool own_next = tn >= 0 && tn < T && tn % G == me;
      const bool ahead = own_next && G > 1;  // this CTA is on the critical path of piece tn
      double pre[16];
      if (ahead)
        {
          // everything that does not depend on piece t's solution is fetched before we wait for it
          load_diag (tn);
          if (FORWARD)
            sweep_block_load<true> (sa, r0, h, (size_t)tn * 64, piece_h (tn), pre);
          else
            sweep_block_load<false> (sa, (size_t)tn * 64, piece_h (tn), r0, h, pre);
        }
      if (t % G == me)
        {
          if (!solved_ahead)
            solve_and_publish (t, false);
          // vs holds the solution of piece t
        }
      else
        {
          sweep_wait (sa.ready + t, sa.epoch);
          if (threadIdx.x < 64)
            vs[threadIdx.x] = threadIdx.x < h ? __ldcg (sa.out + r0 + threadIdx.x) : 0.0;
          __syncthreads ();
        }
      solved_ahead = false;
      if (ahead)
        {
          // neighbour chunk first: finish its accumulator, solve it, publish; then catch up
          sweep_block_apply<FORWARD> (pre, vs, accs[slot_of (tn)], part);
          double vkeep = 0.0, vnext = 0.0;
          if (threadIdx.x < 64)
            vkeep = vs[threadIdx.x];
          __syncthreads ();
          solve_and_publish (tn, true);
          solved_ahead = true;
          if (threadIdx.x < 64)
            {
              vnext = vs[threadIdx.x];
              vs[threadIdx.x] = vkeep;
            }
          __syncthreads ();
          if (FORWARD)
            for (int c = tn + G; c < T; c += G)
              sweep_block_update<true> (sa, r0, h, (size_t)c * 64, piece_h (c), vs, accs[slot_of (c)], part);
          else
            for (int c = tn - G; c >= 0; c -= G)
              sweep_block_update<false> (sa, (size_t)c * 64, 64, r0, h, vs, accs[slot_of (c)], part);
          if (threadIdx.x < 64)
            vs[threadIdx.x] = vnext;
          __syncthreads ();
        }
      else if (FORWARD)
        {
          int c = G == 1 ? t + 1 : t + 1 + ((me - (t + 1)) % G + G) % G;  // first chunk > t owned by me
          for (; c < T; c += G)
            sweep_block_update<true> (sa, r0, h, (size_t)c * 64, piece_h (c), vs, accs[slot_of (c)], part);
        }
      else
        {
          int c = G == 1 ? t - 1 : t - 1 - (((t - 1) - me) % G + G) % G;  // last chunk < t owned by me
          for (; c >= 0; c -= G)
            sweep_block_update<false> (sa, (size_t)c * 64, 64, r0, h, vs, accs[slot_of (c)], part);
        }
    }
}

// ============================================================================ backward sweep, any rank count
// x = R^{-1} D y in ONE launch per rank (compute_x, src/solver.cpp:210-245, without its n gathers).
// Rank p holds block rows I = p, p + P, ...; its local 64-row chunk lc is global piece
// gp(lc) = ((lc / q) P + p) q + lc % q  (q = m / 64 pieces per block row).  Chunk lc belongs to CTA lc % G,
// which keeps that chunk's accumulator  acc = sum_{t > gp} R(chunk, t) x_t  in shared memory.
// For t = T-1 .. 0:
//   * the CTA that owns piece t (on the rank that owns its block row) solves R_tt x_t = D y_t - acc_t,
//     stores x_t into its own x AND into every peer's x (CUDA-IPC mapped pointers: plain stores over
//     NVLink), fences, and raises ready[t] = epoch in every rank's flag array;
//   * every CTA of every rank waits for its LOCAL ready[t], reads x_t from its LOCAL x and applies the
//     64 x 64 blocks R(chunk, t) of the chunks it owns above piece t.
// The chunk on the critical path (piece t-1) is handled first and everything that does not depend on
// x_t — its diagonal block and the block R(t-1, t) — is loaded before the CTA waits.
// No NCCL call and no second launch: at P = 8 the 2 x N_blocks collectives of round 1 (42 ms) are gone.
// Memory access: a row segment of a block is 512 contiguous bytes; thread (row = tid / 4, g = tid % 4)
// loads double2 at columns 8 j + 2 g, j = 0..7, so four neighbouring lanes read 64 contiguous bytes and a
// warp instruction covers 8 rows x 64 B in full sectors; the x_t operands are read as 4 distinct 16-byte
// words per instruction (broadcast) — no bank conflicts (round 1: 67 % conflict wavefronts, 32 sectors per
// request).
// Every wait is bounded (spin_limit clock ticks): a peer that never arrives (crashed rank) makes the
// kernel give up and report through flags[3] instead of hanging the GPU.
constexpr int BSWEEP_MAX_PEERS = 16;

struct bsweep_args
{
  const double *A;
  size_t ld, n;
  const double *y, *d;
  const double *V;  // inverses of this rank's 64 x 64 diagonal pieces, [local chunk][64][64] (bsweep_invert_kernel)
  double *x;       // local x: written by this rank's owners and by the peers' owners
  int *ready;      // local flags, T ints
  double *peer_x[BSWEEP_MAX_PEERS];
  int *peer_ready[BSWEEP_MAX_PEERS];
  int rank, nranks, q;
  int epoch, T, LT;
  double thresh;
  int *flags;
  long long spin_limit;
};

__device__ __forceinline__ bool
bsweep_wait (const int *flag, int epoch, bool sys, long long limit, bool &gave_up, int *flags)
{
  int timed_out = 0;
  if (threadIdx.x == 0 && !gave_up)
    {
      const long long t0 = clock64 ();
      int v;
      for (;;)
        {
          if (sys)
            asm volatile ("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
          else
            asm volatile ("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
          if (v == epoch)
            break;
          if (clock64 () - t0 > limit)
            {
              timed_out = 1;
              flags[3] = 1;
              break;
            }
          __nanosleep (32);
        }
    }
  if (__syncthreads_or (timed_out))
    gave_up = true;  // a peer never arrived: finish the loop without waiting again; the host reports it
  return !gave_up;
}

// 64 x 64 block R(rows lr0.., cols c0..c0+hc): thread (row, g) -> t[2j], t[2j+1] = columns 8j + 2g, +1
__device__ __forceinline__ void
bsweep_block_load (const bsweep_args &sa, size_t lr0, size_t c0, int hc, double (&t)[16])
{
  const int row = threadIdx.x >> 2, g = threadIdx.x & 3;
  const double *rp = sa.A + (lr0 + row) * sa.ld + c0 + 2 * g;
#pragma unroll
  for (int j = 0; j < 8; j++)
    {
      const int col = 8 * j + 2 * g;
      double2 v = make_double2 (0.0, 0.0);
      if (col + 1 < hc)
        v = __ldcs (reinterpret_cast<const double2 *> (rp + 8 * j));
      else if (col < hc)
        v.x = __ldcs (rp + 8 * j);
      t[2 * j] = v.x;
      t[2 * j + 1] = v.y;
    }
}

__device__ __forceinline__ void
bsweep_block_apply (const double (&t)[16], const double *v, double *accs)
{
  const int row = threadIdx.x >> 2, g = threadIdx.x & 3;
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int j = 0; j < 8; j++)
    {
      const double2 xv = *reinterpret_cast<const double2 *> (v + 8 * j + 2 * g);
      s0 = fma (t[2 * j], xv.x, s0);
      s1 = fma (t[2 * j + 1], xv.y, s1);
    }
  double s = s0 + s1;
  s += __shfl_xor_sync (0xffffffffu, s, 1);
  s += __shfl_xor_sync (0xffffffffu, s, 2);
  if (g == 0)
    accs[row] += s;
}

// The 64 x 64 triangles are inverted once, before the sweep (one CTA per local piece, all in parallel): the
// per-piece solve on the sweep's critical path is then a matrix-vector product by the whole CTA (~300 cycles)
// instead of 64 dependent shuffle + FMA steps of one warp (~2500 cycles on each of the n / 64 links of the
// chain).  Same recurrence as reverse_upper (src/matrix_op.cpp:244-285); the pivot test of compute_x
// (|R_ii| < EPS * norm, src/solver.cpp:232-236) happens here.
__global__ void __launch_bounds__ (64)
bsweep_invert_kernel (const double *__restrict__ A, size_t ld, size_t n, size_t m, int rank, int nranks,
                      double *__restrict__ V, double thresh, int *flags)
{
  extern __shared__ double inv_sm[];  // a[64][65] | v[64][65]
  double *a = inv_sm, *v = inv_sm + 64 * 65;
  const size_t lr0 = (size_t)blockIdx.x * 64;
  const size_t lb = lr0 / m;
  const size_t g0 = (lb * (size_t)nranks + (size_t)rank) * m + (lr0 - lb * m);
  const int h = g0 < n ? (int)min ((size_t)64, n - g0) : 0;
  const int tid = threadIdx.x;
  for (int e = tid; e < 64 * 64; e += 64)
    {
      const int i = e >> 6, j = e & 63;
      a[i * 65 + j] = (i < h && j < h && j >= i) ? A[(lr0 + i) * ld + g0 + j] : 0.0;
      v[i * 65 + j] = 0.0;
    }
  __syncthreads ();
  if (tid < h)
    {
      const int j = tid;
      if (fabs (a[j * 65 + j]) < thresh)
        flags[0] = 1;
      for (int i = j; i >= 0; i--)
        {
          double val;
          if (i == j)
            val = 1.0 / a[i * 65 + i];
          else
            {
              double sum = 0;
              for (int k = i + 1; k <= j; k++)
                sum = fma (a[i * 65 + k], v[k * 65 + j], sum);
              val = -sum / a[i * 65 + i];
            }
          v[i * 65 + j] = val;
        }
    }
  __syncthreads ();
  double *out = V + (size_t)blockIdx.x * 64 * 64;
  for (int e = tid; e < 64 * 64; e += 64)
    out[e] = v[(e >> 6) * 65 + (e & 63)];
}

constexpr int BS_RS = 68;  // row stride of the inverse in shared memory: 4 (mod 16) keeps the product's loads at the
                           // unavoidable two wavefronts per request

__global__ void __launch_bounds__ (256, 2)
bsweep_kernel (const bsweep_args sa)
{
  __shared__ double Rs[64 * BS_RS];
  __shared__ __align__ (16) double vs[64];
  __shared__ double rhs_s[64];
  __shared__ double accs[SWEEP_MAX_SLOTS][64];
  const int G = gridDim.x, me = blockIdx.x, T = sa.T, P = sa.nranks, q = sa.q;
  const bool sys = P > 1;
  bool gave_up = false;
  for (int e = threadIdx.x; e < SWEEP_MAX_SLOTS * 64; e += 256)
    (&accs[0][0])[e] = 0.0;
  __syncthreads ();

  auto piece_h = [&] (int t) { return (int)min ((size_t)64, sa.n - (size_t)t * 64); };
  auto owner_of = [&] (int t) { return (t / q) % P; };
  auto lidx_of = [&] (int t) { return ((t / q) / P) * q + t % q; };  // local chunk of an OWNED piece
  // local chunks whose global piece is < t
  auto lcut = [&] (int t) {
    const int I = t / q;
    if (I % P == sa.rank)
      return (I / P) * q + t % q;
    return (I / P + (I % P > sa.rank ? 1 : 0)) * q;
  };

  // inverse of the triangle of (owned) piece t -> Rs: 16 independent, fully coalesced loads per thread
  auto load_diag = [&] (int t) {
    const double *src = sa.V + (size_t)lidx_of (t) * 64 * 64;
    double r[16];
#pragma unroll
    for (int u = 0; u < 16; u++)
      r[u] = src[threadIdx.x + 256 * u];
#pragma unroll
    for (int u = 0; u < 16; u++)
      {
        const int e = threadIdx.x + 256 * u;
        Rs[(e >> 6) * BS_RS + (e & 63)] = r[u];
      }
  };

  // this CTA owns piece t and its accumulator is complete: solve, store x_t everywhere, raise the flags
  // dy: d_i y_i of the piece's rows (threads 0..63), fetched by the caller before it started to wait
  auto solve_and_publish = [&] (int t, bool diag_loaded, double dy) {
    const int h = piece_h (t);
    const size_t r0 = (size_t)t * 64;
    if (!diag_loaded)
      load_diag (t);
    if (threadIdx.x < 64)
      rhs_s[threadIdx.x] = threadIdx.x < h ? dy - accs[lidx_of (t) / G][threadIdx.x] : 0.0;
    __syncthreads ();
    {
      // x_t = V_tt rhs: thread (row, g) takes columns 4 j + g; the zeros below the diagonal do no harm
      const int row = threadIdx.x >> 2, g = threadIdx.x & 3;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int j = 0; j < 16; j += 2)
        {
          s0 = fma (Rs[row * BS_RS + 4 * j + g], rhs_s[4 * j + g], s0);
          s1 = fma (Rs[row * BS_RS + 4 * j + 4 + g], rhs_s[4 * j + 4 + g], s1);
        }
      double sx = s0 + s1;
      sx += __shfl_xor_sync (0xffffffffu, sx, 1);
      sx += __shfl_xor_sync (0xffffffffu, sx, 2);
      if (g == 0)
        vs[row] = row < h ? sx : 0.0;
    }
    __syncthreads ();
    // x_t: own copy by threads 0..63, one peer per further group of 64 threads
    for (int e = threadIdx.x; e < 64 * P; e += 256)
      {
        const int dst = e >> 6, i = e & 63;
        if (i < h)
          {
            double *xp = dst == sa.rank ? sa.x : sa.peer_x[dst];
            xp[r0 + i] = vs[i];
          }
      }
    if (sys)
      __threadfence_system ();
    else
      __threadfence ();
    __syncthreads ();
    if ((int)threadIdx.x < P)
      {
        int *fp = ((int)threadIdx.x == sa.rank ? sa.ready : sa.peer_ready[threadIdx.x]) + t;
        if (sys)
          asm volatile ("st.release.sys.global.s32 [%0], %1;" ::"l"(fp), "r"(sa.epoch) : "memory");
        else
          asm volatile ("st.release.gpu.global.s32 [%0], %1;" ::"l"(fp), "r"(sa.epoch) : "memory");
      }
  };

  bool solved_ahead = false;  // the piece of this iteration was already solved in the previous one
  for (int t = T - 1; t >= 0; t--)
    {
      const int h = piece_h (t);
      const size_t r0 = (size_t)t * 64;
      const int tn = t - 1;
      const bool own_next = tn >= 0 && owner_of (tn) == sa.rank && lidx_of (tn) % G == me;
      const bool mine = owner_of (t) == sa.rank && lidx_of (t) % G == me;
      // this CTA is on the critical path of piece tn; with a single CTA per rank both pieces can be its
      // own, and then piece t needs the shared triangle first
      const bool ahead = own_next && !(mine && !solved_ahead);
      int lc = lcut (t) - 1;            // last local chunk above piece t ...
      if (lc >= 0)
        lc -= ((lc - me) % G + G) % G;  // ... that this CTA owns (== the chunk of piece tn when `ahead`)
      // Everything that does not depend on x_t is fetched BEFORE the CTA waits for it: the blocks
      // R(chunk, t) of its first two chunks (with 4 x 148 chunks per step in flight the HBM latency hides
      // behind the wait, not behind the 64 FMAs), and for the CTA on the critical path also the triangle
      // of piece tn.
      double buf0[16], buf1[16];
      // (the CTA on the critical path keeps only its critical block across the solve: registers)
      const int lc0 = lc, lc1 = (!ahead && lc >= G) ? lc - G : -1;
      double dy = 0.0;  // d y of the piece this CTA solves in this iteration (threads 0..63)
      if (ahead)
        {
          load_diag (tn);
          if (threadIdx.x < 64 && (int)threadIdx.x < piece_h (tn))
            dy = sa.d[(size_t)tn * 64 + threadIdx.x] * sa.y[(size_t)tn * 64 + threadIdx.x];
        }
      else if (mine && !solved_ahead && threadIdx.x < 64 && (int)threadIdx.x < h)
        dy = sa.d[r0 + threadIdx.x] * sa.y[r0 + threadIdx.x];
      if (lc0 >= 0)
        bsweep_block_load (sa, (size_t)lc0 * 64, r0, h, buf0);
      if (lc1 >= 0)
        bsweep_block_load (sa, (size_t)lc1 * 64, r0, h, buf1);
      if (mine)
        {
          if (!solved_ahead)
            solve_and_publish (t, false, dy);  // vs = x_t afterwards
        }
      else
        {
          bsweep_wait (sa.ready + t, sa.epoch, sys, sa.spin_limit, gave_up, sa.flags);
          if (threadIdx.x < 64)
            vs[threadIdx.x] = threadIdx.x < h ? __ldcg (sa.x + r0 + threadIdx.x) : 0.0;
          __syncthreads ();
        }
      solved_ahead = false;
      if (lc0 >= 0)
        bsweep_block_apply (buf0, vs, accs[lc0 / G]);
      if (ahead)
        {
          // neighbour piece first: its accumulator is complete now — solve it, publish, then catch up
          double vkeep = 0.0, vnext = 0.0;
          if (threadIdx.x < 64)
            vkeep = vs[threadIdx.x];
          __syncthreads ();  // accumulator of tn complete, vs saved
          solve_and_publish (tn, true, dy);
          solved_ahead = true;
          if (threadIdx.x < 64)
            {
              vnext = vs[threadIdx.x];
              vs[threadIdx.x] = vkeep;
            }
          __syncthreads ();
          for (lc = lc0 - G; lc >= 0; lc -= G)
            {
              bsweep_block_load (sa, (size_t)lc * 64, r0, h, buf0);
              bsweep_block_apply (buf0, vs, accs[lc / G]);
            }
          __syncthreads ();  // everybody is done with x_t in vs
          if (threadIdx.x < 64)
            vs[threadIdx.x] = vnext;
        }
      else
        {
          if (lc1 >= 0)
            bsweep_block_apply (buf1, vs, accs[lc1 / G]);
          for (lc = lc1 - G; lc >= 0; lc -= G)
            {
              bsweep_block_load (sa, (size_t)lc * 64, r0, h, buf0);
              bsweep_block_apply (buf0, vs, accs[lc / G]);
            }
        }
      __syncthreads ();  // accumulators and vs settled before the next piece rewrites them
    }
}

// returns MS_OK and sets *done = 1 when the single-launch sweep ran, *done = 0 when the caller has to
// fall back to the per-piece launches (m % 64 != 0, too many chunks per CTA; forward: several ranks —
// there the forward solve normally rides along with the factorisation; backward with several ranks:
// peer memory not mapped)
double *ms_p2p_peer_x (const ms_ctx *c, int r);
int *ms_p2p_peer_sweep_ready (const ms_ctx *c, int r);

int
ms_k_sweep (ms_ctx *c, bool forward, double thresh, cudaStream_t s, int *done)
{
  *done = 0;
  if (c->m % MS_SUB != 0 || c->no_sweep)
    return MS_OK;
  const int T = (int)((c->n + 63) / 64);
  static int max_fwd = -1, max_bwd = -1;
  if (max_fwd < 0)
    {
      int a = 0, b = 0;
      MS_CUDA (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&a, sweep_kernel<true>, 256, 0));
      MS_CUDA (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&b, bsweep_kernel, 256, 0));
      max_fwd = std::max (1, std::min (a, 4));
      max_bwd = std::max (1, std::min (b, 4));
    }
  if (forward)
    {
      if (c->nranks != 1)
        return MS_OK;
      int G = std::min (T, max_fwd * c->sm_count);
      if ((T + G - 1) / G > SWEEP_MAX_SLOTS)
        return MS_OK;
      sweep_args sa;
      sa.A = c->A;
      sa.ld = c->ld;
      sa.n = c->n;
      sa.b = c->vec[MS_VEC_B];
      sa.d = c->vec[MS_VEC_D];
      sa.out = c->vec[MS_VEC_Y];
      sa.ready = c->sweep_ready;
      sa.epoch = ++c->sweep_epoch;
      sa.T = T;
      sa.thresh = thresh;
      sa.flags = c->flags;
      void *args[] = {(void *)&sa};
      MS_CUDA (cudaLaunchCooperativeKernel ((void *)sweep_kernel<true>, dim3 (G), dim3 (256), args, 0, s));
      LAUNCH_COUNT (c);
      *done = 1;
      return MS_OK;
    }
  // backward: every rank launches its own kernel; the kernels talk through peer memory
  if (c->nranks > BSWEEP_MAX_PEERS)
    return MS_OK;
  bsweep_args sa = {};
  for (int r = 0; r < c->nranks; r++)
    if (r != c->rank)
      {
        sa.peer_x[r] = ms_p2p_peer_x (c, r);
        sa.peer_ready[r] = ms_p2p_peer_sweep_ready (c, r);
        if (!sa.peer_x[r] || !sa.peer_ready[r])
          return MS_OK;  // transport off on ALL ranks (they agreed at ms_create): per-piece path
      }
  const int LT = (int)(c->rows_local / 64);  // m % 64 == 0: only the last global piece can be short ...
  const int LT_all = (int)((c->rows_local + 63) / 64);  // ... and it is a chunk of its owner
  (void)LT;
  int G = std::max (1, std::min (LT_all, max_bwd * c->sm_count));
  if ((LT_all + G - 1) / G > SWEEP_MAX_SLOTS)
    return MS_OK;  // rank-independent in practice (n / (64 P) chunks against 4 x 148 CTAs x 24 slots)
  if (LT_all > 0 && !c->sweep_inv)
    MS_CUDA (cudaMalloc (&c->sweep_inv, (size_t)LT_all * 64 * 64 * sizeof (double)));
  sa.A = c->A;
  sa.ld = c->ld;
  sa.n = c->n;
  sa.y = c->vec[MS_VEC_Y];
  sa.d = c->vec[MS_VEC_D];
  sa.V = c->sweep_inv;
  sa.x = c->vec[MS_VEC_X];
  sa.ready = c->sweep_ready;
  sa.rank = c->rank;
  sa.nranks = c->nranks;
  sa.q = (int)(c->m / MS_SUB);
  sa.epoch = ++c->sweep_epoch;
  sa.T = T;
  sa.LT = LT_all;
  sa.thresh = thresh;
  sa.flags = c->flags;
  sa.spin_limit = 6000000000ll;  // ~3 s of SM clock: far beyond any healthy sweep
  *done = 1;
  if (LT_all == 0)
    return MS_OK;  // more ranks than block rows: nothing to solve here, the peers still deliver x
  const size_t inv_smem = 2 * 64 * 65 * sizeof (double);
  static bool inv_configured[64] = {};
  if (!inv_configured[c->device & 63])
    {
      MS_CUDA (cudaFuncSetAttribute (bsweep_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)inv_smem));
      inv_configured[c->device & 63] = true;
    }
  bsweep_invert_kernel<<<LT_all, 64, inv_smem, s>>> (c->A, c->ld, c->n, c->m, c->rank, c->nranks, c->sweep_inv, thresh, c->flags);
  LAUNCH_COUNT (c);
  void *args[] = {(void *)&sa};
  MS_CUDA (cudaLaunchCooperativeKernel ((void *)bsweep_kernel, dim3 (G), dim3 (256), args, 0, s));
  LAUNCH_COUNT (c);
  return MS_OK;
}
