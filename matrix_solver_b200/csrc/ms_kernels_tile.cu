// ms_kernels_tile.cu — K3 / K4 / K5: the k-major tile contraction
//
//        C(i, j)  (op)=  sum_k  s_k * A(k, i) * B(k, j)
//
// in two flavours:
//   * tile_dmma_kernel  — the hot path.  64 x 128 output tile per CTA, fp64 tensor cores
//     (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), operands streamed by TMA
//     (cp.async.bulk.tensor, 128-byte swizzle) through a 4-stage mbarrier ring, 8 warps
//     (32 x 32 each, 32 fp64 accumulators per lane), two CTAs per SM so one CTA's
//     read-modify-write epilogue overlaps the other's math.
//   * gemm_generic_kernel — bounds-checked CUDA-core (DFMA) kernel for block sizes the tile
//     kernel does not take (m % 64 != 0).  Same contract, still on the device.
//
// Both serve the two uses of the contraction in the factorisation:
//   MS_MODE_UPDATE (K4 + K5, reference minus_RTDR_o1 / minus_RTDR_l_o1 / minus_RTDRu_ofast /
//       minus_RTDRu_l, src/matrix_op.cpp:4-99,156-242):   C -= P^T D P, A and B both slices of
//       the factored panel P (k = rows of block row I), s_k = d_k.  The diagonal-block (SYRK)
//       case is just the tile whose row and column ranges coincide.
//   MS_MODE_PANEL  (K3 in the reference's own formulation, DRtA / DRtA_l, src/matrix_op.cpp:287-368;
//       selected with ms_set_option("panel_trsm", 0) — the default K3 is the triangular solve in
//       ms_kernels_diag.cu):   the block row is overwritten by D V^T A_row, V = inverse of the piece's
//       factor; A = Wt (= V D, emitted by the diag kernel), B = the piece's rows of the matrix,
//       s_k = 1.  Output also goes to the panel buffer that feeds the updates and the peers.
//
// Fragment mapping (why the shared-memory reads are conflict free) is derived in DESIGN.md
// "tile kernel"; in short a lane (p = lane/4, q = lane%4) loads with ONE 16-byte LDS the two
// neighbouring columns 2p, 2p+1 of k-row 8g+2q+e, which under the 128B swizzle sit in 16-byte
// slot (p ^ (2q+e)) of that row: the 8 lanes of every quarter warp hit 8 distinct slots.
#include <algorithm>
#include <vector>

#include "ms_common.cuh"

#define LAUNCH_COUNT(c) ((c)->stats.kernel_launches++)

// ============================================================================ DMMA tile kernel
constexpr int TILE_MAX_BANDS = 64;

struct tile_args
{
  // TMA coordinates of 16-column chunk `ch` at k-row kk:  x = x0 + ch*xs,  y = y0 + ch*ys + kk
  // (chunk-major operands: xs = 0, ys = chunk height; row-major operands: xs = 16, ys = 0).
  // A chunks are counted from the tile's first output row, B chunks from global column 0.
  int ax0, axs, ay0, ays;
  int bx0, bxs, by0, bys;
  int K;                    // contraction length, multiple of MS_KC
  const double *dsign;      // K signs (+-1.0) or null
  double *C;                // row tile 0, global column 0
  size_t ldc;
  double *C2;               // MODE_PANEL: panel buffer (chunk-major, height c2_H), row offset c2_k0
  size_t c2_H, c2_k0;
  // MODE_UPDATE: row tile rt covers local rows lrow_begin + 64*rt ...; its global row follows from
  // the block-row cyclic map.  MODE_PANEL: single row tile, g0 = global row of the piece.
  size_t lrow_begin, m;
  int rank, nranks;
  int n_row_tiles;
  size_t g0;
  size_t col_tile0;         // MODE_PANEL: first column tile (units of MS_BN)
  size_t col_lo_panel;      // MODE_PANEL: first column to write
  size_t n;                 // columns >= n are never written
  size_t n_rows;            // rows >= n_rows do not exist (n < n_rows + 1 only without the augmented column)
  int tiles_per_cta;        // consecutive tiles one CTA walks before it retires
  // L2-friendly rasterisation (host-built, tile_build_bands): the stored tiles are walked band by band
  // (band = band_w consecutive column tiles), row tile by row tile inside a band.  Band b starts at tile
  // band_start[b]; its first band_r1[b] row tiles own the whole band width, row tiles
  // [band_r1[b], band_r2[b]) start inside the band, later row tiles own nothing of it.
  int band_w, n_bands;
  int band_start[TILE_MAX_BANDS + 1];
  int band_r1[TILE_MAX_BANDS], band_r2[TILE_MAX_BANDS];
};

constexpr int TILE_A_BYTES = (MS_BM / MS_CHUNK) * MS_KC * MS_CHUNK * 8;  // 8 KB
constexpr int TILE_B_BYTES = (MS_BN / MS_CHUNK) * MS_KC * MS_CHUNK * 8;  // 16 KB
constexpr int TILE_STAGE_BYTES = TILE_A_BYTES + TILE_B_BYTES;            // 24 KB
constexpr int TILE_CHUNK_BYTES = MS_KC * MS_CHUNK * 8;                   // 2 KB
constexpr int TILE_MAX_K = 1024;
constexpr int TILE_MAX_RT = 2048;  // row tiles per launch (the host splits larger ranges)
constexpr int TILE_BAND_W = 64;    // column tiles per band: an 8192-column slab of the panel (16 MB at K = 256)
constexpr int TILE_SMEM_BYTES = MS_STAGES * TILE_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/
                                + TILE_MAX_K * 4 + (TILE_MAX_RT + 2) * 4;

// global row of row tile rt / number of column tiles it owns (host and device agree on these)
__host__ __device__ __forceinline__ size_t
tile_row_global (const tile_args &ta, int mode, int rt)
{
  if (mode != MS_MODE_UPDATE)
    return ta.g0;
  const size_t l = ta.lrow_begin + (size_t)rt * MS_BM;
  const size_t lb = l / ta.m;
  return (lb * (size_t)ta.nranks + (size_t)ta.rank) * ta.m + (l - lb * ta.m);
}

__host__ __device__ __forceinline__ size_t
tile_first_col_tile (const tile_args &ta, int mode, size_t g)
{
  return mode == MS_MODE_UPDATE ? g / MS_BN : ta.col_tile0;
}

__host__ __device__ __forceinline__ int
tile_total_col_tiles (const tile_args &ta)
{
  return (int)((ta.n + MS_BN - 1) / MS_BN);
}

// first column tile a row tile owns, == the number of column tiles when it owns none (row tiles past
// the last row).  Non-decreasing in rt: the local rows of a launch ascend in global row order.
__host__ __device__ __forceinline__ int
tile_first_in_row (const tile_args &ta, int mode, int rt)
{
  const size_t g = tile_row_global (ta, mode, rt);
  const int total = tile_total_col_tiles (ta);
  if (g >= ta.n_rows)
    return total;
  const size_t first = tile_first_col_tile (ta, mode, g);
  return first < (size_t)total ? (int)first : total;
}

__host__ __device__ __forceinline__ int
tile_count_in_row (const tile_args &ta, int mode, int rt)
{
  return tile_total_col_tiles (ta) - tile_first_in_row (ta, mode, rt);
}

// Each CTA walks `tiles_per_cta` consecutive entries of the list of stored tiles (same row tile =
// same A operand for neighbours).  The TMA ring keeps running across those tiles, so the next tile's
// operands are already in flight during the epilogue.  CTAs are deliberately NOT persistent for the
// whole launch: retiring every few tiles lets the hardware scheduler hand SM slots to the
// high-priority panel stream (look-ahead) and keeps the two CTAs that share an SM out of phase, so one
// CTA's epilogue overlaps the other's DMMA loop.
//
// Register budget and instruction mix (round 2: this kernel replaced the 9-warp producer/consumer
// variant of round 1 after an A/B on the GPU, profiles/r02_experimental_ab.txt: 2807 vs 3005 ms at
// n = 65536):
//   * 8 warps, 128 registers, two CTAs per SM: no producer warp — thread 0 issues the TMA loads of
//     chunk j + 2 just before the CTA consumes chunk j (it then only ever waits for a stage that every
//     warp left two chunks ago);
//   * UPDATE mode loads through a rank-3 tensor map over the chunk-major panel buffer ({16, hpb,
//     ld/16}, box {16, 16, 4 chunks}): three TMA instructions per stage instead of twelve — this is
//     where the 7 % came from (the 2-D loads, MS_TILE_TMA3=0, measure 3007 ms);
//   * shared memory is addressed through the 32-bit shared window, ld.shared.v2.f64 with immediate
//     offsets: four address registers per lane for the whole k loop;
//   * the stage index is a compile-time constant of the 4x unrolled k loop (K % 64 == 0 on this
//     path, so every tile starts at stage 0);
//   * the fragment loads of k-group G + 1 are issued before the 16 DMMAs of group G by hand (both are
//     volatile asm, the compiler keeps the order).
constexpr int TILE_THREADS = MS_TILE_CONSUMER_WARPS * 32;
constexpr uint32_t TILE_AHEAD = MS_STAGES - 2;
static_assert (MS_STAGES == 4, "the k loop is unrolled over exactly four stages");

__device__ __forceinline__ double2
lds_f64x2 (uint32_t addr)
{
  double2 v;
  asm volatile ("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

struct tile_pos
{
  int rt;
  size_t g_row0, ct;
};

// tile number t of the launch -> (row tile, its global first row, column tile) in the banded order.
// prefF[rt] = sum of tile_first_in_row over row tiles < rt (shared memory, built in the prologue).
//
// Why bands: walked row tile by row tile, every row tile streams the WHOLE panel as its B operand —
// 134 MB at n = 65536, more than the 126 MB L2 — and ncu measured 53.7 GB of DRAM reads for 16.8 GB of C
// in the step-0 launch (profiles/r02_ncu_update_bsweep.txt, 2.09 x the algorithmic bytes).  Inside a band
// the B slab is 64 column tiles (16 MB at K = 256) and stays in L2 while all row tiles pass over it; the
// A slabs (128 KB per row tile) are re-read once per band.
__device__ __forceinline__ int
tile_band_of (const tile_args &ta, int t)
{
  int b = 0, bh = ta.n_bands - 1;
  while (b < bh)
    {
      const int mid = (b + bh + 1) >> 1;
      if (ta.band_start[mid] <= t)
        b = mid;
      else
        bh = mid - 1;
    }
  return b;
}

// prefF[i] = sum of tile_first_in_row over row tiles [rbase, rbase + i): only the row tiles that START inside
// the bands this CTA touches are ever looked up, so the prologue builds the prefix for those (<= 128 per band)
// instead of for every row tile of the launch, and not at all for CTAs whose tiles lie in full-width rows.
template <int MODE>
__device__ __forceinline__ tile_pos
tile_decode (const tile_args &ta, const int *prefF, int rbase, int t)
{
  const int b = tile_band_of (ta, t);
  const int lo_ct = b * ta.band_w;
  const int hi_ct = min (lo_ct + ta.band_w, tile_total_col_tiles (ta));
  const int wfull = hi_ct - lo_ct;
  const int r1 = ta.band_r1[b], r2 = ta.band_r2[b];
  int o = t - ta.band_start[b];
  tile_pos tp;
  if (o < r1 * wfull)
    {
      tp.rt = o / wfull;
      tp.ct = (size_t)(lo_ct + (o - tp.rt * wfull));
    }
  else
    {
      // row tiles [r1, r2) start inside the band: cum(rt) = tiles of [r1, rt) = (rt - r1) hi - (prefF[rt] - prefF[r1])
      o -= r1 * wfull;
      const int *pf = prefF - rbase;
      int lo = r1, hi = r2 - 1;
      while (lo < hi)
        {
          const int mid = (lo + hi + 1) >> 1;
          if ((mid - r1) * hi_ct - (pf[mid] - pf[r1]) <= o)
            lo = mid;
          else
            hi = mid - 1;
        }
      tp.rt = lo;
      const int first = pf[lo + 1] - pf[lo];
      tp.ct = (size_t)(first + (o - ((lo - r1) * hi_ct - (pf[lo] - pf[r1]))));
    }
  tp.g_row0 = tile_row_global (ta, MODE, tp.rt);
  return tp;
}

// one k chunk (16 k rows) out of stage S_: 4 k-groups x 16 DMMAs
template <int MODE, int S_>
__device__ __forceinline__ void
tile_consume (double (&acc)[2][2][2][2][2], uint32_t a_e0, uint32_t a_e1, uint32_t b_e0, uint32_t b_e1,
               const uint32_t *sgn_chunk, bool any_neg)
{
  constexpr uint32_t SO = S_ * TILE_STAGE_BYTES;
  double2 fa[2][2], fb[2][2];  // [buffer][ib or jb]
#pragma unroll
  for (int ib = 0; ib < 2; ib++)
    fa[0][ib] = lds_f64x2 (a_e0 + SO + ib * TILE_CHUNK_BYTES);
#pragma unroll
  for (int jb = 0; jb < 2; jb++)
    fb[0][jb] = lds_f64x2 (b_e0 + SO + jb * TILE_CHUNK_BYTES);
#pragma unroll
  for (int G = 0; G < 4; G++)
    {
      const int cur = G & 1, nxt = cur ^ 1;
      if (G < 3)
        {
          // group G + 1 = (g2, e2): k rows 8*g2 + 2q + e2 — 1024 bytes per g, the e offset is in a_e1/b_e1
          const int g2 = (G + 1) >> 1, e2 = (G + 1) & 1;
          const uint32_t an = (e2 ? a_e1 : a_e0) + SO + g2 * 1024;
          const uint32_t bn = (e2 ? b_e1 : b_e0) + SO + g2 * 1024;
#pragma unroll
          for (int ib = 0; ib < 2; ib++)
            fa[nxt][ib] = lds_f64x2 (an + ib * TILE_CHUNK_BYTES);
#pragma unroll
          for (int jb = 0; jb < 2; jb++)
            fb[nxt][jb] = lds_f64x2 (bn + jb * TILE_CHUNK_BYTES);
        }
      if (MODE == MS_MODE_UPDATE && any_neg)
        {
          const uint32_t sm = sgn_chunk[8 * (G >> 1) + (G & 1)];  // k row 8g + 2q + e (2q is in the pointer)
#pragma unroll
          for (int ib = 0; ib < 2; ib++)
            {
              fa[cur][ib].x = flip_sign (fa[cur][ib].x, sm);
              fa[cur][ib].y = flip_sign (fa[cur][ib].y, sm);
            }
        }
#pragma unroll
      for (int ib = 0; ib < 2; ib++)
#pragma unroll
        for (int jb = 0; jb < 2; jb++)
          {
            dmma_8x8x4 (acc[ib][jb][0][0][0], acc[ib][jb][0][0][1], fa[cur][ib].x, fb[cur][jb].x);
            dmma_8x8x4 (acc[ib][jb][0][1][0], acc[ib][jb][0][1][1], fa[cur][ib].x, fb[cur][jb].y);
            dmma_8x8x4 (acc[ib][jb][1][0][0], acc[ib][jb][1][0][1], fa[cur][ib].y, fb[cur][jb].x);
            dmma_8x8x4 (acc[ib][jb][1][1][0], acc[ib][jb][1][1][1], fa[cur][ib].y, fb[cur][jb].y);
          }
    }
}

// the same chunk for a warp whose 32-column patch sticks out of the matrix on the right (the last
// column tile: with the augmented column at index n only 1 of its 128 columns exists): only the
// 16-column groups jb < njb that hold a real column are multiplied.  Off the hot path — no pipelining.
template <int MODE>
__device__ __forceinline__ void
tile_consume_narrow (double (&acc)[2][2][2][2][2], uint32_t a_e0, uint32_t a_e1, uint32_t b_e0, uint32_t b_e1,
                     uint32_t so, const uint32_t *sgn_chunk, bool any_neg, int njb)
{
#pragma unroll
  for (int G = 0; G < 4; G++)
    {
      const int g = G >> 1, e = G & 1;
      const uint32_t an = (e ? a_e1 : a_e0) + so + g * 1024;
      const uint32_t bn = (e ? b_e1 : b_e0) + so + g * 1024;
      double2 fa[2];
#pragma unroll
      for (int ib = 0; ib < 2; ib++)
        fa[ib] = lds_f64x2 (an + ib * TILE_CHUNK_BYTES);
      if (MODE == MS_MODE_UPDATE && any_neg)
        {
          const uint32_t sm = sgn_chunk[8 * g + e];
#pragma unroll
          for (int ib = 0; ib < 2; ib++)
            {
              fa[ib].x = flip_sign (fa[ib].x, sm);
              fa[ib].y = flip_sign (fa[ib].y, sm);
            }
        }
#pragma unroll
      for (int jb = 0; jb < 2; jb++)
        if (jb < njb)
          {
            const double2 fb = lds_f64x2 (bn + jb * TILE_CHUNK_BYTES);
#pragma unroll
            for (int ib = 0; ib < 2; ib++)
              {
                dmma_8x8x4 (acc[ib][jb][0][0][0], acc[ib][jb][0][0][1], fa[ib].x, fb.x);
                dmma_8x8x4 (acc[ib][jb][0][1][0], acc[ib][jb][0][1][1], fa[ib].x, fb.y);
                dmma_8x8x4 (acc[ib][jb][1][0][0], acc[ib][jb][1][0][1], fa[ib].y, fb.x);
                dmma_8x8x4 (acc[ib][jb][1][1][0], acc[ib][jb][1][1][1], fa[ib].y, fb.y);
              }
          }
    }
}

// the read-modify-write (UPDATE) / store (PANEL) of one warp's 32 x 32 patch; same ownership map as
// per (ib, jb, a) a lane owns row wm*32 + 16*ib + 2p + a, columns wn*32 + 16*jb + 4q .. +3
template <int MODE>
__device__ __forceinline__ void
tile_epilogue_row (const tile_args &ta, const double (&acc)[2][2][2][2][2], double *Ct, int ib, int a, int r,
                    size_t cbase, size_t col_lo)
{
  double *crow = Ct + (size_t)r * ta.ldc;
  double v[2][4];
#pragma unroll
  for (int jb = 0; jb < 2; jb++)
    {
      v[jb][0] = acc[ib][jb][a][0][0];
      v[jb][1] = acc[ib][jb][a][1][0];
      v[jb][2] = acc[ib][jb][a][0][1];
      v[jb][3] = acc[ib][jb][a][1][1];
    }
  const bool fast0 = cbase >= col_lo && cbase + 3 < ta.n;
  const bool fast1 = cbase + 16 >= col_lo && cbase + 19 < ta.n;
  if (fast0 && fast1)
    {
      double2 *p0 = reinterpret_cast<double2 *> (crow + cbase);
      double2 *p1 = reinterpret_cast<double2 *> (crow + cbase + 16);
      if (MODE == MS_MODE_UPDATE)
        {
          double2 o00 = p0[0], o01 = p0[1], o10 = p1[0], o11 = p1[1];
          o00.x -= v[0][0];
          o00.y -= v[0][1];
          o01.x -= v[0][2];
          o01.y -= v[0][3];
          o10.x -= v[1][0];
          o10.y -= v[1][1];
          o11.x -= v[1][2];
          o11.y -= v[1][3];
          p0[0] = o00;
          p0[1] = o01;
          p1[0] = o10;
          p1[1] = o11;
        }
      else
        {
          p0[0] = make_double2 (v[0][0], v[0][1]);
          p0[1] = make_double2 (v[0][2], v[0][3]);
          p1[0] = make_double2 (v[1][0], v[1][1]);
          p1[1] = make_double2 (v[1][2], v[1][3]);
          double2 *q0 = reinterpret_cast<double2 *> (ta.C2 + chunk_major_index (ta.c2_k0 + r, cbase, ta.c2_H));
          double2 *q1 = reinterpret_cast<double2 *> (ta.C2 + chunk_major_index (ta.c2_k0 + r, cbase + 16, ta.c2_H));
          q0[0] = make_double2 (v[0][0], v[0][1]);
          q0[1] = make_double2 (v[0][2], v[0][3]);
          q1[0] = make_double2 (v[1][0], v[1][1]);
          q1[1] = make_double2 (v[1][2], v[1][3]);
        }
    }
  else
    {
#pragma unroll
      for (int jb = 0; jb < 2; jb++)
#pragma unroll
        for (int tt = 0; tt < 4; tt++)
          {
            const size_t col = cbase + 16 * jb + tt;
            if (col >= col_lo && col < ta.n)
              {
                if (MODE == MS_MODE_UPDATE)
                  crow[col] -= v[jb][tt];
                else
                  {
                    crow[col] = v[jb][tt];
                    ta.C2[chunk_major_index (ta.c2_k0 + r, col, ta.c2_H)] = v[jb][tt];
                  }
              }
          }
    }
}

template <int MODE>
__device__ __forceinline__ void
tile_epilogue (const tile_args &ta, const double (&acc)[2][2][2][2][2], double *Ct, int rows_valid, size_t col0,
                size_t col_lo, int wm, int wn, int p, int q)
{
  const size_t cbase = col0 + wn * 32 + 4 * q;
  const bool fast = cbase >= col_lo && cbase + 19 < ta.n;  // both 4-column groups fully inside [col_lo, n)
#pragma unroll
  for (int ib = 0; ib < 2; ib++)
    {
      const int r0 = wm * 32 + 16 * ib + 2 * p;
      if (MODE == MS_MODE_UPDATE && fast && r0 + 1 < rows_valid)
        {
          // interior patch: both rows of the pair in one batch — 8 independent 16-byte loads in flight
          // before the first subtraction, i.e. two memory round trips per tile instead of four
          double2 *pa = reinterpret_cast<double2 *> (Ct + (size_t)r0 * ta.ldc + cbase);
          double2 *pb = reinterpret_cast<double2 *> (Ct + (size_t)(r0 + 1) * ta.ldc + cbase);
          double2 oa[4] = {pa[0], pa[1], pa[8], pa[9]};  // columns cbase .. +3 and cbase + 16 .. +19
          double2 ob[4] = {pb[0], pb[1], pb[8], pb[9]};
#pragma unroll
          for (int jb = 0; jb < 2; jb++)
            {
              oa[2 * jb].x -= acc[ib][jb][0][0][0];
              oa[2 * jb].y -= acc[ib][jb][0][1][0];
              oa[2 * jb + 1].x -= acc[ib][jb][0][0][1];
              oa[2 * jb + 1].y -= acc[ib][jb][0][1][1];
              ob[2 * jb].x -= acc[ib][jb][1][0][0];
              ob[2 * jb].y -= acc[ib][jb][1][1][0];
              ob[2 * jb + 1].x -= acc[ib][jb][1][0][1];
              ob[2 * jb + 1].y -= acc[ib][jb][1][1][1];
            }
          pa[0] = oa[0];
          pa[1] = oa[1];
          pa[8] = oa[2];
          pa[9] = oa[3];
          pb[0] = ob[0];
          pb[1] = ob[1];
          pb[8] = ob[2];
          pb[9] = ob[3];
          continue;
        }
#pragma unroll
      for (int a = 0; a < 2; a++)
        if (r0 + a < rows_valid)
          tile_epilogue_row<MODE> (ta, acc, Ct, ib, a, r0 + a, cbase, col_lo);
    }
}

template <int MODE, bool TM3>
__global__ void __launch_bounds__ (TILE_THREADS, 2)
tile_dmma_kernel (const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tm3, const tile_args ta)
{
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment for the 128-byte swizzle, computed as an OFFSET so that `smem` stays a pointer into
  // the shared window for the compiler (LDS/STS instead of generic accesses)
  const uint32_t raw_u32 = smem_u32 (smem_raw);
  const uint32_t pad = ((raw_u32 + 1023u) & ~1023u) - raw_u32;
  uint8_t *smem = smem_raw + pad;
  const uint32_t sbase = raw_u32 + pad;
  uint64_t *full_bar = (uint64_t *)(smem + MS_STAGES * TILE_STAGE_BYTES);
  uint64_t *empty_bar = full_bar + MS_STAGES;
  uint32_t *sgn = (uint32_t *)(empty_bar + MS_STAGES + 8);
  int *pref = (int *)(sgn + TILE_MAX_K);  // pref[rt] = sum of the first column tiles of row tiles < rt

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = ta.K / MS_KC;  // multiple of 4 (the host checks K % 64 == 0)

  const int total_tiles = ta.band_start[ta.n_bands];
  const int t_begin = blockIdx.x * ta.tiles_per_cta;
  const int t_end = min (total_tiles, t_begin + ta.tiles_per_cta);
  // which row tiles can the decode of THIS CTA's tiles look up?  Only those that start inside the band(s) of
  // its first and last tile, and only if one of the tiles lies beyond the band's full-width rows.
  int rbase = 0, rcnt = 0;
  if (t_begin < t_end)
    {
      const int b0 = tile_band_of (ta, t_begin), b1 = tile_band_of (ta, t_end - 1);
      const int total_ct = tile_total_col_tiles (ta);
      const int w0 = min ((b0 + 1) * ta.band_w, total_ct) - b0 * ta.band_w;
      const int w1 = min ((b1 + 1) * ta.band_w, total_ct) - b1 * ta.band_w;
      const bool part = b0 != b1 || t_begin - ta.band_start[b0] >= ta.band_r1[b0] * w0
                        || t_end - 1 - ta.band_start[b1] >= ta.band_r1[b1] * w1;
      if (part)
        {
          rbase = ta.band_r1[b0];
          rcnt = ta.band_r2[b1] - rbase;
        }
    }
  const uint32_t n_chunks = t_end > t_begin ? (uint32_t)(t_end - t_begin) * (uint32_t)nkb : 0u;

  // producer state, advanced by thread 0 only: next chunk to request and where its operands are
  uint32_t pj = 0;
  int p_t = t_begin, p_kb = 0, p_ax = 0, p_ay = 0, p_bc = 0, p_ac = 0;
  auto producer_seek = [&] () {
    const tile_pos tp = tile_decode<MODE> (ta, pref, rbase, p_t);
    const int achunk0 = MODE == MS_MODE_UPDATE ? (int)(tp.g_row0 / MS_CHUNK) : 0;
    p_ac = achunk0;
    p_ax = ta.ax0 + achunk0 * ta.axs;
    p_ay = ta.ay0 + achunk0 * ta.ays;
    p_bc = (int)(tp.ct * (MS_BN / MS_CHUNK));
  };
  // request every chunk up to number `upto` (thread 0): wait until its stage is free, arm the barrier, issue the loads
  auto request = [&] (uint32_t upto) {
    while (pj < n_chunks && pj <= upto)
      {
        const uint32_t ps = pj & (MS_STAGES - 1);
        mbar_wait (&empty_bar[ps], ((pj >> 2) & 1u) ^ 1u);
        mbar_arrive_expect_tx (&full_bar[ps], TILE_STAGE_BYTES);
        uint8_t *sa = smem + ps * TILE_STAGE_BYTES;
        uint8_t *sb = sa + TILE_A_BYTES;
        const int kk = p_kb * MS_KC;
        if (MODE == MS_MODE_UPDATE && TM3)
          {
            // rank-3 view of the chunk-major panel buffer: coordinates (column in chunk, k row,
            // chunk); a box holds four chunks, so a stage is three instructions, not twelve
            tma_load_3d (sa, &tm3, 0, ta.ay0 + kk, p_ac, &full_bar[ps]);
            tma_load_3d (sb, &tm3, 0, ta.by0 + kk, p_bc, &full_bar[ps]);
            tma_load_3d (sb + 4 * TILE_CHUNK_BYTES, &tm3, 0, ta.by0 + kk, p_bc + 4, &full_bar[ps]);
          }
        else
          {
#pragma unroll
            for (int ci = 0; ci < MS_BM / MS_CHUNK; ci++)
              tma_load_2d (sa + ci * TILE_CHUNK_BYTES, &tmA, p_ax + ci * ta.axs, p_ay + ci * ta.ays + kk, &full_bar[ps]);
#pragma unroll
            for (int cj = 0; cj < MS_BN / MS_CHUNK; cj++)
              tma_load_2d (sb + cj * TILE_CHUNK_BYTES, &tmB, ta.bx0 + (p_bc + cj) * ta.bxs,
                           ta.by0 + (p_bc + cj) * ta.bys + kk, &full_bar[ps]);
          }
        pj++;
        if (++p_kb == nkb)
          {
            p_kb = 0;
            if (++p_t < t_end)
              producer_seek ();
          }
      }
  };
  if (threadIdx.x == 0)
    {
      for (int s = 0; s < MS_STAGES; s++)
        {
          mbar_init (&full_bar[s], 1);
          mbar_init (&empty_bar[s], MS_TILE_CONSUMER_WARPS);
        }
      mbar_fence_init ();
      // The first three chunks are requested NOW, before the rest of the prologue (sign words from global memory,
      // two CTA barriers), so that the ~1 us until the first stage lands overlaps it.  Only CTAs whose tiles need the
      // prefix array have to wait for it.  (Measured neutral with two CTAs per SM — 2819.8 vs 2820.1 ms at n = 65536,
      // 56.5 vs 56.3 ms at n = 16384, m = 128: the other CTA's DMMA loop already covers the fill.)
      if (n_chunks > 0 && rcnt == 0)
        {
          producer_seek ();
          request (TILE_AHEAD);
        }
    }
  int my_neg = 0;
  for (int k = threadIdx.x; k < ta.K; k += TILE_THREADS)
    {
      const bool neg = ta.dsign != nullptr && ta.dsign[k] < 0.0;
      sgn[k] = neg ? 0x80000000u : 0u;
      my_neg |= neg;
    }
  for (int i = threadIdx.x; i < rcnt; i += TILE_THREADS)
    pref[i + 1] = tile_first_in_row (ta, MODE, rbase + i);
  const bool any_neg = __syncthreads_or (my_neg) != 0;
  if (warp == 0 && rcnt > 0)
    {
      // inclusive scan of pref[1..rcnt] by one warp: lane-contiguous segments
      const int seg = (rcnt + 31) / 32;
      const int b = lane * seg, e = min (b + seg, rcnt);
      int sum = 0;
      for (int i = b; i < e; i++)
        sum += pref[i + 1];
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
        {
          int v = __shfl_up_sync (0xffffffffu, incl, o);
          if (lane >= o)
            incl += v;
        }
      int run = incl - sum;
      for (int i = b; i < e; i++)
        {
          int c = pref[i + 1];
          pref[i + 1] = run + c;
          run += c;
        }
      if (lane == 0)
        pref[0] = 0;
    }
  __syncthreads ();

  const int p = lane >> 2, q = lane & 3;
  const int wm = warp >> 2;  // rows  wm*32
  const int wn = warp & 3;   // cols  wn*32
  // fragment addresses inside a stage: k row 8g + 2q + e, 16-byte slot p ^ (2q + e) of the 128-byte row
  // (the swizzle); g adds 1024 bytes, the second 16-column chunk of the warp adds TILE_CHUNK_BYTES
  const uint32_t off_e0 = (uint32_t)((2 * q) * 128 + ((p ^ (2 * q)) << 4));
  const uint32_t off_e1 = (uint32_t)((2 * q + 1) * 128 + ((p ^ (2 * q + 1)) << 4));
  const uint32_t a_base = sbase + (uint32_t)(wm * 2) * TILE_CHUNK_BYTES;
  const uint32_t b_base = sbase + TILE_A_BYTES + (uint32_t)(wn * 2) * TILE_CHUNK_BYTES;
  const uint32_t a_e0 = a_base + off_e0, a_e1 = a_base + off_e1;
  const uint32_t b_e0 = b_base + off_e0, b_e1 = b_base + off_e1;

  if (threadIdx.x == 0 && n_chunks > 0 && rcnt > 0)
    producer_seek ();  // (the other CTAs did this, and requested their first chunks, at the top of the kernel)

  uint32_t it = 0;  // chunks consumed so far by this CTA (all tiles)
  for (int t = t_begin; t < t_end; t++)
    {
      const tile_pos tp = tile_decode<MODE> (ta, pref, rbase, t);
      const size_t g_row0 = tp.g_row0;
      const size_t col0 = tp.ct * MS_BN;
      const int rows_valid = g_row0 + MS_BM <= ta.n_rows ? MS_BM : (int)(ta.n_rows - g_row0);
      const size_t col_lo = (MODE == MS_MODE_UPDATE) ? g_row0 : ta.col_lo_panel;
      double *Ct = ta.C + (size_t)tp.rt * MS_BM * ta.ldc;
      // 16-column groups of this warp's patch that hold at least one column < n (2 everywhere but in
      // the last column tile)
      const size_t wcol0 = col0 + (size_t)wn * 32;
      const int njb = wcol0 + 16 < ta.n ? 2 : (wcol0 < ta.n ? 1 : 0);
      if (MODE == MS_MODE_UPDATE)
        {
          // the C tile is touched exactly once per launch and comes from HBM: start pulling this
          // warp's 32 x 32 patch into L2 now so that the epilogue's loads find it there
#pragma unroll
          for (int i = 0; i < 4; i++)
            {
              const int r = wm * 32 + p * 4 + i;
              const size_t cc = col0 + wn * 32 + q * 8;
              if (r < rows_valid && cc < ta.n)
                asm volatile ("prefetch.global.L2 [%0];" ::"l"(Ct + (size_t)r * ta.ldc + cc));
            }
        }
      double acc[2][2][2][2][2];  // [ib][jb][a][b][c0/c1]
#pragma unroll
      for (int i = 0; i < 32; i++)
        (&acc[0][0][0][0][0])[i] = 0.0;

      for (int kb4 = 0; kb4 < nkb; kb4 += MS_STAGES)
        {
          const uint32_t ph = (it >> 2) & 1u;  // it % 4 == 0 here: all four stages share the phase
#pragma unroll
          for (int s = 0; s < MS_STAGES; s++)
            {
              if (threadIdx.x == 0)
                request (it + (uint32_t)s + TILE_AHEAD);  // three chunks at the very start, then one per chunk
              mbar_wait (&full_bar[s], ph);
              const uint32_t *sgn_chunk = sgn + (kb4 + s) * MS_KC + 2 * q;
              if (njb < 2)
                {
                  if (njb > 0)
                    tile_consume_narrow<MODE> (acc, a_e0, a_e1, b_e0, b_e1, (uint32_t)s * TILE_STAGE_BYTES, sgn_chunk,
                                               any_neg, njb);
                }
              else if (s == 0)
                tile_consume<MODE, 0> (acc, a_e0, a_e1, b_e0, b_e1, sgn_chunk, any_neg);
              else if (s == 1)
                tile_consume<MODE, 1> (acc, a_e0, a_e1, b_e0, b_e1, sgn_chunk, any_neg);
              else if (s == 2)
                tile_consume<MODE, 2> (acc, a_e0, a_e1, b_e0, b_e1, sgn_chunk, any_neg);
              else
                tile_consume<MODE, 3> (acc, a_e0, a_e1, b_e0, b_e1, sgn_chunk, any_neg);
              __syncwarp ();
              if (lane == 0)
                mbar_arrive (&empty_bar[s]);
            }
          it += MS_STAGES;
        }

      // (An epilogue through TMA bulk reductions — C += (-acc) as cp.reduce.async.bulk add.f64 out of the idle
      // operand ring, so that the read-modify-write happens in L2 — was built and measured in round 2: bit-identical,
      // and no faster at K = 256, 128 or 64 (2833 vs 2824 ms, 55.98 vs 56.07 ms, 4.96 vs 4.98 ms).  The per-tile
      // overhead at small K is the pipeline fill, not the epilogue; the variant was removed.)
      tile_epilogue<MODE> (ta, acc, Ct, rows_valid, col0, col_lo, wm, wn, p, q);
    }
}

// ---------------------------------------------------------------------------- tensor maps
typedef CUresult (*encode_fn_t) (CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);

static int
make_map_2d (CUtensorMap *map, void *base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes)
{
  static encode_fn_t encode = nullptr;
  if (!encode)
    {
      void *fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      MS_CUDA (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      if (qres != cudaDriverEntryPointSuccess || fn == nullptr)
        {
          ms_set_error ("cuTensorMapEncodeTiled not available from the driver");
          return MS_ERR_CUDA;
        }
      encode = (encode_fn_t)fn;
    }
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {MS_CHUNK, MS_KC};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode (map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    {
      ms_set_error ("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu rows=%llu stride=%llu)",
                    (int)r, (unsigned long long)inner, (unsigned long long)rows,
                    (unsigned long long)row_stride_bytes);
      return MS_ERR_CUDA;
    }
  return MS_OK;
}

// rank-3 view of a chunk-major panel buffer: (column in chunk, k row, chunk), box = 16 x 16 x 4 chunks
static int
make_map_pb3 (CUtensorMap *map, void *base, uint64_t hpb, uint64_t nchunks)
{
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  MS_CUDA (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || fn == nullptr)
    {
      ms_set_error ("cuTensorMapEncodeTiled not available from the driver");
      return MS_ERR_CUDA;
    }
  cuuint64_t dims[3] = {MS_CHUNK, hpb, nchunks};
  cuuint64_t strides[2] = {MS_CHUNK * 8, hpb * MS_CHUNK * 8};
  cuuint32_t box[3] = {MS_CHUNK, MS_KC, 4};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ((encode_fn_t)fn) (map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    {
      ms_set_error ("cuTensorMapEncodeTiled (rank 3) failed with CUresult %d", (int)r);
      return MS_ERR_CUDA;
    }
  return MS_OK;
}

// encoded on the first update launch; failure is not an error,
// the kernel then takes its 2-D loads
static bool
tile_tm3_ready (ms_ctx *c)
{
  if (c->tm3_state == 0)
    {
      const char *e = getenv ("MS_TILE_TMA3");
      c->tm3_state = (e && !atoi (e)) ? -1 : 1;
      for (int b = 0; b < MS_NPB && c->tm3_state == 1; b++)
        if (make_map_pb3 (&c->tm_PB3[b], c->PB[b], c->hpb, c->ld / MS_CHUNK) != MS_OK)
          c->tm3_state = -1;
    }
  return c->tm3_state == 1;
}

int
ms_tile_setup (ms_ctx *c)
{
  c->have_tm = false;
  if (!c->fast)
    return MS_OK;
  MS_TRY (make_map_2d (&c->tm_A, c->A, c->ld, c->rows_alloc, c->ld * 8));
  for (int b = 0; b < MS_NPB; b++)
    MS_TRY (make_map_2d (&c->tm_PB[b], c->PB[b], MS_CHUNK, (c->ld / MS_CHUNK) * c->hpb, MS_CHUNK * 8));
  MS_TRY (make_map_2d (&c->tm_Wt, c->Wt, MS_CHUNK, (MS_SUB / MS_CHUNK) * MS_SUB, MS_CHUNK * 8));
  MS_CUDA (cudaFuncSetAttribute (tile_dmma_kernel<MS_MODE_UPDATE, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
  MS_CUDA (cudaFuncSetAttribute (tile_dmma_kernel<MS_MODE_UPDATE, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
  MS_CUDA (cudaFuncSetAttribute (tile_dmma_kernel<MS_MODE_PANEL, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
  c->have_tm = true;
  return MS_OK;
}

// fills the band table of `ta` and returns the number of stored tiles of the launch
static int
tile_build_bands (const ms_ctx *c, tile_args &ta, int mode)
{
  const int total_ct = tile_total_col_tiles (ta);
  const int n_rt = ta.n_row_tiles;
  ta.band_w = std::max (c->tile_band_w > 0 ? c->tile_band_w : TILE_BAND_W, (total_ct + TILE_MAX_BANDS - 1) / TILE_MAX_BANDS);
  ta.n_bands = std::max (1, (total_ct + ta.band_w - 1) / ta.band_w);
  std::vector<int> first (n_rt);
  std::vector<long long> pref (n_rt + 1, 0);
  for (int rt = 0; rt < n_rt; rt++)
    {
      first[rt] = tile_first_in_row (ta, mode, rt);
      pref[rt + 1] = pref[rt] + first[rt];
    }
  long long start = 0;
  for (int b = 0; b < ta.n_bands; b++)
    {
      const int lo = b * ta.band_w, hi = std::min (lo + ta.band_w, total_ct);
      // first[] is non-decreasing
      const int r1 = (int)(std::upper_bound (first.begin (), first.end (), lo) - first.begin ());
      const int r2 = (int)(std::lower_bound (first.begin (), first.end (), hi) - first.begin ());
      ta.band_start[b] = (int)std::min<long long> (start, 0x7fffffff);
      ta.band_r1[b] = r1;
      ta.band_r2[b] = r2;
      start += (long long)r1 * (hi - lo) + (long long)(r2 - r1) * hi - (pref[r2] - pref[r1]);
    }
  ta.band_start[ta.n_bands] = (int)std::min<long long> (start, 0x7fffffff);
  return ta.band_start[ta.n_bands];
}

// Tiles per CTA.  Running CTAs are never preempted: the high-priority panel chain (diag, TRSM, the
// next block row's update) only gets an SM slot when a low-priority tile CTA retires, so the CTA
// lifetime is the chain's ramp-up latency.  One tile per CTA everywhere since round 2: with the lean
// prologue (band table in the kernel parameters, prefix built only for the rows a CTA can look up) walking
// several tiles per CTA no longer pays — measured 2836 vs 2862 ms on one GPU and 1420 vs 1445 ms on two
// (profiles/README.md); MS_TILES_PER_CTA overrides.
static int g_tiles_per_cta_max = -1;
static int
tiles_per_cta_for (const ms_ctx *c, int total)
{
  if (g_tiles_per_cta_max < 0)
    {
      const char *e = getenv ("MS_TILES_PER_CTA");
      g_tiles_per_cta_max = e ? std::max (1, atoi (e)) : 0;
    }
  int cap = g_tiles_per_cta_max > 0 ? g_tiles_per_cta_max : 1;
  int t = total / (2 * c->sm_count * 8);
  return std::max (1, std::min (t, cap));
}

// Update local rows [lrow_begin, lrow_end) (64-row tiles of block rows J > I, or the later pieces
// of block row I itself) with rows [k0, k0+K) of panel buffer `pb` (= factored block row I).
int
ms_k_tile_update (ms_ctx *c, int pb, size_t I, size_t lrow_begin, size_t lrow_end, int K, size_t k0,
                  cudaStream_t s)
{
  if (lrow_end <= lrow_begin)
    return MS_OK;
  MS_CHECK_ARG (c->have_tm, "tile kernel not set up");
  MS_CHECK_ARG (K % (MS_KC * MS_STAGES) == 0 && K <= TILE_MAX_K, "K must be a multiple of 64 and <= 1024");
  MS_CHECK_ARG (lrow_begin % MS_BM == 0, "row range must start on a tile boundary");
  for (size_t lr = lrow_begin; lr < lrow_end; lr += (size_t)TILE_MAX_RT * MS_BM)
    {
      size_t lr_end = std::min (lrow_end, lr + (size_t)TILE_MAX_RT * MS_BM);
      tile_args ta = {};
      ta.ax0 = 0; ta.axs = 0; ta.ay0 = (int)k0; ta.ays = (int)c->hpb;
      ta.bx0 = 0; ta.bxs = 0; ta.by0 = (int)k0; ta.bys = (int)c->hpb;
      ta.K = K;
      ta.dsign = c->vec[MS_VEC_D] + I * c->m + k0;
      ta.C = c->A + lr * c->ld;
      ta.ldc = c->ld;
      ta.C2 = nullptr;
      ta.lrow_begin = lr;
      ta.m = c->m;
      ta.rank = c->rank;
      ta.nranks = c->nranks;
      ta.n_row_tiles = (int)((lr_end - lr + MS_BM - 1) / MS_BM);
      ta.n = c->ncols;
      ta.n_rows = c->n;
      int total = tile_build_bands (c, ta, MS_MODE_UPDATE);
      if (total == 0)
        continue;
      ta.tiles_per_cta = tiles_per_cta_for (c, total);
      unsigned grid = (unsigned)((total + ta.tiles_per_cta - 1) / ta.tiles_per_cta);
      if (tile_tm3_ready (c))
        tile_dmma_kernel<MS_MODE_UPDATE, true><<<grid, TILE_THREADS, TILE_SMEM_BYTES, s>>> (
          c->tm_PB[pb], c->tm_PB[pb], c->tm_PB3[pb], ta);
      else
        tile_dmma_kernel<MS_MODE_UPDATE, false><<<grid, TILE_THREADS, TILE_SMEM_BYTES, s>>> (
          c->tm_PB[pb], c->tm_PB[pb], c->tm_PB[pb], ta);
      LAUNCH_COUNT (c);
      c->stats.update_launches++;
      c->stats.used_dmma = 1;
    }
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// Panel of one full piece (h == 64): columns [row0+64, n) of its rows become D V^T A, in place and
// into panel buffer rows [k0, k0+64) where k0 = row0 - I*m.
int
ms_k_tile_panel (ms_ctx *c, int pb, const ms_piece &pc, size_t I, cudaStream_t s)
{
  MS_CHECK_ARG (c->have_tm && pc.h == MS_SUB, "tile panel needs a full 64-row piece");
  size_t col_lo = pc.row0 + MS_SUB;
  if (col_lo >= c->ncols)
    return MS_OK;
  tile_args ta = {};
  ta.ax0 = 0; ta.axs = 0; ta.ay0 = 0; ta.ays = MS_SUB;
  ta.bx0 = 0; ta.bxs = MS_CHUNK; ta.by0 = (int)pc.lrow0; ta.bys = 0;
  ta.K = MS_SUB;
  ta.dsign = nullptr;
  ta.C = c->A + pc.lrow0 * c->ld;
  ta.ldc = c->ld;
  ta.C2 = c->PB[pb];
  ta.c2_H = c->hpb;
  ta.c2_k0 = pc.row0 - I * c->m;
  ta.m = c->m;
  ta.rank = c->rank;
  ta.nranks = c->nranks;
  ta.n_row_tiles = 1;
  ta.g0 = pc.row0;
  ta.col_tile0 = col_lo / MS_BN;
  ta.col_lo_panel = col_lo;
  ta.n = c->ncols;
  ta.n_rows = c->n;
  int total = tile_build_bands (c, ta, MS_MODE_PANEL);
  if (total == 0)
    return MS_OK;
  ta.tiles_per_cta = tiles_per_cta_for (c, total);
  unsigned grid = (unsigned)((total + ta.tiles_per_cta - 1) / ta.tiles_per_cta);
  tile_dmma_kernel<MS_MODE_PANEL, false><<<grid, TILE_THREADS, TILE_SMEM_BYTES, s>>> (c->tm_Wt, c->tm_A, c->tm_A, ta);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// ============================================================================ generic kernel
// 64 x 64 tile, 256 threads, 4 x 4 outputs per thread, k in slabs of 32 through shared memory.
// Row tiles cover rows [64*by, ...) of the launch; every row carries its own A-operand column
// (its global row index in the factorisation), so ranges may straddle block rows (m % 64 != 0).
struct generic_args
{
  ms_operand A, B;
  const double *dsign;
  double *C;          // row 0 of the launch, column 0
  size_t ldc;
  double *C2;         // optional second output (panel buffer, chunk-major)
  size_t c2_H, c2_k0;
  int K;
  size_t rows;        // rows in the launch
  int cyclic;         // 1: row r is LOCAL row lrow_begin + r of the block-row cyclic layout
  size_t lrow_begin;  //    and its A column / diagonal is the matching GLOBAL row
  size_t m;
  int rank, nranks;
  size_t a_cols;      // columns the A operand has (rows whose A column is >= a_cols contribute 0)
  int tri_mask;       // 1: only columns >= the row's own index are written
  int subtract;       // 1: C -= acc, 0: C = acc
  size_t col_begin;   // first column tile start
  size_t col_lo;      // first column written
  size_t n;           // columns of C / B
};

__device__ __forceinline__ double
operand_at (const ms_operand &o, size_t k, size_t j)
{
  return o.ld ? o.p[k * o.ld + j] : o.p[chunk_major_index (k, j, o.H)];
}

__global__ void __launch_bounds__ (256)
gemm_generic_kernel (const generic_args ga)
{
  __shared__ double As[32][65];
  __shared__ double Bs[32][65];
  __shared__ double Ds[32];
  __shared__ size_t grow[64];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const size_t lr_tile = (size_t)blockIdx.y * 64;
  const size_t col_tile = ga.col_begin + (size_t)blockIdx.x * 64;

  if (threadIdx.x < 64)
    {
      size_t lr = lr_tile + threadIdx.x;
      size_t g = ga.a_cols;  // out of range -> contributes zero
      if (lr < ga.rows)
        {
          if (ga.cyclic)
            {
              size_t l = ga.lrow_begin + lr;
              size_t b = l / ga.m;
              g = (b * (size_t)ga.nranks + (size_t)ga.rank) * ga.m + (l - b * ga.m);
            }
          else
            g = lr;
        }
      grow[threadIdx.x] = g;
    }
  __syncthreads ();
  if (ga.tri_mask && col_tile + 64 <= grow[0])
    return;  // whole tile left of its first row's diagonal: nothing stored there

  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int cidx = 0; cidx < 4; cidx++)
      acc[r][cidx] = 0.0;

  for (int k0 = 0; k0 < ga.K; k0 += 32)
    {
      for (int e = threadIdx.x; e < 32 * 64; e += 256)
        {
          int kk = e >> 6, cc = e & 63;
          int k = k0 + kk;
          double va = 0.0, vb = 0.0;
          if (k < ga.K)
            {
              size_t acol = grow[cc];
              if (acol < ga.a_cols)
                va = operand_at (ga.A, (size_t)k, acol);
              size_t col = col_tile + cc;
              if (col < ga.n)
                vb = operand_at (ga.B, (size_t)k, col);
            }
          As[kk][cc] = va;
          Bs[kk][cc] = vb;
        }
      if (threadIdx.x < 32)
        {
          int k = k0 + threadIdx.x;
          Ds[threadIdx.x] = (ga.dsign != nullptr && k < ga.K) ? ga.dsign[k] : 1.0;
        }
      __syncthreads ();
#pragma unroll 8
      for (int kk = 0; kk < 32; kk++)
        {
          double av[4], bv[4];
#pragma unroll
          for (int r = 0; r < 4; r++)
            av[r] = As[kk][ty * 4 + r] * Ds[kk];
#pragma unroll
          for (int cidx = 0; cidx < 4; cidx++)
            bv[cidx] = Bs[kk][tx + 16 * cidx];
#pragma unroll
          for (int r = 0; r < 4; r++)
#pragma unroll
            for (int cidx = 0; cidx < 4; cidx++)
              acc[r][cidx] += av[r] * bv[cidx];
        }
      __syncthreads ();
    }

#pragma unroll
  for (int r = 0; r < 4; r++)
    {
      int rr = ty * 4 + r;
      size_t lr = lr_tile + rr;
      if (lr >= ga.rows)
        continue;
      size_t g = grow[rr];
      double *crow = ga.C + lr * ga.ldc;
#pragma unroll
      for (int cidx = 0; cidx < 4; cidx++)
        {
          size_t col = col_tile + tx + 16 * cidx;
          if (col >= ga.n || col < ga.col_lo || (ga.tri_mask && col < g))
            continue;
          if (ga.subtract)
            crow[col] -= acc[r][cidx];
          else
            {
              crow[col] = acc[r][cidx];
              if (ga.C2)
                ga.C2[chunk_major_index (ga.c2_k0 + lr, col, ga.c2_H)] = acc[r][cidx];
            }
        }
    }
}

// factorisation update of LOCAL rows [lrow_begin, lrow_end) with the K panel rows of `P`
int
ms_k_generic_update (ms_ctx *c, ms_operand P, const double *dsign, size_t lrow_begin, size_t lrow_end,
                     int K, cudaStream_t s)
{
  if (lrow_end <= lrow_begin)
    return MS_OK;
  size_t lb = lrow_begin / c->m;
  size_t g0 = (lb * (size_t)c->nranks + (size_t)c->rank) * c->m + (lrow_begin - lb * c->m);
  if (g0 >= c->n)
    return MS_OK;
  generic_args ga = {};
  ga.A = P;
  ga.B = P;
  ga.dsign = dsign;
  ga.C = c->A + lrow_begin * c->ld;
  ga.ldc = c->ld;
  ga.K = K;
  ga.rows = lrow_end - lrow_begin;
  ga.cyclic = 1;
  ga.lrow_begin = lrow_begin;
  ga.m = c->m;
  ga.rank = c->rank;
  ga.nranks = c->nranks;
  ga.a_cols = c->n;
  ga.tri_mask = 1;
  ga.subtract = 1;
  ga.n = c->ncols;
  ga.col_lo = 0;
  ga.col_begin = (g0 / 64) * 64;
  size_t col_tiles = (c->ncols - ga.col_begin + 63) / 64;
  dim3 grid ((unsigned)col_tiles, (unsigned)((ga.rows + 63) / 64));
  gemm_generic_kernel<<<grid, 256, 0, s>>> (ga);
  LAUNCH_COUNT (c);
  c->stats.update_launches++;
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// out(h x [col_lo, n)) = Wt^T * src.  In place (out_rows == src_rows) is safe only for h <= 64:
// one CTA then owns all rows of its columns and has read them before it writes.
int
ms_k_generic_panel (ms_ctx *c, ms_operand Wt, const double *src_rows, size_t ld_src, double *out_rows,
                    size_t ld_out, int h, size_t col_lo, double *C2, size_t c2_H, size_t c2_k0, size_t n,
                    cudaStream_t s)
{
  if (col_lo >= n)
    return MS_OK;
  MS_CHECK_ARG (h <= 64 || out_rows != src_rows, "in-place panel needs h <= 64");
  generic_args ga = {};
  ga.A = Wt;
  ga.B.p = src_rows;
  ga.B.ld = ld_src;
  ga.B.H = 0;
  ga.dsign = nullptr;
  ga.C = out_rows;
  ga.ldc = ld_out;
  ga.C2 = C2;
  ga.c2_H = c2_H;
  ga.c2_k0 = c2_k0;
  ga.K = h;
  ga.rows = (size_t)h;
  ga.cyclic = 0;
  ga.m = 1;
  ga.nranks = 1;
  ga.a_cols = (size_t)h;
  ga.tri_mask = 0;
  ga.subtract = 0;
  ga.n = n;
  ga.col_lo = col_lo;
  ga.col_begin = (col_lo / 64) * 64;
  size_t col_tiles = (n - ga.col_begin + 63) / 64;
  dim3 grid ((unsigned)col_tiles, (unsigned)((h + 63) / 64));
  gemm_generic_kernel<<<grid, 256, 0, s>>> (ga);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}

// stand-alone block update on plain row-major blocks: A(h x w) -= R1(K x h)^T diag(D) R2(K x w)
int
ms_k_generic_block_update (ms_ctx *c, double *A, size_t lda, const double *R1, size_t ld1, const double *D,
                           const double *R2, size_t ld2, int K, int h, int w, int upper_only,
                           cudaStream_t s)
{
  generic_args ga = {};
  ga.A.p = R1;
  ga.A.ld = ld1;
  ga.B.p = R2;
  ga.B.ld = ld2;
  ga.dsign = D;
  ga.C = A;
  ga.ldc = lda;
  ga.K = K;
  ga.rows = (size_t)h;
  ga.cyclic = 0;
  ga.m = 1;
  ga.nranks = 1;
  ga.a_cols = (size_t)h;
  ga.tri_mask = upper_only;
  ga.subtract = 1;
  ga.n = (size_t)w;
  dim3 grid ((unsigned)((w + 63) / 64), (unsigned)((h + 63) / 64));
  gemm_generic_kernel<<<grid, 256, 0, s>>> (ga);
  LAUNCH_COUNT (c);
  MS_CUDA (cudaGetLastError ());
  return MS_OK;
}
