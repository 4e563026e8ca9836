// ms_kernels_tile.cu — K3 / K4 / K5: the k-major tile contraction
//
//        C(i, j)  (op)=  sum_k  s_k * A(k, i) * B(k, j)
//
// in two flavours:
//   * tile_dmma_kernel  — the hot path.  64 x 128 output tile per CTA, fp64 tensor cores
//     (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), operands streamed by TMA
//     (cp.async.bulk.tensor.2d, 128-byte swizzle) through a 4-stage mbarrier ring by one
//     producer warp, 8 consumer warps (32 x 32 each, 32 fp64 accumulators per lane), two
//     CTAs per SM so one CTA's read-modify-write epilogue overlaps the other's math.
//   * gemm_generic_kernel — bounds-checked CUDA-core (DFMA) kernel for block sizes the tile
//     kernel does not take (m % 64 != 0).  Same contract, still on the device.
//
// Both serve the two uses of the contraction in the factorisation:
//   MS_MODE_UPDATE (K4 + K5, reference minus_RTDR_o1 / minus_RTDR_l_o1 / minus_RTDRu_ofast /
//       minus_RTDRu_l, src/matrix_op.cpp:4-99,156-242):   C -= P^T D P, A and B both slices of
//       the factored panel P (k = rows of block row I), s_k = d_k.  The diagonal-block (SYRK)
//       case is just the tile whose row and column ranges coincide.
//   MS_MODE_PANEL  (K3, reference DRtA / DRtA_l, src/matrix_op.cpp:287-368):   the block row is
//       overwritten by D V^T A_row, V = inverse of the piece's factor; A = Wt (= V D, emitted by
//       the diag kernel), B = the piece's rows of the matrix, s_k = 1.  Output also goes to the
//       panel buffer that feeds the updates / the NCCL broadcast.
//
// Fragment mapping (why the shared-memory reads are conflict free) is derived in DESIGN.md
// "tile kernel"; in short a lane (p = lane/4, q = lane%4) loads with ONE 16-byte LDS the two
// neighbouring columns 2p, 2p+1 of k-row 8g+2q+e, which under the 128B swizzle sit in 16-byte
// slot (p ^ (2q+e)) of that row: the 8 lanes of every quarter warp hit 8 distinct slots.
#include "ms_common.cuh"

#define LAUNCH_COUNT(c) ((c)->stats.kernel_launches++)

// ============================================================================ DMMA tile kernel
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
  size_t g0;
  size_t col_tile0;         // first column tile (units of MS_BN) of the launch
  size_t col_lo_panel;      // MODE_PANEL: first column to write
  size_t n;                 // columns >= n are never written
};

constexpr int TILE_A_BYTES = (MS_BM / MS_CHUNK) * MS_KC * MS_CHUNK * 8;  // 8 KB
constexpr int TILE_B_BYTES = (MS_BN / MS_CHUNK) * MS_KC * MS_CHUNK * 8;  // 16 KB
constexpr int TILE_STAGE_BYTES = TILE_A_BYTES + TILE_B_BYTES;            // 24 KB
constexpr int TILE_CHUNK_BYTES = MS_KC * MS_CHUNK * 8;                   // 2 KB
constexpr int TILE_MAX_K = 2048;
constexpr int TILE_SMEM_BYTES = MS_STAGES * TILE_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/
                                + TILE_MAX_K * 4;

template <int MODE>
__global__ void __launch_bounds__ (MS_TILE_THREADS, 2)
tile_dmma_kernel (const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const tile_args ta)
{
  const int rt = blockIdx.y;                                   // row tile
  const size_t ct = ta.col_tile0 + blockIdx.x;                 // column tile (global, units of 128)
  size_t g_row0 = ta.g0;                                       // global row of tile row 0
  if (MODE == MS_MODE_UPDATE)
    {
      const size_t l = ta.lrow_begin + (size_t)rt * MS_BM;
      const size_t lb = l / ta.m;
      g_row0 = (lb * (size_t)ta.nranks + (size_t)ta.rank) * ta.m + (l - lb * ta.m);
    }
  const size_t col0 = ct * MS_BN;
  if (MODE == MS_MODE_UPDATE && (col0 + MS_BN <= g_row0 || g_row0 >= ta.n))
    return;  // tile entirely left of the diagonal block (not stored) or below the matrix

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *full_bar = (uint64_t *)(smem + MS_STAGES * TILE_STAGE_BYTES);
  uint64_t *empty_bar = full_bar + MS_STAGES;
  uint32_t *sgn = (uint32_t *)(empty_bar + MS_STAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = ta.K / MS_KC;

  if (threadIdx.x == 0)
    {
      for (int s = 0; s < MS_STAGES; s++)
        {
          mbar_init (&full_bar[s], 1);
          mbar_init (&empty_bar[s], MS_TILE_CONSUMER_WARPS);
        }
      mbar_fence_init ();
    }
  for (int k = threadIdx.x; k < ta.K; k += MS_TILE_THREADS)
    sgn[k] = (ta.dsign != nullptr && ta.dsign[k] < 0.0) ? 0x80000000u : 0u;
  __syncthreads ();

  if (warp == MS_TILE_CONSUMER_WARPS)
    {
      // ------------------------------------------------------------------ TMA producer
      if (lane == 0)
        {
          tma_prefetch_desc (&tmA);
          tma_prefetch_desc (&tmB);
          // A chunk 0 = the 16 columns of the panel that match the tile's first output row
          const int achunk0 = MODE == MS_MODE_UPDATE ? (int)(g_row0 / MS_CHUNK) : 0;
          const int ax = ta.ax0 + achunk0 * ta.axs, ay = ta.ay0 + achunk0 * ta.ays;
          const int bchunk0 = (int)(ct * (MS_BN / MS_CHUNK));
          for (int kb = 0; kb < nkb; kb++)
            {
              const int s = kb % MS_STAGES;
              const uint32_t ph = (uint32_t)((kb / MS_STAGES) & 1);
              mbar_wait (&empty_bar[s], ph ^ 1u);
              mbar_arrive_expect_tx (&full_bar[s], TILE_STAGE_BYTES);
              uint8_t *sa = smem + s * TILE_STAGE_BYTES;
              uint8_t *sb = sa + TILE_A_BYTES;
              const int kk = kb * MS_KC;
#pragma unroll
              for (int ci = 0; ci < MS_BM / MS_CHUNK; ci++)
                tma_load_2d (sa + ci * TILE_CHUNK_BYTES, &tmA, ax + ci * ta.axs, ay + ci * ta.ays + kk,
                             &full_bar[s]);
#pragma unroll
              for (int cj = 0; cj < MS_BN / MS_CHUNK; cj++)
                tma_load_2d (sb + cj * TILE_CHUNK_BYTES, &tmB, ta.bx0 + (bchunk0 + cj) * ta.bxs,
                             ta.by0 + (bchunk0 + cj) * ta.bys + kk, &full_bar[s]);
            }
        }
      return;
    }

  // ---------------------------------------------------------------------- DMMA consumers
  const int wm = warp >> 2;  // 0..1 : rows  wm*32
  const int wn = warp & 3;   // 0..3 : cols  wn*32
  const int p = lane >> 2, q = lane & 3;

  double acc[2][2][2][2][2];  // [ib][jb][a][b][e]
#pragma unroll
  for (int i = 0; i < 32; i++)
    (&acc[0][0][0][0][0])[i] = 0.0;

  for (int kb = 0; kb < nkb; kb++)
    {
      const int s = kb % MS_STAGES;
      const uint32_t ph = (uint32_t)((kb / MS_STAGES) & 1);
      mbar_wait (&full_bar[s], ph);
      const uint8_t *sa = smem + s * TILE_STAGE_BYTES + (wm * 2) * TILE_CHUNK_BYTES;
      const uint8_t *sb = smem + s * TILE_STAGE_BYTES + TILE_A_BYTES + (wn * 2) * TILE_CHUNK_BYTES;
#pragma unroll
      for (int g = 0; g < MS_KC / 8; g++)
        {
#pragma unroll
          for (int e = 0; e < 2; e++)
            {
              const int kr = 8 * g + 2 * q + e;                    // k row inside the stage
              const int off = kr * 128 + ((p ^ (kr & 7)) << 4);    // swizzled 16-byte slot
              double2 fa[2], fb[2];
#pragma unroll
              for (int ib = 0; ib < 2; ib++)
                fa[ib] = *reinterpret_cast<const double2 *> (sa + ib * TILE_CHUNK_BYTES + off);
#pragma unroll
              for (int jb = 0; jb < 2; jb++)
                fb[jb] = *reinterpret_cast<const double2 *> (sb + jb * TILE_CHUNK_BYTES + off);
              if (MODE == MS_MODE_UPDATE)
                {
                  const uint32_t sm = sgn[kb * MS_KC + kr];
#pragma unroll
                  for (int ib = 0; ib < 2; ib++)
                    {
                      fa[ib].x = flip_sign (fa[ib].x, sm);
                      fa[ib].y = flip_sign (fa[ib].y, sm);
                    }
                }
#pragma unroll
              for (int ib = 0; ib < 2; ib++)
#pragma unroll
                for (int jb = 0; jb < 2; jb++)
                  {
                    dmma_8x8x4 (acc[ib][jb][0][0][0], acc[ib][jb][0][0][1], fa[ib].x, fb[jb].x);
                    dmma_8x8x4 (acc[ib][jb][0][1][0], acc[ib][jb][0][1][1], fa[ib].x, fb[jb].y);
                    dmma_8x8x4 (acc[ib][jb][1][0][0], acc[ib][jb][1][0][1], fa[ib].y, fb[jb].x);
                    dmma_8x8x4 (acc[ib][jb][1][1][0], acc[ib][jb][1][1][1], fa[ib].y, fb[jb].y);
                  }
            }
        }
      __syncwarp ();
      if (lane == 0)
        mbar_arrive (&empty_bar[s]);
    }

  // ---------------------------------------------------------------------- epilogue
  // lane owns, per (ib, jb, a): row wm*32 + 16*ib + 2p + a, columns wn*32 + 16*jb + 4q .. +3
  const int rows_valid = g_row0 + MS_BM <= ta.n ? MS_BM : (int)(ta.n - g_row0);
  const size_t col_lo = (MODE == MS_MODE_UPDATE) ? g_row0 : ta.col_lo_panel;
  double *Ct = ta.C + (size_t)rt * MS_BM * ta.ldc;
#pragma unroll
  for (int ib = 0; ib < 2; ib++)
#pragma unroll
    for (int a = 0; a < 2; a++)
      {
        const int r = wm * 32 + 16 * ib + 2 * p + a;
        if (r >= rows_valid)
          continue;
        double *crow = Ct + (size_t)r * ta.ldc;
#pragma unroll
        for (int jb = 0; jb < 2; jb++)
          {
            const size_t cc = col0 + wn * 32 + 16 * jb + 4 * q;
            double v[4] = {acc[ib][jb][a][0][0], acc[ib][jb][a][1][0], acc[ib][jb][a][0][1],
                           acc[ib][jb][a][1][1]};
            if (cc >= col_lo && cc + 3 < ta.n)
              {
                double2 *ptr = reinterpret_cast<double2 *> (crow + cc);
                if (MODE == MS_MODE_UPDATE)
                  {
                    double2 o0 = ptr[0], o1 = ptr[1];
                    o0.x -= v[0];
                    o0.y -= v[1];
                    o1.x -= v[2];
                    o1.y -= v[3];
                    ptr[0] = o0;
                    ptr[1] = o1;
                  }
                else
                  {
                    ptr[0] = make_double2 (v[0], v[1]);
                    ptr[1] = make_double2 (v[2], v[3]);
                    double2 *p2 = reinterpret_cast<double2 *> (
                        ta.C2 + chunk_major_index (ta.c2_k0 + r, cc, ta.c2_H));
                    p2[0] = make_double2 (v[0], v[1]);
                    p2[1] = make_double2 (v[2], v[3]);
                  }
              }
            else
              {
#pragma unroll
                for (int t = 0; t < 4; t++)
                  {
                    const size_t col = cc + t;
                    if (col >= col_lo && col < ta.n)
                      {
                        if (MODE == MS_MODE_UPDATE)
                          crow[col] -= v[t];
                        else
                          {
                            crow[col] = v[t];
                            ta.C2[chunk_major_index (ta.c2_k0 + r, col, ta.c2_H)] = v[t];
                          }
                      }
                  }
              }
          }
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

int
ms_tile_setup (ms_ctx *c)
{
  c->have_tm = false;
  if (!c->fast)
    return MS_OK;
  MS_TRY (make_map_2d (&c->tm_A, c->A, c->ld, c->rows_alloc, c->ld * 8));
  for (int b = 0; b < 2; b++)
    MS_TRY (make_map_2d (&c->tm_PB[b], c->PB[b], MS_CHUNK, (c->ld / MS_CHUNK) * c->hpb, MS_CHUNK * 8));
  MS_TRY (make_map_2d (&c->tm_Wt, c->Wt, MS_CHUNK, (MS_SUB / MS_CHUNK) * MS_SUB, MS_CHUNK * 8));
  MS_CUDA (cudaFuncSetAttribute (tile_dmma_kernel<MS_MODE_UPDATE>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
  MS_CUDA (cudaFuncSetAttribute (tile_dmma_kernel<MS_MODE_PANEL>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
  c->have_tm = true;
  return MS_OK;
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
  MS_CHECK_ARG (K % MS_KC == 0 && K <= TILE_MAX_K, "K must be a multiple of 16 and <= 2048");
  MS_CHECK_ARG (lrow_begin % MS_BM == 0, "row range must start on a tile boundary");
  size_t lb = lrow_begin / c->m;
  size_t g0 = (lb * (size_t)c->nranks + (size_t)c->rank) * c->m + (lrow_begin - lb * c->m);
  if (g0 >= c->n)
    return MS_OK;
  tile_args ta = {};
  ta.ax0 = 0; ta.axs = 0; ta.ay0 = (int)k0; ta.ays = (int)c->hpb;
  ta.bx0 = 0; ta.bxs = 0; ta.by0 = (int)k0; ta.bys = (int)c->hpb;
  ta.K = K;
  ta.dsign = c->vec[MS_VEC_D] + I * c->m + k0;
  ta.C = c->A + lrow_begin * c->ld;
  ta.ldc = c->ld;
  ta.C2 = nullptr;
  ta.lrow_begin = lrow_begin;
  ta.m = c->m;
  ta.rank = c->rank;
  ta.nranks = c->nranks;
  ta.g0 = g0;
  ta.col_tile0 = g0 / MS_BN;
  ta.n = c->n;
  size_t col_tiles = (c->n + MS_BN - 1) / MS_BN - ta.col_tile0;
  size_t row_tiles = (lrow_end - lrow_begin + MS_BM - 1) / MS_BM;
  dim3 grid ((unsigned)col_tiles, (unsigned)row_tiles);
  tile_dmma_kernel<MS_MODE_UPDATE><<<grid, MS_TILE_THREADS, TILE_SMEM_BYTES, s>>> (c->tm_PB[pb], c->tm_PB[pb], ta);
  LAUNCH_COUNT (c);
  c->stats.update_launches++;
  c->stats.used_dmma = 1;
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
  if (col_lo >= c->n)
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
  ta.g0 = pc.row0;
  ta.col_tile0 = col_lo / MS_BN;
  ta.col_lo_panel = col_lo;
  ta.n = c->n;
  size_t col_tiles = (c->n + MS_BN - 1) / MS_BN - ta.col_tile0;
  dim3 grid ((unsigned)col_tiles, 1);
  tile_dmma_kernel<MS_MODE_PANEL><<<grid, MS_TILE_THREADS, TILE_SMEM_BYTES, s>>> (c->tm_Wt, c->tm_A, ta);
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
  ga.n = c->n;
  ga.col_lo = 0;
  ga.col_begin = (g0 / 64) * 64;
  size_t col_tiles = (c->n - ga.col_begin + 63) / 64;
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
