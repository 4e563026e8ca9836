// ms_nccl.cu — NCCL-backed exchange steps (see ms_nccl.h).
//
// What replaces what:
//   MPI_Bcast of block column I  (src/solver.cpp:34-35)  -> ncclBroadcast of the factored panel
//        (only columns >= I*m: a contiguous range of the chunk-major panel buffer) + d_I
//   MPI_Bcast of y blocks        (src/solver.cpp:204)    -> ncclReduce of the forward accumulators
//        to the owner per block row, one ncclAllReduce of y at the end
//   MPI_Gatherv per row          (src/init.cpp:401)      -> not reproduced; the backward sweep
//        broadcasts x per block row instead (ms_api.cu)
//   MPI_Allreduce of the flag    (src/init.cpp:573)      -> ncclAllReduce(max) of one int
#include <dlfcn.h>
#include <nccl.h>

#include "ms_nccl.h"

namespace
{
struct nccl_api
{
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId) (ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank) (ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig) (ncclComm_t *, int, ncclUniqueId, int, ncclConfig_t *) = nullptr;
  ncclResult_t (*CommDestroy) (ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast) (const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce) (const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce) (const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather) (const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart) () = nullptr;
  ncclResult_t (*GroupEnd) () = nullptr;
  const char *(*GetErrorString) (ncclResult_t) = nullptr;
};
nccl_api g_nccl;

int
load_nccl ()
{
  if (g_nccl.lib)
    return MS_OK;
  void *h = dlopen ("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h)
    h = dlopen ("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h)
    {
      ms_set_error ("cannot load libnccl.so.2: %s", dlerror ());
      return MS_ERR_NCCL;
    }
#define SYM(field, name)                                                               \
  do                                                                                   \
    {                                                                                  \
      *(void **)(&g_nccl.field) = dlsym (h, name);                                     \
      if (!g_nccl.field)                                                               \
        {                                                                              \
          ms_set_error ("libnccl lacks %s", name);                                     \
          return MS_ERR_NCCL;                                                          \
        }                                                                              \
    }                                                                                  \
  while (0)
  SYM (GetUniqueId, "ncclGetUniqueId");
  SYM (CommInitRank, "ncclCommInitRank");
  SYM (CommDestroy, "ncclCommDestroy");
  SYM (Broadcast, "ncclBroadcast");
  SYM (Reduce, "ncclReduce");
  SYM (AllReduce, "ncclAllReduce");
  SYM (AllGather, "ncclAllGather");
  SYM (GroupStart, "ncclGroupStart");
  SYM (GroupEnd, "ncclGroupEnd");
  SYM (GetErrorString, "ncclGetErrorString");
#undef SYM
  *(void **)(&g_nccl.CommInitRankConfig) = dlsym (h, "ncclCommInitRankConfig");  // optional
  g_nccl.lib = h;
  return MS_OK;
}
}  // namespace

#define MS_NCCL(expr)                                                                  \
  do                                                                                   \
    {                                                                                  \
      ncclResult_t r__ = (expr);                                                       \
      if (r__ != ncclSuccess)                                                          \
        {                                                                              \
          ms_set_error ("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                        g_nccl.GetErrorString (r__));                                  \
          return MS_ERR_NCCL;                                                          \
        }                                                                              \
    }                                                                                  \
  while (0)

static_assert (sizeof (ncclUniqueId) <= MS_UNIQUE_ID_BYTES, "unique id does not fit");

int
ms_nccl_unique_id (void *id128)
{
  MS_TRY (load_nccl ());
  ncclUniqueId id;
  MS_NCCL (g_nccl.GetUniqueId (&id));
  memset (id128, 0, MS_UNIQUE_ID_BYTES);
  memcpy (id128, &id, sizeof (id));
  return MS_OK;
}

int
ms_nccl_init (ms_ctx *c, const void *id128)
{
  MS_TRY (load_nccl ());
  ncclUniqueId id;
  memcpy (&id, id128, sizeof (id));
  ncclComm_t comm;
  // The receivers' broadcast kernels are posted early (look-ahead) and then spin until the owner
  // sends; every CTA they hold is an SM slot the trailing update cannot use.  The panel needs only a
  // few tens of GB/s (134 MB per ~4 ms step at n = 65536 on 8 GPUs), so cap the communicator's CTAs.
  int max_ctas = 0;  // 0 = NCCL default; see profiles/README.md for the sweep
  if (const char *e = getenv ("MS_NCCL_CTAS"))
    max_ctas = atoi (e);
  if (g_nccl.CommInitRankConfig && max_ctas > 0)
    {
      ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
      cfg.minCTAs = 1;
      cfg.maxCTAs = max_ctas;
      MS_NCCL (g_nccl.CommInitRankConfig (&comm, c->nranks, id, c->rank, &cfg));
    }
  else
    MS_NCCL (g_nccl.CommInitRank (&comm, c->nranks, id, c->rank));
  c->nccl = (void *)comm;
  return MS_OK;
}

void
ms_nccl_destroy (ms_ctx *c)
{
  if (c->nccl && g_nccl.CommDestroy)
    g_nccl.CommDestroy ((ncclComm_t)c->nccl);
  c->nccl = nullptr;
}

int
ms_nccl_bcast_panel (ms_ctx *c, size_t I, int pb, int owner, cudaStream_t s)
{
  ncclComm_t comm = (ncclComm_t)c->nccl;
  size_t rows = c->block_rows (I);
  bool has_trailing = (I + 1) * c->m < c->n;
  MS_NCCL (g_nccl.GroupStart ());
  if (has_trailing)
    {
      size_t chunk0 = (I * c->m) / MS_CHUNK;
      size_t off = chunk0 * c->hpb * MS_CHUNK;
      size_t cnt = (c->ld / MS_CHUNK - chunk0) * c->hpb * MS_CHUNK;
      double *p = c->PB[pb] + off;
      MS_NCCL (g_nccl.Broadcast (p, p, cnt, ncclDouble, owner, comm, s));
    }
  double *d = c->vec[MS_VEC_D] + I * c->m;
  MS_NCCL (g_nccl.Broadcast (d, d, rows, ncclDouble, owner, comm, s));
  MS_NCCL (g_nccl.GroupEnd ());
  return MS_OK;
}

int
ms_nccl_bcast (ms_ctx *c, double *p, size_t cnt, int root, cudaStream_t s)
{
  MS_NCCL (g_nccl.Broadcast (p, p, cnt, ncclDouble, root, (ncclComm_t)c->nccl, s));
  return MS_OK;
}

int
ms_nccl_reduce_sum (ms_ctx *c, double *p, size_t cnt, int root, cudaStream_t s)
{
  MS_NCCL (g_nccl.Reduce (p, p, cnt, ncclDouble, ncclSum, root, (ncclComm_t)c->nccl, s));
  return MS_OK;
}

int
ms_nccl_allreduce_sum (ms_ctx *c, double *p, size_t cnt, cudaStream_t s)
{
  MS_NCCL (g_nccl.AllReduce (p, p, cnt, ncclDouble, ncclSum, (ncclComm_t)c->nccl, s));
  return MS_OK;
}

int
ms_nccl_allgather_bytes (ms_ctx *c, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t s)
{
  MS_NCCL (g_nccl.AllGather (send, recv, bytes_per_rank, ncclChar, (ncclComm_t)c->nccl, s));
  return MS_OK;
}

int
ms_nccl_agree_flag (ms_ctx *c, int *flag)
{
  int *d = c->flags + 4;
  MS_CUDA (cudaMemcpyAsync (d, flag, sizeof (int), cudaMemcpyHostToDevice, c->s_main));
  MS_NCCL (g_nccl.AllReduce (d, d, 1, ncclInt, ncclMax, (ncclComm_t)c->nccl, c->s_main));
  MS_CUDA (cudaMemcpyAsync (c->h_flags + 4, d, sizeof (int), cudaMemcpyDeviceToHost, c->s_main));
  MS_CUDA (cudaStreamSynchronize (c->s_main));
  *flag = c->h_flags[4];
  return MS_OK;
}
