/* minimpi.c — fork()+shared-memory implementation of the MPI subset in mpi.h.
 *
 * TEST/BASELINE INFRASTRUCTURE ONLY (see mpi.h).  Ranks = processes forked in
 * MPI_Init (count from $MINIMPI_NP, default 1); rank 0 is the original process and
 * reaps the others in MPI_Finalize.  All collectives are blocking and use one
 * staging buffer in an anonymous MAP_SHARED mapping, moved in chunks, separated by a
 * sense-reversing spin barrier.  Point-to-point uses one mailbox per destination.
 *
 * Semantics relied on by the reference: every rank calls the
 * collectives in the same order; Send/Recv pairs are matched by (source, dest) only.
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <sched.h>
#include <stdint.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#define STAGE_BYTES ((size_t)32 << 20)
#define MBOX_BYTES ((size_t)1 << 20)
#define MAX_RANKS 1024

typedef struct
{
  int state; /* 0 empty, 1 being written, 2 full */
  int src;
  size_t bytes;
  char data[MBOX_BYTES];
} mailbox_t;

typedef struct
{
  int np;
  int bar_count;
  int bar_sense;
  int counts[MAX_RANKS]; /* per-rank scratch for Gatherv */
  char stage[STAGE_BYTES];
  mailbox_t mbox[];
} shared_t;

static shared_t *g_sh = NULL;
static int g_rank = 0;
static int g_np = 1;
static int g_local_sense = 0;
static pid_t g_children[MAX_RANKS];

static size_t
type_size (MPI_Datatype t)
{
  switch (t)
    {
    case MPI_DOUBLE:
      return sizeof (double);
    case MPI_INT:
      return sizeof (int);
    case MPI_UNSIGNED_LONG:
      return sizeof (unsigned long);
    default:
      fprintf (stderr, "minimpi: unsupported datatype %d\n", t);
      abort ();
    }
}

static inline void
relax (unsigned *spins)
{
  __builtin_ia32_pause ();
  if (++*spins > 2000)
    {
      sched_yield ();
      *spins = 0;
    }
}

static void
barrier (void)
{
  if (g_np == 1)
    return;
  g_local_sense ^= 1;
  if (__atomic_add_fetch (&g_sh->bar_count, 1, __ATOMIC_ACQ_REL) == g_np)
    {
      __atomic_store_n (&g_sh->bar_count, 0, __ATOMIC_RELAXED);
      __atomic_store_n (&g_sh->bar_sense, g_local_sense, __ATOMIC_RELEASE);
    }
  else
    {
      unsigned spins = 0;
      while (__atomic_load_n (&g_sh->bar_sense, __ATOMIC_ACQUIRE)
             != g_local_sense)
        relax (&spins);
    }
}

int
MPI_Init (int *argc, char ***argv)
{
  (void)argc;
  (void)argv;
  const char *e = getenv ("MINIMPI_NP");
  g_np = e ? atoi (e) : 1;
  if (g_np < 1)
    g_np = 1;
  if (g_np > MAX_RANKS)
    g_np = MAX_RANKS;
  g_rank = 0;
  if (g_np == 1)
    return MPI_SUCCESS;
  size_t bytes = sizeof (shared_t) + (size_t)g_np * sizeof (mailbox_t);
  void *p = mmap (NULL, bytes, PROT_READ | PROT_WRITE,
                  MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (p == MAP_FAILED)
    {
      perror ("minimpi: mmap");
      abort ();
    }
  g_sh = (shared_t *)p; /* anonymous mappings start zeroed */
  g_sh->np = g_np;
  fflush (NULL);
  for (int r = 1; r < g_np; r++)
    {
      pid_t pid = fork ();
      if (pid < 0)
        {
          perror ("minimpi: fork");
          abort ();
        }
      if (pid == 0)
        {
          g_rank = r;
          break;
        }
      g_children[r] = pid;
    }
  barrier ();
  return MPI_SUCCESS;
}

int
MPI_Finalize (void)
{
  fflush (NULL);
  if (g_np == 1)
    return MPI_SUCCESS;
  barrier ();
  if (g_rank == 0)
    {
      for (int r = 1; r < g_np; r++)
        {
          int st;
          waitpid (g_children[r], &st, 0);
        }
    }
  return MPI_SUCCESS;
}

int
MPI_Comm_rank (MPI_Comm comm, int *rank)
{
  (void)comm;
  *rank = g_rank;
  return MPI_SUCCESS;
}

int
MPI_Comm_size (MPI_Comm comm, int *size)
{
  (void)comm;
  *size = g_np;
  return MPI_SUCCESS;
}

int
MPI_Barrier (MPI_Comm comm)
{
  (void)comm;
  barrier ();
  return MPI_SUCCESS;
}

int
MPI_Bcast (void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
  (void)comm;
  if (g_np == 1)
    return MPI_SUCCESS;
  size_t total = (size_t)count * type_size (type);
  char *p = (char *)buf;
  size_t off = 0;
  do
    {
      size_t len = total - off < STAGE_BYTES ? total - off : STAGE_BYTES;
      if (g_rank == root)
        memcpy (g_sh->stage, p + off, len);
      barrier ();
      if (g_rank != root)
        memcpy (p + off, g_sh->stage, len);
      barrier ();
      off += len;
    }
  while (off < total);
  return MPI_SUCCESS;
}

int
MPI_Send (const void *buf, int count, MPI_Datatype type, int dest, int tag,
          MPI_Comm comm)
{
  (void)tag;
  (void)comm;
  size_t total = (size_t)count * type_size (type);
  const char *p = (const char *)buf;
  mailbox_t *mb = &g_sh->mbox[dest];
  size_t off = 0;
  do
    {
      size_t len = total - off < MBOX_BYTES ? total - off : MBOX_BYTES;
      unsigned spins = 0;
      int expect = 0;
      while (!__atomic_compare_exchange_n (&mb->state, &expect, 1, 0,
                                           __ATOMIC_ACQUIRE, __ATOMIC_RELAXED))
        {
          expect = 0;
          relax (&spins);
        }
      mb->src = g_rank;
      mb->bytes = len;
      memcpy (mb->data, p + off, len);
      __atomic_store_n (&mb->state, 2, __ATOMIC_RELEASE);
      off += len;
    }
  while (off < total);
  return MPI_SUCCESS;
}

int
MPI_Recv (void *buf, int count, MPI_Datatype type, int source, int tag,
          MPI_Comm comm, MPI_Status *status)
{
  (void)tag;
  (void)comm;
  (void)status;
  size_t total = (size_t)count * type_size (type);
  char *p = (char *)buf;
  mailbox_t *mb = &g_sh->mbox[g_rank];
  size_t off = 0;
  do
    {
      unsigned spins = 0;
      while (!(__atomic_load_n (&mb->state, __ATOMIC_ACQUIRE) == 2
               && mb->src == source))
        relax (&spins);
      size_t len = mb->bytes;
      if (off + len > total)
        {
          fprintf (stderr, "minimpi: message truncated\n");
          abort ();
        }
      memcpy (p + off, mb->data, len);
      __atomic_store_n (&mb->state, 0, __ATOMIC_RELEASE);
      off += len;
    }
  while (off < total);
  return MPI_SUCCESS;
}

int
MPI_Gatherv (const void *sendbuf, int sendcount, MPI_Datatype sendtype,
             void *recvbuf, const int *recvcounts, const int *displs,
             MPI_Datatype recvtype, int root, MPI_Comm comm)
{
  (void)comm;
  size_t ts = type_size (sendtype);
  if (g_np == 1)
    {
      memcpy ((char *)recvbuf + (size_t)displs[0] * type_size (recvtype),
              sendbuf, (size_t)sendcount * ts);
      return MPI_SUCCESS;
    }
  g_sh->counts[g_rank] = sendcount;
  barrier ();
  size_t my_off = 0, all = 0;
  for (int r = 0; r < g_np; r++)
    {
      if (r < g_rank)
        my_off += (size_t)g_sh->counts[r];
      all += (size_t)g_sh->counts[r];
    }
  if (all * ts > STAGE_BYTES)
    {
      fprintf (stderr, "minimpi: Gatherv payload exceeds staging buffer\n");
      abort ();
    }
  memcpy (g_sh->stage + my_off * ts, sendbuf, (size_t)sendcount * ts);
  barrier ();
  if (g_rank == root)
    {
      size_t off = 0;
      for (int r = 0; r < g_np; r++)
        {
          size_t c = (size_t)g_sh->counts[r];
          if (c > (size_t)recvcounts[r])
            {
              fprintf (stderr, "minimpi: Gatherv count mismatch\n");
              abort ();
            }
          memcpy ((char *)recvbuf + (size_t)displs[r] * ts,
                  g_sh->stage + off * ts, c * ts);
          off += c;
        }
    }
  barrier ();
  return MPI_SUCCESS;
}

static void
reduce_sum (const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
            int root_or_all)
{
  size_t ts = type_size (type);
  size_t bytes = (size_t)count * ts;
  if (g_np == 1)
    {
      memcpy (recvbuf, sendbuf, bytes);
      return;
    }
  if (bytes * (size_t)g_np > STAGE_BYTES)
    {
      fprintf (stderr, "minimpi: reduction too large\n");
      abort ();
    }
  memcpy (g_sh->stage + (size_t)g_rank * bytes, sendbuf, bytes);
  barrier ();
  if (root_or_all < 0 || root_or_all == g_rank)
    {
      for (int i = 0; i < count; i++)
        {
          if (type == MPI_DOUBLE)
            {
              double s = 0;
              for (int r = 0; r < g_np; r++)
                s += ((double *)(g_sh->stage + (size_t)r * bytes))[i];
              ((double *)recvbuf)[i] = s;
            }
          else if (type == MPI_INT)
            {
              int s = 0;
              for (int r = 0; r < g_np; r++)
                s += ((int *)(g_sh->stage + (size_t)r * bytes))[i];
              ((int *)recvbuf)[i] = s;
            }
          else
            {
              unsigned long s = 0;
              for (int r = 0; r < g_np; r++)
                s += ((unsigned long *)(g_sh->stage + (size_t)r * bytes))[i];
              ((unsigned long *)recvbuf)[i] = s;
            }
        }
    }
  barrier ();
}

int
MPI_Allreduce (const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
               MPI_Op op, MPI_Comm comm)
{
  (void)op;
  (void)comm;
  reduce_sum (sendbuf, recvbuf, count, type, -1);
  return MPI_SUCCESS;
}

int
MPI_Reduce (const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
            MPI_Op op, int root, MPI_Comm comm)
{
  (void)op;
  (void)comm;
  reduce_sum (sendbuf, recvbuf, count, type, root);
  return MPI_SUCCESS;
}
