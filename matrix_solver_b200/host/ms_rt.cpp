// ms_rt.cpp — see ms_rt.h.
#include "ms_rt.h"

#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

namespace
{
struct shared_page
{
  std::atomic<int> arrived;
  std::atomic<int> sense;
  std::atomic<int> id_ready;
  std::atomic<int> abort_flag;             // a rank died: barrier spinners leave instead of waiting forever
  std::atomic<unsigned long long> sum;     // scratch of ms_rt_sum
  char unique_id[MS_UNIQUE_ID_BYTES];
};

shared_page *g_page = nullptr;
int g_rank = 0, g_size = 1, g_local_sense = 0;
bool g_started = false;
std::vector<pid_t> g_kids;

struct problem
{
  ms_ctx *ctx;
  void *reserved;       // PROT_NONE address range standing in for the reference's host body
  size_t reserved_bytes;
};
std::map<double **, problem> g_problems;
volatile sig_atomic_t g_kids_left = 0;

// A rank that dies (ms_rt_die, a crash, an abort inside CUDA/NCCL) must not leave its peers spinning in
// a barrier, inside an NCCL kernel or in cuStreamWaitValue32 for ever: every child dies with its parent
// (PR_SET_PDEATHSIG), and the parent turns the first abnormal child exit into the end of the whole job.
void
kill_kids ()
{
  for (pid_t k : g_kids)
    kill (k, SIGKILL);
}

void
on_child_exit (int)
{
  int st;
  pid_t pid;
  while ((pid = waitpid (-1, &st, WNOHANG)) > 0)
    {
      if (g_kids_left > 0)
        g_kids_left = g_kids_left - 1;
      if (!(WIFEXITED (st) && WEXITSTATUS (st) == 0))
        {
          if (g_page)
            g_page->abort_flag.store (1);
          kill_kids ();
          static const char msg[] = "matrix_solver_b200: a rank terminated abnormally; stopping all ranks\n";
          ssize_t w = write (2, msg, sizeof (msg) - 1);
          (void)w;
          _exit (3);
        }
    }
}

size_t
ref_alloc_elems (size_t n, size_t m, int rank, int size)
{
  size_t cols = n / m + (n % m > 0), total = 0;
  for (size_t J = 0; J < cols; J++)
    {
      size_t w = (J + 1) * m <= n ? m : n - J * m;
      if ((int)(J % (size_t)size) == rank)
        total += J * m * w + w * (w + 1) / 2;
    }
  return total;
}
}  // namespace

void
ms_rt_die (const char *what, int code)
{
  fflush (stdout);
  fprintf (stderr, "matrix_solver_b200: %s (code %d): %s\n", what, code, ms_last_error ());
  fflush (stderr);
  if (g_page)
    g_page->abort_flag.store (1);
  if (g_rank == 0)
    kill_kids ();  // the children of a dead rank 0 would be killed by PDEATHSIG anyway; do not wait for it
  _exit (3);
}

int
ms_rt_init ()
{
  if (g_started)
    return g_rank;
  g_started = true;
  const char *e = getenv ("MS_NGPU");
  g_size = e ? atoi (e) : 1;
  if (g_size < 1)
    g_size = 1;
  if (g_size > 1)
    {
      void *p = mmap (nullptr, sizeof (shared_page), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
      if (p == MAP_FAILED)
        {
          perror ("mmap");
          exit (3);
        }
      g_page = new (p) shared_page ();
      fflush (nullptr);
      struct sigaction sa;
      memset (&sa, 0, sizeof (sa));
      sa.sa_handler = on_child_exit;
      sa.sa_flags = SA_RESTART | SA_NOCLDSTOP;
      sigaction (SIGCHLD, &sa, nullptr);
      const pid_t parent = getpid ();
      for (int r = 1; r < g_size; r++)
        {
          pid_t pid = fork ();  // must happen before the first CUDA call of the process
          if (pid < 0)
            {
              perror ("fork");
              exit (3);
            }
          if (pid == 0)
            {
              g_rank = r;
              g_kids.clear ();
              signal (SIGCHLD, SIG_DFL);
              prctl (PR_SET_PDEATHSIG, SIGKILL);
              if (getppid () != parent)  // the parent died between fork and prctl
                _exit (3);
              break;
            }
          g_kids.push_back (pid);
          g_kids_left = g_kids_left + 1;
        }
    }
  return g_rank;
}

void
ms_rt_barrier ()
{
  if (g_size == 1)
    return;
  g_local_sense ^= 1;
  if (g_page->arrived.fetch_add (1) + 1 == g_size)
    {
      g_page->arrived.store (0);
      g_page->sense.store (g_local_sense);
    }
  else
    while (g_page->sense.load () != g_local_sense)
      {
        if (g_page->abort_flag.load ())
          {
            fflush (stdout);
            _exit (3);  // a peer died (ms_rt_die / crash): nobody will complete this barrier
          }
        sched_yield ();
      }
}

// sum of one value per rank, the same on every rank afterwards (the reference's MPI_Reduce +
// the root's check, test/test_mat.cpp:78-88, without needing a device)
unsigned long long
ms_rt_sum (unsigned long long v)
{
  if (g_size == 1)
    return v;
  ms_rt_barrier ();             // the previous round's reader has left
  if (g_rank == 0)
    g_page->sum.store (0);
  ms_rt_barrier ();
  g_page->sum.fetch_add (v);
  ms_rt_barrier ();
  return g_page->sum.load ();
}

void
ms_rt_release_all ()
{
  for (auto &kv : g_problems)
    {
      ms_destroy (kv.second.ctx);
      if (kv.second.reserved)
        munmap (kv.second.reserved, kv.second.reserved_bytes);
    }
  g_problems.clear ();
}

void
ms_rt_finalize ()
{
  fflush (nullptr);
  ms_rt_release_all ();
  if (g_size > 1)
    {
      ms_rt_barrier ();
      if (g_rank == 0)
        {
          // the SIGCHLD handler reaps; wait until every child is gone
          while (g_kids_left > 0)
            usleep (1000);
          signal (SIGCHLD, SIG_DFL);
          g_kids.clear ();
        }
      else
        {
          fflush (nullptr);
          _exit (0);
        }
    }
  g_started = false;
}

int
ms_rt_rank ()
{
  return g_rank;
}

int
ms_rt_size ()
{
  return g_size;
}

ms_ctx *
ms_rt_ctx (double **rows_p)
{
  auto it = g_problems.find (rows_p);
  return it == g_problems.end () ? nullptr : it->second.ctx;
}

// Creates the device-resident problem for `rows_p` and fills rows_p[J] with addresses inside a
// reserved, inaccessible range laid out like the reference's per-rank body (src/init.cpp:29-58):
// pointer arithmetic such as get_bl() keeps working, dereferencing faults instead of reading junk.
ms_ctx *
ms_rt_bind (double **rows_p, size_t n, size_t m, double **token_out)
{
  ms_rt_init ();
  auto old = g_problems.find (rows_p);
  if (old != g_problems.end ())
    {
      ms_destroy (old->second.ctx);
      if (old->second.reserved)
        munmap (old->second.reserved, old->second.reserved_bytes);
      g_problems.erase (old);
    }
  char id[MS_UNIQUE_ID_BYTES] = {0};
  if (g_size > 1)
    {
      if (g_rank == 0)
        {
          int r = ms_get_unique_id (id);
          if (r < 0)
            ms_rt_die ("ms_get_unique_id", r);
          memcpy (g_page->unique_id, id, sizeof (id));
        }
      ms_rt_barrier ();
      memcpy (id, g_page->unique_id, sizeof (id));
      ms_rt_barrier ();
    }
  int base = getenv ("MS_DEVICE") ? atoi (getenv ("MS_DEVICE")) : 0;
  ms_ctx *ctx = nullptr;
  int r = ms_create (&ctx, n, m, base + g_rank, g_rank, g_size, g_size > 1 ? id : nullptr);
  if (r < 0)
    ms_rt_die ("ms_create", r);

  problem pr;
  pr.ctx = ctx;
  size_t elems = ref_alloc_elems (n, m, g_rank, g_size);
  pr.reserved_bytes = (elems + 1) * sizeof (double);
  pr.reserved = mmap (nullptr, pr.reserved_bytes, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (pr.reserved == MAP_FAILED)
    pr.reserved = nullptr;
  double *cursor = (double *)pr.reserved;
  size_t cols = n / m + (n % m > 0);
  for (size_t J = 0; J < cols; J++)
    {
      size_t w = (J + 1) * m <= n ? m : n - J * m;
      if ((int)(J % (size_t)g_size) == g_rank && cursor)
        {
          rows_p[J] = cursor;
          cursor += J * m * w + w * (w + 1) / 2;
        }
      else
        rows_p[J] = nullptr;
    }
  g_problems[rows_p] = pr;
  if (token_out)
    *token_out = new double[1]();  // what the caller's unique_ptr<double[]> will delete[]
  return ctx;
}
