// Host-side rank runtime (matrix_solver_b200/host/ms_rt.cpp) without a GPU: MS_NGPU ranks are
// forked, agree on rank/size, pass barriers in lock-step and rank 0 reaps the others.
#include <cstdio>
#include <cstdlib>
#include <unistd.h>

#include "ms_rt.h"

int
main ()
{
  int rank = ms_rt_init ();
  int size = ms_rt_size ();
  for (int round = 0; round < 50; round++)
    {
      if (round % size == rank)
        usleep (200);  // stagger: a broken barrier would let ranks run ahead and interleave rounds
      ms_rt_barrier ();
      if (rank == round % size)
        {
          printf ("round %d by rank %d of %d\n", round, rank, size);
          fflush (stdout);
        }
      ms_rt_barrier ();
    }
  ms_rt_finalize ();  // ranks > 0 exit inside; rank 0 returns after reaping them
  printf ("rank 0 done, size %d\n", size);
  return 0;
}
