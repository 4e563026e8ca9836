// Host-side rank runtime (matrix_solver_b200/host/ms_rt.cpp) without a GPU: MS_NGPU ranks are
// forked, agree on rank/size, pass barriers in lock-step, sum a value over the ranks and rank 0 reaps
// the others.  `rt_ranks_test die <rank>`: that rank dies through ms_rt_die half way; its peers are
// parked in a barrier — the whole job must end with exit code 3 instead of hanging.
#include <cstdio>
#include <cstdlib>
#include <unistd.h>

#include "ms_rt.h"

int
main (int argc, char **argv)
{
  int rank = ms_rt_init ();
  int size = ms_rt_size ();
  const int die_rank = argc == 3 ? atoi (argv[2]) : -1;
  for (int round = 0; round < 50; round++)
    {
      if (round == 25 && rank == die_rank)
        ms_rt_die ("test: simulated failure", -1);
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
  for (int round = 0; round < 3; round++)
    {
      unsigned long long total = ms_rt_sum ((unsigned long long)(rank + 1) * (round + 1));
      if (total != (unsigned long long)size * (size + 1) / 2 * (round + 1))
        {
          printf ("rank %d: bad sum %llu in round %d\n", rank, total, round);
          return 4;
        }
    }
  ms_rt_finalize ();  // ranks > 0 exit inside; rank 0 returns after reaping them
  printf ("rank 0 done, size %d\n", size);
  return 0;
}
