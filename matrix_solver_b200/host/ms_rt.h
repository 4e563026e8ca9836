// ms_rt.h — rank runtime of the host mirror.
//
// The reference is an MPI program: `mpirun -np P test_mat ...` gives P ranks and block column J
// lives on rank J % P.  Here a rank is a process that drives ONE GPU; `MS_NGPU=P test_mat ...`
// (or ms_rt_init from any host program) forks P-1 siblings before CUDA is touched, rank r binds
// device r, and the ranks bootstrap one NCCL communicator through a page of shared memory.
// All data-path exchanges then happen inside the CUDA library over NVLink (ms_nccl.cu).
#ifndef MS_RT_H
#define MS_RT_H
#include <cstddef>

#include "ms_b200.h"

int ms_rt_init ();                 // idempotent; returns this process's rank
void ms_rt_finalize ();            // joins the sibling processes on rank 0
int ms_rt_rank ();
int ms_rt_size ();
void ms_rt_barrier ();
unsigned long long ms_rt_sum (unsigned long long v);  // all-reduce(sum) over the ranks, host side
// the problem registered by alloc_rows for this rows_p array (nullptr if unknown)
ms_ctx *ms_rt_ctx (double **rows_p);
ms_ctx *ms_rt_bind (double **rows_p, size_t n, size_t m, double **token_out);
void ms_rt_release_all ();
[[noreturn]] void ms_rt_die (const char *what, int code);
#endif
