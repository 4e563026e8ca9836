// matrix_op.h — block operations, printing and norms of the drop-in API
// (names and signatures of the reference's src/matrix_op.h:6-53; blocks are HOST buffers in the
// reference's layouts, each call runs the matching CUDA kernel on them).
#ifndef MATRIX_OP
#define MATRIX_OP
#include <cstdio>

#include "defines.h"

// A(n x n) -= R1^T diag(D) R2
void minus_RTDR_o1 (mat A, mat R1, vec D, mat R2, int n);
// packed-upper A -= upper(R1^T diag(D) R2)   (the callers pass R2 == R1)
void minus_RTDRu_ofast (mat A, mat R1, vec D, mat R2, int n);
// A(n x l) -= R1(n x n)^T diag(D) R2(n x l)
void minus_RTDR_l_o1 (mat A, mat R1, vec D, mat R2, int n, int l);
// packed-upper A(l x l) -= upper(R1(n x l)^T diag(D) R1)
void minus_RTDRu_l (mat A, mat R1, vec D, int n, int l);
// B = A^{-1} for packed-upper n x n; true = tiny pivot
bool reverse_upper (mat A, mat B, int n, double norma);
// A <- D R^T A with R the packed-upper INVERSE factor (n x n resp. n x l block)
void DRtA (mat R, mat A, vec D, int n);
void DRtA_l (mat R, mat A, vec D, int n, int l);

void print_mat_beauty (int root, size_t n, size_t m, double **&rows_p, size_t len = -1, bool log = 0,
                       int width = 8, int precision = 2);
void print_vec (vec a, size_t n, size_t len = -1UL);
double norma_vec2 (double *r, int n);
double find_error (double *x, int n);
#endif
