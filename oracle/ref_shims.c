/*
 * ref_shims.c -- ORACLE-side glue (test infrastructure, NOT product code).
 *
 * jacobi's kernel in the reference is Fortran (jacobi/jacobi.F90:15-71) called
 * from C through the gfortran by-reference ABI (jacobi/main.c:43-52, call at
 * :245).  gfortran is absent, so the reference's own C driver is linked
 * against the oracle's restatement through this shim.
 */
#include "kgo.h"

void kgo_jacobi_f(int nx, int ny, float c0, float c1, float c2, const float* w0, float* w1);
void kgo_jacobi_d(int nx, int ny, double c0, double c1, double c2, const double* w0, double* w1);

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

#ifndef real
#error "compile with -Dreal=float|double -DKGREF_SHIM_SFX=f|d"
#endif

void jacobi_(int* nx, int* ny, real* c0, real* c1, real* c2, real* w0, real* w1)
{
    CAT(kgo_jacobi_, KGREF_SHIM_SFX)(*nx, *ny, *c0, *c1, *c2, w0, w1);
}
