/*
 * kgo.h -- CPU ORACLE for the kernelgen-perf-tests stencil hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Parity status: pinned against the reference's golden vectors
 * (README.md:113-127 i_mean/f_mean at 512 256 256 10, double) and against the
 * reference sources themselves compiled into oracle/_ref (see Makefile) for
 * the 12 C tests.  jacobi, sincos, matmul are Fortran in the reference
 * (gfortran absent): for those three the oracle is "parity pinned by
 * restatement + the reference's own C driver (jacobi/main.c) only".
 */
#ifndef KGO_H
#define KGO_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Test ids -- same numbering as include/b200_stencil.h (kept in sync by
 * tests/test_abi.py), but deliberately a separate definition: the product
 * never includes this header. */
enum {
    KGO_LAPLACIAN = 0, KGO_WAVE13PT, KGO_DIVERGENCE, KGO_GRADIENT, KGO_UXX1,
    KGO_LAPGSRB, KGO_JACOBI, KGO_GAUSSBLUR, KGO_GAMEOFLIFE, KGO_TRICUBIC,
    KGO_TRICUBIC2, KGO_VECADD, KGO_MATVEC, KGO_SINCOS, KGO_MATMUL, KGO_NTESTS
};
enum { KGO_F32 = 0, KGO_F64 = 1 };

typedef struct {
    const char* name;
    int ndims;      /* 3: <nx> <ny> <ns> <nt>;  2: <nx> <ny> <nt> */
    int narrays;    /* arrays in driver init order */
    int nscalars;   /* rand()-drawn scalars, in draw order */
    int rotation;   /* 0 none, 2 swap slots 0/1, 3 rotate slots 0/1/2 */
} kgo_test_info;

const kgo_test_info* kgo_info(int test);

/* srand(1): puts glibc rand() back into its never-seeded state. */
void kgo_reseed(void);

/* Number of elements of array `slot` (all equal except matvec: A, x, y). */
size_t kgo_array_len(int test, int slot, int nx, int ny, int ns);

/* Draw scalars then fill arrays exactly in the reference driver's rand()
 * order (e.g. laplacian/laplacian.c:141-164).  `arrays` are caller-allocated.
 * Returns the value the driver prints as "initial mean". */
double kgo_init(int test, int dtype, int nx, int ny, int ns,
                double* scalars, void* const* arrays);

/* One sweep, arrays in canonical (driver init) order, no rotation. */
int kgo_sweep(int test, int dtype, int nx, int ny, int ns,
              const double* scalars, void* const* arrays);

/* nt sweeps with the reference's pointer rotation; returns the slot the
 * reference reports its final mean on (-1 on error).  gradient: returns 1
 * (ux) and the mean is over ux+uy+uz. */
int kgo_run(int test, int dtype, int nx, int ny, int ns, int nt,
            const double* scalars, void* const* arrays);

/* "final mean" exactly as the driver computes it, given arrays after
 * kgo_run and the slot it returned. */
double kgo_final_mean(int test, int dtype, int nx, int ny, int ns,
                      void* const* arrays, int slot);

/* Whole driver: alloc + init + nt sweeps + final mean. Returns 0 on success. */
int kgo_driver(int test, int dtype, int nx, int ny, int ns, int nt,
               double* scalars, double* i_mean, double* f_mean);

int kgo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
