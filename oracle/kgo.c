/*
 * kgo.c -- CPU ORACLE dispatcher (test infrastructure, NOT product code).
 * See kgo.h for the contract and kgo_impl.inc for the kernel restatements.
 */
#include "kgo.h"

#include <malloc.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* arrays in driver init order (== the per-element rand() interleave order):
 *  laplacian w0,w1 | wave13pt w0,w1,w2 | divergence u,ux,uy,uz |
 *  gradient u,ux,uy,uz | uxx1 u0,u1,d1,xx,xy,xz | lapgsrb w0,w1 |
 *  jacobi w0,w1 | gaussblur w0,w1 | gameoflife u0,u1 | tricubic[2] u0,u1,a,b,c |
 *  vecadd w0,w1,w2 | matvec A,x,y | sincos x,y,xy | matmul A,B,C */
static const kgo_test_info g_tests[KGO_NTESTS] = {
    { "laplacian",  3, 2, 2, 2 },
    { "wave13pt",   3, 3, 3, 3 },
    { "divergence", 3, 4, 3, 0 },
    { "gradient",   3, 4, 3, 0 },
    { "uxx1",       3, 6, 2, 2 },
    { "lapgsrb",    3, 2, 4, 2 },
    { "jacobi",     2, 2, 3, 2 },
    { "gaussblur",  2, 2, 6, 2 },
    { "gameoflife", 2, 2, 0, 2 },
    { "tricubic",   3, 5, 0, 2 },
    { "tricubic2",  3, 5, 0, 2 },
    { "vecadd",     3, 3, 0, 3 },
    { "matvec",     2, 3, 0, 0 },
    { "sincos",     3, 3, 0, 0 },
    { "matmul",     3, 3, 0, 0 },
};

const kgo_test_info* kgo_info(int test)
{
    if (test < 0 || test >= KGO_NTESTS) return NULL;
    return &g_tests[test];
}

void kgo_reseed(void) { srand(1); }

int kgo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

size_t kgo_array_len(int test, int slot, int nx, int ny, int ns)
{
    const kgo_test_info* ti = kgo_info(test);
    if (!ti || slot < 0 || slot >= ti->narrays) return 0;
    if (test == KGO_MATVEC)
        return slot == 0 ? (size_t)nx * ny : slot == 1 ? (size_t)nx : (size_t)ny;
    if (test == KGO_MATMUL)     /* A nx*ny, B ny*ns, C nx*ns  (matmul/main.c:80-87) */
        return slot == 0 ? (size_t)nx * ny : slot == 1 ? (size_t)ny * ns : (size_t)nx * ns;
    return ti->ndims == 3 ? (size_t)nx * ny * ns : (size_t)nx * ny;
}

#define REAL float
#define SFX _f
#include "kgo_impl.inc"
#undef REAL
#undef SFX

#define REAL double
#define SFX _d
#include "kgo_impl.inc"
#undef REAL
#undef SFX

double kgo_init(int test, int dtype, int nx, int ny, int ns,
                double* scalars, void* const* arrays)
{
    if (!kgo_info(test)) return NAN;
    return dtype == KGO_F32 ? kgo_init_f(test, nx, ny, ns, scalars, arrays)
                            : kgo_init_d(test, nx, ny, ns, scalars, arrays);
}

int kgo_sweep(int test, int dtype, int nx, int ny, int ns,
              const double* scalars, void* const* arrays)
{
    double sc[8] = { 0 };
    const kgo_test_info* ti = kgo_info(test);
    if (!ti) return -1;
    for (int q = 0; q < ti->nscalars; q++) sc[q] = scalars[q];
    return dtype == KGO_F32 ? kgo_sweep_f(test, nx, ny, ns, sc, arrays)
                            : kgo_sweep_d(test, nx, ny, ns, sc, arrays);
}

int kgo_run(int test, int dtype, int nx, int ny, int ns, int nt,
            const double* scalars, void* const* arrays)
{
    const kgo_test_info* ti = kgo_info(test);
    if (!ti) return -1;
    void* cur[8];
    int idxs[3] = { 0, 1, 2 };
    for (int q = 0; q < ti->narrays; q++) cur[q] = arrays[q];
    for (int it = 0; it < nt; it++)
    {
        if (kgo_sweep(test, dtype, nx, ny, ns, scalars, cur)) return -1;
        if (ti->rotation == 2)
        {   /* laplacian.c:299-300 */
            void* w = cur[0]; cur[0] = cur[1]; cur[1] = w;
            int t = idxs[0]; idxs[0] = idxs[1]; idxs[1] = t;
        }
        else if (ti->rotation == 3)
        {   /* wave13pt.c:919-920 */
            void* w = cur[0]; cur[0] = cur[1]; cur[1] = cur[2]; cur[2] = w;
            int t = idxs[0]; idxs[0] = idxs[1]; idxs[1] = idxs[2]; idxs[2] = t;
        }
    }
    if (ti->rotation) return idxs[1];             /* laplacian.c:307-309 */
    switch (test)
    {
    case KGO_DIVERGENCE: return 0;                /* u */
    case KGO_GRADIENT:   return 1;                /* ux (+uy+uz) */
    default:             return 2;                /* matvec y, sincos xy, matmul C */
    }
}

double kgo_final_mean(int test, int dtype, int nx, int ny, int ns,
                      void* const* arrays, int slot)
{
    if (!kgo_info(test)) return NAN;
    return dtype == KGO_F32 ? kgo_final_mean_f(test, nx, ny, ns, arrays, slot)
                            : kgo_final_mean_d(test, nx, ny, ns, arrays, slot);
}

int kgo_driver(int test, int dtype, int nx, int ny, int ns, int nt,
               double* scalars, double* i_mean, double* f_mean)
{
    const kgo_test_info* ti = kgo_info(test);
    if (!ti) return -1;
    size_t esz = dtype == KGO_F32 ? sizeof(float) : sizeof(double);
    void* arrays[8] = { 0 };
    double sc[8] = { 0 };
    int rc = 0;
    for (int q = 0; q < ti->narrays; q++)
    {
        arrays[q] = memalign(4096, kgo_array_len(test, q, nx, ny, ns) * esz);
        if (!arrays[q]) { rc = -2; goto done; }
    }
    *i_mean = kgo_init(test, dtype, nx, ny, ns, sc, arrays);
    if (scalars) memcpy(scalars, sc, sizeof(sc));
    int slot = kgo_run(test, dtype, nx, ny, ns, nt, sc, arrays);
    if (slot < 0) { rc = -3; goto done; }
    *f_mean = kgo_final_mean(test, dtype, nx, ny, ns, arrays, slot);
done:
    for (int q = 0; q < ti->narrays; q++) free(arrays[q]);
    return rc;
}
