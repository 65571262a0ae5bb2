/* kg_init.h -- input initialisation of the b200 drivers: the reference's rand() fill, serial (default,
 * literally the reference's loops, e.g. laplacian/laplacian.c:158-165) or in parallel with the same draw
 * order (B200_INIT_THREADS = N > 1, kg_rand.h).  Include after `real` is defined.
 *
 * The parallel fill produces bit-identical arrays.  The "initial mean" is a sum in `real`: the parallel
 * path adds per-thread partial sums in thread order, which can change its last printed digits for float
 * (the reference's own vectorised -ffast-math build re-associates that sum as well). */
#ifndef KG_INIT_H
#define KG_INIT_H

#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

#include "kg_rand.h"

/* the reference's input generator, a double expression (laplacian.c:112) */
#define kg_real_from(r) (((real)((r) / (double)RAND_MAX) - 0.5) * 2)

static uint64_t kg_ndraws = 0;                 /* rand() calls made so far through kg_counted_rand() */
static inline int kg_counted_rand(void) { kg_ndraws++; return rand(); }
#define real_rand() kg_real_from(kg_counted_rand())

typedef struct {
	real** a;          /* na arrays filled element-interleaved: a[0][i], a[1][i], ... then i + 1 */
	int na;
	size_t lo, hi;     /* this worker's elements */
	uint64_t draw0;    /* draws made before element 0 of the group */
	real sum;          /* sum over the worker's elements of (a[0][i] + a[1][i] + ...) */
} kg_job;

static void* kg_worker(void* p)
{
	kg_job* j = (kg_job*)p;
	kg_rand_t g;
	kg_rand_seek(&g, j->draw0 + (uint64_t)j->na * j->lo);
	real mean = 0.0f;
	for (size_t i = j->lo; i < j->hi; i++)
	{
		j->a[0][i] = kg_real_from(kg_rand_next(&g));
		real s = j->a[0][i];
		for (int q = 1; q < j->na; q++) { j->a[q][i] = kg_real_from(kg_rand_next(&g)); s = s + j->a[q][i]; }
		mean += s;
	}
	j->sum = mean;
	return NULL;
}

/* Fill a group of na interleaved arrays of n elements each; returns the sum of all values drawn.
 * nthreads <= 1: the reference's serial loop on rand() itself. */
static real kg_fill(real** a, int na, size_t n, int nthreads)
{
	real mean = 0.0f;
	if (nthreads <= 1)
	{
		for (size_t i = 0; i < n; i++)
		{
			a[0][i] = real_rand();
			real s = a[0][i];
			for (int q = 1; q < na; q++) { a[q][i] = real_rand(); s = s + a[q][i]; }
			mean += s;
		}
		return mean;
	}
	if ((size_t)nthreads > n / 4096 + 1) nthreads = (int)(n / 4096 + 1);
	if (nthreads > 256) nthreads = 256;
	kg_job job[256];
	pthread_t th[256];
	int started[256];
	for (int t = 0; t < nthreads; t++)
	{
		job[t].a = a; job[t].na = na; job[t].draw0 = kg_ndraws; job[t].sum = 0.0f;
		job[t].lo = n / nthreads * t + (n % nthreads < (size_t)t ? n % nthreads : (size_t)t);
		job[t].hi = n / nthreads * (t + 1) + (n % nthreads < (size_t)(t + 1) ? n % nthreads : (size_t)(t + 1));
		started[t] = pthread_create(&th[t], NULL, kg_worker, &job[t]) == 0;
		if (!started[t]) kg_worker(&job[t]);       /* no thread to be had: do the block here */
	}
	for (int t = 0; t < nthreads; t++)
	{
		if (started[t]) pthread_join(th[t], NULL);
		mean += job[t].sum;
	}
	kg_ndraws += (uint64_t)na * n;                 /* the stream position moves on, as rand() would have */
	return mean;
}

static inline int kg_init_threads(void)
{
	const char* e = getenv("B200_INIT_THREADS");
	const int n = e ? atoi(e) : 1;
	return n < 1 ? 1 : n;
}

#endif
