/* kg_rand.h -- glibc's default rand() stream with random access, for order-preserving PARALLEL input
 * initialisation in the b200 drivers (B200_INIT_THREADS > 1).
 *
 * The reference fills its inputs with unseeded rand() calls, serially, 2-6 draws per grid point
 * (e.g. laplacian/laplacian.c:112,158-165): at 1024x1024x512 that is minutes of one host core, far more
 * than the sweeps take on a B200.  The inputs are DEFINED by the draw order, so a faster
 * generator must reproduce the very same stream.  glibc's rand() (TYPE_3 additive feedback, random_r.c) is
 * the linear recurrence  z[n] = z[n-31] + z[n-3]  (mod 2^32), output z[n] >> 1, started from an LCG-filled
 * table with 310 values discarded.  Being linear, it can be advanced by k steps with x^k mod (x^31 - x^28 - 1)
 * over Z/2^32 -- O(31^2 log k) -- so every thread seeks to the first draw of its block of grid points and
 * generates from there: bit-identical arrays, any number of threads.
 *
 * Restated from the published algorithm (glibc stdlib/random_r.c); checked against rand() itself by
 * tests/test_driver_init.py. */
#ifndef KG_RAND_H
#define KG_RAND_H

#include <stddef.h>
#include <stdint.h>

typedef struct {
	uint32_t z[31];   /* the last 31 values of the recurrence, circular */
	int pos;          /* slot of z[n-31], the one the next value overwrites */
} kg_rand_t;

/* Position the generator so that the next kg_rand_next() returns what the (ndraws + 1)-th call of rand()
 * returns in a process that never called srand() (seed 1). */
void kg_rand_seek(kg_rand_t* g, uint64_t ndraws);

static inline int kg_rand_next(kg_rand_t* g)
{
	int p = g->pos, q = p + 28;
	if (q >= 31) q -= 31;
	const uint32_t v = g->z[p] + g->z[q];
	g->z[p] = v;
	g->pos = (p + 1 == 31) ? 0 : p + 1;
	return (int)(v >> 1);
}

#endif
