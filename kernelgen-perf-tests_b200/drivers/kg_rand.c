/* kg_rand.c -- see kg_rand.h */
#include "kg_rand.h"

#include <string.h>

#define DEG 31
#define SEP 3
#define DISCARD 310       /* glibc: 10 * DEG values are thrown away after seeding */

/* a(x) * b(x) mod (x^31 - x^28 - 1), coefficients mod 2^32 */
static void polymul(const uint32_t* a, const uint32_t* b, uint32_t* out)
{
	uint32_t t[2 * DEG - 1];
	memset(t, 0, sizeof(t));
	for (int i = 0; i < DEG; i++)
		for (int j = 0; j < DEG; j++) t[i + j] += a[i] * b[j];
	/* x^(31+e) = x^(28+e) + x^e, highest power first */
	for (int e = 2 * DEG - 2; e >= DEG; e--)
	{
		t[e - 3] += t[e];
		t[e - DEG] += t[e];
	}
	memcpy(out, t, DEG * sizeof(uint32_t));
}

void kg_rand_seek(kg_rand_t* g, uint64_t ndraws)
{
	/* the table srand(1) builds: r[0] = 1, r[i] = 16807 r[i-1] mod (2^31 - 1)  (Schrage's form) */
	int32_t r[DEG];
	r[0] = 1;
	for (int i = 1; i < DEG; i++)
	{
		const int32_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
		int32_t w = 16807 * lo - 2836 * hi;
		if (w < 0) w += 2147483647;
		r[i] = w;
	}
	/* As a sequence: step n computes y[n] = (old content of slot (n+3) mod 31) + y[n-3].  With
	 * z[j] = r[(j + 34) mod 31] for j = -31..-1 this is z[n] = z[n-31] + z[n-3] for every n >= 0. */
	uint32_t w0[2 * DEG - 1];                      /* z[-31] .. z[29] */
	for (int j = 0; j < DEG; j++) w0[j] = (uint32_t)r[(j + 3) % DEG];
	for (int j = DEG; j < 2 * DEG - 1; j++) w0[j] = w0[j - DEG] + w0[j - SEP];

	/* c(x) = x^k mod P, k = DISCARD + ndraws: z[m + k] = sum_j c[j] z[m + j] */
	uint64_t k = (uint64_t)DISCARD + ndraws;
	uint32_t c[DEG], b[DEG];
	memset(c, 0, sizeof(c));
	memset(b, 0, sizeof(b));
	c[0] = 1;
	b[1] = 1;
	while (k)
	{
		if (k & 1) polymul(c, b, c);
		polymul(b, b, b);
		k >>= 1;
	}
	/* window z[k-31] .. z[k-1] */
	for (int m = 0; m < DEG; m++)
	{
		uint32_t s = 0;
		for (int j = 0; j < DEG; j++) s += c[j] * w0[m + j];
		g->z[m] = s;
	}
	g->pos = 0;
}
