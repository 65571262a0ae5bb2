/* init_check.c -- CPU-only check of the drivers' input initialisation (tests/test_driver_init.py):
 *     init_check <na> <n> <ndraws_before> <threads>
 * draws <ndraws_before> values (the coefficients), fills na interleaved arrays of n elements serially with
 * rand() -- the reference's loop -- and again with <threads> threads through kg_rand, and compares the arrays
 * byte for byte.  Prints "identical <serial sum> <parallel sum>" or "DIFFERENT ..." (exit 1). */
#include <stdio.h>
#include <string.h>

#ifndef real
#define real double
#endif
#include "kg_init.h"

int main(int argc, char** argv)
{
	if (argc != 5) { printf("Usage: %s <na> <n> <ndraws_before> <threads>\n", argv[0]); return 2; }
	const int na = atoi(argv[1]);
	const size_t n = (size_t)atoll(argv[2]);
	const int before = atoi(argv[3]), threads = atoi(argv[4]);
	if (na < 1 || na > 8 || threads < 2) return 2;      /* threads = 1 is the serial rand() path itself */
	real *a[8], *b[8];
	for (int q = 0; q < na; q++)
	{
		a[q] = (real*)malloc(n * sizeof(real) + 16);
		b[q] = (real*)malloc(n * sizeof(real) + 16);
		if (!a[q] || !b[q]) return 2;
	}
	volatile real sink = 0;
	for (int i = 0; i < before; i++) sink += real_rand();
	const uint64_t mark = kg_ndraws;
	const real s1 = kg_fill(a, na, n, 1);
	const int next_serial = rand();                 /* the draw after the fill */
	kg_ndraws = mark;
	const real s2 = kg_fill(b, na, n, threads);
	kg_rand_t g;
	kg_rand_seek(&g, kg_ndraws);
	const int next_parallel = kg_rand_next(&g);     /* stream position after the parallel fill */
	int same = next_serial == next_parallel;
	for (int q = 0; q < na; q++) same = same && memcmp(a[q], b[q], n * sizeof(real)) == 0;
	printf("%s %.17g %.17g\n", same ? "identical" : "DIFFERENT", (double)s1, (double)s2);
	return same ? 0 : 1;
}
