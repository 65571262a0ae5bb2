/* kg_rand_check.c -- CPU-only check of kg_rand against glibc's rand() itself (tests/test_driver_init.py):
 * the first 100000 draws, seeks to scattered offsets up to 1e8 (reached by really calling rand()), and
 * seek(a + b) == seek(a) advanced b times at an offset beyond 2^35 (larger than any grid of the suite). */
#include <stdio.h>
#include <stdlib.h>

#include "kg_rand.h"

int main(void)
{
	kg_rand_t g, h;
	kg_rand_seek(&g, 0);
	for (int i = 0; i < 100000; i++)
	{
		const int a = rand(), b = kg_rand_next(&g);
		if (a != b) { printf("mismatch at draw %d: rand %d, kg_rand %d\n", i, a, b); return 1; }
	}
	const unsigned long long offs[] = { 100001ull, 12345678ull, 99999999ull };
	unsigned long long cur = 100000;
	for (int t = 0; t < 3; t++)
	{
		while (cur < offs[t]) { rand(); cur++; }
		kg_rand_seek(&g, offs[t]);
		for (int i = 0; i < 1000; i++, cur++)
			if (rand() != kg_rand_next(&g)) { printf("mismatch after seek to %llu (+%d)\n", offs[t], i); return 1; }
	}
	kg_rand_seek(&g, (1ull << 35) + 12345);
	kg_rand_seek(&h, (1ull << 35) + 12345 - 777);
	for (int i = 0; i < 777; i++) kg_rand_next(&h);
	for (int i = 0; i < 100; i++)
		if (kg_rand_next(&g) != kg_rand_next(&h)) { printf("seek composition mismatch\n"); return 1; }
	printf("kg_rand OK\n");
	return 0;
}
