/*
 * driver.c -- the per-test host driver of the `b200` target.
 *
 * One source, compiled once per test and precision:
 *     gcc -O3 -ffast-math -march=native -std=c99 -D_GNU_SOURCE \
 *         -DB200_TEST_ID=B200_LAPLACIAN -Dreal=double driver.c timing.c -lb200stencil
 * It replaces main() of the reference's <test>/<test>.c (e.g. laplacian/laplacian.c:114-385;
 * jacobi/main.c, sincos/main.c for the Fortran tests) for the new target and keeps that
 * driver's contract to the letter, because `benchmark` (benchmark:146-265), mktable and mkchart
 * parse its stdout:
 *   - arguments <nx> <ny> <ns> <nt> (3D) or <nx> <ny> <nt> (2D), same usage / "invalid" messages;
 *   - the same rand() draw order: coefficients first, then the arrays element-interleaved;
 *   - the same lines: coefficient header, "initial mean", "init time", "device buffer alloc time",
 *     "data load time", "compute time", "data save time", "device buffer free time",
 *     "final mean" (+ "<k> regcount" / "<k> kernel time" when PROFILING_FNAME is set, as
 *     <test>/cuda/cuda_profiling.cu:214-251 prints them);
 *   - the same buffer rotation per iteration and the same choice of the buffer the final mean
 *     is taken over (the reference's idxs[] remap, laplacian.c:307-313).
 * All GPU work goes through the C ABI of libb200stencil.so (include/b200_stencil.h); there
 * is no CPU fallback: any library error is printed like CUDA_SAFE_CALL does and exits with -1.
 *
 * Extra environment: B200_NGPUS=<1..8> cuts the grid into z-slabs over that many GPUs;
 * B200_INIT_THREADS=<N> fills the host arrays with N threads in the same rand() draw order (kg_rand.h);
 * B200_PINNED_HOST=0 uses memalign instead of page-locked host arrays; B200_VERIFY=1 runs the job twice and compares; B200_HBM_GBS=<GB/s> is the ceiling the
 * throughput line is compared with.  PROFILING_LINENO is accepted and unused, as in the
 * reference's cuda target (cuda_profiling.cu:21-26 stores it and never reads it).
 */
#include <malloc.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200_stencil.h"
#include "timing.h"

#ifndef real
#error "compile with -Dreal=float or -Dreal=double"
#endif
#ifndef B200_TEST_ID
#error "compile with -DB200_TEST_ID=B200_<TEST>"
#endif

#define MEMALIGN 4096
#define TEST B200_TEST_ID

/* the reference's input generator real_rand() (laplacian.c:112) and the array fill, serial or -- with
 * B200_INIT_THREADS=N -- parallel in the same draw order */
#include "kg_init.h"

#define parse_arg(name, arg) \
	int name = atoi(arg); \
	if (name < 0) \
	{ \
		printf("Value for " #name " is invalid: %d\n", name); \
		exit(1); \
	}

#define B200_SAFE_CALL(x) \
	do { int rc__ = (x); if (rc__ != B200_OK) { \
		fprintf(stderr, "Error \"%s\" at %s:%d\n", b200_last_error(), __FILE__, __LINE__); exit(-1); } \
	} while (0)

/* tests whose reference driver prints "initial mean" even under NO_TIMING
 * (wave13pt.c:756, gaussblur.c:163, gameoflife.c:152, jacobi/main.c:115, sincos/main.c:107) */
#define IMEAN_ALWAYS (TEST == B200_WAVE13PT || TEST == B200_GAUSSBLUR || TEST == B200_GAMEOFLIFE || \
                      TEST == B200_JACOBI || TEST == B200_SINCOS)

int main(int argc, char* argv[])
{
	const b200_test_info* ti = b200_get_test_info(TEST);
	const int ndims = ti->ndims;

	if (argc != ndims + 2)
	{
		if (ndims == 3) printf("Usage: %s <nx> <ny> <ns> <nt>\n", argv[0]);
		else printf("Usage: %s <nx> <ny> <nt>\n", argv[0]);
		exit(1);
	}

	const char* no_timing = getenv("NO_TIMING");
	const char* profiling_fname = getenv("PROFILING_FNAME");

	parse_arg(nx, argv[1]);
	parse_arg(ny, argv[2]);
	int ns = 1;
	if (ndims == 3)
	{
		ns = atoi(argv[3]);
		if (ns < 0)
		{
			printf("Value for ns is invalid: %d\n", ns);
			exit(1);
		}
	}
	parse_arg(nt, argv[ndims + 1]);

	/* ---- coefficients: drawn first, in the reference's order and with its divisors ---- */
	real sc[B200_MAX_SCALARS];
	memset(sc, 0, sizeof(sc));
	switch (TEST)
	{
	case B200_LAPLACIAN:                                  /* laplacian.c:141-144 */
		sc[0] = real_rand(); sc[1] = real_rand();
		printf("alpha = %f, beta = %f\n", sc[0], sc[1]);
		break;
	case B200_WAVE13PT:                                   /* wave13pt.c:729-733 */
		sc[0] = real_rand(); sc[1] = real_rand() / 6.; sc[2] = real_rand() / 6.;
		printf("m0 = %f, m1 = %f, m2 = %f\n", sc[0], sc[1], sc[2]);
		break;
	case B200_DIVERGENCE: case B200_GRADIENT:             /* divergence.c:140-144, gradient.c:141-145 */
		sc[0] = real_rand(); sc[1] = real_rand(); sc[2] = real_rand();
		printf("alpha = %f, beta = %f, gamma = %f\n", sc[0], sc[1], sc[2]);
		break;
	case B200_UXX1:                                       /* uxx1.c:149-152 */
		sc[0] = real_rand(); sc[1] = real_rand();
		printf("c1 = %f, c2 = %f\n", sc[0], sc[1]);
		break;
	case B200_LAPGSRB:                                    /* lapgsrb.c:161-166 */
		sc[0] = real_rand(); sc[1] = real_rand() / 6.; sc[2] = real_rand() / 12.; sc[3] = real_rand() / 6.;
		printf("c0 = %f, c1 = %f, c2 = %f, c3 = %f\n", sc[0], sc[1], sc[2], sc[3]);
		break;
	case B200_JACOBI:                                     /* jacobi/main.c:90-94 */
		sc[0] = real_rand(); sc[1] = real_rand() / 4.; sc[2] = real_rand() / 4.;
		printf("c0 = %f, c1 = %f, c2 = %f\n", sc[0], sc[1], sc[2]);
		break;
	case B200_GAUSSBLUR:                                  /* gaussblur.c:134-142 */
		for (int q = 0; q < 6; q++) sc[q] = real_rand();
		printf("s0 = %f, s1 = %f, s2 = %f\n", sc[0], sc[1], sc[2]);
		printf("s4 = %f, s5 = %f, s8 = %f\n", sc[3], sc[4], sc[5]);
		break;
	default:
		break;
	}
	double scd[B200_MAX_SCALARS];
	for (int q = 0; q < B200_MAX_SCALARS; q++) scd[q] = (double)sc[q];

	/* ---- host arrays ---- */
	const int na = ti->narrays;
	size_t len[B200_MAX_ARRAYS];
	size_t szarray = (ndims == 3) ? (size_t)nx * ny * ns : (size_t)nx * ny;
	for (int q = 0; q < na; q++) len[q] = szarray;
	if (TEST == B200_MATVEC) { len[0] = (size_t)nx * ny; len[1] = (size_t)nx; len[2] = (size_t)ny; }
	if (TEST == B200_MATMUL) { len[0] = (size_t)nx * ny; len[1] = (size_t)ny * ns; len[2] = (size_t)nx * ns; }  /* matmul/main.c:80-85 */
	size_t szarrayb = szarray * sizeof(real);

	/* Host arrays are page-locked (b200_host_alloc) instead of the reference's memalign (laplacian.c:149-150), so that
	 * "data load time" / "data save time" run at PCIe rate (34-39 / 39-42 GB/s) instead of through the CUDA driver's
	 * staging of pageable memory (9 / 16 GB/s, profiles/r1x_driver_pinned_host.txt).  Page-locking needs the CUDA
	 * context, so the device initialisation -- what the reference's "init time" reports (laplacian.c:192-199) -- is done
	 * and TIMED here, before the host allocations; its line is printed at the reference's place in the output.
	 * B200_PINNED_HOST=0: memalign, as in the reference. */
	const char* pinned_env = getenv("B200_PINNED_HOST");
	const int pinned = !(pinned_env && atoi(pinned_env) == 0);
	volatile struct timespec t0, t1;
	b200_ctx* ctx = NULL;
	/* B200_VERIFY=1: self-check of the GPU path -- the whole job (load, nt sweeps, save) is run a second time on an
	 * independent context from a copy of the inputs and must reproduce the result bit for bit: nothing in the kernels
	 * (mbarrier rings, halo pushes and flags between GPUs, item scheduling) may depend on timing. */
	const char* verify_env = getenv("B200_VERIFY");
	const int verify = verify_env && atoi(verify_env) != 0;
	real* a_copy[B200_MAX_ARRAYS] = { 0 };
	double init_t = 0.0;
	if (pinned)
	{
		get_time(&t0);
		B200_SAFE_CALL(b200_init(&ctx, 0));
		get_time(&t1);
		init_t = get_time_diff((struct timespec*)&t0, (struct timespec*)&t1);
	}
	real* a[B200_MAX_ARRAYS] = { 0 };
	int a_pinned[B200_MAX_ARRAYS] = { 0 };
	int ok = 1;
	for (int q = 0; q < na; q++)
	{
		if (pinned)
		{
			/* page-locking tens of GB can be refused (locked-memory limits): such an array falls back to memalign */
			void* ptr = NULL;
			if (b200_host_alloc(&ptr, len[q] * sizeof(real) + 16) == B200_OK && ptr) { a[q] = (real*)ptr; a_pinned[q] = 1; }
		}
		if (!a[q]) a[q] = (real*)memalign(MEMALIGN, len[q] * sizeof(real) + 16);
		if (!a[q]) ok = 0;
	}
	if (!ok)
	{
		printf("Error allocating memory for arrays:");
		for (int q = 0; q < na; q++) printf(" %p%s", (void*)a[q], q + 1 < na ? "," : "\n");
		exit(1);
	}

	/* ---- init, element-interleaved across arrays (laplacian.c:158-165); matvec fills A, x, y
	 * in three loops (matvec.c:119-134) ---- */
	const int init_threads = kg_init_threads();
	real mean = 0.0f;
	if (TEST == B200_MATVEC)
	{
		const real amean = kg_fill(&a[0], 1, len[0], init_threads);
		const real xmean = kg_fill(&a[1], 1, (size_t)nx, init_threads);
		const real ymean = kg_fill(&a[2], 1, (size_t)ny, init_threads);
		if (!no_timing) printf("initial mean = %f\n", amean / (nx * ny) + xmean / nx + ymean / ny);
	}
	else if (TEST == B200_MATMUL)
	{
		/* matmul/main.c:98-110: A, then B; printed unconditionally.  The reference never
		 * initialises C (fresh memalign pages read as zero); it is zeroed explicitly here. */
		const real meanA = kg_fill(&a[0], 1, len[0], init_threads);
		const real meanB = kg_fill(&a[1], 1, len[1], init_threads);
		memset(a[2], 0, len[2] * sizeof(real));
		printf("initial mean = %f\n", (meanA / len[0] + meanB / len[1]));
	}
	else
	{
		mean = kg_fill(a, na, szarray, init_threads);
		if (IMEAN_ALWAYS || !no_timing) printf("initial mean = %f\n", mean / szarray / na);
	}

	if (verify)
		for (int q = 0; q < na; q++)
		{
			a_copy[q] = (real*)memalign(MEMALIGN, len[q] * sizeof(real) + 16);
			if (!a_copy[q]) { printf("Error allocating memory for the B200_VERIFY copies\n"); exit(1); }
			memcpy(a_copy[q], a[q], len[q] * sizeof(real));
		}

	/* 1) device / context initialisation  (reference: cudaGetDeviceCount probe, laplacian.c:192-199): done and timed
	 *    above when the host arrays are page-locked (they need the context), here otherwise; reported here, where the
	 *    reference reports it */
	if (!pinned)
	{
		get_time(&t0);
		B200_SAFE_CALL(b200_init(&ctx, 0));
		get_time(&t1);
		init_t = get_time_diff((struct timespec*)&t0, (struct timespec*)&t1);
	}
	if (!no_timing) printf("init time = %f sec\n", init_t);

	/* 2) device buffers  (laplacian.c:223-231) */
	get_time(&t0);
	B200_SAFE_CALL(b200_plan(ctx, TEST, sizeof(real) == 4 ? B200_F32 : B200_F64, nx, ny, ns, scd, ti->nscalars));
	B200_SAFE_CALL(b200_alloc(ctx));
	get_time(&t1);
	if (!no_timing) printf("device buffer alloc time = %f sec\n", get_time_diff((struct timespec*)&t0, (struct timespec*)&t1));

	/* 3) host -> device  (laplacian.c:255-262) */
	get_time(&t0);
	size_t loaded = 0;
	for (int q = 0; q < na; q++)
	{
		/* output buffers: the first sweep overwrites their interior before anything reads it, so
		 * only the boundary shell has to travel (the reference's cuda target copies them whole,
		 * laplacian.c:257-258); with nt = 0 the untouched array is the result, so it goes whole */
		if (nt >= 1 && b200_slot_interior_dead(TEST, q))
			B200_SAFE_CALL(b200_load_shell(ctx, q, a[q]));
		else
		{
			B200_SAFE_CALL(b200_load(ctx, q, a[q]));
			loaded += len[q] * sizeof(real);
		}
	}
	get_time(&t1);
	double load_t = get_time_diff((struct timespec*)&t0, (struct timespec*)&t1);
	if (!no_timing) printf("data load time = %f sec (%f GB/sec)\n", load_t, loaded / (load_t * 1024 * 1024 * 1024));

	/* 4) the nt sweeps, data resident on the device, rotation inside the library
	 *    (laplacian.c:287-301; no per-iteration synchronisation) */
	b200_stats st;
	get_time(&t0);
	B200_SAFE_CALL(b200_run(ctx, nt, &st));
	get_time(&t1);
	double compute_t = get_time_diff((struct timespec*)&t0, (struct timespec*)&t1);
	if (!no_timing) printf("compute time = %f sec\n", compute_t);
	if (profiling_fname)
	{
		/* what __wrap_cudaLaunchKernel prints per launch (cuda_profiling.cu:229-249); the time
		 * is the CUDA-event time per sweep */
		printf("%s regcount = %d\n", st.kernel_name, st.regs_per_thread);
		printf("%s kernel time = %f\n", st.kernel_name, st.kernel_ms_per_sweep * 1e-3);
	}

	/* the reference's idxs[] remap: which buffer the final mean is over (laplacian.c:307-313) */
	const int slot = b200_result_slot(ctx);

	/* 5) device -> host  (laplacian.c:334-340; gradient copies ux, uy, uz back) */
	get_time(&t0);
	size_t saved = 0;
	if (TEST == B200_GRADIENT)
	{
		for (int q = 1; q <= 3; q++) { B200_SAFE_CALL(b200_save(ctx, q, a[q])); saved += len[q] * sizeof(real); }
	}
	else
	{
		B200_SAFE_CALL(b200_save(ctx, slot, a[slot]));
		saved += len[slot] * sizeof(real);
	}
	get_time(&t1);
	double save_t = get_time_diff((struct timespec*)&t0, (struct timespec*)&t1);
	if (!no_timing) printf("data save time = %f sec (%f GB/sec)\n", save_t, saved / (save_t * 1024 * 1024 * 1024));

	/* 6) release  (laplacian.c:362-369) */
	get_time(&t0);
	B200_SAFE_CALL(b200_free(ctx));
	get_time(&t1);
	if (!no_timing) printf("device buffer free time = %f sec\n", get_time_diff((struct timespec*)&t0, (struct timespec*)&t1));

	if (verify)
	{
		b200_ctx* c2 = NULL;
		b200_stats st2;
		B200_SAFE_CALL(b200_init(&c2, 0));
		B200_SAFE_CALL(b200_plan(c2, TEST, sizeof(real) == 4 ? B200_F32 : B200_F64, nx, ny, ns, scd, ti->nscalars));
		B200_SAFE_CALL(b200_alloc(c2));
		for (int q = 0; q < na; q++) B200_SAFE_CALL(b200_load(c2, q, a_copy[q]));      /* whole arrays: also checks the shell-only loads */
		B200_SAFE_CALL(b200_run(c2, nt, &st2));
		size_t checked = 0;
		int bad = b200_result_slot(c2) != slot;
		for (int q = 0; q < na && !bad; q++)
		{
			const int is_out = (TEST == B200_GRADIENT) ? (q >= 1 && q <= 3) : (q == slot);
			if (!is_out) continue;
			B200_SAFE_CALL(b200_save(c2, q, a_copy[q]));
			if (memcmp(a_copy[q], a[q], len[q] * sizeof(real)) != 0) bad = 1;
			checked += len[q] * sizeof(real);
		}
		B200_SAFE_CALL(b200_free(c2));
		b200_destroy(c2);
		for (int q = 0; q < na; q++) free(a_copy[q]);
		if (bad) { fprintf(stderr, "b200 verify: FAILED -- the second run differs from the first\n"); exit(-1); }
		printf("b200 verify: second independent run bit-identical (%zu bytes compared)\n", checked);
	}

	/* extra, unparsed by benchmark: throughput against the algorithmic byte count */
	if (!no_timing && nt > 0 && st.kernel_ms_per_sweep > 0)
	{
		double lups = (double)b200_interior_points(TEST, nx, ny, ns);
		double sec = st.kernel_ms_per_sweep * 1e-3;
		double bytes = lups * (ti->nread + ti->nwritten) * sizeof(real);
		if (TEST == B200_MATMUL)
			printf("b200: %d GPU(s), %.3f TFLOP/s (2*nx*ny*ns flops per sweep)\n", st.ngpus, 2 * lups / sec * 1e-12);
		else
		{
			/* % of the HBM roofline: against B200_HBM_GBS (the measured device-copy ceiling, exported by the b200
			 * makefile from MEASURED_PEAKS.json) x the number of GPUs; default = the profiling guide's fallback figure */
			const char* pk = getenv("B200_HBM_GBS");
			double peak = pk && atof(pk) > 0 ? atof(pk) : 6650.0;
			printf("b200: %d GPU(s), %.3f GLUP/s, %.1f GB/s algorithmic (%d+%d arrays x %d bytes per LUP), %.1f %% of the HBM roofline (%.1f GB/s per GPU%s)\n",
				st.ngpus, lups / sec * 1e-9, bytes / sec * 1e-9, ti->nread, ti->nwritten, (int)sizeof(real),
				100.0 * bytes / sec * 1e-9 / (peak * st.ngpus), peak, pk ? "" : ", fallback figure");
		}
	}

	/* final mean, over the buffer the reference would report (laplacian.c:374-377) */
	if (TEST == B200_MATVEC)
	{
		real ymean = 0.0f;
		for (int i = 0; i < ny; i++) ymean += a[2][i];
		printf("final mean = %f\n", ymean / ny);
	}
	else if (TEST == B200_MATMUL)
	{
		real meanC = 0.0f;                                 /* matmul/main.c:311-314 */
		for (size_t i = 0; i < len[2]; i++) meanC += a[2][i];
		printf("final mean = %f\n", meanC / len[2]);
	}
	else if (TEST == B200_GRADIENT)
	{
		mean = 0.0f;
		for (size_t i = 0; i < szarray; i++) mean += a[1][i] + a[2][i] + a[3][i];
		printf("final mean = %f\n", mean / szarray / 3);
	}
	else
	{
		mean = 0.0f;
		for (size_t i = 0; i < szarray; i++) mean += a[slot][i];
		printf("final mean = %f\n", mean / szarray);
	}
	(void)szarrayb;

	b200_destroy(ctx);
	for (int q = 0; q < na; q++)
	{
		if (a_pinned[q]) b200_host_free(a[q]);
		else free(a[q]);
	}
	fflush(stdout);
	return 0;
}
