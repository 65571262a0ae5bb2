/*
 * b200_stencil.h -- C ABI of libb200stencil.so: the B200 (sm_100a) target for the
 * kernelgen-perf-tests stencil hot path.
 *
 * This is the drop-in boundary.  Everything a `b200` per-test driver (the
 * replacement for main() in <test>/<test>.c of the reference) needs is here:
 * plain C, plain pointers and sizes, int return codes (0 = ok; the message is
 * in b200_last_error()).  No CUDA or torch types appear in any signature;
 * streams are passed as void* (a cudaStream_t, NULL = default stream).
 *
 * Two layers:
 *
 *  (1) Sweep launch API -- stateless: one sweep of one test over caller-owned
 *      DEVICE buffers.  Replaces the kernel call site of the reference:
 *        CPU   laplacian(nx, ny, ns, alpha, beta, w0p, w1p)         laplacian/laplacian.c:290
 *        CUDA  laplacian<<<grid,block,shmem>>>(nx,ny,ns,config,...)  laplacian/laplacian.c:292-295
 *      plus kernelgen_cuda_configure_gird()                           <test>/cuda/cuda_profiling.cu:36-111
 *      (launch geometry is planned inside the library).
 *
 *  (2) Context API -- what the reference's cuda-target driver does around the
 *      kernel, phase by phase, so the driver can print the same timing lines:
 *        b200_init   <-> cudaGetDeviceCount probe      laplacian/laplacian.c:192-199  ("init time")
 *        b200_plan + b200_alloc <-> cudaMalloc xN      laplacian/laplacian.c:223-231  ("device buffer alloc time")
 *        b200_load   <-> cudaMemcpy H2D                laplacian/laplacian.c:255-262  ("data load time")
 *        b200_run    <-> the nt-loop with swap         laplacian/laplacian.c:287-301  ("compute time")
 *        b200_result_slot <-> idxs remap               laplacian/laplacian.c:307-313
 *        b200_save   <-> cudaMemcpy D2H                laplacian/laplacian.c:334-340  ("data save time")
 *        b200_free   <-> cudaFree xN                   laplacian/laplacian.c:362-369  ("device buffer free time")
 *      and the per-launch profiler lines ("<k> regcount = N", "<k> kernel time = T")
 *      of __wrap_cudaLaunchKernel, <test>/cuda/cuda_profiling.cu:214-251, are
 *      served by b200_stats.
 *
 * Array order ("slots") for every test is the reference driver's init order:
 *   laplacian w0,w1 | wave13pt w0,w1,w2 | divergence u,ux,uy,uz | gradient u,ux,uy,uz |
 *   uxx1 u0,u1,d1,xx,xy,xz | lapgsrb w0,w1 | jacobi w0,w1 | gaussblur w0,w1 |
 *   gameoflife u0,u1 | tricubic,tricubic2 u0,u1,a,b,c | vecadd w0,w1,w2 |
 *   matvec A,x,y | sincos x,y,xy | matmul A,B,C
 * Layout: x fastest, index = i + nx*(j + ny*k).  2D tests use (nx, ny), ns = 1.
 * matmul (matmul/matmul.F90:56-68, matmul/main.c:80-87): column-major A nx*ny, B ny*ns, C nx*ns;
 * every sweep ACCUMULATES C += A*B (matmul/main.c:232-244).  Multi-GPU: columns of B and C are
 * split, A is replicated; out_begin/out_end select a column range.
 *
 * There is NO CPU fallback: every entry point that computes fails with
 * B200_ERR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef B200_STENCIL_H
#define B200_STENCIL_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_API_VERSION 1

typedef enum {
    B200_LAPLACIAN = 0, B200_WAVE13PT, B200_DIVERGENCE, B200_GRADIENT, B200_UXX1,
    B200_LAPGSRB, B200_JACOBI, B200_GAUSSBLUR, B200_GAMEOFLIFE, B200_TRICUBIC,
    B200_TRICUBIC2, B200_VECADD, B200_MATVEC, B200_SINCOS, B200_MATMUL, B200_NTESTS
} b200_test_t;

typedef enum { B200_F32 = 0, B200_F64 = 1 } b200_dtype_t;

enum {
    B200_OK = 0,
    B200_ERR_ARG = 1,        /* bad argument (sizes, ids, NULL) */
    B200_ERR_CUDA = 2,       /* a CUDA runtime/driver call failed */
    B200_ERR_NO_DEVICE = 3,  /* no usable sm_100 device: there is no CPU fallback */
    B200_ERR_STATE = 4,      /* context call out of order */
    B200_ERR_NOMEM = 5
};

#define B200_MAX_ARRAYS 8
#define B200_MAX_SCALARS 8

/* Static description of a test (mirrors what each reference driver hard-codes). */
typedef struct {
    const char* name;      /* "laplacian", ... == directory / kernel name in the suite */
    int ndims;             /* 3: <nx> <ny> <ns> <nt>;  2: <nx> <ny> <nt> */
    int narrays;           /* arrays, in driver init order */
    int nscalars;          /* rand()-drawn coefficients */
    int rotation;          /* 0 none; 2 swap slots 0,1 per sweep; 3 rotate slots 0,1,2 */
    int lo[3], hi[3];      /* interior: lo[d] <= idx < n[d]-hi[d]  (x,y,z) */
    int nread, nwritten;   /* arrays touched per sweep -> algorithmic bytes/LUP = (nread+nwritten)*sizeof(real) */
    int zghost_lo, zghost_hi; /* ghost depth needed below/above in the slab-split dimension (z; y for 2D) */
    int exchange_slot;     /* slot whose ghost planes must be refreshed after each sweep (-1: none) */
} b200_test_info;

const b200_test_info* b200_get_test_info(int test);
int b200_test_by_name(const char* name);            /* -1 if unknown */
const char* b200_last_error(void);
int b200_api_version(void);

/* Interior lattice-point updates per sweep for the given extents (0 if degenerate).
 * matvec: matrix elements; matmul: multiply-adds (nx*ny*ns, i.e. flops / 2). */
unsigned long long b200_interior_points(int test, int nx, int ny, int ns);

/* ---- device / environment ------------------------------------------------ */
int b200_device_count(int* count);                  /* usable sm_100 devices */

/* ---- (1) sweep launch API ------------------------------------------------ */
typedef struct {
    int test;                       /* b200_test_t */
    int dtype;                      /* b200_dtype_t */
    int nx, ny, ns;                 /* extents of the arrays as laid out in memory (a z-slab incl. ghosts) */
    double scalars[B200_MAX_SCALARS];
    /* Output range in the slab-split dimension (z for 3D tests, y for 2D tests),
     * in local array coordinates, half-open.  Must lie inside the interior.
     * out_begin == out_end == 0 selects the whole interior. */
    int out_begin, out_end;
    /* Fused halo push (multi-GPU): optional peer-mapped pointers to the lower /
     * upper neighbour's copy of the OUTPUT array (same layout, the neighbour's
     * local coordinates).  The kernel copies the planes the neighbour needs as
     * ghosts directly into peer memory over NVLink while the rest of the grid is still being swept.
     * push_lo_dst_plane: first ghost plane index in the lower neighbour that
     * receives our planes [push_lo_src_plane, +zghost_hi); likewise for hi. */
    void* push_lo; int push_lo_src_plane; int push_lo_dst_plane; int push_lo_count;
    void* push_hi; int push_hi_src_plane; int push_hi_dst_plane; int push_hi_count;
    /* 1: walk the work items back to front.  Alternating 0/1 between consecutive sweeps makes a
     * sweep start on the data its predecessor touched last, which is still in L2. */
    int reverse_order;
    /* Ordering between neighbour ranks, done INSIDE the sweep kernel (only with push_lo/push_hi).  A pushing launch
     * walks the END units of the split dimension first (the z-chunks / tile rows that read a ghost plane or produce a
     * plane a neighbour needs); before a CTA loads its first such item it spins until *wait_flag[i] >= wait_value
     * (acquire, system scope; NULL = no wait), every finished end item copies its share of the ghost planes into the
     * neighbours' arrays, and when the last end item of the grid has done so signal_value is stored to signal_flag[i]
     * (release, system scope; normally a flag in the neighbour's memory) -- early in the sweep; interior items take no
     * part in the ordering. */
    const void* wait_flag[2];
    unsigned long long wait_value;
    void* signal_flag[2];
    unsigned long long signal_value;
} b200_sweep_desc;

/* One sweep: arrays[] are DEVICE pointers in slot order (current rotation
 * already applied by the caller).  Asynchronous on `stream`. */
int b200_sweep(const b200_sweep_desc* desc, void* const* arrays, void* stream);

/* `niters` sweeps enqueued back to back with the reference driver's buffer rotation applied
 * between them (the nt-loop of laplacian/laplacian.c:287-301 without its per-iteration
 * cudaDeviceSynchronize).  arrays[] is rotated IN PLACE: on return it holds the pointers in
 * the order the next sweep would see them.  Not for sweeps that push halos (those need the
 * per-sweep b200_wait / b200_signal ordering). */
int b200_sweep_loop(const b200_sweep_desc* desc, void** arrays, int niters, void* stream);

/* Temporal blocking: TWO sweeps in one pass over memory, for the tests where it is implemented
 * (b200_sweep2_supported: jacobi, gaussblur, gameoflife).  arrays = { w0 (state t), w1 (only its
 * boundary shell is read: the values state t+1 has on the global boundary), out (receives state
 * t+2 in the interior; must already hold w0's boundary shell) }.  Per point the arithmetic is
 * that of two b200_sweep calls, so the result is bit-identical; state t+1 is never written.
 * Replaces two iterations of the nt-loop (e.g. gameoflife/gameoflife.c:276-289).  Single GPU. */
int b200_sweep2_supported(int test);
/* 1 when b200_run / b200_sweep_loop2 would use the fused kernel for this test and row length (a measured
 * policy: today jacobi with nx >= 1024; B200_FUSE=1 / 0 forces it on / off, as does B200_TBLOCK=2 / 1,
 * the number of sweeps per pass). */
int b200_sweep2_profitable(int test, int nx);
int b200_sweep2(const b200_sweep_desc* desc, void* const* arrays, void* stream);

/* b200_sweep_loop with a scratch buffer (same size as the arrays, holding arrays[0]'s boundary
 * shell): runs (niters-2)/2 fused pairs followed by the remaining single sweeps, so that the last
 * two states (the ones the reference driver reports, laplacian/laplacian.c:307-313) are both in
 * memory.  After a pair the buffer that held state t becomes the scratch: arrays[] AND *scratch
 * are updated in place.  Falls back to b200_sweep_loop when b200_sweep2_profitable() says no,
 * scratch is NULL or niters < 4. */
int b200_sweep_loop2(const b200_sweep_desc* desc, void** arrays, void** scratch, int niters, void* stream);

/* The nt-loop of one z-slab whose neighbours are other processes / GPUs: `niters` sweeps back to
 * back, each pushing its boundary planes into the neighbours' copy of the array it writes
 * (peer_lo[] / peer_hi[]: the neighbours' buffers in the SAME rotation order as arrays[], peer
 * mapped; NULL entries = no neighbour on that side) and ordered against the neighbours purely on
 * the device: sweep number n (counted from first_sweep) waits for flag value n on
 * desc->wait_flag[] and publishes n+1 on desc->signal_flag[].  desc carries the plane ranges
 * (push_*_src_plane / dst_plane / count).  arrays[], peer_lo[], peer_hi[] are rotated in place. */
int b200_slab_loop(const b200_sweep_desc* desc, void** arrays, void** peer_lo, void** peer_hi,
                   int niters, unsigned long long first_sweep, void* stream);

/* Registers per thread / kernel symbol of the kernel b200_sweep would launch. */
int b200_kernel_info(int test, int dtype, int* regs_per_thread, const char** kernel_name);

/* Number of kernel launches issued by this library since load (for gpu_launches). */
unsigned long long b200_launch_count(void);

/* ---- (2) context API ----------------------------------------------------- */
typedef struct b200_ctx b200_ctx;

typedef struct {
    double kernel_ms_per_sweep;     /* CUDA-event time of the nt-loop / nt */
    double kernel_ms_total;
    int regs_per_thread;
    int launches;                   /* kernel launches issued by b200_run */
    const char* kernel_name;
    int ngpus;
} b200_stats;

int b200_init(b200_ctx** ctx, int ngpus);           /* ngpus <= 0: env B200_NGPUS or 1 */
int b200_plan(b200_ctx* ctx, int test, int dtype, int nx, int ny, int ns,
              const double* scalars, int nscalars);
int b200_alloc(b200_ctx* ctx);
int b200_load(b200_ctx* ctx, int slot, const void* host);
/* Arrays whose interior the first sweep overwrites before anything reads it (the output buffers:
 * laplacian w1, wave13pt w2, divergence u, gradient ux/uy/uz, ... -- b200_slot_interior_dead)
 * only need their boundary shell on the device: the reference's cuda target copies them whole
 * (laplacian/laplacian.c:257-258), the sweeps never touch the shell, the result keeps it.
 * b200_load_shell copies exactly the points outside the interior box (a few strided copies), so
 * "data load time" shrinks by one array (laplacian: 2 -> 1, gradient: 4 -> 1).  Only valid when
 * at least one sweep follows (with nt = 0 the reference reports the untouched array). */
int b200_slot_interior_dead(int test, int slot);
int b200_load_shell(b200_ctx* ctx, int slot, const void* host);
int b200_run(b200_ctx* ctx, int niters, b200_stats* stats);
int b200_result_slot(const b200_ctx* ctx);          /* slot (original numbering) the reference reports f_mean on */
int b200_save(b200_ctx* ctx, int slot, void* host);
/* Asynchronous mode, SINGLE-GPU contexts only (b200_set_async(ctx, 1) on a multi-GPU context returns B200_ERR_STATE: there
 * the ordering between neighbouring slabs' loads, halo pushes and saves relies on the synchronisation this mode removes).
 * New: the reference's drivers synchronise after every phase, laplacian.c:255-262,303-305, 334-340.  With b200_set_async(ctx, 1) the phase calls b200_load / b200_load_shell / b200_run / b200_save only
 * ENQUEUE their copies and sweeps and return; b200_sync(ctx) waits for all of it.  The sweeps run on the context's own
 * stream; the copies of all asynchronous contexts of a device run on two shared streams, one per direction, each copy
 * ordered after everything the context enqueued before it and before everything it enqueues next (copies left on the
 * contexts' own streams did not overlap across contexts: profiles/r2_e2e_pipeline.txt; B200_COPY_STREAMS=0 puts them back).
 * Host arrays must be pinned (b200_host_alloc) and stay untouched until b200_sync; b200_run fills no times
 * (stats: launches, regs, name only).  Two contexts driven alternately overlap one job's device->host copy and
 * sweeps with the next job's host->device copy (PCIe is full duplex): that is how bench.py's e2e leg and a
 * driver that streams many grids through the GPU should call the library. */
/* b200_rewind: start another job (load / run / save) on the buffers already planned and allocated: the rotation
 * state returns to that of a fresh plan, so slot q means the reference's array q again (the reference starts a
 * new process per job; its idxs[] starts at {0,1,2}, laplacian.c:269). */
int b200_rewind(b200_ctx* ctx);
int b200_set_async(b200_ctx* ctx, int on);
int b200_sync(b200_ctx* ctx);
int b200_free(b200_ctx* ctx);                       /* releases device buffers; ctx stays valid for re-plan */
int b200_destroy(b200_ctx* ctx);

/* Pinned host staging buffers (so t_load/t_save and the e2e bench run at PCIe rate). */
int b200_host_alloc(void** ptr, size_t bytes);
int b200_host_free(void* ptr);

/* ---- one-process-per-GPU slabs: peer memory + device-side ordering -------------------------
 * For launchers that run one process per GPU (torchrun): device buffers that can be exported
 * to the neighbour ranks over CUDA IPC, so that b200_sweep's push_lo / push_hi pointers reach
 * the neighbour's ghost planes over NVLink, and a flag per neighbour for ordering: after its
 * sweep a rank signals the neighbours (b200_signal stores `value` to a flag in THEIR memory,
 * release at system scope); before the next sweep it waits (b200_wait spins on its own flag on
 * the device, acquire at system scope) -- no host round trip, no collective on the data path.
 * b200_wait traps after ~20 s instead of hanging. */
/* b200_load_shell for a slab the CALLER owns (one process per GPU): copies the points outside the interior box from the host
 * copy of the slab -- planes (3D) / rows (2D) [mem_lo, mem_hi) of a grid whose split dimension has split_n units, both
 * pointers starting at unit mem_lo -- asynchronously on `stream`.  For the output buffers of a job
 * (b200_slot_interior_dead): ghost planes need not travel either, the neighbours' first sweep pushes them.  Replaces the
 * whole-array cudaMemcpy of laplacian/laplacian.c:257-258 for those arrays. */
int b200_load_shell_slab(int test, int dtype, int nx, int ny, int split_n, int mem_lo, int mem_hi, void* dev_slab,
                         const void* host_slab, void* stream);

#define B200_IPC_HANDLE_BYTES 64
int b200_device_alloc(void** ptr, size_t bytes);
int b200_device_free(void* ptr);
int b200_ipc_export(void* dev_ptr, void* handle /* B200_IPC_HANDLE_BYTES */);
int b200_ipc_import(const void* handle, void** peer_ptr);
int b200_ipc_close(void* peer_ptr);
int b200_signal(void* flag, unsigned long long value, void* stream);
int b200_wait(const void* flag, unsigned long long value, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_STENCIL_H */
