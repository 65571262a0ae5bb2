// b200_api.cu -- the C ABI of libb200stencil.so (see include/b200_stencil.h).
//
// Layer 1: b200_sweep()  -- one sweep on caller-owned device buffers.
// Layer 2: b200_ctx      -- the phases of the reference's cuda-target driver (init / alloc /
//          load / nt-loop with rotation / save / free), with the grid cut into z-slabs
//          (y-slabs for the 2D tests) over 1..8 GPUs of one box.  Ghost planes of the evolving
//          field are refreshed by the sweep kernel itself: it stores the planes its neighbours
//          need straight into their memory (peer pointers over NVLink), so the transfer overlaps
//          the sweep; cross-GPU ordering is one event wait per neighbour per sweep.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "b200_internal.h"

namespace b200 {

// ------------------------------------------------------------------------------------------
// errors, counters
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    // same shape as the reference's CUDA_SAFE_CALL message (<test>/cuda/cuda_profiling.h:11-15)
    set_error("Error \"%s\" at %s:%d (%s)", cudaGetErrorString(e), file, line, what);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return B200_ERR_NO_DEVICE;
    return B200_ERR_CUDA;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------
// test table (mirrors what each reference driver hard-codes; cited in include/b200_stencil.h)
// ------------------------------------------------------------------------------------------
//   name        nd na nsc rot  lo          hi          nr nw  glo ghi xslot
static const b200_test_info g_tests[B200_NTESTS] = {
    { "laplacian",  3, 2, 2, 2, {1, 1, 1}, {1, 1, 1}, 1, 1, 1, 1, 1 },   // laplacian/laplacian.c:81-99
    { "wave13pt",   3, 3, 3, 3, {2, 2, 2}, {2, 2, 2}, 2, 1, 2, 2, 2 },   // wave13pt/wave13pt.c:143-160,495
    { "divergence", 3, 4, 3, 0, {1, 1, 1}, {1, 1, 1}, 3, 1, 1, 1, -1 },  // divergence/divergence.c:81-98
    { "gradient",   3, 4, 3, 0, {1, 1, 1}, {1, 1, 1}, 1, 3, 1, 1, -1 },  // gradient/gradient.c:82-99
    { "uxx1",       3, 6, 2, 2, {2, 2, 2}, {1, 1, 1}, 5, 1, 2, 1, -1 },  // uxx1/uxx1.c:83-107 (u is point-wise)
    { "lapgsrb",    3, 2, 4, 2, {2, 2, 2}, {2, 2, 2}, 1, 1, 2, 2, 1 },   // lapgsrb/lapgsrb.c:81-118
    { "jacobi",     2, 2, 3, 2, {1, 1, 0}, {1, 1, 0}, 1, 1, 1, 1, 1 },   // jacobi/jacobi.F90:60-69
    { "gaussblur",  2, 2, 6, 2, {2, 2, 0}, {2, 2, 0}, 1, 1, 2, 2, 1 },   // gaussblur/gaussblur.c:78-93
    { "gameoflife", 2, 2, 0, 2, {1, 1, 0}, {1, 1, 0}, 1, 1, 1, 1, 1 },   // gameoflife/gameoflife.c:72-92
    { "tricubic",   3, 5, 0, 2, {1, 1, 1}, {2, 2, 2}, 4, 1, 1, 2, 1 },   // tricubic/tricubic.c:138-155,509
    { "tricubic2",  3, 5, 0, 2, {2, 2, 2}, {2, 2, 2}, 4, 1, 1, 2, 1 },   // tricubic2/tricubic2.c:67-79
    { "vecadd",     3, 3, 0, 3, {0, 0, 0}, {0, 0, 0}, 2, 1, 0, 0, -1 },  // vecadd/vecadd.c:66-83
    { "matvec",     2, 3, 0, 0, {0, 0, 0}, {0, 0, 0}, 1, 0, 0, 0, -1 },  // matvec/matvec.c:59-68
    { "sincos",     3, 3, 0, 0, {0, 0, 0}, {0, 0, 0}, 2, 1, 0, 0, -1 },  // sincos/sincos.F90:60-72
    { "matmul",     3, 3, 0, 0, {0, 0, 0}, {0, 0, 0}, 3, 1, 0, 0, -1 },  // matmul/matmul.F90:56-68 (C is read and written)
};

typedef int (*launch_fn)(int, const HostArgs&);
typedef int (*info_fn)(int, KernelInfo*);
static const launch_fn g_launch[B200_NTESTS] = {
    launch_laplacian, launch_wave13pt, launch_divergence, launch_gradient, launch_uxx1, launch_lapgsrb,
    launch_jacobi, launch_gaussblur, launch_gameoflife, launch_tricubic, launch_tricubic /* tricubic2: same kernel, other bounds */,
    launch_vecadd, launch_matvec, launch_sincos, launch_matmul };
static const info_fn g_info[B200_NTESTS] = {
    info_laplacian, info_wave13pt, info_divergence, info_gradient, info_uxx1, info_lapgsrb,
    info_jacobi, info_gaussblur, info_gameoflife, info_tricubic, info_tricubic,
    info_vecadd, info_matvec, info_sincos, info_matmul };

// ------------------------------------------------------------------------------------------
// device bookkeeping
// ------------------------------------------------------------------------------------------
struct DeviceInfo { bool probed, usable; int num_sms; };
static DeviceInfo g_dev[16];
static std::mutex g_dev_mu;

static int probe_device(int dev, DeviceInfo* out)
{
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (dev < 0 || dev >= 16) { set_error("device index %d out of range", dev); return B200_ERR_ARG; }
    if (!g_dev[dev].probed) {
        cudaDeviceProp p;
        B200_CUDA(cudaGetDeviceProperties(&p, dev));
        g_dev[dev].probed = true;
        g_dev[dev].usable = (p.major == 10);      // sm_100a cubin only: no PTX, no other architecture
        g_dev[dev].num_sms = p.multiProcessorCount;
    }
    *out = g_dev[dev];
    if (!out->usable) {
        set_error("device %d is not an sm_100 (B200) device; libb200stencil has no other code path", dev);
        return B200_ERR_NO_DEVICE;
    }
    return B200_OK;
}

// ------------------------------------------------------------------------------------------
// TMA descriptors
// ------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn g_encode = nullptr;

struct TmaEntry { TmaBoxKey key; CUtensorMap map; };
static std::vector<TmaEntry> g_tma_cache;
static std::mutex g_tma_mu;

int get_tensor_map(const TmaBoxKey& key, void* out_map)
{
    std::lock_guard<std::mutex> lk(g_tma_mu);
    for (const TmaEntry& e : g_tma_cache)
        if (!memcmp(&e.key, &key, sizeof(key))) { memcpy(out_map, &e.map, sizeof(CUtensorMap)); return B200_OK; }
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        B200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available"); return B200_ERR_CUDA; }
        g_encode = (encode_tiled_fn)fn;
    }
    TmaEntry e;
    memset(&e, 0, sizeof(e));
    e.key = key;
    const cuuint64_t dims[3] = { (cuuint64_t)key.nx, (cuuint64_t)key.ny, (cuuint64_t)key.ns };
    const cuuint64_t px = key.px ? key.px : key.nx, py = key.py ? key.py : key.ny;
    const cuuint64_t strides[2] = { px * key.esz, px * py * key.esz };
    const cuuint32_t box[3] = { (cuuint32_t)key.bw, (cuuint32_t)key.bh, 1u };
    const cuuint32_t estr[3] = { 1u, 1u, 1u };
    const CUresult r = g_encode(&e.map, key.esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                                3, const_cast<void*>(key.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for %dx%dx%d box %dx%d", (int)r, key.nx, key.ny, key.ns, key.bw, key.bh);
        return B200_ERR_CUDA;
    }
    if (g_tma_cache.size() >= 256) g_tma_cache.erase(g_tma_cache.begin(), g_tma_cache.begin() + 128);
    g_tma_cache.push_back(e);
    memcpy(out_map, &e.map, sizeof(CUtensorMap));
    return B200_OK;
}

static unsigned int* g_done_counter[16] = {};
int get_done_counter(int device, unsigned int** counter)
{
    std::lock_guard<std::mutex> lk(g_tma_mu);
    if (device < 0 || device >= 16) { set_error("device index %d out of range", device); return B200_ERR_ARG; }
    if (!g_done_counter[device]) {
        B200_CUDA(cudaMalloc(&g_done_counter[device], 256));
        B200_CUDA(cudaMemset(g_done_counter[device], 0, 256));
    }
    *counter = g_done_counter[device];
    return B200_OK;
}

static void drop_tensor_maps_for(const void* lo, const void* hi)
{
    std::lock_guard<std::mutex> lk(g_tma_mu);
    for (size_t i = 0; i < g_tma_cache.size();) {
        const char* p = (const char*)g_tma_cache[i].key.ptr;
        if (p >= (const char*)lo && p < (const char*)hi) g_tma_cache.erase(g_tma_cache.begin() + i);
        else i++;
    }
}

static int check_sweep_args(const b200_sweep_desc* d, void* const* arrays)
{
    if (!d || !arrays) { set_error("NULL argument"); return B200_ERR_ARG; }
    if (d->test < 0 || d->test >= B200_NTESTS) { set_error("unknown test id %d", d->test); return B200_ERR_ARG; }
    if (d->dtype != B200_F32 && d->dtype != B200_F64) { set_error("unknown dtype %d", d->dtype); return B200_ERR_ARG; }
    if (d->nx < 0 || d->ny < 0 || d->ns < 0) { set_error("negative extent"); return B200_ERR_ARG; }
    return B200_OK;
}

}  // namespace b200

using namespace b200;

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

const b200_test_info* b200_get_test_info(int test)
{
    if (test < 0 || test >= B200_NTESTS) return nullptr;
    return &g_tests[test];
}

int b200_test_by_name(const char* name)
{
    if (!name) return -1;
    for (int t = 0; t < B200_NTESTS; t++)
        if (!strcmp(name, g_tests[t].name)) return t;
    return -1;
}

const char* b200_last_error(void) { return g_err; }
int b200_api_version(void) { return B200_API_VERSION; }
unsigned long long b200_launch_count(void) { return g_launches.load(); }

unsigned long long b200_interior_points(int test, int nx, int ny, int ns)
{
    const b200_test_info* ti = b200_get_test_info(test);
    if (!ti) return 0;
    if (test == B200_MATVEC) return (unsigned long long)nx * ny;
    if (test == B200_MATMUL) return (unsigned long long)nx * ny * (unsigned long long)ns;
    const long long ex = nx - ti->lo[0] - ti->hi[0], ey = ny - ti->lo[1] - ti->hi[1];
    const long long ez = ti->ndims == 3 ? ns - ti->lo[2] - ti->hi[2] : 1;
    if (ex <= 0 || ey <= 0 || ez <= 0) return 0;
    return (unsigned long long)(ex * ey * ez);
}

int b200_device_count(int* count)
{
    if (!count) { set_error("NULL argument"); return B200_ERR_ARG; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
    int usable = 0;
    for (int d = 0; d < n && d < 16; d++) {
        DeviceInfo di;
        if (probe_device(d, &di) == B200_OK) usable++;
    }
    *count = usable;
    if (!usable) { set_error("no sm_100 (B200) device found; libb200stencil has no CPU fallback"); return B200_ERR_NO_DEVICE; }
    return B200_OK;
}

int b200_sweep(const b200_sweep_desc* desc, void* const* arrays, void* stream)
{
    if (int rc = check_sweep_args(desc, arrays)) return rc;
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    DeviceInfo di;
    if (int rc = probe_device(dev, &di)) return rc;
    HostArgs a{desc, arrays, (cudaStream_t)stream, dev, di.num_sms};
    return g_launch[desc->test](desc->dtype, a);
}

int b200_sweep_loop(const b200_sweep_desc* desc, void** arrays, int niters, void* stream)
{
    if (int rc = check_sweep_args(desc, arrays)) return rc;
    if (niters < 0) { set_error("negative iteration count"); return B200_ERR_ARG; }
    if (desc->push_lo || desc->push_hi) { set_error("b200_sweep_loop: halo push needs per-sweep ordering, use b200_sweep"); return B200_ERR_ARG; }
    const b200_test_info* ti = &g_tests[desc->test];
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    DeviceInfo di;
    if (int rc = probe_device(dev, &di)) return rc;
    b200_sweep_desc d = *desc;
    HostArgs a{&d, arrays, (cudaStream_t)stream, dev, di.num_sms};
    for (int it = 0; it < niters; it++) {
        d.reverse_order = (desc->reverse_order + it) & 1;
        if (int rc = g_launch[desc->test](desc->dtype, a)) return rc;
        // the reference driver's rotation: laplacian.c:299-300 (swap), wave13pt.c:919-920 (3-cycle)
        if (ti->rotation == 2) { void* w = arrays[0]; arrays[0] = arrays[1]; arrays[1] = w; }
        else if (ti->rotation == 3) { void* w = arrays[0]; arrays[0] = arrays[1]; arrays[1] = arrays[2]; arrays[2] = w; }
    }
    return B200_OK;
}

// Fused two-sweep kernels exist for the three 2D stencils; whether a loop USES them is a measured policy:
// the fused pass halves the DRAM traffic but its consumer warps do two sweeps' arithmetic per tile, and with 12
// warps per SM that is latency-bound.  Measured on B200 (profiles/README.md): jacobi at nx >= 1024 gains 13-19 %,
// jacobi at nx = 512 loses (5 overlapped x-tiles for 4), gaussblur (-17 %) and gameoflife (FP64 pipe, -25 %) lose.
// B200_FUSE=1 forces the fused path wherever it exists, B200_FUSE=0 disables it (tests, A/B runs).
static launch_fn fused_kernel(int test)
{
    switch (test) {
    case B200_JACOBI: return launch_jacobi2;
    case B200_GAUSSBLUR: return launch_gaussblur2;
    case B200_GAMEOFLIFE: return launch_gameoflife2;
    default: return nullptr;
    }
}
static launch_fn fused_launch(int test, int nx)
{
    const char* e = getenv("B200_FUSE");           // read per call: tests toggle it
    if (e) return atoi(e) ? fused_kernel(test) : nullptr;
    // B200_TBLOCK=<sweeps per pass> (env row): 1 = one sweep per pass, 2 = the two-sweep kernels wherever they exist
    if (const char* tb = getenv("B200_TBLOCK")) return atoi(tb) >= 2 ? fused_kernel(test) : nullptr;
    return (test == B200_JACOBI && nx >= 1024) ? fused_kernel(test) : nullptr;
}

int b200_sweep2_supported(int test) { return fused_kernel(test) != nullptr; }
int b200_sweep2_profitable(int test, int nx) { return fused_launch(test, nx) != nullptr; }

int b200_sweep2(const b200_sweep_desc* desc, void* const* arrays, void* stream)
{
    if (int rc = check_sweep_args(desc, arrays)) return rc;
    launch_fn fn = fused_kernel(desc->test);
    if (!fn) { set_error("%s has no two-sweep kernel", g_tests[desc->test].name); return B200_ERR_ARG; }
    if (desc->push_lo || desc->push_hi || desc->out_begin || desc->out_end) { set_error("b200_sweep2: whole grid on one GPU only"); return B200_ERR_ARG; }
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    DeviceInfo di;
    if (int rc = probe_device(dev, &di)) return rc;
    HostArgs a{desc, arrays, (cudaStream_t)stream, dev, di.num_sms};
    return fn(desc->dtype, a);
}

int b200_sweep_loop2(const b200_sweep_desc* desc, void** arrays, void** scratch, int niters, void* stream)
{
    if (int rc = check_sweep_args(desc, arrays)) return rc;
    launch_fn fn = fused_launch(desc->test, desc->nx);
    const bool whole = !(desc->push_lo || desc->push_hi || desc->out_begin || desc->out_end);
    int pairs = (fn && whole && scratch && *scratch && niters >= 4) ? (niters - 2) / 2 : 0;
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    DeviceInfo di;
    if (int rc = probe_device(dev, &di)) return rc;
    b200_sweep_desc d = *desc;
    for (int p = 0; p < pairs; p++) {
        void* three[3] = { arrays[0], arrays[1], *scratch };
        d.reverse_order = (desc->reverse_order + p) & 1;
        HostArgs a{&d, three, (cudaStream_t)stream, dev, di.num_sms};
        if (int rc = fn(desc->dtype, a)) return rc;
        void* w = arrays[0]; arrays[0] = *scratch; *scratch = w;      // state t+2 is the new w0
    }
    d = *desc;
    d.reverse_order = (desc->reverse_order + pairs) & 1;
    return b200_sweep_loop(&d, arrays, niters - 2 * pairs, stream);
}

int b200_slab_loop(const b200_sweep_desc* desc, void** arrays, void** peer_lo, void** peer_hi,
                   int niters, unsigned long long first_sweep, void* stream)
{
    if (int rc = check_sweep_args(desc, arrays)) return rc;
    if (niters < 0 || !peer_lo || !peer_hi) { set_error("b200_slab_loop: bad arguments"); return B200_ERR_ARG; }
    const b200_test_info* ti = &g_tests[desc->test];
    if (ti->exchange_slot < 0 || !ti->rotation) { set_error("%s has no exchanged array", ti->name); return B200_ERR_ARG; }
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    DeviceInfo di;
    if (int rc = probe_device(dev, &di)) return rc;
    b200_sweep_desc d = *desc;
    HostArgs a{&d, arrays, (cudaStream_t)stream, dev, di.num_sms};
    const int rot = ti->rotation, out_pos = rot == 3 ? 2 : 1;
    for (int it = 0; it < niters; it++) {
        const unsigned long long n = first_sweep + (unsigned long long)it;
        d.reverse_order = (int)(n & 1ull);
        d.push_lo = peer_lo[out_pos];
        d.push_hi = peer_hi[out_pos];
        d.wait_value = n;                         // flags start at 0: sweep 0 does not wait
        d.signal_value = n + 1;
        if (!d.push_lo) { d.push_lo_count = 0; }
        if (!d.push_hi) { d.push_hi_count = 0; }
        if (!d.push_lo && !d.push_hi) { set_error("b200_slab_loop: no neighbour"); return B200_ERR_ARG; }
        if (int rc = g_launch[desc->test](desc->dtype, a)) return rc;
        void** sets[3] = { arrays, peer_lo, peer_hi };
        for (void** w : sets) {
            if (rot == 2) { void* t = w[0]; w[0] = w[1]; w[1] = t; }
            else { void* t = w[0]; w[0] = w[1]; w[1] = w[2]; w[2] = t; }
        }
    }
    return B200_OK;
}

int b200_kernel_info(int test, int dtype, int* regs_per_thread, const char** kernel_name)
{
    if (test < 0 || test >= B200_NTESTS) { set_error("unknown test id %d", test); return B200_ERR_ARG; }
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    DeviceInfo di;
    if (int rc = probe_device(dev, &di)) return rc;
    KernelInfo ki{};
    if (int rc = g_info[test](dtype, &ki)) return rc;
    if (regs_per_thread) *regs_per_thread = ki.regs;
    if (kernel_name) *kernel_name = g_tests[test].name;
    return B200_OK;
}

int b200_host_alloc(void** ptr, size_t bytes)
{
    if (!ptr) { set_error("NULL argument"); return B200_ERR_ARG; }
    B200_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return B200_OK;
}

int b200_host_free(void* ptr)
{
    if (ptr) B200_CUDA(cudaFreeHost(ptr));
    return B200_OK;
}

// ------------------------------------------------------------------------------------------
// one-process-per-GPU helpers: IPC-exportable buffers, device-side signal / wait
// ------------------------------------------------------------------------------------------
}  // extern "C"

__global__ void b200_signal_kernel(unsigned long long* flag, unsigned long long value)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

__global__ void b200_wait_kernel(const unsigned long long* flag, unsigned long long value, unsigned long long timeout_ns)
{
    unsigned long long t0, now, cur;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(cur) : "l"(flag) : "memory");
        if (cur >= value) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > timeout_ns) __trap();       // a neighbour died: fail instead of hanging the box
        __nanosleep(200);
    }
}

extern "C" {

int b200_device_alloc(void** ptr, size_t bytes)
{
    if (!ptr) { set_error("NULL argument"); return B200_ERR_ARG; }
    B200_CUDA(cudaMalloc(ptr, bytes ? bytes : 16));
    return B200_OK;
}

int b200_device_free(void* ptr)
{
    if (ptr) B200_CUDA(cudaFree(ptr));
    return B200_OK;
}

int b200_ipc_export(void* dev_ptr, void* handle)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == B200_IPC_HANDLE_BYTES, "IPC handle size");
    if (!dev_ptr || !handle) { set_error("NULL argument"); return B200_ERR_ARG; }
    B200_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle, dev_ptr));
    return B200_OK;
}

int b200_ipc_import(const void* handle, void** peer_ptr)
{
    if (!handle || !peer_ptr) { set_error("NULL argument"); return B200_ERR_ARG; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    B200_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return B200_OK;
}

int b200_ipc_close(void* peer_ptr)
{
    if (peer_ptr) B200_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return B200_OK;
}

int b200_signal(void* flag, unsigned long long value, void* stream)
{
    if (!flag) { set_error("NULL argument"); return B200_ERR_ARG; }
    b200_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)flag, value);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

int b200_wait(const void* flag, unsigned long long value, void* stream)
{
    if (!flag) { set_error("NULL argument"); return B200_ERR_ARG; }
    b200_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const unsigned long long*)flag, value, 20000000000ull);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

// ------------------------------------------------------------------------------------------
// context API
// ------------------------------------------------------------------------------------------
struct b200_slab {
    int dev;
    int own_lo, own_hi;         // owned range in the split dimension (global coordinates)
    int mem_lo, mem_hi;         // stored range (owned + ghosts), global coordinates
    void* arr[B200_MAX_ARRAYS]; // slot order (ORIGINAL numbering; rotation is applied per sweep)
    void* scratch;              // third buffer of the 2-buffer tests that have a fused two-sweep kernel (1 GPU)
    int scratch_shell_slot;     // the slot (arr[] numbering) whose boundary shell the scratch buffer carries, -1: none yet
    cudaStream_t stream;
    cudaEvent_t done[2];        // sweep-complete events, alternating
    cudaEvent_t t0, t1;         // timing
    cudaEvent_t fork, join;     // asynchronous mode: hand a copy to a per-direction copy stream and take it back
};

struct b200_ctx {
    int ngpus;
    int devs[8];
    bool planned, allocated;
    int test, dtype, nx, ny, ns;
    double sc[B200_MAX_SCALARS];
    int split_n;                // extent of the split dimension
    size_t unit;                // elements per plane (3D) / row (2D) of the split dimension
    b200_slab slab[8];
    int idxs[3];                // rotation state, like the reference's idxs[] (laplacian.c:269,300)
    bool peer_enabled;
    bool async_mode;            // b200_set_async: the phase calls enqueue only, b200_sync waits
};

// wait for the slabs' streams -- unless the context is in asynchronous mode (b200_set_async)
static int sync_streams(b200_ctx* c, bool force = false)
{
    if (c->async_mode && !force) return B200_OK;
    for (int g = 0; g < c->ngpus; g++) {
        B200_CUDA(cudaSetDevice(c->slab[g].dev));
        B200_CUDA(cudaStreamSynchronize(c->slab[g].stream));
    }
    return B200_OK;
}

static size_t esz_of(int dtype) { return dtype == B200_F32 ? 4 : 8; }

// Asynchronous mode (b200_set_async) runs the host<->device copies of ALL contexts of a device on two shared streams, one
// per direction, instead of on each context's own stream: with the copies of two alternating contexts on their own streams
// the return copy of one job did not overlap the upload of the next (measured: profiles/r2_e2e_pipeline.txt, 14.1 ms per
// wave13pt job against 10.7 ms for the same copies issued on one stream per direction).  The context's stream stays its one
// timeline: a copy forks from it (the copy stream waits for everything enqueued so far) and joins back (the context's
// stream waits for the copy), so b200_sync and the order of the phase calls mean what they meant.
// B200_COPY_STREAMS=0: copies on the context's stream as in synchronous mode.
static cudaStream_t g_copy_stream[16][2] = {};
static std::mutex g_copy_mu;
enum { COPY_UP = 0, COPY_DOWN = 1 };
static bool copy_streams_enabled()
{
    static const bool on = [] { const char* e = getenv("B200_COPY_STREAMS"); return !(e && e[0] == '0'); }();
    return on;
}
static int copy_fork(b200_ctx* c, b200_slab& s, int dir, cudaStream_t* st)
{
    *st = s.stream;
    if (!c->async_mode || !copy_streams_enabled() || s.dev < 0 || s.dev >= 16) return B200_OK;
    {
        std::lock_guard<std::mutex> lk(g_copy_mu);
        if (!g_copy_stream[s.dev][dir]) B200_CUDA(cudaStreamCreateWithFlags(&g_copy_stream[s.dev][dir], cudaStreamNonBlocking));
    }
    *st = g_copy_stream[s.dev][dir];
    B200_CUDA(cudaEventRecord(s.fork, s.stream));
    B200_CUDA(cudaStreamWaitEvent(*st, s.fork, 0));
    return B200_OK;
}
static int copy_join(b200_slab& s, cudaStream_t st)
{
    if (st == s.stream) return B200_OK;
    B200_CUDA(cudaEventRecord(s.join, st));
    B200_CUDA(cudaStreamWaitEvent(s.stream, s.join, 0));
    return B200_OK;
}

int b200_init(b200_ctx** out, int ngpus)
{
    if (!out) { set_error("NULL argument"); return B200_ERR_ARG; }
    int avail = 0;
    if (int rc = b200_device_count(&avail)) return rc;
    if (ngpus <= 0) {
        const char* e = getenv("B200_NGPUS");
        ngpus = e ? atoi(e) : 1;
        if (ngpus <= 0) ngpus = 1;
    }
    if (ngpus > avail || ngpus > 8) { set_error("%d GPUs requested, %d usable", ngpus, avail); return B200_ERR_ARG; }
    b200_ctx* c = (b200_ctx*)calloc(1, sizeof(b200_ctx));
    if (!c) { set_error("out of host memory"); return B200_ERR_NOMEM; }
    c->ngpus = ngpus;
    for (int g = 0; g < ngpus; g++) c->devs[g] = g;
    // warm the context of every device now, so that it lands in "init time" like the reference's probe
    for (int g = 0; g < ngpus; g++) {
        B200_CUDA(cudaSetDevice(c->devs[g]));
        B200_CUDA(cudaFree(0));
    }
    if (ngpus > 1) {
        for (int g = 0; g < ngpus; g++) {
            B200_CUDA(cudaSetDevice(c->devs[g]));
            for (int n = g - 1; n <= g + 1; n += 2) {
                if (n < 0 || n >= ngpus) continue;
                int can = 0;
                B200_CUDA(cudaDeviceCanAccessPeer(&can, c->devs[g], c->devs[n]));
                if (!can) { set_error("no peer access between GPU %d and %d", g, n); free(c); return B200_ERR_CUDA; }
                cudaError_t e = cudaDeviceEnablePeerAccess(c->devs[n], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) { free(c); return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__); }
            }
        }
        c->peer_enabled = true;
        B200_CUDA(cudaSetDevice(c->devs[0]));
    }
    *out = c;
    return B200_OK;
}

int b200_plan(b200_ctx* c, int test, int dtype, int nx, int ny, int ns, const double* scalars, int nscalars)
{
    if (!c) { set_error("NULL context"); return B200_ERR_ARG; }
    if (c->allocated) { set_error("b200_plan: free the previous plan first"); return B200_ERR_STATE; }
    const b200_test_info* ti = b200_get_test_info(test);
    if (!ti) { set_error("unknown test id %d", test); return B200_ERR_ARG; }
    if (dtype != B200_F32 && dtype != B200_F64) { set_error("unknown dtype %d", dtype); return B200_ERR_ARG; }
    if (nx < 0 || ny < 0 || ns < 0) { set_error("negative extent"); return B200_ERR_ARG; }
    if (nscalars != ti->nscalars || (nscalars > 0 && !scalars)) {
        set_error("%s takes %d scalars, got %d", ti->name, ti->nscalars, nscalars);
        return B200_ERR_ARG;
    }
    c->test = test; c->dtype = dtype; c->nx = nx; c->ny = ny; c->ns = ti->ndims == 3 ? ns : 1;
    memset(c->sc, 0, sizeof(c->sc));
    for (int q = 0; q < nscalars; q++) c->sc[q] = scalars[q];
    c->split_n = ti->ndims == 3 ? c->ns : ny;
    c->unit = ti->ndims == 3 ? (size_t)nx * ny : (size_t)nx;
    if (test == B200_MATVEC) { c->split_n = ny; c->unit = (size_t)nx; }
    if (test == B200_MATMUL) { c->split_n = ns; c->unit = (size_t)nx; }     // columns of B and C are split, A replicated
    // contiguous, near-equal slabs of the whole extent (boundary planes belong to the end slabs)
    const int G = c->ngpus;
    for (int g = 0; g < G; g++) {
        b200_slab& s = c->slab[g];
        s.dev = c->devs[g];
        s.own_lo = (int)((long long)c->split_n * g / G);
        s.own_hi = (int)((long long)c->split_n * (g + 1) / G);
        s.mem_lo = s.own_lo - ti->zghost_lo; if (s.mem_lo < 0) s.mem_lo = 0;
        s.mem_hi = s.own_hi + ti->zghost_hi; if (s.mem_hi > c->split_n) s.mem_hi = c->split_n;
        if (G > 1 && s.own_hi - s.own_lo < ti->zghost_lo + ti->zghost_hi + 1) {
            set_error("%s: extent %d too small for %d slabs", ti->name, c->split_n, G);
            return B200_ERR_ARG;
        }
    }
    c->idxs[0] = 0; c->idxs[1] = 1; c->idxs[2] = 2;
    c->planned = true;
    return B200_OK;
}

// elements of array `slot` stored on a slab
static size_t slab_elems(const b200_ctx* c, const b200_slab& s, int slot)
{
    if (c->test == B200_MATVEC) {
        if (slot == 1) return (size_t)c->nx;                        // x is replicated
        if (slot == 2) return (size_t)(s.mem_hi - s.mem_lo);        // y rows
    }
    if (c->test == B200_MATMUL) {
        if (slot == 0) return (size_t)c->nx * c->ny;                // A is replicated
        return (size_t)(slot == 1 ? c->ny : c->nx) * (size_t)(s.mem_hi - s.mem_lo);   // columns of B / C
    }
    return c->unit * (size_t)(s.mem_hi - s.mem_lo);
}
static size_t slab_unit(const b200_ctx* c, int slot)
{
    if (c->test == B200_MATVEC) return slot == 0 ? (size_t)c->nx : slot == 1 ? 0 : 1;
    if (c->test == B200_MATMUL) return slot == 0 ? 0 : slot == 1 ? (size_t)c->ny : (size_t)c->nx;
    return c->unit;
}

int b200_alloc(b200_ctx* c)
{
    if (!c || !c->planned) { set_error("b200_alloc: no plan"); return B200_ERR_STATE; }
    if (c->allocated) { set_error("b200_alloc: already allocated"); return B200_ERR_STATE; }
    const b200_test_info* ti = b200_get_test_info(c->test);
    const size_t esz = esz_of(c->dtype);
    for (int g = 0; g < c->ngpus; g++) {
        b200_slab& s = c->slab[g];
        B200_CUDA(cudaSetDevice(s.dev));
        for (int q = 0; q < ti->narrays; q++) {
            size_t bytes = slab_elems(c, s, q) * esz;
            B200_CUDA(cudaMalloc(&s.arr[q], bytes ? bytes : 16));
            // B200_POISON=1 (tests): start from NaN patterns so that anything a shell-only load or a sweep
            // fails to write shows up in the comparison
            if (getenv("B200_POISON")) B200_CUDA(cudaMemset(s.arr[q], 0xFF, bytes ? bytes : 16));
        }
        s.scratch = nullptr;
        s.scratch_shell_slot = -1;
        if (c->ngpus == 1 && fused_launch(c->test, c->nx)) {
            size_t bytes = slab_elems(c, s, 0) * esz;
            B200_CUDA(cudaMalloc(&s.scratch, bytes ? bytes : 16));
            if (getenv("B200_POISON")) B200_CUDA(cudaMemset(s.scratch, 0xFF, bytes ? bytes : 16));
            KernelInfo k2{};
            if (c->test == B200_JACOBI) { if (int rc = info_jacobi2(c->dtype, &k2)) return rc; }
            else if (c->test == B200_GAUSSBLUR) { if (int rc = info_gaussblur2(c->dtype, &k2)) return rc; }
            else if (c->test == B200_GAMEOFLIFE) { if (int rc = info_gameoflife2(c->dtype, &k2)) return rc; }
        }
        B200_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        B200_CUDA(cudaEventCreateWithFlags(&s.done[0], cudaEventDisableTiming));
        B200_CUDA(cudaEventCreateWithFlags(&s.done[1], cudaEventDisableTiming));
        B200_CUDA(cudaEventCreate(&s.t0));
        B200_CUDA(cudaEventCreate(&s.t1));
        B200_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        B200_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
        // one-time per-device kernel setup (module load, shared-memory attribute, occupancy) belongs
        // to the allocation phase, not to the first timed sweep
        KernelInfo ki{};
        if (int rc = g_info[c->test](c->dtype, &ki)) return rc;
        unsigned int* dc = nullptr;
        if (c->ngpus > 1) { if (int rc = get_done_counter(s.dev, &dc)) return rc; }
    }
    B200_CUDA(cudaSetDevice(c->slab[0].dev));
    c->allocated = true;
    return B200_OK;
}

static int upload_shell(b200_ctx* c, int slot, const void* host, bool to_scratch);

int b200_load(b200_ctx* c, int slot, const void* host)
{
    if (!c || !c->allocated) { set_error("b200_load: not allocated"); return B200_ERR_STATE; }
    const b200_test_info* ti = b200_get_test_info(c->test);
    if (slot < 0 || slot >= ti->narrays || !host) { set_error("b200_load: bad slot/pointer"); return B200_ERR_ARG; }
    const size_t esz = esz_of(c->dtype);
    for (int g = 0; g < c->ngpus; g++) {
        b200_slab& s = c->slab[g];
        B200_CUDA(cudaSetDevice(s.dev));
        const size_t off = slab_unit(c, slot) * (size_t)s.mem_lo * esz;
        cudaStream_t st;
        if (int rc = copy_fork(c, s, COPY_UP, &st)) return rc;
        B200_CUDA(cudaMemcpyAsync(s.arr[slot], (const char*)host + off, slab_elems(c, s, slot) * esz,
                                  cudaMemcpyHostToDevice, st));
        if (int rc = copy_join(s, st)) return rc;
    }
    if (int rc = sync_streams(c)) return rc;
    B200_CUDA(cudaSetDevice(c->slab[0].dev));
    // the scratch buffer of the fused two-sweep kernels takes over w0's role: it needs w0's boundary shell
    if (slot == 0 && c->slab[0].scratch) {
        c->slab[0].scratch_shell_slot = 0;
        return upload_shell(c, 0, host, true);
    }
    return B200_OK;
}

int b200_slot_interior_dead(int test, int slot)
{
    switch (test) {
    case B200_LAPLACIAN: case B200_UXX1: case B200_LAPGSRB: case B200_JACOBI: case B200_GAUSSBLUR:
    case B200_GAMEOFLIFE: case B200_TRICUBIC: case B200_TRICUBIC2: return slot == 1;
    case B200_WAVE13PT: case B200_VECADD: case B200_MATVEC: case B200_SINCOS: return slot == 2;
    case B200_DIVERGENCE: return slot == 0;
    case B200_GRADIENT: return slot >= 1 && slot <= 3;
    default: return 0;                      // matmul: C is read (it accumulates)
    }
}

// Copies the points outside the interior box of a test from a host slab into the device copy of the same slab: planes
// (3D) / rows (2D) [mem_lo, mem_hi) of a grid whose split dimension has split_n units; `src` and `dst` both start at unit
// mem_lo.  A few strided copies on `stream`.
static int shell_copy(const b200_test_info* ti, size_t esz, int nx, int ny, int split_n, int mem_lo, int mem_hi,
                      char* dst, const char* src, cudaStream_t stream)
{
    const bool d3 = ti->ndims == 3;
    const int lox = ti->lo[0], hix = ti->hi[0], loy = ti->lo[1], hiy = ti->hi[1];
    const int loz = d3 ? ti->lo[2] : 0, hiz = d3 ? ti->hi[2] : 0;
    const size_t row_b = (size_t)nx * esz, plane_b = row_b * (size_t)ny;
    const int n_split = mem_hi - mem_lo;                   // planes (3D) / rows (2D) stored in the slab
    const size_t rows = d3 ? (size_t)ny * n_split : (size_t)n_split;
    if (rows == 0) return B200_OK;
    const bool no_interior = nx <= lox + hix || (d3 && ny <= loy + hiy);
    if (no_interior) {                                       // degenerate: everything is shell
        B200_CUDA(cudaMemcpyAsync(dst, src, rows * row_b, cudaMemcpyHostToDevice, stream));
        return B200_OK;
    }
    // (1) x edges of every row: the right edge of row r and the left edge of row r+1 are adjacent
    if (lox + hix > 0) {
        if (lox) B200_CUDA(cudaMemcpyAsync(dst, src, (size_t)lox * esz, cudaMemcpyHostToDevice, stream));
        if (hix) B200_CUDA(cudaMemcpyAsync(dst + rows * row_b - (size_t)hix * esz, src + rows * row_b - (size_t)hix * esz,
                                           (size_t)hix * esz, cudaMemcpyHostToDevice, stream));
        if (rows > 1)
            B200_CUDA(cudaMemcpy2DAsync(dst + row_b - (size_t)hix * esz, row_b, src + row_b - (size_t)hix * esz, row_b,
                                        (size_t)(lox + hix) * esz, rows - 1, cudaMemcpyHostToDevice, stream));
    }
    // (2) whole rows / planes outside the interior in the split dimension (global coordinates)
    const int glo = d3 ? loz : loy, ghi = split_n - (d3 ? hiz : hiy);         // interior [glo, ghi) of the split dim
    const size_t unit_b = d3 ? plane_b : row_b;
    int a0 = mem_lo, a1 = mem_hi < glo ? mem_hi : glo;                         // below the interior
    if (a1 > a0) B200_CUDA(cudaMemcpyAsync(dst + (size_t)(a0 - mem_lo) * unit_b, src + (size_t)(a0 - mem_lo) * unit_b,
                                           (size_t)(a1 - a0) * unit_b, cudaMemcpyHostToDevice, stream));
    a0 = mem_lo > ghi ? mem_lo : ghi; a1 = mem_hi;                             // above the interior
    if (a1 > a0) B200_CUDA(cudaMemcpyAsync(dst + (size_t)(a0 - mem_lo) * unit_b, src + (size_t)(a0 - mem_lo) * unit_b,
                                           (size_t)(a1 - a0) * unit_b, cudaMemcpyHostToDevice, stream));
    // (3) 3D: the y-shell rows of every plane
    if (d3) {
        if (loy) B200_CUDA(cudaMemcpy2DAsync(dst, plane_b, src, plane_b, (size_t)loy * row_b, n_split, cudaMemcpyHostToDevice, stream));
        if (hiy) B200_CUDA(cudaMemcpy2DAsync(dst + plane_b - (size_t)hiy * row_b, plane_b, src + plane_b - (size_t)hiy * row_b, plane_b,
                                             (size_t)hiy * row_b, n_split, cudaMemcpyHostToDevice, stream));
    }
    return B200_OK;
}

// Context form: `host` is a whole array; every slab takes its part (to_scratch: into the slabs' scratch buffers instead).
static int upload_shell(b200_ctx* c, int slot, const void* host, bool to_scratch)
{
    const b200_test_info* ti = b200_get_test_info(c->test);
    const size_t esz = esz_of(c->dtype);
    for (int g = 0; g < c->ngpus; g++) {
        b200_slab& s = c->slab[g];
        B200_CUDA(cudaSetDevice(s.dev));
        char* dst = (char*)(to_scratch ? s.scratch : s.arr[slot]);
        if (!dst) continue;
        const char* src = (const char*)host + c->unit * (size_t)s.mem_lo * esz;
        cudaStream_t st;
        if (int rc = copy_fork(c, s, COPY_UP, &st)) return rc;
        if (int rc = shell_copy(ti, esz, c->nx, c->ny, c->split_n, s.mem_lo, s.mem_hi, dst, src, st)) return rc;
        if (int rc = copy_join(s, st)) return rc;
    }
    if (int rc = sync_streams(c)) return rc;
    B200_CUDA(cudaSetDevice(c->slab[0].dev));
    return B200_OK;
}

// Stateless form for launchers that own the slab buffers (one process per GPU): see include/b200_stencil.h.
int b200_load_shell_slab(int test, int dtype, int nx, int ny, int split_n, int mem_lo, int mem_hi, void* dev_slab,
                         const void* host_slab, void* stream)
{
    const b200_test_info* ti = b200_get_test_info(test);
    if (!ti || (dtype != B200_F32 && dtype != B200_F64)) { set_error("b200_load_shell_slab: bad test / dtype"); return B200_ERR_ARG; }
    if (!dev_slab || !host_slab || nx < 0 || ny < 0 || mem_lo < 0 || mem_hi < mem_lo || mem_hi > split_n) {
        set_error("b200_load_shell_slab: bad arguments");
        return B200_ERR_ARG;
    }
    if (test == B200_MATVEC || test == B200_MATMUL || test == B200_VECADD || test == B200_SINCOS) return B200_OK;   // no shell
    return shell_copy(ti, esz_of(dtype), nx, ny, split_n, mem_lo, mem_hi, (char*)dev_slab, (const char*)host_slab, (cudaStream_t)stream);
}

int b200_load_shell(b200_ctx* c, int slot, const void* host)
{
    if (!c || !c->allocated) { set_error("b200_load_shell: not allocated"); return B200_ERR_STATE; }
    const b200_test_info* ti = b200_get_test_info(c->test);
    if (slot < 0 || slot >= ti->narrays || !host) { set_error("b200_load_shell: bad slot/pointer"); return B200_ERR_ARG; }
    if (!b200_slot_interior_dead(c->test, slot)) return b200_load(c, slot, host);
    if (c->test == B200_MATVEC || c->test == B200_VECADD || c->test == B200_SINCOS) return B200_OK;   // no shell at all
    return upload_shell(c, slot, host, false);
}

int b200_run(b200_ctx* c, int niters, b200_stats* stats)
{
    if (!c || !c->allocated) { set_error("b200_run: not allocated"); return B200_ERR_STATE; }
    if (niters < 0) { set_error("negative iteration count"); return B200_ERR_ARG; }
    const b200_test_info* ti = b200_get_test_info(c->test);
    const int G = c->ngpus;
    const bool exchange = G > 1 && ti->exchange_slot >= 0;
    const unsigned long long launches0 = b200_launch_count();

    for (int g = 0; g < G; g++) {
        B200_CUDA(cudaSetDevice(c->slab[g].dev));
        B200_CUDA(cudaEventRecord(c->slab[g].t0, c->slab[g].stream));
    }
    const int out_pos = ti->rotation == 3 ? 2 : 1;     // position of the written array in rotated order
    // temporal blocking (1 GPU, tests with a fused kernel): (niters-2)/2 two-sweep passes, then single sweeps so
    // that the last two states -- the ones the reference reports -- are both in memory
    // A fused pass writes state t+2 into the scratch buffer, which must carry the boundary shell of the array in the w0
    // role (arr[idxs[0]]).  b200_load seeds it with slot 0's shell; after an ODD number of sweeps the roles are swapped
    // (idxs = {1,0}), so a later b200_run first does ONE single sweep to realign them (`lead`), then the pairs.
    int lead = 0, pairs = 0;
    launch_fn fused = nullptr;
    if (G == 1 && c->slab[0].scratch && c->slab[0].scratch_shell_slot >= 0 && (fused = fused_launch(c->test, c->nx))) {
        lead = c->idxs[0] == c->slab[0].scratch_shell_slot ? 0 : 1;
        pairs = niters - lead >= 4 ? (niters - lead - 2) / 2 : 0;
    }
    for (int it = 0; it < niters; it++) {
        if (it == lead && pairs > 0) {
            b200_slab& s = c->slab[0];
            B200_CUDA(cudaSetDevice(s.dev));
            for (int p = 0; p < pairs; p++) {
                b200_sweep_desc d;
                memset(&d, 0, sizeof(d));
                d.test = c->test; d.dtype = c->dtype;
                d.nx = c->nx; d.ny = c->ny; d.ns = c->ns;
                memcpy(d.scalars, c->sc, sizeof(d.scalars));
                d.reverse_order = (lead + p) & 1;
                void* three[3] = { s.arr[c->idxs[0]], s.arr[c->idxs[1]], s.scratch };
                DeviceInfo di;
                if (int rc = probe_device(s.dev, &di)) return rc;
                HostArgs a{&d, three, s.stream, s.dev, di.num_sms};
                if (int rc = fused(c->dtype, a)) return rc;
                void* w = s.arr[c->idxs[0]]; s.arr[c->idxs[0]] = s.scratch; s.scratch = w;   // two swaps = the same roles
            }
            it += 2 * pairs;
            pairs = 0;
            if (it >= niters) break;      // cannot happen (two single sweeps always follow), kept for safety
        }
        for (int g = 0; g < G; g++) {
            b200_slab& s = c->slab[g];
            B200_CUDA(cudaSetDevice(s.dev));
            if (exchange && it > 0) {
                // ghosts of this sweep's input were pushed by the neighbours' previous sweep
                if (g > 0)     B200_CUDA(cudaStreamWaitEvent(s.stream, c->slab[g - 1].done[(it - 1) & 1], 0));
                if (g < G - 1) B200_CUDA(cudaStreamWaitEvent(s.stream, c->slab[g + 1].done[(it - 1) & 1], 0));
            }
            b200_sweep_desc d;
            memset(&d, 0, sizeof(d));
            d.test = c->test; d.dtype = c->dtype;
            d.nx = c->nx; d.ny = c->ny; d.ns = c->ns;
            const int mem_n = s.mem_hi - s.mem_lo;
            if (c->test == B200_MATVEC) d.ny = mem_n;
            else if (ti->ndims == 3) d.ns = mem_n;
            else d.ny = mem_n;
            memcpy(d.scalars, c->sc, sizeof(d.scalars));
            d.reverse_order = it & 1;
            // output range: owned planes inside the global interior, in local coordinates
            const int split_dim = ti->ndims == 3 ? 2 : 1;
            int lo = s.own_lo, hi = s.own_hi;
            const int ilo = ti->lo[split_dim], ihi = c->split_n - ti->hi[split_dim];
            if (lo < ilo) lo = ilo;
            if (hi > ihi) hi = ihi;
            void* arrays[B200_MAX_ARRAYS];
            for (int q = 0; q < ti->narrays; q++) arrays[q] = s.arr[q];
            if (ti->rotation) for (int q = 0; q < ti->rotation; q++) arrays[q] = s.arr[c->idxs[q]];
            if (hi > lo) {
                d.out_begin = lo - s.mem_lo;
                d.out_end = hi - s.mem_lo;
                if (exchange) {
                    // our lowest owned planes are the upper ghosts of slab g-1; our highest owned
                    // planes are the lower ghosts of slab g+1
                    if (g > 0) {
                        const b200_slab& n = c->slab[g - 1];
                        const int cnt = n.mem_hi - n.own_hi;              // = zghost_hi
                        d.push_lo = n.arr[c->idxs[out_pos]];
                        d.push_lo_src_plane = s.own_lo - s.mem_lo;
                        d.push_lo_dst_plane = n.own_hi - n.mem_lo;
                        d.push_lo_count = cnt;
                    }
                    if (g < G - 1) {
                        const b200_slab& n = c->slab[g + 1];
                        const int cnt = n.own_lo - n.mem_lo;              // = zghost_lo
                        d.push_hi = n.arr[c->idxs[out_pos]];
                        d.push_hi_src_plane = (s.own_hi - cnt) - s.mem_lo;
                        d.push_hi_dst_plane = 0;
                        d.push_hi_count = cnt;
                    }
                }
                HostArgs a{&d, arrays, s.stream, s.dev, 0};
                DeviceInfo di;
                if (int rc = probe_device(s.dev, &di)) return rc;
                a.num_sms = di.num_sms;
                if (int rc = g_launch[c->test](c->dtype, a)) return rc;
            }
            if (exchange) B200_CUDA(cudaEventRecord(s.done[it & 1], s.stream));
        }
        // rotate exactly like the reference driver (laplacian.c:299-300, wave13pt.c:919-920)
        if (ti->rotation == 2) { int t = c->idxs[0]; c->idxs[0] = c->idxs[1]; c->idxs[1] = t; }
        else if (ti->rotation == 3) { int t = c->idxs[0]; c->idxs[0] = c->idxs[1]; c->idxs[1] = c->idxs[2]; c->idxs[2] = t; }
    }
    float ms_max = 0.f;
    for (int g = 0; g < G; g++) {
        B200_CUDA(cudaSetDevice(c->slab[g].dev));
        B200_CUDA(cudaEventRecord(c->slab[g].t1, c->slab[g].stream));
    }
    for (int g = 0; g < G && !c->async_mode; g++) {      // asynchronous mode: no wait, no times (b200_sync + own events)
        B200_CUDA(cudaSetDevice(c->slab[g].dev));
        B200_CUDA(cudaStreamSynchronize(c->slab[g].stream));
        float ms = 0.f;
        B200_CUDA(cudaEventElapsedTime(&ms, c->slab[g].t0, c->slab[g].t1));
        if (ms > ms_max) ms_max = ms;
    }
    B200_CUDA(cudaSetDevice(c->slab[0].dev));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->kernel_ms_total = ms_max;
        stats->kernel_ms_per_sweep = niters > 0 ? ms_max / niters : 0.0;
        stats->launches = (int)(b200_launch_count() - launches0);
        stats->ngpus = G;
        stats->kernel_name = ti->name;
        KernelInfo ki{};
        if (g_info[c->test](c->dtype, &ki) == B200_OK) stats->regs_per_thread = ki.regs;
    }
    return B200_OK;
}

int b200_result_slot(const b200_ctx* c)
{
    if (!c || !c->planned) return -1;
    const b200_test_info* ti = b200_get_test_info(c->test);
    if (ti->rotation) return c->idxs[1];          // laplacian.c:307-313, wave13pt.c:928-932
    switch (c->test) {
    case B200_DIVERGENCE: return 0;               // u       divergence.c:378-381
    case B200_GRADIENT:   return 1;               // ux (+uy,uz: slots 2,3)  gradient.c:388-391
    default:              return 2;               // matvec y, sincos xy, matmul C
    }
}

int b200_save(b200_ctx* c, int slot, void* host)
{
    if (!c || !c->allocated) { set_error("b200_save: not allocated"); return B200_ERR_STATE; }
    const b200_test_info* ti = b200_get_test_info(c->test);
    if (slot < 0 || slot >= ti->narrays || !host) { set_error("b200_save: bad slot/pointer"); return B200_ERR_ARG; }
    const size_t esz = esz_of(c->dtype);
    for (int g = 0; g < c->ngpus; g++) {
        b200_slab& s = c->slab[g];
        B200_CUDA(cudaSetDevice(s.dev));
        const size_t unit = slab_unit(c, slot);
        if (unit == 0 && g != 0) continue;      // replicated (matvec x): slab 0 holds the whole thing
        cudaStream_t st;
        if (int rc = copy_fork(c, s, COPY_DOWN, &st)) return rc;
        if (unit == 0) {
            B200_CUDA(cudaMemcpyAsync(host, s.arr[slot], slab_elems(c, s, slot) * esz, cudaMemcpyDeviceToHost, st));
        } else {
            const size_t src_off = unit * (size_t)(s.own_lo - s.mem_lo) * esz;
            const size_t dst_off = unit * (size_t)s.own_lo * esz;
            B200_CUDA(cudaMemcpyAsync((char*)host + dst_off, (const char*)s.arr[slot] + src_off,
                                      unit * (size_t)(s.own_hi - s.own_lo) * esz, cudaMemcpyDeviceToHost, st));
        }
        if (int rc = copy_join(s, st)) return rc;
    }
    if (int rc = sync_streams(c)) return rc;
    B200_CUDA(cudaSetDevice(c->slab[0].dev));
    return B200_OK;
}

int b200_rewind(b200_ctx* c)
{
    if (!c || !c->planned) { set_error("b200_rewind: not planned"); return B200_ERR_STATE; }
    c->idxs[0] = 0; c->idxs[1] = 1; c->idxs[2] = 2;      // enqueue-order state only: legal while work is in flight
    return B200_OK;
}

int b200_set_async(b200_ctx* c, int on)
{
    if (!c) { set_error("NULL context"); return B200_ERR_ARG; }
    // Multi-GPU contexts order neighbouring slabs' loads, halo pushes and saves through the host synchronisation that
    // asynchronous mode removes (a neighbour's sweep-0 push could race a b200_load of the same buffer): single GPU only.
    if (on && c->ngpus > 1) { set_error("b200_set_async: single-GPU contexts only (ngpus = %d)", c->ngpus); return B200_ERR_STATE; }
    if (!on && c->async_mode && c->allocated) { if (int rc = sync_streams(c, true)) return rc; }
    c->async_mode = on != 0;
    return B200_OK;
}

int b200_sync(b200_ctx* c)
{
    if (!c) { set_error("NULL context"); return B200_ERR_ARG; }
    if (!c->allocated) return B200_OK;
    if (int rc = sync_streams(c, true)) return rc;
    B200_CUDA(cudaSetDevice(c->slab[0].dev));
    return B200_OK;
}

int b200_free(b200_ctx* c)
{
    if (!c) { set_error("NULL context"); return B200_ERR_ARG; }
    if (!c->allocated) return B200_OK;
    if (int rc = sync_streams(c, true)) return rc;       // asynchronous mode: nothing may still be in flight
    const b200_test_info* ti = b200_get_test_info(c->test);
    const size_t esz = esz_of(c->dtype);
    for (int g = 0; g < c->ngpus; g++) {
        b200_slab& s = c->slab[g];
        B200_CUDA(cudaSetDevice(s.dev));
        for (int q = 0; q < ti->narrays; q++) {
            if (s.arr[q]) {
                drop_tensor_maps_for(s.arr[q], (const char*)s.arr[q] + slab_elems(c, s, q) * esz + 1);
                B200_CUDA(cudaFree(s.arr[q]));
            }
            s.arr[q] = nullptr;
        }
        if (s.scratch) {
            drop_tensor_maps_for(s.scratch, (const char*)s.scratch + slab_elems(c, s, 0) * esz + 1);
            B200_CUDA(cudaFree(s.scratch));
            s.scratch = nullptr;
        }
        B200_CUDA(cudaStreamDestroy(s.stream));
        B200_CUDA(cudaEventDestroy(s.done[0]));
        B200_CUDA(cudaEventDestroy(s.done[1]));
        B200_CUDA(cudaEventDestroy(s.t0));
        B200_CUDA(cudaEventDestroy(s.t1));
        B200_CUDA(cudaEventDestroy(s.fork));
        B200_CUDA(cudaEventDestroy(s.join));
    }
    B200_CUDA(cudaSetDevice(c->slab[0].dev));
    c->allocated = false;
    return B200_OK;
}

int b200_destroy(b200_ctx* c)
{
    if (!c) return B200_OK;
    int rc = b200_free(c);
    free(c);
    return rc;
}

}  // extern "C"
