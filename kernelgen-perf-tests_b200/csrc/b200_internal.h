// b200_internal.h -- host-side internals shared by the translation units of libb200stencil.so.
#pragma once

#include <cuda_runtime.h>

#include "../../include/b200_stencil.h"

namespace b200 {

// Everything a launcher needs for one sweep.
struct HostArgs {
    const b200_sweep_desc* desc;
    void* const* arrays;
    cudaStream_t stream;
    int device;
    int num_sms;
};

struct KernelInfo {
    int regs;
    int smem_bytes;
    int blocks_per_sm;
    const char* name;
};

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch();

#define B200_CUDA(call)                                                          \
    do {                                                                         \
        cudaError_t e__ = (call);                                                \
        if (e__ != cudaSuccess) return ::b200::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// Streaming-engine stencils: one launcher + one info query per test (defined in k_<test>.cu).
#define B200_DECLARE_OP(name)                              \
    int launch_##name(int dtype, const HostArgs& a);       \
    int info_##name(int dtype, KernelInfo* ki);

B200_DECLARE_OP(laplacian)
B200_DECLARE_OP(wave13pt)
B200_DECLARE_OP(divergence)
B200_DECLARE_OP(gradient)
B200_DECLARE_OP(uxx1)
B200_DECLARE_OP(lapgsrb)
B200_DECLARE_OP(jacobi)
B200_DECLARE_OP(gaussblur)
B200_DECLARE_OP(gameoflife)
B200_DECLARE_OP(tricubic)
// two sweeps in one pass (temporal blocking, b200_ops2d.cuh: Fused2D); arrays = { w0, w1, out }
B200_DECLARE_OP(jacobi2)
B200_DECLARE_OP(gaussblur2)
B200_DECLARE_OP(gameoflife2)
// bandwidth kernels (k_pointwise.cu)
B200_DECLARE_OP(vecadd)
B200_DECLARE_OP(matvec)
B200_DECLARE_OP(sincos)
// the one dense contraction (k_matmul.cu)
B200_DECLARE_OP(matmul)

// Tensor-map (TMA descriptor) creation with a small cache; returns 0 on success.
struct TmaBoxKey {
    const void* ptr;
    int nx, ny, ns, esz, bw, bh;   // tensor extents (elements), element size, box
    int px, py;                    // pitches: elements per row, rows per plane (0 = nx, ny)
};
int get_tensor_map(const TmaBoxKey& key, void* out_map /* CUtensorMap* */);

// Per-device CTA completion counter used by PUSH launches to elect the signalling CTA.
int get_done_counter(int device, unsigned int** counter);

}  // namespace b200
