// tools/probes/fp64_probe.cu -- what the FP64 pipe and the shared-memory pipe of one B200 SM really sustain
// (the two ceilings of the tricubic kernels).  Standalone: nvcc -arch=sm_100a -O3 fp64_probe.cu
//   mode 0: independent DFMA chains (ILP 8) -> DFMA per clk per SM, for 4..16 warps per SM
//   mode 1: LDS.128 conflict-free -> wavefronts (128 B) per clk per SM
//   mode 2: the tricubic mix: 7 FP64 per LDS.128, interleaved in one instruction stream
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(double* out, long long* cycles, int iters)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 1.0 + threadIdx.x * 1e-6 + k;
    const double m = 1.0000001, c = 1e-9;
    const double2* p = reinterpret_cast<const double2*>(sm) + threadIdx.x % 256;
    double2 acc = {0.0, 0.0};
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int k = 0; k < 8; k++) a[k] = a[k] * m + c;
        } else if (MODE == 1) {
#pragma unroll
            for (int r = 0; r < 16; r++) {
                double2 v = p[(r * 256 + it) & 1023];
                acc.x += v.x; acc.y += v.y;           // 2 DADD per LDS.128 (kept small relative to the loads)
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                double2 v = p[(r * 256 + it) & 1023];
#pragma unroll
                for (int k = 0; k < 7; k++) a[k] = a[k] * v.x + v.y;
            }
        }
    }
    const long long t1 = clock64();
    double s = acc.x + acc.y;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int warps, double ops_per_thread_iter, const char* unit)
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int threads = warps * 32, iters = 4000;
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * nsm * threads);
    cudaMalloc(&cyc, sizeof(long long) * nsm);
    probe<MODE><<<nsm, threads, 4096 * 8>>>(out, cyc, 10);
    probe<MODE><<<nsm, threads, 4096 * 8>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nsm; i++) avg += (double)h[i];
    avg /= nsm;
    printf("%-28s warps/SM %2d : %7.2f %s per clk per SM\n", name, warps, ops_per_thread_iter * iters * threads / avg, unit);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w : {4, 8, 12, 16, 24, 32}) run<0>("DFMA, 8 chains/thread", w, 64.0, "DFMA (lanes)");
    for (int w : {4, 8, 12, 16}) run<1>("LDS.128 conflict-free", w, 16.0 * 16 / 128, "wavefronts(128B)");
    for (int w : {8, 12, 16}) run<2>("7 DFMA : 1 LDS.128", w, 56.0, "DFMA (lanes)");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
