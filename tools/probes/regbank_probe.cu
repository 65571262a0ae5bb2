// tools/probes/regbank_probe.cu -- does the FP64 pipe of a B200 SM sustain 1 DFMA per 2 clk per SMSP when all three
// source operands are distinct registers?  (tricubic's DFMAs are weight x window + accumulator: three live registers.)
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(double* out, long long* cycles, int iters, const double* in)
{
    double a[8], b[8], c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = in[threadIdx.x + k]; b[k] = in[threadIdx.x + 8 + k]; c[k] = in[threadIdx.x + 16 + k]; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int g = 0; g < 8; g++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (MODE == 0) a[k] = fma(a[k], b[k], c[k]);            // 3 distinct register pairs per instruction
                if (MODE == 1) a[k] = fma(a[k], b[0], c[0]);            // two operands shared by consecutive instructions
                if (MODE == 2) a[k] = fma(a[k], b[k], c[0]);
                if (MODE == 3) a[k] = fma(b[k], c[(k + g) & 7], a[k]);  // accumulate form: acc += w * u, rotating pairs
                if (MODE == 4) a[k] = a[k] * b[k];                      // DMUL, 2 operands
                if (MODE == 5) a[k] = a[k] + b[k];                      // DADD
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k] + b[k] + c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name)
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int warps = 12, threads = warps * 32, iters = 2000;
    double *out, *in; long long* cyc;
    cudaMalloc(&out, sizeof(double) * nsm * threads);
    cudaMalloc(&in, sizeof(double) * (threads + 32));
    cudaMemset(in, 0, sizeof(double) * (threads + 32));
    cudaMalloc(&cyc, sizeof(long long) * nsm);
    probe<MODE><<<nsm, threads>>>(out, cyc, 10, in);
    probe<MODE><<<nsm, threads>>>(out, cyc, iters, in);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nsm; i++) avg += (double)h[i];
    avg /= nsm;
    printf("%-52s : %5.2f clk per FP64 warp instruction per SMSP (2.00 = pipe peak)\n", name, avg / (iters * 64.0 * 3.0));
    cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main()
{
    run<0>("DFMA a = a*b[k] + c[k]   (3 distinct registers)");
    run<1>("DFMA a = a*b0 + c0       (2 shared)");
    run<2>("DFMA a = a*b[k] + c0");
    run<3>("DFMA a = b[k]*c[(k+g)&7] + a");
    run<4>("DMUL a = a*b[k]");
    run<5>("DADD a = a+b[k]");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
