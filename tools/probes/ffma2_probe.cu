// tools/probes/ffma2_probe.cu -- FP32 FMA throughput of a B200 SM sub-partition: scalar FFMA vs packed FFMA2
// (fma.rn.f32x2), with three distinct register operands and with shared operands (operand-reuse cache).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void probe(float* out, long long* cycles, int iters, const float* in)
{
    float a[8], b[8], c[8];
    u64 A[8], B[8], C[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        a[k] = in[threadIdx.x + k]; b[k] = in[threadIdx.x + 8 + k]; c[k] = in[threadIdx.x + 16 + k];
        A[k] = pack(a[k], b[k]); B[k] = pack(b[k], c[k]); C[k] = pack(c[k], a[k]);
    }
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int g = 0; g < 8; g++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (MODE == 0) a[k] = fmaf(a[k], b[k], c[k]);
                if (MODE == 1) a[k] = fmaf(a[k], b[0], c[0]);
                if (MODE == 2) A[k] = fma2(A[k], B[k], C[k]);
                if (MODE == 3) A[k] = fma2(A[k], B[0], C[0]);
                if (MODE == 4) A[k] = fma2(B[k], C[(k + g) & 7], A[k]);
                if (MODE == 5) A[k] = fma2(B[k], C[0], A[k]);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k] + b[k] + c[k] + (float)(A[k] & 0xffff) + (float)(A[k] >> 48);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name)
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int warps = 12, threads = warps * 32, iters = 2000;
    float *out, *in; long long* cyc;
    cudaMalloc(&out, sizeof(float) * nsm * threads);
    cudaMalloc(&in, sizeof(float) * (threads + 32));
    cudaMemset(in, 0, sizeof(float) * (threads + 32));
    cudaMalloc(&cyc, sizeof(long long) * nsm);
    probe<MODE><<<nsm, threads>>>(out, cyc, 10, in);
    probe<MODE><<<nsm, threads>>>(out, cyc, iters, in);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nsm; i++) avg += (double)h[i];
    avg /= nsm;
    printf("%-52s : %5.2f clk per warp instruction per SMSP\n", name, avg / (iters * 64.0 * 3.0));
    cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main()
{
    run<0>("FFMA  a = a*b[k] + c[k]  (3 distinct registers)");
    run<1>("FFMA  a = a*b0 + c0      (2 shared)");
    run<2>("FFMA2 A = A*B[k] + C[k]  (3 distinct pairs)");
    run<3>("FFMA2 A = A*B0 + C0      (2 shared)");
    run<4>("FFMA2 A = B[k]*C[(k+g)&7] + A");
    run<5>("FFMA2 A = B[k]*C0 + A    (1 shared)");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
