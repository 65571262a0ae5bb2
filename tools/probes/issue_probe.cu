// tools/probes/issue_probe.cu -- how much FP64 issue time the OTHER instructions of a DFMA-bound loop cost on a B200 SM.
// Each variant runs G groups of: 8 independent DFMA + `extra` instructions of one kind; reports clk per group per SMSP
// (8 DFMA alone = 16 clk if the pipe takes one warp instruction every 2 clk).  12 warps per SM (3 per SMSP).
#include <cstdio>
#include <cuda_runtime.h>

enum { K_NONE, K_IMAD, K_FFMA, K_LDS32, K_LDS64, K_LDS128, K_ISETP, K_LDS128_C2, K_STG128, K_SHFL };

template <int KIND, int EXTRA>
__global__ void probe(double* out, long long* cycles, int iters, double* gbuf)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 1.0 + threadIdx.x * 1e-6 + k;
    const double m = 1.0000001, c = 1e-9;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 255) * 16;
    unsigned base2 = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 255) * 32;     // 2-way conflicting stride
    int x0 = threadIdx.x, x1 = 3, x2 = 5;
    float f0 = 1.f, f1 = 1.0001f, f2 = 0.5f;
    double* gp = gbuf + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int g = 0; g < 8; g++) {
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = a[k] * m + c;
#pragma unroll
            for (int e = 0; e < EXTRA; e++) {
                if (KIND == K_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x0) : "r"(x1), "r"(x2));
                if (KIND == K_ISETP) asm volatile("{ .reg .pred p; setp.lt.s32 p, %0, %1; }" :: "r"(x0), "r"(x1));
                if (KIND == K_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f0) : "f"(f1), "f"(f2));
                if (KIND == K_LDS32) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (g * 8 + e) * 4096 % 32768)); }
                if (KIND == K_LDS64) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + (g * 8 + e) * 4096 % 32768)); }
                if (KIND == K_LDS128) { double v, w; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v), "=d"(w) : "r"(base + (g * 8 + e) * 4096 % 32768)); }
                if (KIND == K_LDS128_C2) { double v, w; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v), "=d"(w) : "r"(base2 + (g * 8 + e) * 8192 % 32768)); }
                if (KIND == K_STG128) asm volatile("st.global.v2.f64 [%0], {%1,%2};" :: "l"(gp), "d"(a[0]), "d"(a[1]) : "memory");
                if (KIND == K_SHFL) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(x0));
            }
        }
    }
    const long long t1 = clock64();
    double s = x0 + f0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND, int EXTRA> void run(const char* name)
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int warps = 12, threads = warps * 32, iters = 2000;
    double *out, *gbuf; long long* cyc;
    cudaMalloc(&out, sizeof(double) * nsm * threads);
    cudaMalloc(&gbuf, sizeof(double) * nsm * threads * 2);
    cudaMalloc(&cyc, sizeof(long long) * nsm);
    cudaFuncSetAttribute(probe<KIND, EXTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    probe<KIND, EXTRA><<<nsm, threads, 65536>>>(out, cyc, 10, gbuf);
    probe<KIND, EXTRA><<<nsm, threads, 65536>>>(out, cyc, iters, gbuf);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nsm; i++) avg += (double)h[i];
    avg /= nsm;
    // per SMSP: 3 warps, each iters*8 groups
    const double clk_per_group = avg / (iters * 8.0 * 3.0);
    printf("%-34s extra %d : %6.2f clk per (8 DFMA + extra) per SMSP  -> %5.2f clk per extra instruction\n", name, EXTRA,
           clk_per_group, EXTRA ? (clk_per_group - 16.0) / EXTRA : 0.0);
    cudaFree(out); cudaFree(cyc); cudaFree(gbuf);
}

int main()
{
    run<K_NONE, 0>("8 DFMA");
    run<K_IMAD, 2>("+ IMAD"); run<K_IMAD, 4>("+ IMAD"); run<K_IMAD, 8>("+ IMAD");
    run<K_ISETP, 4>("+ ISETP"); run<K_ISETP, 8>("+ ISETP");
    run<K_FFMA, 4>("+ FFMA"); run<K_FFMA, 8>("+ FFMA");
    run<K_LDS32, 1>("+ LDS.32"); run<K_LDS32, 4>("+ LDS.32");
    run<K_LDS64, 1>("+ LDS.64"); run<K_LDS64, 4>("+ LDS.64");
    run<K_LDS128, 1>("+ LDS.128"); run<K_LDS128, 2>("+ LDS.128"); run<K_LDS128, 4>("+ LDS.128");
    run<K_LDS128_C2, 1>("+ LDS.128 stride 32 B (2-way)"); run<K_LDS128_C2, 2>("+ LDS.128 stride 32 B (2-way)");
    run<K_STG128, 1>("+ STG.128"); run<K_STG128, 2>("+ STG.128");
    run<K_SHFL, 2>("+ SHFL"); run<K_SHFL, 4>("+ SHFL");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
