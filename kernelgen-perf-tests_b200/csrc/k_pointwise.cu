// k_pointwise.cu -- the pure-bandwidth kernels of the suite: vecadd, sincos, matvec.
//   vecadd  w2 = w0 + w1 over all points                      vecadd/vecadd.c:66-83
//   sincos  xy = sin(x) + cos(y) over all points               sincos/sincos.F90:60-72
//   matvec  y[j] = sum_i A[i + nx*j] * x[i]                    matvec/matvec.c:59-68
// No data reuse, so no shared-memory staging: 16-byte coalesced streaming loads/stores,
// four independent vectors in flight per thread, persistent grid of SMs x 8 CTAs.
#include <math.h>

#include "b200_common.cuh"
#include "b200_internal.h"

namespace b200 {

constexpr int PW_THREADS = 256;
constexpr int PW_UNROLL = 4;

template <typename T> struct AddF    { B200_DEV static T f(T a, T b) { return a + b; } };
// sin / cos on |x| <= 1 -- the range of the suite's inputs, real_rand() in [-1, 1] (sincos/main.c:102-104) -- by
// their Taylor polynomials in Horner form (truncation error below 1e-17 / 3e-9, i.e. under half an ulp; the
// result is within 1-2 ulp of libm's), without the range reduction and quadrant selection of the general
// routine: 22 FP64 instructions per point instead of ~45, which is what makes the double test a bandwidth
// kernel (0.61 -> HBM-bound).  Any other argument takes the libdevice routine.
B200_DEV double sin_unit(double x)
{
    const double z = x * x;
    double p = -1.0 / 355687428096000.0;               // -1/17!
    p = fma(p, z, 1.0 / 1307674368000.0);              //  1/15!
    p = fma(p, z, -1.0 / 6227020800.0);                // -1/13!
    p = fma(p, z, 1.0 / 39916800.0);                   //  1/11!
    p = fma(p, z, -1.0 / 362880.0);                    // -1/9!
    p = fma(p, z, 1.0 / 5040.0);                       //  1/7!
    p = fma(p, z, -1.0 / 120.0);                       // -1/5!
    p = fma(p, z, 1.0 / 6.0);                          //  1/3!  (sign folded below)
    p = fma(p, -z, 1.0);
    // p = 1 - z/3! + z^2/5! - ...   (alternating signs: the chain above carries them from the top term down)
    return x * p;
}
B200_DEV double cos_unit(double x)
{
    const double z = x * x;
    double p = 1.0 / 6402373705728000.0;               //  1/18!
    p = fma(p, z, -1.0 / 20922789888000.0);            // -1/16!
    p = fma(p, z, 1.0 / 87178291200.0);                //  1/14!
    p = fma(p, z, -1.0 / 479001600.0);                 // -1/12!
    p = fma(p, z, 1.0 / 3628800.0);                    //  1/10!
    p = fma(p, z, -1.0 / 40320.0);                     // -1/8!
    p = fma(p, z, 1.0 / 720.0);                        //  1/6!
    p = fma(p, z, -1.0 / 24.0);                        // -1/4!
    p = fma(p, z, 0.5);                                //  1/2!
    return fma(p, -z, 1.0);
}
B200_DEV float sin_unit(float x)
{
    const float z = x * x;
    float p = -1.0f / 39916800.0f;                     // -1/11!
    p = fmaf(p, z, 1.0f / 362880.0f);
    p = fmaf(p, z, -1.0f / 5040.0f);
    p = fmaf(p, z, 1.0f / 120.0f);
    p = fmaf(p, z, -1.0f / 6.0f);
    return fmaf(x * z, p, x);
}
B200_DEV float cos_unit(float x)
{
    const float z = x * x;
    float p = 1.0f / 479001600.0f;                     //  1/12!
    p = fmaf(p, z, -1.0f / 3628800.0f);
    p = fmaf(p, z, 1.0f / 40320.0f);
    p = fmaf(p, z, -1.0f / 720.0f);
    p = fmaf(p, z, 1.0f / 24.0f);
    p = fmaf(p, z, -0.5f);
    return fmaf(p, z, 1.0f);
}
template <typename T> struct SinCosF;
template <> struct SinCosF<float> {
    B200_DEV static float f(float a, float b)
    {
        const float s = fabsf(a) <= 1.0f ? sin_unit(a) : sinf(a);
        const float c = fabsf(b) <= 1.0f ? cos_unit(b) : cosf(b);
        return s + c;
    }
};
template <> struct SinCosF<double> {
    B200_DEV static double f(double a, double b)
    {
        const double s = fabs(a) <= 1.0 ? sin_unit(a) : sin(a);
        const double c = fabs(b) <= 1.0 ? cos_unit(b) : cos(b);
        return s + c;
    }
};

template <typename T, template <typename> class F>
__global__ void __launch_bounds__(PW_THREADS)
binary_stream_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ c, size_t n, int vec_ok)
{
    constexpr int V = 16 / sizeof(T);
    const size_t tid = (size_t)blockIdx.x * PW_THREADS + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * PW_THREADS;
    if (vec_ok) {
        const size_t nvec = n / V;
        size_t i = tid;
        for (; i + (PW_UNROLL - 1) * nthreads < nvec; i += PW_UNROLL * nthreads) {
            VReg<T> va[PW_UNROLL], vb[PW_UNROLL];
#pragma unroll
            for (int u = 0; u < PW_UNROLL; u++) {
                va[u] = ldv_stream(a + (i + u * nthreads) * V);
                vb[u] = ldv_stream(b + (i + u * nthreads) * V);
            }
#pragma unroll
            for (int u = 0; u < PW_UNROLL; u++) {
                VReg<T> r;
#pragma unroll
                for (int v = 0; v < V; v++) r[v] = F<T>::f(va[u][v], vb[u][v]);
                *reinterpret_cast<uint4*>(c + (i + u * nthreads) * V) = *reinterpret_cast<const uint4*>(r.v);
            }
        }
        for (; i < nvec; i += nthreads) {
            const VReg<T> va = ldv_stream(a + i * V), vb = ldv_stream(b + i * V);
            VReg<T> r;
#pragma unroll
            for (int v = 0; v < V; v++) r[v] = F<T>::f(va[v], vb[v]);
            *reinterpret_cast<uint4*>(c + i * V) = *reinterpret_cast<const uint4*>(r.v);
        }
        for (size_t j = nvec * V + tid; j < n; j += nthreads) c[j] = F<T>::f(a[j], b[j]);
    } else {
        for (size_t j = tid; j < n; j += nthreads) c[j] = F<T>::f(a[j], b[j]);
    }
}

// One warp per row; lanes stride over 16-byte vectors of the row, then a shuffle tree.
// (The reference sums each row sequentially; the tree changes rounding only.)
template <typename T>
__global__ void __launch_bounds__(PW_THREADS)
matvec_kernel(const T* __restrict__ A, const T* __restrict__ x, T* __restrict__ y, int nx, int row0, int row1, int vec_ok)
{
    constexpr int V = 16 / sizeof(T);
    const int lane = threadIdx.x & 31;
    const int warps_per_block = PW_THREADS / 32;
    const int warp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_block;
    for (int j = row0 + warp; j < row1; j += nwarps) {
        const T* row = A + (size_t)j * nx;
        T s0 = 0, s1 = 0;
        if (vec_ok) {
            int i = lane * V;
            for (; i + 32 * V < nx; i += 64 * V) {
                const VReg<T> a0 = ldv_stream(row + i), a1 = ldv_stream(row + i + 32 * V);
                const VReg<T> x0 = ldv(x + i), x1 = ldv(x + i + 32 * V);
#pragma unroll
                for (int v = 0; v < V; v++) { s0 += a0[v] * x0[v]; s1 += a1[v] * x1[v]; }
            }
            for (; i < nx; i += 32 * V) {
                const VReg<T> a0 = ldv_stream(row + i);
                const VReg<T> x0 = ldv(x + i);
#pragma unroll
                for (int v = 0; v < V; v++) s0 += a0[v] * x0[v];
            }
        } else {
            for (int i = lane; i < nx; i += 32) s0 += row[i] * x[i];
        }
        T s = s0 + s1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[j] = s;
    }
}

template <typename T, template <typename> class F>
static int launch_binary(const HostArgs& a, int sa, int sb, int sc)
{
    const b200_sweep_desc& d = *a.desc;
    const size_t plane = (size_t)d.nx * d.ny;
    size_t off = 0, n = plane * d.ns;
    if (d.out_begin != 0 || d.out_end != 0) {
        if (d.out_begin < 0 || d.out_end > d.ns || d.out_begin > d.out_end) { set_error("bad plane range"); return B200_ERR_ARG; }
        off = plane * d.out_begin;
        n = plane * (size_t)(d.out_end - d.out_begin);
    }
    if (n == 0) return B200_OK;
    const T* pa = (const T*)a.arrays[sa] + off;
    const T* pb = (const T*)a.arrays[sb] + off;
    T* pc = (T*)a.arrays[sc] + off;
    const int vec_ok = !(((uintptr_t)pa | (uintptr_t)pb | (uintptr_t)pc) & 15);
    const size_t want = (n / (16 / sizeof(T)) + PW_THREADS * PW_UNROLL - 1) / (PW_THREADS * PW_UNROLL);
    const int cap = a.num_sms * 8;
    const int grid = (int)(want < 1 ? 1 : (want < (size_t)cap ? want : (size_t)cap));
    binary_stream_kernel<T, F><<<grid, PW_THREADS, 0, a.stream>>>(pa, pb, pc, n, vec_ok);
    B200_CUDA(cudaGetLastError());
    count_launch();
    return B200_OK;
}

int launch_vecadd(int dtype, const HostArgs& a)
{
    return dtype == B200_F32 ? launch_binary<float, AddF>(a, 0, 1, 2) : launch_binary<double, AddF>(a, 0, 1, 2);
}
int launch_sincos(int dtype, const HostArgs& a)
{
    return dtype == B200_F32 ? launch_binary<float, SinCosF>(a, 0, 1, 2) : launch_binary<double, SinCosF>(a, 0, 1, 2);
}

template <typename T> static int launch_matvec_t(const HostArgs& a)
{
    const b200_sweep_desc& d = *a.desc;
    int r0 = 0, r1 = d.ny;
    if (d.out_begin != 0 || d.out_end != 0) {
        if (d.out_begin < 0 || d.out_end > d.ny || d.out_begin > d.out_end) { set_error("bad row range"); return B200_ERR_ARG; }
        r0 = d.out_begin; r1 = d.out_end;
    }
    if (r1 <= r0 || d.nx <= 0) return B200_OK;
    const T* A = (const T*)a.arrays[0];
    const T* x = (const T*)a.arrays[1];
    T* y = (T*)a.arrays[2];
    const int vec_ok = !(((uintptr_t)A | (uintptr_t)x) & 15) && ((size_t)d.nx * sizeof(T)) % 16 == 0;
    const int rows = r1 - r0;
    const int want = (rows + PW_THREADS / 32 - 1) / (PW_THREADS / 32);
    const int cap = a.num_sms * 8;
    matvec_kernel<T><<<want < cap ? want : cap, PW_THREADS, 0, a.stream>>>(A, x, y, d.nx, r0, r1, vec_ok);
    B200_CUDA(cudaGetLastError());
    count_launch();
    return B200_OK;
}
int launch_matvec(int dtype, const HostArgs& a)
{
    return dtype == B200_F32 ? launch_matvec_t<float>(a) : launch_matvec_t<double>(a);
}

template <typename K> static int info_of(K kernel, KernelInfo* ki, const char* name)
{
    cudaFuncAttributes fa;
    B200_CUDA(cudaFuncGetAttributes(&fa, kernel));
    ki->regs = fa.numRegs;
    ki->smem_bytes = 0;
    ki->blocks_per_sm = 8;
    ki->name = name;
    return B200_OK;
}
int info_vecadd(int dtype, KernelInfo* ki)
{
    return dtype == B200_F32 ? info_of(binary_stream_kernel<float, AddF>, ki, "vecadd")
                             : info_of(binary_stream_kernel<double, AddF>, ki, "vecadd");
}
int info_sincos(int dtype, KernelInfo* ki)
{
    return dtype == B200_F32 ? info_of(binary_stream_kernel<float, SinCosF>, ki, "sincos")
                             : info_of(binary_stream_kernel<double, SinCosF>, ki, "sincos");
}
int info_matvec(int dtype, KernelInfo* ki)
{
    return dtype == B200_F32 ? info_of(matvec_kernel<float>, ki, "matvec")
                             : info_of(matvec_kernel<double>, ki, "matvec");
}

}  // namespace b200
