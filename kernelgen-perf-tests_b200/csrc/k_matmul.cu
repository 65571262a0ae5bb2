// k_matmul.cu -- the suite's one dense contraction: C(i,j) += sum_k A(i,k)*B(k,j), column-major,
// A nx x ny, B ny x ns, C nx x ns, accumulating over the nt sweeps (matmul/matmul.F90:56-68,
// matmul/main.c:232-244).  The only test of the suite that belongs on tensor cores.
//
//   double : hand-written DGEMM on the FP64 tensor cores -- mma.sync.m8n8k4.f64 (DMMA; tcgen05 /
//            UMMA has no FP64 kind).  CTA tile 128x128x16, 16 warps of 32x32, 4-stage cp.async ring.
//   float  : "3xTF32" -- every operand is split into a TF32 head and a TF32 tail, D += At*Bh + Ah*Bt + Ah*Bh, which keeps
//            FP32-level accuracy (the 1e-5 parity bar; a single TF32 pass does not).  Problems TMA can address (16-byte
//            aligned bases, nx and ny multiples of 4) run on the 5th-generation tensor cores: k_matmul_tc05.cu (tcgen05.mma
//            kind::tf32, TMA-fed shared-memory operands, TMEM accumulators).  Everything else takes the kernel in this
//            file: the same three products on mma.sync.m16n8k8.tf32, operands split in registers, CTA tile 128x128x16,
//            8 warps of 64x32.
//
// Shared-memory tiles keep the memory order of the operands (A: m contiguous, B: k contiguous) so
// that they are filled with 16-byte cp.async straight from the column-major arrays; the row
// pitches are padded so that the fragment loads of a warp hit 32 distinct banks.  Extents that are
// not multiples of the 16-byte vector (odd nx / ny) take the same kernel with element-sized
// cp.async.  Diagnostics library only (-DB200_DIAG): B200_MATMUL=cublas selects cuBLAS (dlopen'ed) as a baseline,
// B200_MATMUL_TC05=0 the mma.sync float kernel for aligned problems too.
#include <dlfcn.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "b200_common.cuh"
#include "b200_internal.h"

namespace b200 {

// ------------------------------------------------------------------------------------------
// cp.async helpers
// ------------------------------------------------------------------------------------------
template <int BYTES> B200_DEV void cp_async_zfill(void* smem_dst, const void* gsrc, bool valid)
{
    const uint32_t d = smem_u32(smem_dst);
    const int src = valid ? BYTES : 0;              // src-size 0: the destination is zero-filled
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(d), "l"(gsrc), "n"(BYTES), "r"(src) : "memory");
}
B200_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> B200_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int MM_BM = 128, MM_BN = 128;
#ifndef B200_MM_BK
#define B200_MM_BK 32
#endif
constexpr int MM_BK = B200_MM_BK;                 // k extent of a stage (one __syncthreads per stage)
template <typename T> constexpr int mm_stages() { return MM_BK == 32 ? (sizeof(T) == 8 ? 3 : 4) : 4; }

template <typename T> struct MmGeo {
    // pitches (in elements): A tile [BK][LDA] (m contiguous), B tile [BN][LDB] (k contiguous)
    //   double: fragment loads are 8 bytes, served per half-warp: LDA % 16 == 4, LDB % 16 == 4
    //   float : 4-byte loads, whole warp:                            LDA % 32 == 8, LDB % 32 == 4
    static constexpr int LDA = MM_BM + (sizeof(T) == 8 ? 4 : 8);
    static constexpr int LDB = MM_BK + 4;
    static constexpr int A_ELEMS = MM_BK * LDA, B_ELEMS = MM_BN * LDB;
    static constexpr int STAGE_BYTES = (A_ELEMS + B_ELEMS) * (int)sizeof(T);
    static constexpr int STAGES = mm_stages<T>();
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES;
    static constexpr int THREADS = sizeof(T) == 8 ? 512 : 256;
};

// Fill one stage: A(m0.., k0..) and B(k0.., n0..) with zero fill outside the matrices.
// VE = elements per cp.async (16 bytes when the pitches allow it, else 1).
template <typename T, int VE>
B200_DEV void mm_load_stage(T* As, T* Bs, const T* __restrict__ A, const T* __restrict__ B, int M, int N, int K,
                            int lda, int ldb, int m0, int n0, int k0, int tid)
{
    using G = MmGeo<T>;
    constexpr int BYTES = VE * (int)sizeof(T);
    constexpr int AV = MM_BM / VE;                  // vectors per k-row of the A tile
#pragma unroll
    for (int v = tid; v < MM_BK * AV; v += G::THREADS) {
        const int k = v / AV, mv = (v - k * AV) * VE;
        const bool ok = (k0 + k) < K && (m0 + mv) < M;
        const T* src = ok ? A + (size_t)(k0 + k) * lda + (m0 + mv) : A;
        cp_async_zfill<BYTES>(As + k * G::LDA + mv, src, ok);
    }
    constexpr int BV = MM_BK / VE;                  // vectors per column of the B tile
#pragma unroll
    for (int v = tid; v < MM_BN * BV; v += G::THREADS) {
        const int n = v / BV, kv = (v - n * BV) * VE;
        const bool ok = (n0 + n) < N && (k0 + kv) < K;
        const T* src = ok ? B + (size_t)(n0 + n) * ldb + (k0 + kv) : B;
        cp_async_zfill<BYTES>(Bs + n * G::LDB + kv, src, ok);
    }
}

// Tiles are numbered so that 16 consecutive CTAs share a B panel and 16 A panels stay hot in L2.
B200_DEV void mm_tile_coords(int tile, int mt, int nt, int& tm, int& tn)
{
    constexpr int GROUP = 16;
    const int per_group = GROUP * nt;
    const int g = tile / per_group, r = tile - g * per_group;
    const int rows = min(GROUP, mt - g * GROUP);
    tm = g * GROUP + r % rows;
    tn = r / rows;
}

B200_DEV void dmma_m8n8k4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------
// double: DMMA
// ------------------------------------------------------------------------------------------
template <int VE>
__global__ void __launch_bounds__(512, 1)
matmul_f64_kernel(int M, int N, int K, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                  double* __restrict__ C, int ldc, int mt, int nt)
{
    using G = MmGeo<double>;
    extern __shared__ __align__(16) unsigned char mm_smem[];
    double* smem = reinterpret_cast<double*>(mm_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int nk = (K + MM_BK - 1) / MM_BK;

    for (int tile = blockIdx.x; tile < mt * nt; tile += gridDim.x) {
        int tm, tn;
        mm_tile_coords(tile, mt, nt, tm, tn);
        const int m0 = tm * MM_BM, n0 = tn * MM_BN;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

        __syncthreads();                            // the previous tile's readers are done with the ring
#pragma unroll
        for (int s = 0; s < G::STAGES - 1; s++) {
            if (s < nk) {
                double* st = smem + (size_t)s * (G::A_ELEMS + G::B_ELEMS);
                mm_load_stage<double, VE>(st, st + G::A_ELEMS, A, B, M, N, K, lda, ldb, m0, n0, s * MM_BK, tid);
            }
            cp_async_commit();
        }
        for (int kt = 0; kt < nk; kt++) {
            cp_async_wait<G::STAGES - 2>();
            __syncthreads();
            {
                const int kn = kt + G::STAGES - 1;
                if (kn < nk) {
                    double* st = smem + (size_t)(kn % G::STAGES) * (G::A_ELEMS + G::B_ELEMS);
                    mm_load_stage<double, VE>(st, st + G::A_ELEMS, A, B, M, N, K, lda, ldb, m0, n0, kn * MM_BK, tid);
                }
                cp_async_commit();
            }
            const double* As = smem + (size_t)(kt % G::STAGES) * (G::A_ELEMS + G::B_ELEMS);
            const double* Bs = As + G::A_ELEMS;
#pragma unroll
            for (int ks = 0; ks < MM_BK; ks += 4) {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = As[(ks + t) * G::LDA + wm + i * 8 + g];
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = Bs[(wn + j * 8 + g) * G::LDB + ks + t];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        cp_async_wait<0>();
        // C += acc   (C(i,j) accumulates over the sweeps: matmul.F90:62-66)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int row = m0 + wm + i * 8 + g;
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int col = n0 + wn + j * 8 + 2 * t + e;
                    if (row < M && col < N) {
                        double* p = C + (size_t)col * ldc + row;
                        *p = *p + acc[i][j][e];
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// float: 3xTF32
// ------------------------------------------------------------------------------------------
// x = head + tail with head = x truncated to TF32 (10 mantissa bits) and tail = x - head, exact in
// FP32.  The tensor core ignores the low 13 mantissa bits of a TF32 operand, so the tail is passed
// as it is.  (cvt.rna.tf32 would expand to ~6 instructions per element and make the splits, not
// the MMAs, the bottleneck.)  Dropped: tail*tail (2^-22 relative) and the tail's own truncation.
B200_DEV void split_tf32(float x, uint32_t& head, uint32_t& tail)
{
    head = __float_as_uint(x) & 0xffffe000u;
    tail = __float_as_uint(x - __uint_as_float(head));
}
B200_DEV void mma_tf32_m16n8k8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int VE>
__global__ void __launch_bounds__(256, 1)
matmul_f32_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                  float* __restrict__ C, int ldc, int mt, int nt)
{
    using G = MmGeo<float>;
    extern __shared__ __align__(16) unsigned char mm_smem[];
    float* smem = reinterpret_cast<float*>(mm_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
    const int nk = (K + MM_BK - 1) / MM_BK;

    for (int tile = blockIdx.x; tile < mt * nt; tile += gridDim.x) {
        int tm, tn;
        mm_tile_coords(tile, mt, nt, tm, tn);
        const int m0 = tm * MM_BM, n0 = tn * MM_BN;
        float acc[4][4][4];                         // [m16 tile][n8 tile][fragment]
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[i][j][e] = 0.f;

        __syncthreads();
#pragma unroll
        for (int s = 0; s < G::STAGES - 1; s++) {
            if (s < nk) {
                float* st = smem + (size_t)s * (G::A_ELEMS + G::B_ELEMS);
                mm_load_stage<float, VE>(st, st + G::A_ELEMS, A, B, M, N, K, lda, ldb, m0, n0, s * MM_BK, tid);
            }
            cp_async_commit();
        }
        for (int kt = 0; kt < nk; kt++) {
            cp_async_wait<G::STAGES - 2>();
            __syncthreads();
            {
                const int kn = kt + G::STAGES - 1;
                if (kn < nk) {
                    float* st = smem + (size_t)(kn % G::STAGES) * (G::A_ELEMS + G::B_ELEMS);
                    mm_load_stage<float, VE>(st, st + G::A_ELEMS, A, B, M, N, K, lda, ldb, m0, n0, kn * MM_BK, tid);
                }
                cp_async_commit();
            }
            const float* As = smem + (size_t)(kt % G::STAGES) * (G::A_ELEMS + G::B_ELEMS);
            const float* Bs = As + G::A_ELEMS;
            // The tensor core truncates when it adds into its FP32 accumulator, a bias that grows with the
            // length of the chain (measured: 6e-5 normwise at K = 8192).  So a stage accumulates from zero in
            // `part` (6 MMAs deep) and is then added to `acc` with a correctly rounded FADD.
            float part[4][4][4];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int e = 0; e < 4; e++) part[i][j][e] = 0.f;
#pragma unroll
            for (int ks = 0; ks < MM_BK; ks += 8) {
                uint32_t bh[4][2], bt[4][2];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float* p = Bs + (wn + j * 8 + g) * G::LDB + ks + t;
                    split_tf32(p[0], bh[j][0], bt[j][0]);           // (k = t,   n = g)
                    split_tf32(p[4], bh[j][1], bt[j][1]);           // (k = t+4, n = g)
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t ah[4], at[4];
                    const float* p = As + (ks + t) * G::LDA + wm + i * 16 + g;
                    split_tf32(p[0], ah[0], at[0]);                 // (row g,   col t)
                    split_tf32(p[8], ah[1], at[1]);                 // (row g+8, col t)
                    split_tf32(p[4 * G::LDA], ah[2], at[2]);        // (row g,   col t+4)
                    split_tf32(p[4 * G::LDA + 8], ah[3], at[3]);    // (row g+8, col t+4)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        mma_tf32_m16n8k8(part[i][j], at, bh[j]);    // small terms first
                        mma_tf32_m16n8k8(part[i][j], ah, bt[j]);
                        mma_tf32_m16n8k8(part[i][j], ah, bh[j]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[i][j][e] += part[i][j][e];
        }
        cp_async_wait<0>();
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int row = m0 + wm + i * 16 + g + (e >> 1) * 8;
                    const int col = n0 + wn + j * 8 + 2 * t + (e & 1);
                    if (row < M && col < N) {
                        float* p = C + (size_t)col * ldc + row;
                        *p = *p + acc[i][j][e];
                    }
                }
            }
        }
    }
}

#ifdef B200_DIAG      // libb200stencil_diag.so only: the library GEMM as a baseline (B200_MATMUL=cublas)
// ------------------------------------------------------------------------------------------
// cuBLAS baseline (B200_MATMUL=cublas), resolved with dlopen at first use
// ------------------------------------------------------------------------------------------
typedef void* cublasHandle_t_;
typedef int (*fn_create)(cublasHandle_t_*);
typedef int (*fn_set_stream)(cublasHandle_t_, cudaStream_t);
typedef int (*fn_dgemm)(cublasHandle_t_, int, int, int, int, int, const double*, const double*, int, const double*, int,
                        const double*, double*, int);
typedef int (*fn_sgemm)(cublasHandle_t_, int, int, int, int, int, const float*, const float*, int, const float*, int,
                        const float*, float*, int);

static struct {
    std::mutex mu;
    void* lib = nullptr;
    fn_create create = nullptr;
    fn_set_stream set_stream = nullptr;
    fn_dgemm dgemm = nullptr;
    fn_sgemm sgemm = nullptr;
    cublasHandle_t_ handle[16] = {};
} g_blas;

static int blas_handle(int device, cublasHandle_t_* h)
{
    std::lock_guard<std::mutex> lk(g_blas.mu);
    if (!g_blas.lib) {
        const char* names[] = { "libcublas.so.12", "libcublas.so", "/usr/local/cuda/lib64/libcublas.so.12" };
        for (const char* n : names) {
            g_blas.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_blas.lib) break;
        }
        if (!g_blas.lib) { set_error("matmul: cannot load libcublas (%s)", dlerror()); return B200_ERR_CUDA; }
        g_blas.create = (fn_create)dlsym(g_blas.lib, "cublasCreate_v2");
        g_blas.set_stream = (fn_set_stream)dlsym(g_blas.lib, "cublasSetStream_v2");
        g_blas.dgemm = (fn_dgemm)dlsym(g_blas.lib, "cublasDgemm_v2");
        g_blas.sgemm = (fn_sgemm)dlsym(g_blas.lib, "cublasSgemm_v2");
        if (!g_blas.create || !g_blas.set_stream || !g_blas.dgemm || !g_blas.sgemm) {
            set_error("matmul: libcublas lacks the v2 GEMM entry points");
            return B200_ERR_CUDA;
        }
    }
    if (device < 0 || device >= 16) { set_error("device index %d out of range", device); return B200_ERR_ARG; }
    if (!g_blas.handle[device]) {
        if (g_blas.create(&g_blas.handle[device]) != 0) { set_error("cublasCreate failed"); return B200_ERR_CUDA; }
    }
    *h = g_blas.handle[device];
    return B200_OK;
}

static int launch_cublas(int dtype, const HostArgs& a, int c0, int c1)
{
    const b200_sweep_desc& d = *a.desc;
    cublasHandle_t_ h;
    if (int rc = blas_handle(a.device, &h)) return rc;
    if (g_blas.set_stream(h, a.stream) != 0) { set_error("cublasSetStream failed"); return B200_ERR_CUDA; }
    int st;
    if (dtype == B200_F64) {
        const double one = 1.0;
        st = g_blas.dgemm(h, 0, 0, d.nx, c1 - c0, d.ny, &one, (const double*)a.arrays[0], d.nx,
                          (const double*)a.arrays[1] + (size_t)c0 * d.ny, d.ny, &one,
                          (double*)a.arrays[2] + (size_t)c0 * d.nx, d.nx);
    } else {
        const float one = 1.0f;
        st = g_blas.sgemm(h, 0, 0, d.nx, c1 - c0, d.ny, &one, (const float*)a.arrays[0], d.nx,
                          (const float*)a.arrays[1] + (size_t)c0 * d.ny, d.ny, &one,
                          (float*)a.arrays[2] + (size_t)c0 * d.nx, d.nx);
    }
    if (st != 0) { set_error("cuBLAS GEMM failed with status %d", st); return B200_ERR_CUDA; }
    count_launch();
    return B200_OK;
}

#endif  // B200_DIAG

// ------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------
template <typename T, typename K> static int mm_prepare(K kernel)
{
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MmGeo<T>::SMEM_BYTES));
    return B200_OK;
}

template <typename T> static int launch_tc(const HostArgs& a, int c0, int c1)
{
    using G = MmGeo<T>;
    constexpr int V = 16 / (int)sizeof(T);
    const b200_sweep_desc& d = *a.desc;
    const int M = d.nx, K = d.ny, N = c1 - c0;
    const T* A = (const T*)a.arrays[0];
    const T* B = (const T*)a.arrays[1] + (size_t)c0 * d.ny;
    T* C = (T*)a.arrays[2] + (size_t)c0 * d.nx;
    const bool vec = (M % V == 0) && (K % V == 0) && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0);
    const int mt = (M + MM_BM - 1) / MM_BM, nt = (N + MM_BN - 1) / MM_BN;
    const long long tiles = (long long)mt * nt;
    const int grid = (int)(tiles < a.num_sms ? tiles : a.num_sms);
    if constexpr (sizeof(T) == 8) {
        auto kern = vec ? matmul_f64_kernel<2> : matmul_f64_kernel<1>;
        if (int rc = mm_prepare<T>(kern)) return rc;      // per device, cheap
        kern<<<grid, G::THREADS, G::SMEM_BYTES, a.stream>>>(M, N, K, A, M, B, K, C, M, mt, nt);
    } else {
        auto kern = vec ? matmul_f32_kernel<4> : matmul_f32_kernel<1>;
        if (int rc = mm_prepare<T>(kern)) return rc;      // per device, cheap
        kern<<<grid, G::THREADS, G::SMEM_BYTES, a.stream>>>(M, N, K, A, M, B, K, C, M, mt, nt);
    }
    B200_CUDA(cudaGetLastError());
    count_launch();
    return B200_OK;
}

// k_matmul_tc05.cu: the float GEMM on tcgen05 (TMA-fed shared-memory operands, TMEM accumulator)
bool matmul_tc05_eligible(const float* A, const float* B, const float* C, int M, int N, int K);
int launch_matmul_tc05(const float* A, const float* B, float* C, int M, int N, int K, int num_sms, cudaStream_t stream);
int info_matmul_tc05(KernelInfo* ki);

int launch_matmul(int dtype, const HostArgs& a)
{
    const b200_sweep_desc& d = *a.desc;
    int c0 = 0, c1 = d.ns;                       // columns of B and C handled by this launch
    if (d.out_begin != 0 || d.out_end != 0) {
        if (d.out_begin < 0 || d.out_end > d.ns || d.out_begin > d.out_end) { set_error("bad column range"); return B200_ERR_ARG; }
        c0 = d.out_begin; c1 = d.out_end;
    }
    if (c1 <= c0 || d.nx <= 0 || d.ny <= 0) return B200_OK;
#ifdef B200_DIAG
    const char* mode = getenv("B200_MATMUL");
    if (mode && !strcmp(mode, "cublas")) return launch_cublas(dtype, a, c0, c1);
#endif
    if (dtype == B200_F32) {
        const float* A = (const float*)a.arrays[0];
        const float* B = (const float*)a.arrays[1] + (size_t)c0 * d.ny;
        float* C = (float*)a.arrays[2] + (size_t)c0 * d.nx;
        bool tc05 = matmul_tc05_eligible(A, B, C, d.nx, c1 - c0, d.ny);
#ifdef B200_DIAG
        if (const char* e = getenv("B200_MATMUL_TC05")) tc05 = tc05 && atoi(e) != 0;     // 0: the mma.sync kernel (A/B runs)
#endif
        if (tc05) return launch_matmul_tc05(A, B, C, d.nx, c1 - c0, d.ny, a.num_sms, a.stream);
    }
    return dtype == B200_F32 ? launch_tc<float>(a, c0, c1) : launch_tc<double>(a, c0, c1);
}

int info_matmul(int dtype, KernelInfo* ki)
{
    if (dtype == B200_F32) return info_matmul_tc05(ki);          // the kernel aligned float problems run (others: matmul_f32_kernel)
    cudaFuncAttributes fa;
    if (dtype == B200_F32) B200_CUDA(cudaFuncGetAttributes(&fa, matmul_f32_kernel<4>));
    else B200_CUDA(cudaFuncGetAttributes(&fa, matmul_f64_kernel<2>));
    ki->regs = fa.numRegs;
    ki->smem_bytes = dtype == B200_F32 ? MmGeo<float>::SMEM_BYTES : MmGeo<double>::SMEM_BYTES;
    ki->blocks_per_sm = 1;
    ki->name = "matmul";
    return B200_OK;
}

}  // namespace b200
