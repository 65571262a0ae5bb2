// k_matmul_tc05.cu -- the float matmul of the suite on the 5th-generation tensor cores (tcgen05 / UMMA), hand-written.
//
//   C(i,j) += sum_k A(i,k) * B(k,j)      column-major, A nx x ny, B ny x ns, C nx x ns      matmul/matmul.F90:56-68
//
// FP32 accuracy from TF32 tensor-core math ("3xTF32"): every operand is split once per sweep into a TF32 head and a TF32
// tail (a = ah + at exactly, both representable in TF32 up to the tail's last bits), and
//     D += At*Bh + Ah*Bt + Ah*Bh
// is issued as three tcgen05.mma.kind::tf32 per k-step of 8 (the At*Bt term is 2^-22 relative and dropped).
//
// Structure: one CTA per 128 x 128 tile of C, 6 warps with fixed roles
//   warp 0   TMA producer: per k-block of 32, cp.async.bulk.tensor.2d of the four operand tiles (Ah, At: 4 boxes of
//            32 m x 32 k each, M-contiguous "MN-major", 128-byte swizzle with 32-byte atoms; Bh, Bt: one box of 32 k x 128 n, "K-major", 128-byte swizzle),
//            into a 3-stage shared-memory ring, completion on the stage's `full` mbarrier;
//   warp 1   MMA issuer: allocates 128 TMEM columns (the 128 x 128 FP32 accumulator), waits for `full`, issues
//            4 k-steps x 3 tcgen05.mma (M = 128, N = 128, K = 8) from shared-memory descriptors, and hands the stage back
//            with tcgen05.commit -> `empty` mbarrier; after the last k-block tcgen05.commit -> `acc_full`;
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) of the accumulator, C += acc with coalesced
//            128-byte stores (lane = row m, C is m-contiguous).
// Out-of-range parts of edge tiles are zero-filled by the TMA unit (they add 0) and masked in the epilogue.
// TMA needs 16-byte aligned bases and row pitches (nx, ny multiples of 4): other shapes take the mma.sync kernel of
// k_matmul.cu.  SASS: UTCMMA / UTMALDG / LDTM / SYNCS.
#include <cuda.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "b200_common.cuh"
#include "b200_internal.h"

namespace b200 {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32;           // BK * 4 bytes = 128 = one swizzle row
constexpr int TC_STAGES = 3;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;                  // 16 KB: 4 boxes of 32 m x 32 k
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;                  // 16 KB: 128 rows (n) of 32 k
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;
constexpr int TC_SMEM_BYTES = 1024 /*alignment*/ + TC_STAGES * TC_STAGE_BYTES + 256 /*barriers, tmem address*/;
constexpr int TC_THREADS = 192;
constexpr int TC_TMEM_COLS = 256;                             // two 128-column accumulators (double buffered K-chunks)
// The tensor core truncates when it adds into its FP32 accumulator: a bias that grows with the length of the chain
// (measured on this kernel: 4e-7 normwise at K = 32, 2.6e-6 at 256, 8.4e-6 at 1024, 6.2e-5 at 8192).  So K is cut into
// chunks of TC_KCHUNK: each chunk is accumulated from zero in one TMEM buffer while the epilogue warps add the previous
// chunk (other buffer) into FP32 running sums in registers with correctly rounded FADDs.
constexpr int TC_KCHUNK_BLOCKS = 4;                           // k-blocks of 32 per chunk: 128

struct alignas(64) TcMaps { CUtensorMap ah, at, bh, bt; };

// ---- PTX wrappers -----------------------------------------------------------------------------
B200_DEV void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
B200_DEV void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
B200_DEV void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
B200_DEV void tc_commit(uint64_t* bar)      // arrives on the mbarrier when every tcgen05.mma issued so far has completed
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], M x N x 8 TF32, issued by one thread for the whole CTA
B200_DEV void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout): start address, leading / stride byte offsets in
// 16-byte units, version 1 (Blackwell), 128-byte swizzle
B200_DEV uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2u)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                          // version_
    d |= (uint64_t)layout_type << 61;                // layout_type_: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor layout): D = F32, A = B = TF32, A MN-major, B K-major
constexpr uint32_t tc_idesc(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- operand split ----------------------------------------------------------------------------
// hi = a with the 13 low mantissa bits cleared (exactly a TF32 number); lo = a - hi (exact), rounded to nearest TF32 so
// that the tensor core's own truncation of its inputs is exact
__global__ void __launch_bounds__(256) tc_split_kernel(const float4* __restrict__ src, float4* __restrict__ hi, float4* __restrict__ lo, size_t n4)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldcs(src + i);
        float a[4] = { v.x, v.y, v.z, v.w }, h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            h[q] = __uint_as_float(__float_as_uint(a[q]) & 0xFFFFE000u);
            const float t = a[q] - h[q];
            l[q] = __uint_as_float((__float_as_uint(t) + 0x1000u) & 0xFFFFE000u);
        }
        hi[i] = make_float4(h[0], h[1], h[2], h[3]);
        lo[i] = make_float4(l[0], l[1], l[2], l[3]);
    }
}

// ---- the GEMM ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
matmul_tc05_kernel(const __grid_constant__ TcMaps maps, float* __restrict__ C, int M, int N, int K, int ldc, int mt)
{
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    unsigned char* tiles = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* acc_full = empty + TC_STAGES;          // [2] chunk accumulator complete (tcgen05.commit)
    uint64_t* acc_empty = acc_full + 2;              // [2] chunk accumulator read out by the epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tm = blockIdx.x % mt, tn = blockIdx.x / mt;
    const int m0 = tm * TC_BM, n0 = tn * TC_BN;
    const int nkb = (K + TC_BK - 1) / TC_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
        mbar_fence_init();
    }
    if (warp == 1) {
        // the accumulator: 128 lanes x 128 columns of TMEM, allocated (and later freed) by this warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            tma_prefetch_desc(&maps.ah); tma_prefetch_desc(&maps.at); tma_prefetch_desc(&maps.bh); tma_prefetch_desc(&maps.bt);
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                unsigned char* st = tiles + s * TC_STAGE_BYTES;
                mbar_arrive_expect_tx(&full[s], TC_STAGE_BYTES);
                const int k0 = kb * TC_BK;
#pragma unroll
                for (int b = 0; b < TC_BM / 32; b++) {                  // A tiles: 4 boxes of 32 m x 32 k
                    tma_load_2d(st + b * 4096, &maps.ah, &full[s], m0 + 32 * b, k0);
                    tma_load_2d(st + TC_A_BYTES + b * 4096, &maps.at, &full[s], m0 + 32 * b, k0);
                }
                tma_load_2d(st + 2 * TC_A_BYTES, &maps.bh, &full[s], k0, n0);
                tma_load_2d(st + 2 * TC_A_BYTES + TC_B_BYTES, &maps.bt, &full[s], k0, n0);
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc(TC_BM, TC_BN);
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1u;
                const int chunk = kb / TC_KCHUNK_BLOCKS, buf = chunk & 1;
                const bool first = kb % TC_KCHUNK_BLOCKS == 0, last = (kb + 1) % TC_KCHUNK_BLOCKS == 0 || kb + 1 == nkb;
                if (first) {
                    mbar_wait(&acc_empty[buf], (((uint32_t)chunk >> 1) & 1u) ^ 1u);      // the epilogue has read this buffer's previous chunk
                    tc_fence_after();
                }
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t st = smem_u32(tiles + s * TC_STAGE_BYTES);
                const uint32_t acc = tmem_acc + (uint32_t)(buf * TC_BN);
#pragma unroll
                for (int j = 0; j < TC_BK / 8; j++) {
                    // A (MN-major TF32: the only legal layout is the 128-byte swizzle with 32-byte atoms, "SWIZZLE_128B_BASE32B"):
                    // m-blocks of 32 are 4096 bytes apart (LBO), k-groups of 4 rows are 512 bytes apart (SBO); a k-step of 8
                    // is two groups = 1024 bytes
                    const uint64_t ah = tc_smem_desc(st + j * 1024, 4096, 512, 1u);
                    const uint64_t at = tc_smem_desc(st + TC_A_BYTES + j * 1024, 4096, 512, 1u);
                    // B (K-major, 128-byte swizzle): rows of 128 bytes, groups of 8 rows 1024 bytes apart (SBO); k-step = 32 bytes
                    const uint64_t bh = tc_smem_desc(st + 2 * TC_A_BYTES + j * 32, 16, 1024);
                    const uint64_t bt = tc_smem_desc(st + 2 * TC_A_BYTES + TC_B_BYTES + j * 32, 16, 1024);
                    tc_mma_tf32(acc, at, bh, idesc, (first && j == 0) ? 0u : 1u);
                    tc_mma_tf32(acc, ah, bt, idesc, 1u);
                    tc_mma_tf32(acc, ah, bh, idesc, 1u);
                }
                tc_commit(&empty[s]);                // the stage is free once these MMAs have read it
                if (last) tc_commit(&acc_full[buf]); // the chunk's accumulator is complete
            }
        }
    } else {
        // ------------------------------ epilogue: running sums += chunk accumulators; C += sums ------------------------------
        const int q = warp & 3;                      // this warp reads TMEM lanes [32 q, 32 q + 32)
        const int m = m0 + 32 * q + lane;
        float sum[TC_BN];
#pragma unroll
        for (int c = 0; c < TC_BN; c++) sum[c] = 0.f;
        const int nchunks = (nkb + TC_KCHUNK_BLOCKS - 1) / TC_KCHUNK_BLOCKS;
        for (int chunk = 0; chunk < nchunks; chunk++) {
            const int buf = chunk & 1;
            mbar_wait(&acc_full[buf], ((uint32_t)chunk >> 1) & 1u);
            tc_fence_after();
#pragma unroll
            for (int cb = 0; cb < TC_BN / 32; cb++) {
                uint32_t r[32];
                const uint32_t taddr = tmem_acc + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * TC_BN + cb * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 32; c++) sum[cb * 32 + c] += __uint_as_float(r[c]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);      // this warp has read the buffer out
        }
        if (m < M) {
#pragma unroll
            for (int c = 0; c < TC_BN; c++) {
                const int n = n0 + c;
                if (n < N) {
                    float* p = C + (size_t)n * ldc + m;
                    *p = *p + sum[c];
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TC_TMEM_COLS) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int tc_make_map(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_inner,
                       uint32_t box_outer, CUtensorMapSwizzle swz)
{
    static tc_encode_fn enc = nullptr;
    static std::mutex mu;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!enc) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            B200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
            if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available"); return B200_ERR_CUDA; }
            enc = (tc_encode_fn)fn;
        }
    }
    const cuuint64_t dims[2] = { inner, outer };
    const cuuint64_t strides[1] = { pitch_elems * 4 };
    const cuuint32_t box[2] = { box_inner, box_outer };
    const cuuint32_t estr[2] = { 1u, 1u };
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("matmul: cuTensorMapEncodeTiled failed (%d)", (int)r); return B200_ERR_CUDA; }
    return B200_OK;
}

// can the tcgen05 path take this problem?  (TMA: 16-byte aligned bases and pitches)
bool matmul_tc05_eligible(const float* A, const float* B, const float* C, int M, int N, int K)
{
    (void)C;
    return M > 0 && N > 0 && K > 0 && M % 4 == 0 && K % 4 == 0 && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0);
}

// C (M x N, ld M) += A (M x K, ld M) * B (K x N, ld K), column-major, on `stream`
int launch_matmul_tc05(const float* A, const float* B, float* C, int M, int N, int K, int num_sms, cudaStream_t stream)
{
    static bool prepared = false;
    if (!prepared) {
        B200_CUDA(cudaFuncSetAttribute(matmul_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        prepared = true;
    }
    // split workspace, stream-ordered: Ah, At (M*K each), Bh, Bt (K*N each)
    const size_t na = (size_t)M * K, nb = (size_t)K * N;
    float* ws = nullptr;
    B200_CUDA(cudaMallocAsync((void**)&ws, 2 * (na + nb) * sizeof(float), stream));
    float *ah = ws, *at = ws + na, *bh = ws + 2 * na, *bt = ws + 2 * na + nb;
    const int sgrid = num_sms * 8;
    tc_split_kernel<<<sgrid, 256, 0, stream>>>((const float4*)A, (float4*)ah, (float4*)at, na / 4);
    tc_split_kernel<<<sgrid, 256, 0, stream>>>((const float4*)B, (float4*)bh, (float4*)bt, nb / 4);
    B200_CUDA(cudaGetLastError());
    count_launch();
    count_launch();
    TcMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (int rc = tc_make_map(&maps.ah, ah, (uint64_t)M, (uint64_t)K, (uint64_t)M, 32, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    if (int rc = tc_make_map(&maps.at, at, (uint64_t)M, (uint64_t)K, (uint64_t)M, 32, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    if (int rc = tc_make_map(&maps.bh, bh, (uint64_t)K, (uint64_t)N, (uint64_t)K, TC_BK, TC_BN, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = tc_make_map(&maps.bt, bt, (uint64_t)K, (uint64_t)N, (uint64_t)K, TC_BK, TC_BN, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    const int mt = (M + TC_BM - 1) / TC_BM, nt = (N + TC_BN - 1) / TC_BN;
    matmul_tc05_kernel<<<mt * nt, TC_THREADS, TC_SMEM_BYTES, stream>>>(maps, C, M, N, K, M, mt);
    B200_CUDA(cudaGetLastError());
    count_launch();
    B200_CUDA(cudaFreeAsync(ws, stream));
    return B200_OK;
}

int info_matmul_tc05(KernelInfo* ki)
{
    cudaFuncAttributes fa;
    B200_CUDA(cudaFuncGetAttributes(&fa, matmul_tc05_kernel));
    ki->regs = fa.numRegs;
    ki->smem_bytes = TC_SMEM_BYTES;
    ki->blocks_per_sm = 1;
    ki->name = "matmul";
    return B200_OK;
}

}  // namespace b200
