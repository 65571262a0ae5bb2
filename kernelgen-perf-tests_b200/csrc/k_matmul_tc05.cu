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
//   warp 0   TMA producer: per k-block of 32, two 32 KB bulk copies (cp.async.bulk, the TMA unit's linear mode) bring the four
//            operand tiles (Ah, At | Bh, Bt) into a 3-stage shared-memory ring, completion on the stage's `full` mbarrier.
//            The tiles are contiguous in memory because the split pass (below) writes the operands PACKED: tile by tile, in
//            exactly the shared-memory image the tensor core wants -- both operands "K-major": 128 rows (m or n) x 32 k,
//            128-byte swizzle, zero-padded to whole tiles (A is transposed by the packing pass: fed m-contiguous, which for
//            TF32 requires the 128-byte swizzle on 32-byte atoms, the MMA runs at 42 % of its rate).  (First version: tensor-map
//            loads straight from the split arrays -- 640 requests of 128 bytes at 16 KB strides per k-block; the tensor pipe
//            sat at 40 % waiting for them, profiles/r2_ncu_matmul_tc05.txt.)
//   warp 1   MMA issuer: allocates 128 TMEM columns (the 128 x 128 FP32 accumulator), waits for `full`, issues
//            4 k-steps x 3 tcgen05.mma (M = 128, N = 128, K = 8) from shared-memory descriptors, and hands the stage back
//            with tcgen05.commit -> `empty` mbarrier; after the last k-block tcgen05.commit -> `acc_full`;
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) of the accumulator, C += acc with coalesced
//            128-byte stores (lane = row m, C is m-contiguous).
// Out-of-range parts of edge tiles are zero in the packed operands (they add 0) and masked in the epilogue, so every shape
// is eligible.  SASS: UTCHMMA / UBLKCP / LDTM / SYNCS.
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


// ---- PTX wrappers -----------------------------------------------------------------------------
// linear bulk copy global -> shared through the TMA unit, completion counted in bytes on an mbarrier
B200_DEV void tma_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
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
// instruction descriptor (cute::UMMA::InstrDescriptor layout): D = F32, A = B = TF32, both operands K-major
constexpr uint32_t tc_idesc(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- operand split + packing ------------------------------------------------------------------
// hi = a with the 13 low mantissa bits cleared (exactly a TF32 number); lo = a - hi (exact), rounded to nearest TF32 so
// that the tensor core's own truncation of its inputs is exact
B200_DEV void tc_split(float a, float& h, float& l)
{
    h = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
    const float t = a - h;
    l = __uint_as_float((__float_as_uint(t) + 0x1000u) & 0xFFFFE000u);
}
// A (M x K, column-major, ld = M) -> packed tiles [tm][kb]{ hi 16 KB | lo 16 KB }, TRANSPOSED to "K-major": row = m, 128 bytes
// (32 k) per row, 16-byte chunks XOR-swizzled with (m & 7)  (Swizzle<3,4,3> = SWIZZLE_128B), the same image as B's tiles.
// (tcgen05 also accepts A m-contiguous -- "MN-major", for TF32 only with the 128-byte swizzle on 32-byte atoms -- but then the
// MMA itself runs at 42 % of the TF32 rate: measured with the loads taken out of the loop, 139 TFLOP/s at 8192^3.)
// One block per tile: coalesced reads along m, transpose through shared memory, coalesced 128-byte row writes.
__global__ void __launch_bounds__(256) tc_pack_a_kernel(const float* __restrict__ A, int M, int K, unsigned char* __restrict__ ws, int mt, int nkb, int vec_ok)
{
    __shared__ float hi[TC_BK][TC_BM + 1], lo[TC_BK][TC_BM + 1];
    for (int tile = blockIdx.x; tile < mt * nkb; tile += gridDim.x) {
        const int tm = tile / nkb, kb = tile - tm * nkb;
        const int m0 = tm * TC_BM, k0 = kb * TC_BK;
        __syncthreads();
        for (int e = threadIdx.x; e < TC_BK * (TC_BM / 4); e += 256) {
            const int kk = e / (TC_BM / 4), m = 4 * (e % (TC_BM / 4));
            float a[4] = { 0.f, 0.f, 0.f, 0.f };
            if (k0 + kk < K) {
                const float* src = A + (size_t)(k0 + kk) * M + m0 + m;
                if (vec_ok && m0 + m + 3 < M) { const float4 t = __ldcs(reinterpret_cast<const float4*>(src)); a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w; }
                else {
#pragma unroll
                    for (int q = 0; q < 4; q++) if (m0 + m + q < M) a[q] = src[q];
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) tc_split(a[q], hi[kk][m + q], lo[kk][m + q]);
        }
        __syncthreads();
        unsigned char* out = ws + (size_t)tile * (2 * TC_A_BYTES);
        for (int e = threadIdx.x; e < TC_BM * (TC_BK / 4); e += 256) {
            const int mm = e / (TC_BK / 4), c = e % (TC_BK / 4);           // row m, 16-byte chunk c = k / 4
            const int chunk = c ^ (mm & 7);
            float4 h, l;
            h.x = hi[4 * c][mm]; h.y = hi[4 * c + 1][mm]; h.z = hi[4 * c + 2][mm]; h.w = hi[4 * c + 3][mm];
            l.x = lo[4 * c][mm]; l.y = lo[4 * c + 1][mm]; l.z = lo[4 * c + 2][mm]; l.w = lo[4 * c + 3][mm];
            *reinterpret_cast<float4*>(out + mm * 128 + chunk * 16) = h;
            *reinterpret_cast<float4*>(out + TC_A_BYTES + mm * 128 + chunk * 16) = l;
        }
    }
}
// B (K x N, column-major, ld = K) -> packed tiles [tn][kb]{ hi 16 KB | lo 16 KB }: row nn = n, 128 bytes (32 k) per row,
// 16-byte chunks XOR-swizzled with (nn & 7)  (Swizzle<3,4,3> = SWIZZLE_128B)
__global__ void __launch_bounds__(256) tc_pack_b_kernel(const float* __restrict__ B, int K, int N, unsigned char* __restrict__ ws, int nt, int nkb, int vec_ok)
{
    const size_t kq = (size_t)nkb * (TC_BK / 4), total = kq * (size_t)nt * TC_BN;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
        const int k = 4 * (int)(v % kq), n = (int)(v / kq);
        float a[4] = { 0.f, 0.f, 0.f, 0.f };
        if (n < N) {
            const float* src = B + (size_t)n * K + k;
            if (vec_ok && k + 3 < K) { const float4 t = __ldcs(reinterpret_cast<const float4*>(src)); a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w; }
            else {
#pragma unroll
                for (int q = 0; q < 4; q++) if (k + q < K) a[q] = src[q];
            }
        }
        float h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; q++) tc_split(a[q], h[q], l[q]);
        const int tn = n / TC_BN, kb = k / TC_BK, nn = n % TC_BN, kk = k % TC_BK;
        const int chunk = (kk >> 2) ^ (nn & 7);
        unsigned char* tile = ws + ((size_t)tn * nkb + kb) * (2 * TC_B_BYTES) + nn * 128 + chunk * 16;
        *reinterpret_cast<float4*>(tile) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(tile + TC_B_BYTES) = make_float4(l[0], l[1], l[2], l[3]);
    }
}

// ---- the GEMM ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
matmul_tc05_kernel(const unsigned char* __restrict__ wsA, const unsigned char* __restrict__ wsB, float* __restrict__ C, int M, int N, int K, int ldc, int mt)
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
            const unsigned char* srcA = wsA + (size_t)tm * nkb * (2 * TC_A_BYTES);
            const unsigned char* srcB = wsB + (size_t)tn * nkb * (2 * TC_B_BYTES);
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                unsigned char* st = tiles + s * TC_STAGE_BYTES;
                mbar_arrive_expect_tx(&full[s], TC_STAGE_BYTES);
                tma_bulk_load(st, srcA + (size_t)kb * (2 * TC_A_BYTES), 2 * TC_A_BYTES, &full[s]);                 // Ah | At
                tma_bulk_load(st + 2 * TC_A_BYTES, srcB + (size_t)kb * (2 * TC_B_BYTES), 2 * TC_B_BYTES, &full[s]); // Bh | Bt
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
                    // A and B tiles alike (K-major, 128-byte swizzle): rows of 128 bytes (32 k), groups of 8 rows 1024 bytes apart
                    // (SBO); a k-step of 8 = 32 bytes along the row
                    const uint64_t ah = tc_smem_desc(st + j * 32, 16, 1024);
                    const uint64_t at = tc_smem_desc(st + TC_A_BYTES + j * 32, 16, 1024);
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
// every float problem is eligible: the packing pass pads to whole tiles and reads unaligned operands element-wise
bool matmul_tc05_eligible(const float* A, const float* B, const float* C, int M, int N, int K)
{
    (void)A; (void)B; (void)C;
    return M > 0 && N > 0 && K > 0;
}

// C (M x N, ld M) += A (M x K, ld M) * B (K x N, ld K), column-major, on `stream`
int launch_matmul_tc05(const float* A, const float* B, float* C, int M, int N, int K, int num_sms, cudaStream_t stream)
{
    {   // function attributes are per device (multi-GPU contexts launch on several)
        static bool prepared[16] = {};
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        int dev = 0;
        B200_CUDA(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 16 && !prepared[dev]) {
            B200_CUDA(cudaFuncSetAttribute(matmul_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
            prepared[dev] = true;
        }
    }
    const int mt = (M + TC_BM - 1) / TC_BM, nt = (N + TC_BN - 1) / TC_BN, nkb = (K + TC_BK - 1) / TC_BK;
    // packed, split operands (stream-ordered workspace): A tiles mt x nkb x 32 KB, B tiles nt x nkb x 32 KB
    const size_t bytesA = (size_t)mt * nkb * 2 * TC_A_BYTES, bytesB = (size_t)nt * nkb * 2 * TC_B_BYTES;
    unsigned char* ws = nullptr;
    B200_CUDA(cudaMallocAsync((void**)&ws, bytesA + bytesB, stream));
    const int pgrid = num_sms * 8;
    tc_pack_a_kernel<<<(mt * nkb < pgrid * 4 ? mt * nkb : pgrid * 4), 256, 0, stream>>>(A, M, K, ws, mt, nkb, (M % 4 == 0) && ((uintptr_t)A % 16 == 0));
    tc_pack_b_kernel<<<pgrid, 256, 0, stream>>>(B, K, N, ws + bytesA, nt, nkb, (K % 4 == 0) && ((uintptr_t)B % 16 == 0));
    B200_CUDA(cudaGetLastError());
    count_launch();
    count_launch();
    matmul_tc05_kernel<<<mt * nt, TC_THREADS, TC_SMEM_BYTES, stream>>>(ws, ws + bytesA, C, M, N, K, M, mt);
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
