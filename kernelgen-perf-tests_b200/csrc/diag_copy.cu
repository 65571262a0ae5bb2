// diag_copy.cu -- DIAGNOSTICS, compiled only into libb200stencil_diag.so (make diag, -DB200_DIAG); never part of
// libb200stencil.so.
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

// ---- engine diagnostics (not part of the product path) ---------------------------------------
// B200_DEBUG_COPY=1 makes the laplacian entry point run a point-wise copy w1 = w0 through the same
// TMA ring / consumer / store machinery: the engine's own streaming ceiling, measured by
// tools/gap_probe.py.  B200_DEBUG_COPY=2: same with the laplacian's halo'd tile (loads the halos,
// ignores them).
namespace b200 {
template <typename T, int HX, int HY, int TXV, int TYD, int STG> struct EngineCopyOp : NoTmaStore {
    using real = T;
    static constexpr int NC = pick_nc<T>(384);
    static constexpr int TX = TXV, TY = pick_ty<T>(sizeof(T) == 4 ? 2 * TYD : TYD, NC, TXV), STAGES = STG, HOLD = 0, WARM = 0, PERIOD = 1;
    static constexpr bool STREAM_OUT = false;
    static constexpr int NSTAGED = 1;
    static constexpr StagedSpec spec(int) { return StagedSpec{0, HX ? 1 : 0, HY, HY, 0, 0}; }
#ifdef B200_EXP_TS
    static constexpr int NOUT = 1;
    static constexpr int out_slot(int) { return 1; }
    static constexpr int out_dpl(int) { return 0; }
#endif
    using G = Geo<EngineCopyOp>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { };
    B200_DEV EngineCopyOp(const StreamParams&) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State&)
    {
#pragma unroll
        for (int c = 0; c < CPT; c++) {
            const int row = ctx.ty + G::LY * c;
            const VReg<T> v = ldv(ctx.template tile<0>(row));
            T o[V];
#pragma unroll
            for (int i = 0; i < V; i++) o[i] = v[i];
            ctx.template store<1>(row, o);
        }
    }
};
template <int HX, int HY, int TXV, int TYD, int STG> static int launch_copy_variant(int dtype, const HostArgs& a)
{
    return dtype == B200_F32 ? launch_stream<EngineCopyOp<float, HX, HY, TXV, TYD, STG>>(a)
                             : launch_stream<EngineCopyOp<double, HX, HY, TXV, TYD, STG>>(a);
}
// B200_DEBUG_COPY=3 / 4: no consumers at all -- a pure TMA copy (global -> shared ring -> global, one loading and
// one storing thread per CTA, persistent grid, the engine's item order) of w0 into w1 (3) or of u into 3 arrays
// (4, via the gradient entry point).  Box = B200_TC_TX x B200_TC_TY elements.  The ceiling of a TMA-store path.
__global__ void __launch_bounds__(64) tma_copy_kernel(const __grid_constant__ CUtensorMap in, const __grid_constant__ TensorMaps out,
                                                       int nout, int tx, int ty, int ntx, int nty, int nz, int zc, int stage_bytes, int S)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 8;
    unsigned char* stages = smem + 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    const int nzc = (nz + zc - 1) / zc, nitems = ntx * nty * nzc;
    if (threadIdx.x == 0) {
        uint32_t g = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int c = item / (ntx * nty), t = item - c * ntx * nty, yi = t / ntx, xi = t - yi * ntx;
            for (int z = c * zc; z < min(nz, c * zc + zc); z++, g++) {
                const uint32_t st = g % S, ph = (g / S) & 1u;
                mbar_wait(&empty[st], ph ^ 1u);
                mbar_arrive_expect_tx(&full[st], (uint32_t)stage_bytes);
                tma_load_3d(stages + st * stage_bytes, &in, &full[st], xi * tx, yi * ty, z);
            }
        }
    } else if (threadIdx.x == 32) {
        constexpr int K = 2;                       // stores in flight
        uint32_t g = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int c = item / (ntx * nty), t = item - c * ntx * nty, yi = t / ntx, xi = t - yi * ntx;
            for (int z = c * zc; z < min(nz, c * zc + zc); z++, g++) {
                const uint32_t st = g % S, ph = (g / S) & 1u;
                mbar_wait(&full[st], ph);
                for (int q = 0; q < nout; q++)
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(&out.m[q])), "r"(smem_u32(stages + st * stage_bytes)),
                                   "r"(xi * tx), "r"(yi * ty), "r"(z) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(K) : "memory");
                if (g >= (uint32_t)K) mbar_arrive(&empty[(g - K) % S]);
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int launch_tma_copy(int dtype, const HostArgs& a, int nout)
{
    const b200_sweep_desc& d = *a.desc;
    const int esz = dtype == B200_F32 ? 4 : 8;
    const int tx = getenv("B200_TC_TX") ? atoi(getenv("B200_TC_TX")) : 128;
    const int ty = getenv("B200_TC_TY") ? atoi(getenv("B200_TC_TY")) : 24;
    const int S = getenv("B200_TC_S") ? atoi(getenv("B200_TC_S")) : 6;
    const int zc = getenv("B200_TC_ZC") ? atoi(getenv("B200_TC_ZC")) : 32;
    CUtensorMap in;
    TensorMaps out{};
    TmaBoxKey key{a.arrays[0], d.nx, d.ny, d.ns, esz, tx, ty};
    if (int rc = get_tensor_map(key, &in)) return rc;
    for (int q = 0; q < nout; q++) {
        TmaBoxKey ko{a.arrays[1 + q], d.nx, d.ny, d.ns, esz, tx, ty};
        if (int rc = get_tensor_map(ko, &out.m[q])) return rc;
    }
    const int stage_bytes = (tx * ty * esz + 127) / 128 * 128;
    const int smem = 256 + S * stage_bytes;
    B200_CUDA(cudaFuncSetAttribute(tma_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tma_copy_kernel<<<a.num_sms, 64, smem, a.stream>>>(in, out, nout, tx, ty, (d.nx + tx - 1) / tx, (d.ny + ty - 1) / ty, d.ns, zc,
                                                       tx * ty * esz, S);
    B200_CUDA(cudaGetLastError());
    count_launch();
    return B200_OK;
}

int launch_debug_copy(int dtype, const HostArgs& a, int mode)
{
    if (mode == 3) return launch_tma_copy(dtype, a, 1);
    switch (mode) {         // tile-shape / halo study (tools/quick.sh, profiles/README.md)
    case 10: return launch_copy_variant<0, 1, 128, 24, 6>(dtype, a);    // y halo only
    case 11: return launch_copy_variant<1, 0, 128, 24, 6>(dtype, a);    // x halo only
    case 12: return launch_copy_variant<1, 1, 128, 48, 4>(dtype, a);    // taller tile
    case 13: return launch_copy_variant<1, 1, 64, 48, 6>(dtype, a);     // narrower, taller
    case 14: return launch_copy_variant<1, 2, 128, 12, 6>(dtype, a);    // wave13pt / lapgsrb geometry
    case 15: return launch_copy_variant<1, 2, 64, 24, 6>(dtype, a);     // same point count, squarer
    case 16: return launch_copy_variant<1, 2, 128, 24, 4>(dtype, a);    // radius 2, taller
    case 17: return launch_copy_variant<1, 1, 128, 24, 4>(dtype, a);    // shallower ring
    default: break;
    }
    if (mode == 2) return launch_copy_variant<1, 1, 128, 24, 6>(dtype, a);
    return launch_copy_variant<0, 0, 128, 24, 6>(dtype, a);
}
}  // namespace b200
