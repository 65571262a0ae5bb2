// diag_fanout.cu -- DIAGNOSTICS, compiled only into libb200stencil_diag.so (make diag, -DB200_DIAG); never part of
// libb200stencil.so.
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

namespace b200 {

// ---- diagnostics (not part of the product path): B200_DEBUG_FANOUT=1 replaces the sweep by a plain
// grid-stride kernel with gradient's memory pattern (1 array read, 3 arrays written, 16-byte accesses),
// i.e. the ceiling of that pattern without any stencil machinery (tools/quick.sh gradient double).
template <typename T>
__global__ void __launch_bounds__(256) fanout_kernel(const T* __restrict__ a, T* __restrict__ b, T* __restrict__ c,
                                                     T* __restrict__ d, size_t nvec)
{
    constexpr int V = 16 / sizeof(T);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
        const VReg<T> v = ldv_stream(a + i * V);
        *reinterpret_cast<uint4*>(b + i * V) = *reinterpret_cast<const uint4*>(v.v);
        *reinterpret_cast<uint4*>(c + i * V) = *reinterpret_cast<const uint4*>(v.v);
        *reinterpret_cast<uint4*>(d + i * V) = *reinterpret_cast<const uint4*>(v.v);
    }
}

// same traffic, but in the engine's spatial pattern: one CTA per SM, 384 threads, items = (128-element x-tile,
// 24/48-row y-tile, z-chunk of 16 planes) dealt round-robin, each CTA marching over the planes of its item
template <typename T>
__global__ void __launch_bounds__(384) fanout_tiles_kernel(const T* __restrict__ a, T* __restrict__ b, T* __restrict__ c,
                                                           T* __restrict__ d, int nx, int ny, int ns)
{
    constexpr int V = 16 / sizeof(T), TX = 128, LX = TX / V, LY = 384 / LX, TY = sizeof(T) == 4 ? 48 : 24, CPT = TY / LY, ZC = 16;
    const int ntx = nx / TX, nty = (ny + TY - 1) / TY, nzc = (ns + ZC - 1) / ZC;
    const int tx = threadIdx.x % LX, ty = threadIdx.x / LX;
    for (int item = blockIdx.x; item < ntx * nty * nzc; item += gridDim.x) {
        const int zc = item / (ntx * nty), t = item - zc * ntx * nty, tyi = t / ntx, txi = t - tyi * ntx;
        for (int z = zc * ZC; z < min(ns, zc * ZC + ZC); z++) {
#pragma unroll
            for (int r = 0; r < CPT; r++) {
                const int y = tyi * TY + ty + LY * r;
                if (y >= ny) continue;
                const size_t off = ((size_t)z * ny + y) * nx + txi * TX + tx * V;
                const VReg<T> v = ldv_stream(a + off);
                *reinterpret_cast<uint4*>(b + off) = *reinterpret_cast<const uint4*>(v.v);
                *reinterpret_cast<uint4*>(c + off) = *reinterpret_cast<const uint4*>(v.v);
                *reinterpret_cast<uint4*>(d + off) = *reinterpret_cast<const uint4*>(v.v);
            }
        }
    }
}

template <typename T> static int launch_fanout(const HostArgs& a, int mode)
{
    const b200_sweep_desc& d = *a.desc;
    const size_t nvec = (size_t)d.nx * d.ny * d.ns / (16 / sizeof(T));
    if (mode == 3)
        fanout_tiles_kernel<T><<<a.num_sms, 384, 0, a.stream>>>((const T*)a.arrays[0], (T*)a.arrays[1], (T*)a.arrays[2],
                                                               (T*)a.arrays[3], d.nx, d.ny, d.ns);
    else if (mode == 2)
        fanout_kernel<T><<<a.num_sms, 256, 0, a.stream>>>((const T*)a.arrays[0], (T*)a.arrays[1], (T*)a.arrays[2],
                                                           (T*)a.arrays[3], nvec);
    else
    fanout_kernel<T><<<a.num_sms * 8, 256, 0, a.stream>>>((const T*)a.arrays[0], (T*)a.arrays[1], (T*)a.arrays[2],
                                                           (T*)a.arrays[3], nvec);
    B200_CUDA(cudaGetLastError());
    count_launch();
    return B200_OK;
}

// modes 4 / 5: the engine itself (TMA ring, 12 consumer warps, Ctx::store) doing a point-wise copy of u into
// ux, uy, uz -- without (4) / with (5) gradient's halo'd tile: the engine's ceiling for 1 read + 3 writes
template <typename T, int HALO> struct EngineFanoutOp : NoTmaStore {
    using real = T;
    static constexpr int NC = pick_nc<T>(384);
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? 24 : 12, NC, 128), STAGES = 6, HOLD = 0, WARM = 0, PERIOD = 1;
    static constexpr bool STREAM_OUT = true;
    static constexpr int NSTAGED = 1;
    static constexpr StagedSpec spec(int) { return StagedSpec{0, HALO, HALO, HALO, 0, 0}; }
#ifdef B200_EXP_TS
    static constexpr int NOUT = 3;
    static constexpr int out_slot(int q) { return 1 + q; }
    static constexpr int out_dpl(int) { return 0; }
#endif
    using G = Geo<EngineFanoutOp>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { };
    B200_DEV EngineFanoutOp(const StreamParams&) {}
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
            ctx.template store<2>(row, o);
            ctx.template store<3>(row, o);
        }
    }
};

int launch_tma_copy(int dtype, const HostArgs& a, int nout);
int launch_gradient_diag(int dtype, const HostArgs& a, int dbg)
{
    if (dbg == 6) return launch_tma_copy(dtype, a, 3);
    if (dbg == 4) return dtype == B200_F32 ? launch_stream<EngineFanoutOp<float, 0>>(a) : launch_stream<EngineFanoutOp<double, 0>>(a);
    if (dbg == 5) return dtype == B200_F32 ? launch_stream<EngineFanoutOp<float, 1>>(a) : launch_stream<EngineFanoutOp<double, 1>>(a);
    return dtype == B200_F32 ? launch_fanout<float>(a, dbg) : launch_fanout<double>(a, dbg);
}
}  // namespace b200
