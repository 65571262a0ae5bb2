// k_laplacian.cu -- instantiates the laplacian stencil (float + double) of the tile-streaming engine.
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

namespace b200 {
int launch_debug_copy(int dtype, const HostArgs& a, int mode);
static int launch_laplacian_real(int dtype, const HostArgs& a)
{
    return dtype == B200_F32 ? launch_stream<LaplacianOp<float>>(a) : launch_stream<LaplacianOp<double>>(a);
}
int launch_laplacian(int dtype, const HostArgs& a)
{
    static const int dbg = getenv("B200_DEBUG_COPY") ? atoi(getenv("B200_DEBUG_COPY")) : 0;
    return dbg ? launch_debug_copy(dtype, a, dbg) : launch_laplacian_real(dtype, a);
}
int info_laplacian(int dtype, KernelInfo* ki)
{
    return dtype == B200_F32 ? info_stream<LaplacianOp<float>>(ki, "laplacian") : info_stream<LaplacianOp<double>>(ki, "laplacian");
}
}  // namespace b200

// ---- engine diagnostics (not part of the product path) ---------------------------------------
// B200_DEBUG_COPY=1 makes the laplacian entry point run a point-wise copy w1 = w0 through the same
// TMA ring / consumer / store machinery: the engine's own streaming ceiling, measured by
// tools/gap_probe.py.  B200_DEBUG_COPY=2: same with the laplacian's halo'd tile (loads the halos,
// ignores them).
namespace b200 {
template <typename T, int HALO> struct EngineCopyOp {
    using real = T;
    static constexpr int NC = pick_nc<T>(384);
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? 48 : 24, NC, 128), STAGES = 6, HOLD = 0, WARM = 0, PERIOD = 1;
    static constexpr bool STREAM_OUT = false;
    static constexpr int NSTAGED = 1;
    static constexpr StagedSpec spec(int) { return StagedSpec{0, HALO, HALO, HALO, 0, 0}; }
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
int launch_debug_copy(int dtype, const HostArgs& a, int mode)
{
    if (mode == 2) return dtype == B200_F32 ? launch_stream<EngineCopyOp<float, 1>>(a) : launch_stream<EngineCopyOp<double, 1>>(a);
    return dtype == B200_F32 ? launch_stream<EngineCopyOp<float, 0>>(a) : launch_stream<EngineCopyOp<double, 0>>(a);
}
}  // namespace b200
