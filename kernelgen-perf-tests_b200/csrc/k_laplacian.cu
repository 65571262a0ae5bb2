// k_laplacian.cu -- instantiates the laplacian stencil (float + double) of the tile-streaming engine.
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

namespace b200 {
#ifdef B200_DIAG      // libb200stencil_diag.so only (diag_copy.cu): the engine-copy ceiling behind the laplacian entry point
int launch_debug_copy(int dtype, const HostArgs& a, int mode);
#endif
int launch_laplacian(int dtype, const HostArgs& a)
{
#ifdef B200_DIAG
    static const int dbg = getenv("B200_DEBUG_COPY") ? atoi(getenv("B200_DEBUG_COPY")) : 0;
    if (dbg) return launch_debug_copy(dtype, a, dbg);
#endif
    return dtype == B200_F32 ? launch_by_tile_policy<LaplacianOp<float>, LaplacianOp<float, 24>>(a) : launch_stream<LaplacianOp<double>>(a);
}
int info_laplacian(int dtype, KernelInfo* ki)
{
    return dtype == B200_F32 ? info_stream<LaplacianOp<float>>(ki, "laplacian") : info_stream<LaplacianOp<double>>(ki, "laplacian");
}
}  // namespace b200

