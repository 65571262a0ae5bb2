// k_gradient.cu -- instantiates the gradient stencil (float + double) of the tile-streaming engine.
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

namespace b200 {

#ifdef B200_DIAG      // libb200stencil_diag.so only (diag_fanout.cu): plain kernels with gradient's memory pattern
int launch_gradient_diag(int dtype, const HostArgs& a, int dbg);
#endif
int launch_gradient(int dtype, const HostArgs& a)
{
#ifdef B200_DIAG
    static const int dbg = getenv("B200_DEBUG_FANOUT") ? atoi(getenv("B200_DEBUG_FANOUT")) : 0;
    if (dbg) return launch_gradient_diag(dtype, a, dbg);
    // forms A/B (profiles/r2_gradient_forms.txt): 1 plain stores, 2 branchy stores, 3 taller tiles, 4 16 warps, 5 = 1 + 3
    static const int form = getenv("B200_GRAD_FORM") ? atoi(getenv("B200_GRAD_FORM")) : 0;
#define GF(...) return dtype == B200_F32 ? launch_stream<GradientOp<float, __VA_ARGS__>>(a) : launch_stream<GradientOp<double, __VA_ARGS__>>(a)
    switch (form) {
    case 1: GF(24, 12, false, true);
    case 2: GF(24, 12, true, false);
    case 3: GF(48, 24, true, true);
    case 4: GF(32, 16, true, true, 512);
    case 5: GF(48, 24, false, true);
    case 6: GF(24, 12, false, false);
    default: break;
    }
#undef GF
#endif
    return dtype == B200_F32 ? launch_by_tile_policy<GradientOp<float>, GradientOp<float, 12>>(a) : launch_stream<GradientOp<double>>(a);
}
int info_gradient(int dtype, KernelInfo* ki)
{
    return dtype == B200_F32 ? info_stream<GradientOp<float>>(ki, "gradient") : info_stream<GradientOp<double>>(ki, "gradient");
}
}  // namespace b200
