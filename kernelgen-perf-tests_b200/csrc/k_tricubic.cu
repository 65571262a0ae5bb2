// k_tricubic.cu -- instantiates the tricubic stencil (float + double) of the tile-streaming engine.
//
// Two forms of the same arithmetic (b200_ops3d.cuh): TricubicOp (12 warps, one row of V points per thread)
// and TricubicRowsOp<T, 2> (8 warps, two adjacent rows per thread: 37.5 % less shared-memory traffic).
// B200_TRICUBIC_ROWS[_F64|_F32]=1|2 overrides the per-precision default (A/B measurements: tools/tricubic_ab.py).
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

namespace b200 {
#ifndef B200_TRICUBIC_ROWS_F64
#define B200_TRICUBIC_ROWS_F64 1
#endif
#ifndef B200_TRICUBIC_ROWS_F32
#define B200_TRICUBIC_ROWS_F32 1
#endif
static int tricubic_rows(int dtype)
{
    static const int env = getenv("B200_TRICUBIC_ROWS") ? atoi(getenv("B200_TRICUBIC_ROWS")) : 0;
    static const int env64 = getenv("B200_TRICUBIC_ROWS_F64") ? atoi(getenv("B200_TRICUBIC_ROWS_F64")) : 0;
    static const int env32 = getenv("B200_TRICUBIC_ROWS_F32") ? atoi(getenv("B200_TRICUBIC_ROWS_F32")) : 0;
    if (env == 1 || env == 2) return env;
    const int e = dtype == B200_F32 ? env32 : env64;
    if (e == 1 || e == 2) return e;
    return dtype == B200_F32 ? B200_TRICUBIC_ROWS_F32 : B200_TRICUBIC_ROWS_F64;
}
int launch_tricubic(int dtype, const HostArgs& a)
{
    if (tricubic_rows(dtype) == 2)
        return dtype == B200_F32 ? launch_stream<TricubicRowsOp<float, 2>>(a) : launch_stream<TricubicRowsOp<double, 2>>(a);
    return dtype == B200_F32 ? launch_stream<TricubicOp<float>>(a) : launch_stream<TricubicOp<double>>(a);
}
int info_tricubic(int dtype, KernelInfo* ki)
{
    if (tricubic_rows(dtype) == 2)
        return dtype == B200_F32 ? info_stream<TricubicRowsOp<float, 2>>(ki, "tricubic") : info_stream<TricubicRowsOp<double, 2>>(ki, "tricubic");
    return dtype == B200_F32 ? info_stream<TricubicOp<float>>(ki, "tricubic") : info_stream<TricubicOp<double>>(ki, "tricubic");
}
}  // namespace b200
