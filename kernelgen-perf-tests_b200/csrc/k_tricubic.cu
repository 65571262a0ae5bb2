// k_tricubic.cu -- instantiates the tricubic stencil (float + double) of the tile-streaming engine.
//
// Four forms of the same arithmetic (b200_ops3d.cuh), bit-identical to each other:
//   1  TricubicOp: 12 warps, one row of V points per thread (shared-memory bound);
//   2  TricubicRowsOp<T, 2>: 8 warps at 168 registers, two adjacent rows per thread (37.5 % less shared-memory
//      traffic) -- default for float;
//   3  the same with 10 warps and a 5-stage ring (measured slower: kept as the measurement);
//   4  form 2 with a, b, c in the engine's transient ring and 8 stages of u0 -- default for double.
// B200_TRICUBIC_ROWS[_F64|_F32]=1..4 overrides the per-precision default (A/B: tools/r1r_run.sh, r1s_run.sh,
// r1t_run.sh -> profiles/r1r_tricubic_ab.txt, r1t_tricubic_ab.txt).
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

namespace b200 {
#ifndef B200_TRICUBIC_ROWS_F64
#define B200_TRICUBIC_ROWS_F64 4
#endif
#ifndef B200_TRICUBIC_ROWS_F32
#define B200_TRICUBIC_ROWS_F32 2
#endif
static int tricubic_rows(int dtype)
{
    static const int env = getenv("B200_TRICUBIC_ROWS") ? atoi(getenv("B200_TRICUBIC_ROWS")) : 0;
    static const int env64 = getenv("B200_TRICUBIC_ROWS_F64") ? atoi(getenv("B200_TRICUBIC_ROWS_F64")) : 0;
    static const int env32 = getenv("B200_TRICUBIC_ROWS_F32") ? atoi(getenv("B200_TRICUBIC_ROWS_F32")) : 0;
    if (env >= 1 && env <= 4) return env;
    const int e = dtype == B200_F32 ? env32 : env64;
    if (e >= 1 && e <= 4) return e;
    return dtype == B200_F32 ? B200_TRICUBIC_ROWS_F32 : B200_TRICUBIC_ROWS_F64;
}
template <typename T> using Rows2 = TricubicRowsOp<T, 2>;                 // 8 consumer warps, 6-stage ring
template <typename T> using Rows2W10 = TricubicRowsOp<T, 2, 320, 5>;      // 10 consumer warps, 5-stage ring
template <typename T> using Rows2S8 = TricubicRowsOp<T, 2, 256, 8, true>; // 8 warps, 8 stages of u0 + 5 slots of a, b, c
int launch_tricubic(int dtype, const HostArgs& a)
{
    switch (tricubic_rows(dtype)) {
    case 2: return dtype == B200_F32 ? launch_stream<Rows2<float>>(a) : launch_stream<Rows2<double>>(a);
    case 3: return dtype == B200_F32 ? launch_stream<Rows2W10<float>>(a) : launch_stream<Rows2W10<double>>(a);
    case 4: return dtype == B200_F32 ? launch_stream<Rows2S8<float>>(a) : launch_stream<Rows2S8<double>>(a);
    default: return dtype == B200_F32 ? launch_stream<TricubicOp<float>>(a) : launch_stream<TricubicOp<double>>(a);
    }
}
int info_tricubic(int dtype, KernelInfo* ki)
{
    switch (tricubic_rows(dtype)) {
    case 2: return dtype == B200_F32 ? info_stream<Rows2<float>>(ki, "tricubic") : info_stream<Rows2<double>>(ki, "tricubic");
    case 3: return dtype == B200_F32 ? info_stream<Rows2W10<float>>(ki, "tricubic") : info_stream<Rows2W10<double>>(ki, "tricubic");
    case 4: return dtype == B200_F32 ? info_stream<Rows2S8<float>>(ki, "tricubic") : info_stream<Rows2S8<double>>(ki, "tricubic");
    default: return dtype == B200_F32 ? info_stream<TricubicOp<float>>(ki, "tricubic") : info_stream<TricubicOp<double>>(ki, "tricubic");
    }
}
}  // namespace b200
