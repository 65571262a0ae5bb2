// k_tricubic.cu -- instantiates the tricubic stencil (float + double) of the tile-streaming engine.
//
// Product library: ONE form, TricubicOuterOp<T, 2> (b200_ops3d.cuh): two adjacent rows per thread,
// 8 consumer warps at 168 registers, the window row consumed as a rank-1 update so that consecutive FMAs share an operand
// (the B200 register file feeds one 64-bit operand per lane per clk: a DFMA with three distinct register operands takes
// 3 clk, tools/probes/regbank_probe.cu), a, b, c in the engine's transient ring, 8 stages of u0.
//
// Diagnostics library (make diag, -DB200_DIAG): B200_TRICUBIC_ROWS[_F64|_F32]=1..10 selects one of the forms measured on
// the way there (profiles/r1r_tricubic_ab.txt, r1t_tricubic_ab.txt, r2_tricubic_ab.txt); all are parity-tested
// (tests/test_gpu_parity.py::test_tricubic_row_variants), forms 1-6 are bit-identical to each other, forms 7-10 to each other:
//   1  TricubicOp: 12 warps, one row of V points per thread (shared-memory bound);
//   2  TricubicRowsOp<T, 2>: 8 warps at 168 registers, two adjacent rows per thread;
//   3  the same with 10 warps and a 5-stage ring;   4  form 2 + transient ring, 8 stages (round-1 default for double);
//   5  form 4 with 10 warps, 7 stages;   6  form 4 with 12 warps at 128 registers, 6 stages;
//   7  TricubicOuterOp<T, 2>, 8 warps (THE PRODUCT FORM);   8  with 12 warps at 128 registers;
//   9  with 64-wide tiles (TY = 16);   10 with 10 warps.
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

namespace b200 {
template <typename T> using Outer2S8 = TricubicOuterOp<T, 2, 256, 8, true>;  // outer-product form, 8 warps (168 regs)

#ifdef B200_DIAG
static int tricubic_rows(int dtype)
{
    static const int env = getenv("B200_TRICUBIC_ROWS") ? atoi(getenv("B200_TRICUBIC_ROWS")) : 0;
    static const int env64 = getenv("B200_TRICUBIC_ROWS_F64") ? atoi(getenv("B200_TRICUBIC_ROWS_F64")) : 0;
    static const int env32 = getenv("B200_TRICUBIC_ROWS_F32") ? atoi(getenv("B200_TRICUBIC_ROWS_F32")) : 0;
    if (env >= 1 && env <= 10) return env;
    const int e = dtype == B200_F32 ? env32 : env64;
    if (e >= 1 && e <= 10) return e;
    return 7;
}
template <typename T> using Rows2 = TricubicRowsOp<T, 2>;                 // 8 consumer warps, 6-stage ring
template <typename T> using Rows2W10 = TricubicRowsOp<T, 2, 320, 5>;      // 10 consumer warps, 5-stage ring
template <typename T> using Rows2S8 = TricubicRowsOp<T, 2, 256, 8, true>; // 8 warps, 8 stages of u0 + 5 slots of a, b, c
template <typename T> using Rows2W10S7 = TricubicRowsOp<T, 2, 320, 7, true>; // 10 warps (168 regs), 7 stages + 4 slots
template <typename T> using Rows2W12S6 = TricubicRowsOp<T, 2, 384, 6, true>; // 12 warps (128 regs), 6 stages + 3 slots
template <typename T> using Outer2W12 = TricubicOuterOp<T, 2, 384, 6, true>; // outer-product form, 12 warps (128 regs)
template <typename T> using Outer2X64 = TricubicOuterOp<T, 2, 256, 8, true, 64>;   // 64-wide tiles: 8 thread rows, TY = 16 (double)
template <typename T> using Outer2W10 = TricubicOuterOp<T, 2, 320, 7, true>;       // 10 warps
#define B200_TRI_FORMS(F) \
    switch (tricubic_rows(dtype)) { \
    case 1: F(TricubicOp) case 2: F(Rows2) case 3: F(Rows2W10) case 4: F(Rows2S8) case 5: F(Rows2W10S7) \
    case 6: F(Rows2W12S6) case 8: F(Outer2W12) case 9: F(Outer2X64) case 10: F(Outer2W10) default: F(Outer2S8) }
#else
#define B200_TRI_FORMS(F) F(Outer2S8)
#endif

int launch_tricubic(int dtype, const HostArgs& a)
{
#define F(Op) return dtype == B200_F32 ? launch_stream<Op<float>>(a) : launch_stream<Op<double>>(a);
    B200_TRI_FORMS(F)
#undef F
}
int info_tricubic(int dtype, KernelInfo* ki)
{
#define F(Op) return dtype == B200_F32 ? info_stream<Op<float>>(ki, "tricubic") : info_stream<Op<double>>(ki, "tricubic");
    B200_TRI_FORMS(F)
#undef F
}
}  // namespace b200
