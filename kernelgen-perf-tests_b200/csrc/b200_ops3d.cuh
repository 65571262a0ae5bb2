// b200_ops3d.cuh -- the 3D stencils as Ops of the tile-streaming engine (b200_stream.cuh).
//
// Every Op marches in z (2.5D): at the step that emits output plane s the ring stage holds
// plane s+lead of each staged input.  In-plane (x,y) neighbours are read from shared memory the
// moment a plane arrives and folded into per-thread partial sums; z neighbours live in
// per-thread REGISTER RINGS indexed by the compile-time phase PH = (step in item) mod PERIOD, so
// advancing the queue costs no register moves.  Reference formulas are cited per Op (file:line
// under the reference tree); evaluation order follows the reference where that is free,
// otherwise the re-association is noted (covered by the stated normwise tolerance).
#pragma once

#include "b200_stream.cuh"

namespace b200 {

#define B200_UNROLL _Pragma("unroll")

// ------------------------------------------------------------------------------------------
// laplacian: w1 = alpha*w0 + beta*(x+1 + x-1 + y+1 + y-1 + z+1 + z-1)     laplacian/laplacian.c:93-97
// Accumulate-forward form: when plane p arrives, acc_p = alpha*C_p + beta*(inplane_p + C_{p-1}) is
// formed and out_{p-1} = acc_{p-1} + beta*C_p is emitted: two live values per point, no queue.
// (beta is distributed over the z+1 term: a re-association.)
// ------------------------------------------------------------------------------------------
// TYF: tile height of the float form.  The second value each 3D Op is instantiated with (half the default) is the
// small-grid form of b200_launch.cuh: launch_by_tile_policy (more tiles, hence fewer and longer z-chunks).
template <typename T, int TYF = 48> struct LaplacianOp : NoTmaStore {
    using real = T;
    static constexpr int NC = pick_nc<T>(384);
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? TYF : 24, NC, 128), STAGES = 6, HOLD = 0, WARM = 2, PERIOD = 1;
    static constexpr bool STREAM_OUT = false;
    static constexpr int NSTAGED = 1;
    static constexpr StagedSpec spec(int) { return StagedSpec{0, 1, 1, 1, 1, 1}; }
#ifdef B200_EXP_TS      // experiment build: w1 through the TMA-store path (measured slower, profiles/README.md)
    static constexpr int NOUT = 1;
    static constexpr int out_slot(int) { return 1; }
    static constexpr int out_dpl(int) { return 0; }
#endif
    using G = Geo<LaplacianOp>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { T acc[CPT][V], cp[CPT][V]; };
    T alpha, beta;
    B200_DEV LaplacianOp(const StreamParams& P) : alpha((T)P.sc[0]), beta((T)P.sc[1]) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State& S)
    {
        constexpr int BW = G::bw(0);
        B200_UNROLL
        for (int c = 0; c < CPT; c++) {
            const int row = ctx.ty + G::LY * c;
            const T* p = ctx.template tile<0>(row);          // plane s+1
            Window<1, 1, T> w;
            w.load(p);
            const VReg<T> up = ldv(p + BW), dn = ldv(p - BW);
            if (ctx.rel >= 0) {
                T o[V];
                B200_UNROLL
                for (int v = 0; v < V; v++) o[v] = S.acc[c][v] + beta * w.at(v, 0);
                ctx.template store<1>(row, o);
            }
            B200_UNROLL
            for (int v = 0; v < V; v++) {
                const T ip = ((w.at(v, 1) + w.at(v, -1)) + up[v]) + dn[v];
                S.acc[c][v] = alpha * w.at(v, 0) + beta * (ip + S.cp[c][v]);
                S.cp[c][v] = w.at(v, 0);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// wave13pt: w2 = m0*w1 - w0 + m1*(6 radius-1 nbrs of w1) + m2*(6 radius-2 nbrs of w1)
//                                                                  wave13pt/wave13pt.c:640-651
// Accumulate-forward form (m1, m2 distributed over their sums: a re-association).  When plane
// p = s+2 of w1 arrives:
//     out_s      = acc_s + m2*C_p - w0_s                                   (emitted)
//     acc_{s+1} += m1*C_p
//     acc_p      = m0*C_p + inplane_p + m1*C_{p-1} + m2*C_{p-2}
// rings of two: acc[PH&1] = acc_s, acc[(PH+1)&1] = acc_{s+1}; Cq[PH&1] = C_s, Cq[(PH+1)&1] = C_{s+1}.
// ------------------------------------------------------------------------------------------
template <typename T, int TYF = 24> struct Wave13ptOp : NoTmaStore {
    using real = T;
    static constexpr int NC = pick_nc<T>(384);
#ifdef B200_EXP_WAVE_TY18
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? 36 : 18, NC, 128), STAGES = 5, HOLD = 0, WARM = 4, PERIOD = 2;
#else
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? TYF : 12, NC, 128), STAGES = 6, HOLD = 0, WARM = 4, PERIOD = 2;
#endif
    static constexpr bool STREAM_OUT = false;
    static constexpr bool PRED_STORE = sizeof(T) == 4;      // measured: float +8 %, double -2 %
    static constexpr int NSTAGED = 2;
    static constexpr StagedSpec spec(int a)
    {
        return a == 0 ? StagedSpec{1, 1, 2, 2, 2, 2}      // w1: radius-2 halo, arrives 2 planes ahead
                      : StagedSpec{0, 0, 0, 0, 0, 0, 1};  // w0: point-wise, plane s; overwritten by the next sweep
    }
    using G = Geo<Wave13ptOp>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { T acc[2][CPT][V]; T Cq[2][CPT][V]; };
    T m0, m1, m2;
    B200_DEV Wave13ptOp(const StreamParams& P) : m0((T)P.sc[0]), m1((T)P.sc[1]), m2((T)P.sc[2]) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State& S)
    {
        constexpr int BW = G::bw(0);
        constexpr int A0 = PH & 1, A1 = (PH + 1) & 1;
        B200_UNROLL
        for (int c = 0; c < CPT; c++) {
            const int row = ctx.ty + G::LY * c;
            const T* p = ctx.template tile<0>(row);          // w1 plane s+2
            Window<2, 2, T> w;
            w.load(p);
            const VReg<T> u1 = ldv(p + BW), d1 = ldv(p - BW), u2 = ldv(p + 2 * BW), d2 = ldv(p - 2 * BW);
            if (ctx.rel >= 0) {
                const VReg<T> w0 = ldv(ctx.template tile<1>(row));   // w0 plane s
                T o[V];
                B200_UNROLL
                for (int v = 0; v < V; v++) o[v] = (S.acc[A0][c][v] + m2 * w.at(v, 0)) - w0[v];
                ctx.template store<2>(row, o);
            }
            B200_UNROLL
            for (int v = 0; v < V; v++) {
                const T cp = w.at(v, 0);
                const T in1 = ((w.at(v, 1) + w.at(v, -1)) + u1[v]) + d1[v];
                const T in2 = ((w.at(v, 2) + w.at(v, -2)) + u2[v]) + d2[v];
                const T fresh = m0 * cp + m1 * (in1 + S.Cq[A1][c][v]) + m2 * (in2 + S.Cq[A0][c][v]);
                S.acc[A1][c][v] += m1 * cp;
                S.acc[A0][c][v] = fresh;
                S.Cq[A0][c][v] = cp;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// divergence: u = alpha*(ux[x+1]-ux[x-1]) + beta*(uy[y+1]-uy[y-1]) + gamma*(uz[z+1]-uz[z-1])
//                                                                  divergence/divergence.c:93-96
// ring z[2]: z[PH] = uz plane s-1, z[PH^1] = plane s; plane s+1 arrives.
// u is rewritten every sweep and never read: streamed past the L2.
// ------------------------------------------------------------------------------------------
template <typename T, int TYF = 24> struct DivergenceOp : NoTmaStore {
    using real = T;
    static constexpr int NC = 384;            // 3 staged arrays: a 16-warp tile would not fit the shared memory
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? TYF : 12, NC, 128), STAGES = 5, HOLD = 0, WARM = 2, PERIOD = 2;
    static constexpr bool STREAM_OUT = true;
    static constexpr int NSTAGED = 3;
    static constexpr StagedSpec spec(int a)
    {
        return a == 0 ? StagedSpec{1, 1, 0, 0, 0, 0}     // ux: x halo, plane s
             : a == 1 ? StagedSpec{2, 0, 1, 1, 0, 0}     // uy: y halo, plane s
                      : StagedSpec{3, 0, 0, 0, 1, 1};    // uz: planes s-1 .. s+1 through the ring
    }
    using G = Geo<DivergenceOp>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { T z[2][CPT][V]; };
    T alpha, beta, gamma;
    B200_DEV DivergenceOp(const StreamParams& P) : alpha((T)P.sc[0]), beta((T)P.sc[1]), gamma((T)P.sc[2]) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State& S)
    {
        constexpr int BW1 = G::bw(1);
        B200_UNROLL
        for (int c = 0; c < CPT; c++) {
            const int row = ctx.ty + G::LY * c;
            const VReg<T> zp = ldv(ctx.template tile<2>(row));      // uz plane s+1
            if (ctx.rel >= 0) {
                Window<1, 1, T> wx;
                wx.load(ctx.template tile<0>(row));
                const T* py = ctx.template tile<1>(row);
                const VReg<T> up = ldv(py + BW1), dn = ldv(py - BW1);
                T o[V];
                B200_UNROLL
                for (int v = 0; v < V; v++)
                    o[v] = alpha * (wx.at(v, 1) - wx.at(v, -1)) + beta * (up[v] - dn[v]) +
                           gamma * (zp[v] - S.z[PH][c][v]);
                ctx.template store<0>(row, o);
            }
            B200_UNROLL
            for (int v = 0; v < V; v++) S.z[PH][c][v] = zp[v];
        }
    }
};

// ------------------------------------------------------------------------------------------
// gradient: ux = alpha*(u[x+1]-u[x-1]); uy = beta*(u[y+1]-u[y-1]); uz = gamma*(u[z+1]-u[z-1])
//                                                                  gradient/gradient.c:95-97
// When plane p = s+1 of u arrives, ux and uy OF PLANE p are stored at once (they only need that
// plane) and uz of plane s = gamma*(C_p - C_{s-1}).  ring Cq[2]: Cq[PH&1] = C_{s-1}, Cq[(PH+1)&1] = C_s.
// The three outputs are rewritten every sweep and never read: streamed past the L2.
// ------------------------------------------------------------------------------------------
template <typename T, int TYF = 24, int TYD = 12, bool STREAM = true, bool PRED = true, int NCV = 384, int STG = 6> struct GradientOp : NoTmaStore {
    using real = T;
    static constexpr int NC = pick_nc<T>(NCV);
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? TYF : TYD, NC, 128), STAGES = STG, HOLD = 0, WARM = 2, PERIOD = 2;
    static constexpr bool STREAM_OUT = STREAM;
    static constexpr bool PRED_STORE = PRED;
    static constexpr int NSTAGED = 1;
    static constexpr StagedSpec spec(int) { return StagedSpec{0, 1, 1, 1, 1, 1}; }
#ifdef B200_EXP_TS      // experiment build: ux, uy (plane s+1) and uz (plane s) through the TMA-store path
    static constexpr int NOUT = 3;
    static constexpr int out_slot(int q) { return 1 + q; }
    static constexpr int out_dpl(int q) { return q < 2 ? 1 : 0; }
#endif
    using G = Geo<GradientOp>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { T Cq[2][CPT][V]; };
    T alpha, beta, gamma;
    B200_DEV GradientOp(const StreamParams& P) : alpha((T)P.sc[0]), beta((T)P.sc[1]), gamma((T)P.sc[2]) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State& S)
    {
        constexpr int BW = G::bw(0);
        constexpr int A0 = PH & 1;
        const bool xy_plane = ctx.rel >= -1 && ctx.s + 1 < ctx.zb;      // plane s+1 is an output plane of this item
        B200_UNROLL
        for (int c = 0; c < CPT; c++) {
            const int row = ctx.ty + G::LY * c;
            const T* p = ctx.template tile<0>(row);          // u plane s+1
            Window<1, 1, T> w;
            w.load(p);
            if (xy_plane) {
                const VReg<T> up = ldv(p + BW), dn = ldv(p - BW);
                T ox[V], oy[V];
                B200_UNROLL
                for (int v = 0; v < V; v++) {
                    ox[v] = alpha * (w.at(v, 1) - w.at(v, -1));
                    oy[v] = beta * (up[v] - dn[v]);
                }
                ctx.template store<1>(row, ox, 1);
                ctx.template store<2>(row, oy, 1);
            }
            if (ctx.rel >= 0) {
                T oz[V];
                B200_UNROLL
                for (int v = 0; v < V; v++) oz[v] = gamma * (w.at(v, 0) - S.Cq[A0][c][v]);
                ctx.template store<3>(row, oz);
            }
            B200_UNROLL
            for (int v = 0; v < V; v++) S.Cq[A0][c][v] = w.at(v, 0);
        }
    }
};

// ------------------------------------------------------------------------------------------
// uxx1: d = 0.25*(d1[c]+d1[j-1]+d1[k-1]+d1[j-1,k-1]);  u1 = u0 + (dth/d)*( c1*(xx[c]-xx[i-1]) +
//       c2*(xx[i+1]-xx[i-2]) + c1*(xy[c]-xy[j-1]) + c2*(xy[j+1]-xy[j-2]) + c1*(xz[c]-xz[k-1]) +
//       c2*(xz[k+1]-xz[k-2]) ),  dth = 1./nx                           uxx1/uxx1.c:70,95-105
// The d sum keeps the reference's textual order (it is ill-conditioned); one IEEE division.
// ring xq[3]: xq[(PH+k)%3] = xz plane s-2+k (k=0..2); plane s+1 arrives and replaces s-2.
// ------------------------------------------------------------------------------------------
template <typename T> struct Uxx1Op : NoTmaStore {
    using real = T;
    static constexpr int NC = pick_nc<T>(384);
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? 12 : 6, NC, 128), STAGES = 5, HOLD = 0, WARM = 3, PERIOD = 3;
    static constexpr bool STREAM_OUT = false;
    static constexpr int NSTAGED = 5;
    static constexpr StagedSpec spec(int a)
    {
        return a == 0 ? StagedSpec{0, 0, 0, 0, 0, 0, 1}  // u0: point-wise, plane s; overwritten by the next sweep
             : a == 1 ? StagedSpec{2, 0, 1, 0, 0, 1}     // d1: rows j-1..j, planes s-1..s
             : a == 2 ? StagedSpec{3, 1, 0, 0, 0, 0}     // xx: x-2..x+1, plane s
             : a == 3 ? StagedSpec{4, 0, 2, 1, 0, 0}     // xy: rows j-2..j+1, plane s
                      : StagedSpec{5, 0, 0, 0, 1, 2};    // xz: planes s-2..s+1 through the ring
    }
    using G = Geo<Uxx1Op>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { T xq[3][CPT][V]; T dp[2][CPT][V]; };
    T c1, c2, dth;
    B200_DEV Uxx1Op(const StreamParams& P) : c1((T)P.sc[0]), c2((T)P.sc[1]), dth((T)(1. / P.nx)) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State& S)
    {
        constexpr int BWD = G::bw(1), BWY = G::bw(3);
        constexpr int X0 = PH % 3, X1 = (PH + 1) % 3, X2 = (PH + 2) % 3;
        B200_UNROLL
        for (int c = 0; c < CPT; c++) {
            const int row = ctx.ty + G::LY * c;
            const VReg<T> zp = ldv(ctx.template tile<4>(row));      // xz plane s+1
            VReg<T> dc, dm;
            if (ctx.rel >= -1) {
                const T* pd = ctx.template tile<1>(row);            // d1 plane s
                dc = ldv(pd);
                dm = ldv(pd - BWD);
            }
            if (ctx.rel >= 0) {
                Window<2, 1, T> wx;
                wx.load(ctx.template tile<2>(row));
                const T* py = ctx.template tile<3>(row);
                const VReg<T> y0 = ldv(py), ym1 = ldv(py - BWY), yp1 = ldv(py + BWY), ym2 = ldv(py - 2 * BWY);
                const VReg<T> u0 = ldv(ctx.template tile<0>(row));
                T o[V];
                B200_UNROLL
                for (int v = 0; v < V; v++) {
                    const T d = (T)0.25 * (((dc[v] + dm[v]) + S.dp[0][c][v]) + S.dp[1][c][v]);
                    const T t = c1 * (wx.at(v, 0) - wx.at(v, -1)) + c2 * (wx.at(v, 1) - wx.at(v, -2)) +
                                c1 * (y0[v] - ym1[v]) + c2 * (yp1[v] - ym2[v]) +
                                c1 * (S.xq[X2][c][v] - S.xq[X1][c][v]) + c2 * (zp[v] - S.xq[X0][c][v]);
                    o[v] = u0[v] + (dth / d) * t;
                }
                ctx.template store<1>(row, o);
            }
            B200_UNROLL
            for (int v = 0; v < V; v++) {
                S.xq[X0][c][v] = zp[v];
                if (ctx.rel >= -1) {
                    S.dp[0][c][v] = dc[v];
                    S.dp[1][c][v] = dm[v];
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// lapgsrb: w1 = c0*w0 + c1*(6 faces) + c2*(12 edge diagonals) + c3*(6 radius-2)   lapgsrb/lapgsrb.c:93-117
// Accumulate-forward z-scatter form.  For an arriving plane p = s+2 the in-plane sums are formed
// once: C_p centre, F_p = x+-1 + y+-1, D_p = 4 in-plane diagonals, G_p = x+-2 + y+-2 (the edge
// diagonals that live in planes k+-1 of an output k are exactly F of those planes), then
//     out_s      = acc_s + c3*C_p                                           (emitted)
//     acc_{s+1} += c1*C_p + c2*F_p
//     acc_p      = c0*C_p + c1*F_p + c2*D_p + c3*G_p + c1*C_{p-1} + c2*F_{p-1} + c3*C_{p-2}
// Five live values per point: rings of two for acc and C, one F.  Re-associated.
// ------------------------------------------------------------------------------------------
template <typename T, int TYF = 24> struct LapgsrbOp : NoTmaStore {
    using real = T;
    static constexpr int NC = pick_nc<T>(384);
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? TYF : 12, NC, 128), STAGES = 8, HOLD = 0, WARM = 4, PERIOD = 2;
    static constexpr bool STREAM_OUT = false;
    static constexpr int NSTAGED = 1;
    static constexpr StagedSpec spec(int) { return StagedSpec{0, 1, 2, 2, 2, 2}; }
    using G = Geo<LapgsrbOp>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { T acc[2][CPT][V]; T Cq[2][CPT][V]; T Fp[CPT][V]; };
    T c0, c1, c2, c3;
    B200_DEV LapgsrbOp(const StreamParams& P) : c0((T)P.sc[0]), c1((T)P.sc[1]), c2((T)P.sc[2]), c3((T)P.sc[3]) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State& S)
    {
        constexpr int BW = G::bw(0);
        constexpr int A0 = PH & 1, A1 = (PH + 1) & 1;
        B200_UNROLL
        for (int c = 0; c < CPT; c++) {
            const int row = ctx.ty + G::LY * c;
            const T* p = ctx.template tile<0>(row);          // w0 plane s+2
            Window<2, 2, T> w;
            Window<1, 1, T> wu, wd;
            w.load(p);
            wu.load(p + BW);
            wd.load(p - BW);
            const VReg<T> u2 = ldv(p + 2 * BW), d2 = ldv(p - 2 * BW);
            if (ctx.rel >= 0) {
                T o[V];
                B200_UNROLL
                for (int v = 0; v < V; v++) o[v] = S.acc[A0][c][v] + c3 * w.at(v, 0);
                ctx.template store<1>(row, o);
            }
            B200_UNROLL
            for (int v = 0; v < V; v++) {
                const T cp = w.at(v, 0);
                const T Fn = ((w.at(v, 1) + w.at(v, -1)) + wu.at(v, 0)) + wd.at(v, 0);
                const T Dn = ((wu.at(v, 1) + wd.at(v, 1)) + wu.at(v, -1)) + wd.at(v, -1);
                const T Gn = ((w.at(v, 2) + w.at(v, -2)) + u2[v]) + d2[v];
                const T fresh = c0 * cp + c1 * (Fn + S.Cq[A1][c][v]) + c2 * (Dn + S.Fp[c][v]) + c3 * (Gn + S.Cq[A0][c][v]);
                S.acc[A1][c][v] += c1 * cp + c2 * Fn;
                S.acc[A0][c][v] = fresh;
                S.Cq[A0][c][v] = cp;
                S.Fp[c][v] = Fn;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// tricubic / tricubic2: 4x4x4 box with cubic Lagrange weights taken from a,b,c at the output point
//                          tricubic/tricubic.c:1027-1145 ; tricubic2/tricubic2.c:81-102
// Evaluated separably (x, then y, then z): 84 FMA + weights instead of the reference's 64
// three-factor products -- a re-association.  The four u0 planes s-1..s+2 stay in the ring
// (HOLD = 3); a,b,c are point-wise planes staged by TMA in the same ring (direct global loads
// exposed a DRAM latency per step: measured 8 stall cycles per issue on long scoreboard).
// tricubic and tricubic2 differ only in the interior bounds (b200_test_info).
// ------------------------------------------------------------------------------------------
template <typename T> B200_DEV void cubic_weights(T t, T (&w)[4])
{
    // 7 multiply-class operations + 3 adds for the four weights instead of 12 + 3: the weights are 30 of the ~114 FP64
    // instructions per point of a pipe-bound kernel.  s = t (t+1) / 6 and h = (t-1) (t+2) / 2 = 3 s - 1 (one FMA).
    const T sixth = (T)(1.0 / 6.0);
    const T tm1 = t - (T)1, tp1 = t + (T)1, tp2 = t + (T)2;
    const T s = (sixth * t) * tp1;          // t (t+1) / 6
    const T h = (T)3 * s - (T)1;            // (t-1) (t+2) / 2
    w[0] = s * tp2;
    w[1] = -(h * tp1);
    w[2] = h * t;
    w[3] = -(s * tm1);
}

template <typename T> struct TricubicOp : NoTmaStore {
    using real = T;
    static constexpr int NC = pick_nc<T>(384);
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? 12 : 6, NC, 128), STAGES = 6, HOLD = 3, WARM = 3, PERIOD = 1;
    static constexpr bool STREAM_OUT = false;
    static constexpr int NSTAGED = 4;
    static constexpr StagedSpec spec(int a)
    {
        return a == 0 ? StagedSpec{0, 1, 1, 2, 2, 1}      // u0: x -1..+2, y -1..+2, planes s-1..s+2 (held in the ring)
                      : StagedSpec{a + 1, 0, 0, 0, 0, 0}; // a, b, c (slots 2,3,4): point-wise, plane s
    }
    using G = Geo<TricubicOp>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { };
    B200_DEV TricubicOp(const StreamParams&) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State&)
    {
        if (ctx.rel < 0) return;
        constexpr int BW = G::bw(0);
        B200_UNROLL
        for (int c = 0; c < CPT; c++) {
            const int row = ctx.ty + G::LY * c;
            const VReg<T> va = ldv(ctx.template tile<1>(row)), vb = ldv(ctx.template tile<2>(row)),
                          vc = ldv(ctx.template tile<3>(row));
            T o[V];
            T wa[V][4], wb[V][4], wc[V][4];
            B200_UNROLL
            for (int v = 0; v < V; v++) {
                cubic_weights(va[v], wa[v]);
                cubic_weights(vb[v], wb[v]);
                cubic_weights(vc[v], wc[v]);
                o[v] = (T)0;
            }
            B200_UNROLL
            for (int kk = 0; kk < 4; kk++) {
                const T* p = ctx.template tile<0>(row, 3 - kk);     // u0 plane s+kk-1
                T pz[V];
                B200_UNROLL
                for (int v = 0; v < V; v++) pz[v] = (T)0;
                B200_UNROLL
                for (int jj = 0; jj < 4; jj++) {
                    Window<1, 2, T> w;
                    w.load(p + (jj - 1) * BW);
                    B200_UNROLL
                    for (int v = 0; v < V; v++) {
                        const T px = ((wa[v][0] * w.at(v, -1) + wa[v][1] * w.at(v, 0)) + wa[v][2] * w.at(v, 1)) +
                                     wa[v][3] * w.at(v, 2);
                        pz[v] += wb[v][jj] * px;
                    }
                }
                B200_UNROLL
                for (int v = 0; v < V; v++) o[v] += wc[v][kk] * pz[v];
            }
            ctx.template store<1>(row, o);
        }
    }
};

// Row-sharing form of the same stencil: a thread owns R ADJACENT rows (R * V points), so the u0 window
// rows it reads from shared memory -- R + 3 per plane instead of 4 R -- are shared between its points.
// Why: the one-row form is bound by shared-memory bandwidth (ncu: 85 % of the pipe's wavefronts; 16 window
// rows of 40 bytes per 2 points = 2.7 clk per point per SM), not by HBM.  With R = 2 the window traffic drops
// by 37.5 %.  The weights of R * V points (12 each) need more than 128 registers, so the Op runs 8 consumer
// warps (+ the producer: 3+2+2+2 over the sub-partitions, 168 registers) -- the same number of points in
// flight per SM as 12 warps x 1 row.  Per point the arithmetic and its order are those of TricubicOp
// (x, then y in increasing jj, then z): bit-identical results.
// SPLIT = true: a, b, c (read only at the step they arrive with) live in the shallower transient ring
// (b200_stream.cuh: Op::transient), which pays for two more stages of u0 in flight.
template <typename T, int R, int NCV = 256, int STG = 6, bool SPLIT = false> struct TricubicRowsOp : NoTmaStore {
    using real = T;
    static constexpr int NC = NCV;
    static constexpr int TX = 128, TY = R * (NC / (TX / (16 / (int)sizeof(T)))), STAGES = STG, HOLD = 3, WARM = 3, PERIOD = 1;
    static constexpr bool STREAM_OUT = false;
    static constexpr int NSTAGED = 4;
    static constexpr StagedSpec spec(int a)
    {
        return a == 0 ? StagedSpec{0, 1, 1, 2, 2, 1} : StagedSpec{a + 1, 0, 0, 0, 0, 0};
    }
    static constexpr bool transient(int a) { return SPLIT && a > 0; }
    using G = Geo<TricubicRowsOp>;
    static constexpr int V = G::V;
    static_assert(G::CPT == R, "R adjacent rows per thread");
    struct State { };
    B200_DEV TricubicRowsOp(const StreamParams&) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State&)
    {
        if (ctx.rel < 0) return;
        constexpr int BW = G::bw(0);
        const int row0 = R * ctx.ty;
        T wa[R][V][4], wb[R][V][4], wc[R][V][4], o[R][V];
        B200_UNROLL
        for (int q = 0; q < R; q++) {
            const VReg<T> va = ldv(ctx.template tile<1>(row0 + q)), vb = ldv(ctx.template tile<2>(row0 + q)),
                          vc = ldv(ctx.template tile<3>(row0 + q));
            B200_UNROLL
            for (int v = 0; v < V; v++) {
                cubic_weights(va[v], wa[q][v]);
                cubic_weights(vb[v], wb[q][v]);
                cubic_weights(vc[v], wc[q][v]);
                o[q][v] = (T)0;
            }
        }
        B200_UNROLL
        for (int kk = 0; kk < 4; kk++) {
            const T* p = ctx.template tile<0>(row0, 3 - kk);        // u0 plane s+kk-1, first owned row
            T pz[R][V];
            B200_UNROLL
            for (int q = 0; q < R; q++) {
                B200_UNROLL
                for (int v = 0; v < V; v++) pz[q][v] = (T)0;
            }
            B200_UNROLL
            for (int r = 0; r < R + 3; r++) {                       // window row row0 - 1 + r
                Window<1, 2, T> w;
                w.load(p + (r - 1) * BW);
                B200_UNROLL
                for (int q = 0; q < R; q++) {
                    const int jj = r - q;                           // this window row is row jj of point row q
                    if (jj >= 0 && jj <= 3) {
                        B200_UNROLL
                        for (int v = 0; v < V; v++) {
                            const T px = ((wa[q][v][0] * w.at(v, -1) + wa[q][v][1] * w.at(v, 0)) + wa[q][v][2] * w.at(v, 1)) +
                                         wa[q][v][3] * w.at(v, 2);
                            pz[q][v] += wb[q][v][jj] * px;
                        }
                    }
                }
            }
            B200_UNROLL
            for (int q = 0; q < R; q++) {
                B200_UNROLL
                for (int v = 0; v < V; v++) o[q][v] += wc[q][v][kk] * pz[q][v];
            }
        }
        B200_UNROLL
        for (int q = 0; q < R; q++) ctx.template store<1>(row0 + q, o[q]);
    }
};

// Outer-product form of the same stencil (round 2).  Why: on B200 a DFMA whose three source operands are
// three distinct registers occupies the FP64 path of an SM sub-partition for 3 clk, not 2 (the register file delivers one
// 64-bit operand per lane per clk; measured, tools/probes/regbank_probe.cu: 3.00 clk per warp instruction, 2.17 with one
// operand held in the operand-reuse cache, DMUL / DADD 2.00).  The separable x-y-z form above is made of such DFMAs
// (weight x window value + accumulator, all different) and ptxas finds a reusable operand for one in five.
// Here the sums are re-ordered so that the work of one window row is a RANK-1 UPDATE:
//     T[q][v][i] += wzy[q][v] * W[v + i]          q: point row, v: point in the vector, i = 0..3, W: the row's window
// with wzy = wc[kk] * wb[jj] formed once per (plane, row).  All the FMAs of a row are independent (sixteen accumulators per
// point row pair), consecutive ones share either the weight or the window value, and the x weights are applied once at
// the very end: out = sum_i wa[i] * T[i].  Same 84 multiply-adds per point (64 + 16 products wzy + 4), and the wa weights
// are not live during the march (they are computed at the end from a).  Re-associated with respect to TricubicOp
// (z and y before x): covered by the stated tolerance, not bit-identical to the other forms.
template <typename T, int R, int NCV = 256, int STG = 8, bool SPLIT = true, int TXV = 128> struct TricubicOuterOp : NoTmaStore {
    using real = T;
    static constexpr int NC = NCV;
    static constexpr int TX = TXV, TY = R * (NC / (TX / (16 / (int)sizeof(T)))), STAGES = STG, HOLD = 3, WARM = 3, PERIOD = 1;
    static constexpr bool STREAM_OUT = false;
    static constexpr int NSTAGED = 4;
    static constexpr StagedSpec spec(int a)
    {
        return a == 0 ? StagedSpec{0, 1, 1, 2, 2, 1} : StagedSpec{a + 1, 0, 0, 0, 0, 0};
    }
    static constexpr bool transient(int a) { return SPLIT && a > 0; }
    using G = Geo<TricubicOuterOp>;
    static constexpr int V = G::V;
    static_assert(G::CPT == R, "R adjacent rows per thread");
    struct State { };
    B200_DEV TricubicOuterOp(const StreamParams&) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State&)
    {
        if (ctx.rel < 0) return;
        constexpr int BW = G::bw(0);
        const int row0 = R * ctx.ty;
        T av[R][V], wb[R][V][4], wc[R][V][4], acc[R][V][4];
        B200_UNROLL
        for (int q = 0; q < R; q++) {
            const VReg<T> va = ldv(ctx.template tile<1>(row0 + q)), vb = ldv(ctx.template tile<2>(row0 + q)),
                          vc = ldv(ctx.template tile<3>(row0 + q));
            B200_UNROLL
            for (int v = 0; v < V; v++) {
                av[q][v] = va[v];
                cubic_weights(vb[v], wb[q][v]);
                cubic_weights(vc[v], wc[q][v]);
            }
        }
        B200_UNROLL
        for (int kk = 0; kk < 4; kk++) {
            const T* p = ctx.template tile<0>(row0, 3 - kk);        // u0 plane s+kk-1, first owned row
            B200_UNROLL
            for (int r = 0; r < R + 3; r++) {                       // window row row0 - 1 + r
                Window<1, 2, T> w;
                w.load(p + (r - 1) * BW);
                B200_UNROLL
                for (int q = 0; q < R; q++) {
                    const int jj = r - q;                           // this window row is row jj of point row q
                    if (jj >= 0 && jj <= 3) {
                        B200_UNROLL
                        for (int v = 0; v < V; v++) {
                            const T wzy = wc[q][v][kk] * wb[q][v][jj];
                            B200_UNROLL
                            for (int i = 0; i < 4; i++) {
                                if (kk == 0 && jj == 0) acc[q][v][i] = wzy * w.w[v + i];
                                else acc[q][v][i] = wzy * w.w[v + i] + acc[q][v][i];
                            }
                        }
                    }
                }
            }
        }
        B200_UNROLL
        for (int q = 0; q < R; q++) {
            T o[V];
            B200_UNROLL
            for (int v = 0; v < V; v++) {
                T wa[4];
                cubic_weights(av[q][v], wa);
                o[v] = ((wa[0] * acc[q][v][0] + wa[1] * acc[q][v][1]) + wa[2] * acc[q][v][2]) + wa[3] * acc[q][v][3];
            }
            ctx.template store<1>(row0 + q, o);
        }
    }
};

}  // namespace b200
