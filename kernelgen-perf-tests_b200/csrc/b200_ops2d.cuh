// b200_ops2d.cuh -- the 2D stencils (nx x ny, ny huge) as Ops of the tile-streaming engine.
//
// A work item is one TX x TY tile (one ring stage, no z-march).  Inside the tile each thread
// owns CPT consecutive rows of one 16-byte x-vector and marches over them in y with a sliding
// register window of x-windows, so every shared-memory row window is read once per thread.
#pragma once

#include "b200_stream.cuh"

namespace b200 {

#define B200_UNROLL _Pragma("unroll")

// ------------------------------------------------------------------------------------------
// jacobi: w1 = c0*w0 + c1*(W + S + E + N) + c2*(SW + NW + SE + NE)          jacobi/jacobi.F90:60-69
// ------------------------------------------------------------------------------------------
template <typename T> struct JacobiOp : NoTmaStore {
    using real = T;
    static constexpr int NC = pick_nc<T>(sizeof(T) == 8 ? 512 : 384);    // measured: 16 warps +3 % (double), -1 % (float)
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? 96 : 48, NC, 128), STAGES = 4, HOLD = 0, WARM = 0, PERIOD = 1;
    static constexpr bool STREAM_OUT = false;
    static constexpr int NSTAGED = 1;
    static constexpr StagedSpec spec(int) { return StagedSpec{0, 1, 1, 1, 0, 0}; }
    using G = Geo<JacobiOp>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { };
    T c0, c1, c2;
    B200_DEV JacobiOp(const StreamParams& P) : c0((T)P.sc[0]), c1((T)P.sc[1]), c2((T)P.sc[2]) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State&)
    {
        constexpr int BW = G::bw(0);
        const int r0 = ctx.ty * CPT;
        const T* p = ctx.template tile<0>(r0);
        Window<1, 1, T> wm, wc, wp;
        wm.load(p - BW);
        wc.load(p);
        B200_UNROLL
        for (int r = 0; r < CPT; r++) {
            wp.load(p + (r + 1) * BW);
            T o[V];
            B200_UNROLL
            for (int v = 0; v < V; v++)
                o[v] = c0 * wc.at(v, 0) +
                       c1 * (((wc.at(v, -1) + wm.at(v, 0)) + wc.at(v, 1)) + wp.at(v, 0)) +
                       c2 * (((wm.at(v, -1) + wp.at(v, -1)) + wm.at(v, 1)) + wp.at(v, 1));
            ctx.template store<1>(r0 + r, o);
            wm = wc;
            wc = wp;
        }
    }
};

// ------------------------------------------------------------------------------------------
// gaussblur: 5x5, six weights, normalised by f = 1./(s0 + 4*(s1+s2+s4+s8) + 8*s5)
//                                                                  gaussblur/gaussblur.c:65,85-92
// ------------------------------------------------------------------------------------------
template <typename T> struct GaussblurOp : NoTmaStore {
    using real = T;
    static constexpr int NC = pick_nc<T>(sizeof(T) == 4 ? 512 : 384);    // float: issue-bound, 16 warps +29 %; double needs 127 registers
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? 96 : 48, NC, 128), STAGES = 4, HOLD = 0, WARM = 0, PERIOD = 1;
    static constexpr bool STREAM_OUT = false;
    static constexpr int NSTAGED = 1;
    static constexpr StagedSpec spec(int) { return StagedSpec{0, 1, 2, 2, 0, 0}; }
    using G = Geo<GaussblurOp>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { };
    T s0, s1, s2, s4, s5, s8, f;
    B200_DEV GaussblurOp(const StreamParams& P)
        : s0((T)P.sc[0]), s1((T)P.sc[1]), s2((T)P.sc[2]), s4((T)P.sc[3]), s5((T)P.sc[4]), s8((T)P.sc[5])
    {
        f = (T)(1. / (double)(s0 + 4 * (s1 + s2 + s4 + s8) + 8 * s5));
    }
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State&)
    {
        constexpr int BW = G::bw(0);
        const int r0 = ctx.ty * CPT;
        const T* p = ctx.template tile<0>(r0);
        Window<2, 2, T> a, b, c, d, e;      // rows j-2, j-1, j, j+1, j+2
        a.load(p - 2 * BW);
        b.load(p - BW);
        c.load(p);
        d.load(p + BW);
        B200_UNROLL
        for (int r = 0; r < CPT; r++) {
            e.load(p + (r + 2) * BW);
            T o[V];
            B200_UNROLL
            for (int v = 0; v < V; v++)
                o[v] = f * (
                    s0 * c.at(v, 0) +
                    s1 * (((c.at(v, -1) + c.at(v, 1)) + b.at(v, 0)) + d.at(v, 0)) +
                    s2 * (((b.at(v, -1) + b.at(v, 1)) + d.at(v, -1)) + d.at(v, 1)) +
                    s4 * (((c.at(v, -2) + c.at(v, 2)) + a.at(v, 0)) + e.at(v, 0)) +
                    s5 * (((((((b.at(v, -2) + a.at(v, -1)) + a.at(v, 1)) + b.at(v, 2)) +
                             d.at(v, -2)) + e.at(v, -1)) + e.at(v, 1)) + d.at(v, 2)) +
                    s8 * (((a.at(v, -2) + a.at(v, 2)) + e.at(v, -2)) + e.at(v, 2)));
            ctx.template store<1>(r0 + r, o);
            a = b;
            b = c;
            c = d;
            d = e;
        }
    }
};

// ------------------------------------------------------------------------------------------
// gameoflife: L = sum of 8 neighbours (in real, textual order);
//             u1 = 1. / (1. + (u0 + L - 3.) * (L - 3.) * C),  C = 1e20 stored in real
//                                                                  gameoflife/gameoflife.c:82-91
// The rule has double literals: it is evaluated in double even when real = float, with
// correctly rounded, un-contracted operations, so the result is bit-identical to a strict-IEEE
// build of the reference.
// ------------------------------------------------------------------------------------------
template <typename T> struct GameoflifeOp : NoTmaStore {
    using real = T;
    static constexpr int NC = pick_nc<T>(sizeof(T) == 4 ? 512 : 384);    // float: FP64-pipe latency, 16 warps +15 %
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? 96 : 48, NC, 128), STAGES = 4, HOLD = 0, WARM = 0, PERIOD = 1;
    static constexpr bool STREAM_OUT = false;
    static constexpr int NSTAGED = 1;
    static constexpr StagedSpec spec(int) { return StagedSpec{0, 1, 1, 1, 0, 0}; }
    using G = Geo<GameoflifeOp>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { };
    double Cbig;
    B200_DEV GameoflifeOp(const StreamParams&) { Cbig = (double)(T)100000000000000000000.; }
    template <class C> B200_DEV void pre(const C&, State&) {}
    B200_DEV static T add(T a, T b)
    {
        if constexpr (sizeof(T) == 4) return __fadd_rn(a, b);
        else return __dadd_rn(a, b);
    }
    template <int PH, class C> B200_DEV void step(const C& ctx, State&)
    {
        constexpr int BW = G::bw(0);
        const int r0 = ctx.ty * CPT;
        const T* p = ctx.template tile<0>(r0);
        Window<1, 1, T> wm, wc, wp;
        wm.load(p - BW);
        wc.load(p);
        B200_UNROLL
        for (int r = 0; r < CPT; r++) {
            wp.load(p + (r + 1) * BW);
            T o[V];
            B200_UNROLL
            for (int v = 0; v < V; v++) {
                T L = add(wm.at(v, -1), wm.at(v, 0));
                L = add(L, wm.at(v, 1));
                L = add(L, wc.at(v, -1));
                L = add(L, wc.at(v, 1));
                L = add(L, wp.at(v, -1));
                L = add(L, wp.at(v, 0));
                L = add(L, wp.at(v, 1));
                const double x = __dadd_rn((double)add(wc.at(v, 0), L), -3.);
                const double y = __dadd_rn((double)L, -3.);
                const double den = __dadd_rn(1., __dmul_rn(__dmul_rn(x, y), Cbig));
                o[v] = (T)__ddiv_rn(1., den);
            }
            ctx.template store<1>(r0 + r, o);
            wm = wc;
            wc = wp;
        }
    }
};

}  // namespace b200
