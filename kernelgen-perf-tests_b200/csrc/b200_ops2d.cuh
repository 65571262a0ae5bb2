// b200_ops2d.cuh -- the 2D stencils (nx x ny, ny huge) as Ops of the tile-streaming engine.
//
// A work item is one TX x TY tile (one ring stage, no z-march).  Inside the tile each thread
// owns CPT consecutive rows of one 16-byte x-vector and marches over them in y with a sliding
// register window of x-windows, so every shared-memory row window is read once per thread.
//
// Every stencil is written once as a row kernel -- rows<PITCH, CPT>(p, emit): p points at this
// thread's vector in the first of its CPT rows inside a shared-memory tile of row pitch PITCH --
// and used by
//   * Plain2D<K>:  one sweep  (input tile -> global), and
//   * Fused2D<K>:  TWO sweeps in one pass over memory (temporal blocking): the first sweep's values
//     of the thread-owned TX x TY box go to a second shared-memory tile, the second sweep is
//     evaluated from it on the box shrunk by the stencil radius (overlapped tiling: tile pitch
//     TX-2V x TY-2R).  Per point the arithmetic is the single-sweep one, so results are bit-identical
//     to two plain sweeps.  The intermediate state is never written to memory; its values on the
//     global boundary are the shell of the buffer the reference would have written it to (w1),
//     read from there.
#pragma once

#include "b200_stream.cuh"

namespace b200 {

#define B200_UNROLL _Pragma("unroll")

// ------------------------------------------------------------------------------------------
// jacobi: w1 = c0*w0 + c1*(W + S + E + N) + c2*(SW + NW + SE + NE)          jacobi/jacobi.F90:60-69
// ------------------------------------------------------------------------------------------
template <typename T> struct JacobiK {
    using real = T;
    static constexpr int R = 1;
    static constexpr int NC_PLAIN = sizeof(T) == 8 ? 512 : 384;    // measured: 16 warps +3 % (double), -1 % (float)
    static constexpr bool PRED_STORE = true;
    static constexpr int V = 16 / (int)sizeof(T);
    T c0, c1, c2;
    B200_DEV JacobiK(const StreamParams& P) : c0((T)P.sc[0]), c1((T)P.sc[1]), c2((T)P.sc[2]) {}
    template <int PITCH, int CPT, class Emit> B200_DEV void rows(const T* p, Emit&& emit) const
    {
        Window<1, 1, T> wm, wc, wp;
        wm.load(p - PITCH);
        wc.load(p);
        B200_UNROLL
        for (int r = 0; r < CPT; r++) {
            wp.load(p + (r + 1) * PITCH);
            T o[V];
            B200_UNROLL
            for (int v = 0; v < V; v++)
                o[v] = c0 * wc.at(v, 0) +
                       c1 * (((wc.at(v, -1) + wm.at(v, 0)) + wc.at(v, 1)) + wp.at(v, 0)) +
                       c2 * (((wm.at(v, -1) + wp.at(v, -1)) + wm.at(v, 1)) + wp.at(v, 1));
            emit(r, o);
            wm = wc;
            wc = wp;
        }
    }
};

// ------------------------------------------------------------------------------------------
// gaussblur: 5x5, six weights, normalised by f = 1./(s0 + 4*(s1+s2+s4+s8) + 8*s5)
//                                                                  gaussblur/gaussblur.c:65,85-92
// ------------------------------------------------------------------------------------------
template <typename T> struct GaussblurK {
    using real = T;
    static constexpr int R = 2;
    static constexpr int NC_PLAIN = sizeof(T) == 4 ? 512 : 384;    // float: issue-bound, 16 warps +29 %; double needs 127 registers
    static constexpr bool PRED_STORE = sizeof(T) == 8;             // float at the 96-register cap of 16 warps: -14 %
    static constexpr int V = 16 / (int)sizeof(T);
    T s0, s1, s2, s4, s5, s8, f;
    B200_DEV GaussblurK(const StreamParams& P)
        : s0((T)P.sc[0]), s1((T)P.sc[1]), s2((T)P.sc[2]), s4((T)P.sc[3]), s5((T)P.sc[4]), s8((T)P.sc[5])
    {
        f = (T)(1. / (double)(s0 + 4 * (s1 + s2 + s4 + s8) + 8 * s5));
    }
    template <int PITCH, int CPT, class Emit> B200_DEV void rows(const T* p, Emit&& emit) const
    {
        Window<2, 2, T> a, b, c, d, e;      // rows j-2, j-1, j, j+1, j+2
        a.load(p - 2 * PITCH);
        b.load(p - PITCH);
        c.load(p);
        d.load(p + PITCH);
        B200_UNROLL
        for (int r = 0; r < CPT; r++) {
            e.load(p + (r + 2) * PITCH);
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
            emit(r, o);
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
template <typename T> struct GameoflifeK {
    using real = T;
    static constexpr int R = 1;
    static constexpr int NC_PLAIN = sizeof(T) == 4 ? 512 : 384;    // float: FP64-pipe latency, 16 warps +15 %
    static constexpr bool PRED_STORE = sizeof(T) == 8;             // float: -3 %
    static constexpr int V = 16 / (int)sizeof(T);
    double Cbig;
    B200_DEV GameoflifeK(const StreamParams&) { Cbig = (double)(T)100000000000000000000.; }
    B200_DEV static T add(T a, T b)
    {
        if constexpr (sizeof(T) == 4) return __fadd_rn(a, b);
        else return __dadd_rn(a, b);
    }
    template <int PITCH, int CPT, class Emit> B200_DEV void rows(const T* p, Emit&& emit) const
    {
        Window<1, 1, T> wm, wc, wp;
        wm.load(p - PITCH);
        wc.load(p);
        B200_UNROLL
        for (int r = 0; r < CPT; r++) {
            wp.load(p + (r + 1) * PITCH);
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
#ifdef B200_EXP_GOL_DDIV
                o[v] = (T)__ddiv_rn(1., den);
#else
                o[v] = (T)__drcp_rn(den);          // correctly rounded 1/den == __ddiv_rn(1., den), fewer instructions
#endif
            }
            emit(r, o);
            wm = wc;
            wc = wp;
        }
    }
};

// ------------------------------------------------------------------------------------------
// one sweep: input tile (halo R) -> global
// ------------------------------------------------------------------------------------------
template <class K> struct Plain2D : NoTmaStore {
    using real = typename K::real;
    using T = real;
    static constexpr int NC = pick_nc<T>(K::NC_PLAIN);
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? 96 : 48, NC, 128), STAGES = 4, HOLD = 0, WARM = 0, PERIOD = 1;
    static constexpr bool STREAM_OUT = false;
    static constexpr bool PRED_STORE = K::PRED_STORE;
    static constexpr int NSTAGED = 1;
    static constexpr StagedSpec spec(int) { return StagedSpec{0, 1, K::R, K::R, 0, 0}; }
    using G = Geo<Plain2D>;
    static constexpr int V = G::V, CPT = G::CPT;
    struct State { };
    K k;
    B200_DEV Plain2D(const StreamParams& P) : k(P) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    template <int PH, class C> B200_DEV void step(const C& ctx, State&)
    {
        const int r0 = ctx.ty * CPT;
        k.template rows<G::bw(0), CPT>(ctx.template tile<0>(r0), [&](int r, const T (&o)[V]) { ctx.template store<1>(r0 + r, o); });
    }
};

// ------------------------------------------------------------------------------------------
// two sweeps in one pass: arrays = { w0 (state t), w1 (only its boundary shell is read: the
// values of state t+1 on the global boundary), out (slot 2: receives state t+2 in the interior;
// must already hold w0's shell) }.  Single GPU (no halo push).
// ------------------------------------------------------------------------------------------
template <class K> struct Fused2D : NoTmaStore {
    using real = typename K::real;
    using T = real;
    static constexpr int R = K::R;
    static constexpr int NC = pick_nc<T>(384);
    static constexpr int TX = 128, TY = pick_ty<T>(sizeof(T) == 4 ? 96 : 48, NC, 128), STAGES = 2, HOLD = 0, WARM = 0, PERIOD = 1;
    static constexpr int V = 16 / (int)sizeof(T);
    static_assert(R <= V, "x shrink is one vector");
    static constexpr int PX = TX - 2 * V, PY = TY - 2 * R, OX = -V, OY = -R;       // overlapped tiling
    static constexpr int MW = TX + 2 * V, MH = TY + 2 * R;                         // intermediate tile (padded)
    static constexpr int M_ELEMS = MW * MH;
    static constexpr int EXTRA_SMEM = 2 * M_ELEMS * (int)sizeof(T);                 // double-buffered: one barrier per item
    static constexpr int EXTRA_ARRAYS = 1;                                          // slot 2: the output buffer
    static constexpr bool STREAM_OUT = false;
    static constexpr int NSTAGED = 1;
    static constexpr StagedSpec spec(int) { return StagedSpec{0, 1, R, R, 0, 0}; }
    using G = Geo<Fused2D>;
    static constexpr int CPT = G::CPT;
    struct State { };
    K k;
    int par;                        // which intermediate tile this item uses
    B200_DEV Fused2D(const StreamParams& P) : k(P), par(0) {}
    template <class C> B200_DEV void pre(const C&, State&) {}
    B200_DEV static void sync_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory"); }

    template <int PH, class C> B200_DEV void step(const C& ctx, State&)
    {
        const StreamParams& P = ctx.P;
        const int r0 = ctx.ty * CPT;
        T* const m = reinterpret_cast<T*>(ctx.extra) + par * M_ELEMS + (r0 + R) * MW + V + V * ctx.tx;   // this thread's vector, row r0
        par ^= 1;
        // ---- sweep 1 on every owned point: state t+1 -> intermediate tile
        k.template rows<G::bw(0), CPT>(ctx.template tile<0>(r0), [&](int r, const T (&o)[V]) {
            VReg<T> q;
            B200_UNROLL
            for (int v = 0; v < V; v++) q[v] = o[v];
            *reinterpret_cast<uint4*>(m + r * MW) = *reinterpret_cast<const uint4*>(q.v);
        });
        // ---- points of the global boundary are not swept: state t+1 there is w1's shell (trap 2)
        {
            const int gx = ctx.x, gy0 = ctx.Y0 + r0;
            const bool edge_x = gx < R || gx + V > P.nx - R;
            const bool edge_y = gy0 < R || gy0 + CPT > P.ny - R;
            if (edge_x || edge_y) {
                const T* w1 = reinterpret_cast<const T*>(P.arr[1]);
                B200_UNROLL
                for (int r = 0; r < CPT; r++) {
                    const int gy = gy0 + r;
                    if (gy < 0 || gy >= P.ny) continue;
                    const bool row_b = gy < R || gy >= P.ny - R;
                    B200_UNROLL
                    for (int v = 0; v < V; v++) {
                        const int x = gx + v;
                        if (x >= 0 && x < P.nx && (row_b || x < R || x >= P.nx - R)) m[r * MW + v] = w1[(size_t)gy * P.nx + x];
                    }
                }
            }
        }
        sync_consumers();
        // ---- sweep 2 from the intermediate tile; valid on the owned box shrunk by (V, R)
        const bool col_ok = ctx.tx >= 1 && ctx.tx < G::LX - 1;
        k.template rows<MW, CPT>(m, [&](int r, const T (&o)[V]) {
            const int row = r0 + r;
            if (col_ok && row >= R && row < TY - R) ctx.template store<2>(row, o);
        });
        // no second barrier: the next item writes the OTHER intermediate tile, and nobody can reach the item after
        // that (which reuses this one) before everybody has passed the next item's barrier, i.e. finished reading here
    }
};

template <typename T> using JacobiOp = Plain2D<JacobiK<T>>;
template <typename T> using GaussblurOp = Plain2D<GaussblurK<T>>;
template <typename T> using GameoflifeOp = Plain2D<GameoflifeK<T>>;
template <typename T> using Jacobi2Op = Fused2D<JacobiK<T>>;
template <typename T> using Gaussblur2Op = Fused2D<GaussblurK<T>>;
template <typename T> using Gameoflife2Op = Fused2D<GameoflifeK<T>>;

}  // namespace b200
