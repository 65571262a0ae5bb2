// b200_stream.cuh -- the tile-streaming stencil engine for sm_100a.
//
// One persistent, warp-specialised kernel template drives every stencil:
//
//   * the grid is cut into work items (x-tile, y-tile, z-chunk); CTA b takes
//     items b, b+grid, ... so that CTAs resident at the same time work on
//     neighbouring tiles (their halo overlap is then an L2 hit, not HBM);
//   * a producer warp walks the item's planes and brings one plane-tile of
//     every staged input (with its x/y halo) into a shared-memory ring with
//     TMA (cp.async.bulk.tensor.3d + mbarrier complete_tx); out-of-range
//     halo is zero-filled by the TMA unit;
//   * 8 - 16 consumer warps (Op::NC) wait on the stage's "full" mbarrier, run Op::step()
//     -- the stencil proper: in-plane neighbours from shared memory with
//     16-byte loads, z neighbours from per-thread register queues (2.5D
//     z-march) -- store results with 16-byte coalesced stores and hand the
//     stage back through the "empty" mbarrier;
//   * optional fused halo push (PUSH launches): the items that touch the slab's ends are walked first, each finished
//     one copies the planes a z-neighbour GPU needs as ghosts straight into that GPU's memory (peer pointer over
//     NVLink), and the neighbour ordering (flags) involves only those items.
//
// When a row pitch is not a multiple of 16 bytes TMA cannot be used; the
// producer warp then fills the same ring with bounds-checked scalar loads.
#pragma once

#include "b200_common.cuh"

namespace b200 {

// Consumer threads per CTA (one CTA per SM, + 1 producer warp) are chosen per Op (Op::NC):
//   384 (12 warps): 13 warps = 4+3+3+3 over the four SM sub-partitions, 128 registers per thread;
//   512 (16 warps): 17 warps put 5 on one sub-partition -> 96 registers per thread; for the Ops that
//                   fit, a third more warps hide more latency and keep more stores in flight.
// Experiment forms (-DB200_EXP_*) exist only in the diagnostics library (make diag: -DB200_DIAG): the product build refuses them.
#if !defined(B200_DIAG) && (defined(B200_EXP_TS) || defined(B200_EXP_WAVE_TY18) || defined(B200_EXP_GOL_DDIV) || \
                            defined(B200_EXP_NC32) || defined(B200_EXP_NC64))
#error "B200_EXP_* forms are diagnostics: build them with `make diag EXTRA=-DB200_EXP_...`"
#endif
#ifndef B200_EXP_NC32
#define B200_EXP_NC32 0
#endif
#ifndef B200_EXP_NC64
#define B200_EXP_NC64 0
#endif
// consumer threads of an Op: its own choice unless an experiment build overrides it (-DB200_EXP_NC32/64=384|512)
template <typename T> constexpr int pick_nc(int own) { return sizeof(T) == 4 ? (B200_EXP_NC32 ? B200_EXP_NC32 : own) : (B200_EXP_NC64 ? B200_EXP_NC64 : own); }
// tile height: the Op's preferred TY rounded up to a whole number of thread rows (NC / (TX / V))
template <typename T> constexpr int pick_ty(int want, int nc, int tx) { return (want + nc / (tx / (16 / (int)sizeof(T))) - 1) / (nc / (tx / (16 / (int)sizeof(T)))) * (nc / (tx / (16 / (int)sizeof(T)))); }
constexpr int MAX_STAGED = 5;
constexpr int MAX_STAGES = 8;           // 2*MAX_STAGES mbarriers fit the 128-byte header

// One staged (TMA-loaded) input of an Op.
struct StagedSpec {
    int slot;   // array slot (driver order) it is loaded from
    int hx;     // 1: tile carries an x halo of V elements (16 bytes) on each side
    int ylo;    // rows of halo below / above
    int yhi;
    int lead;   // at the step that emits output plane s, plane s+lead of this array arrives
    int zlo;    // deepest plane below the output plane that is needed (s - zlo)
    int dead;   // 1: the buffer is overwritten by the NEXT sweep (rotation), so what this sweep reads
                //    from it is dead afterwards: load with L2 evict-first and leave the L2 to the
                //    arrays the next sweep will read again
};

struct StreamParams {
    int nx, ny, ns;                 // array extents
    long long nxny;                 // nx * ny (elements per plane)
    int xlo, xhi, ylo, yhi;         // output box (half-open), x and y
    int z0, z1;                     // output planes
    int ntx, nty, nzc, zc_len;      // work decomposition
    int nitems;
    int use_tma;                    // 0: scalar fallback loader
    int vec_ok;                     // 16-byte stores legal
    void* arr[8];                   // device pointers, slot order
    double sc[8];                   // scalars
    // fused halo push of the exchanged output array
    void* push_lo; int push_lo_src, push_lo_dst, push_lo_cnt;
    void* push_hi; int push_hi_src, push_hi_dst, push_hi_cnt;
    int push_slot;
    int push_dim;                   // 2: planes (3D tests), 1: rows (2D tests)
    // in-kernel ordering between neighbour ranks (PUSH launches): before the sweep spin until
    // *wait_flag[i] >= wait_value; after it the last CTA to finish stores signal_value to signal_flag[i]
    const unsigned long long* wait_flag[2];
    unsigned long long wait_value;
    unsigned long long* signal_flag[2];
    unsigned long long signal_value;
    unsigned int* done_counter;     // per-device completion counter of the END items (library-owned, self-resetting)
    int reverse;                    // 1: walk the items in reverse order (serpentine sweeps)
    // PUSH launches walk the "end" units of the split dimension first (z-chunks for the 3D tests, tile rows for the 2D
    // tests: the e_lo lowest and e_hi highest of split_n units, split_stride items each).  Only those items read ghost
    // planes or push planes to a neighbour, so only they take part in the neighbour ordering: the producer waits for the
    // neighbours' flags before the first of them, and the flags are published when the last of them has completed --
    // early in the sweep, long before a neighbour's next sweep asks for them (was: every CTA spun at kernel start on a flag
    // the neighbour's LAST CTA released, so every sweep inherited the slowest neighbour's tail).
    int split_n, split_stride, e_lo, e_hi, end_items;
};

struct alignas(64) TensorMaps {
    CUtensorMap m[MAX_STAGED];
};

// Output side (TMA stores).  An Op that lists its outputs (NOUT > 0) has its results staged in shared
// memory and written by cp.async.bulk.tensor from a dedicated store warp: whole tiles leave the SM as bulk
// transactions, the consumer warps never wait on a global store.  The tensor map of an output covers
// exactly the part of the interior that whole 16-byte vectors can reach -- x in [V, nx-V), rows
// [ylo, yhi) -- so the boundary shell is protected by the TMA unit's own clipping; the first and last
// vector of a row (which contain boundary elements) are stored by their threads, element-predicated.
constexpr int MAX_OUT = 3;
// (A TMA store must not start at a negative coordinate -- it raises an illegal-instruction error -- so
// the first x-tile, whose first vector is excluded, uses a second map with a box narrower by V and a
// staging tile packed accordingly.)
struct alignas(64) OutMaps {
    CUtensorMap m[MAX_OUT];        // box TX x TY, x-tiles 1..
    CUtensorMap m0[MAX_OUT];       // box (TX - V) x TY, x-tile 0
};
struct NoTmaStore {                       // defaults of an Op; the register store path
    // work-item grid: tile (i, j) has its thread-owned box at x = i*PX + OX, y = ylo + j*PY + OY.  Plain Ops
    // own exactly what they write (PX = TX, PY = TY); the fused two-sweep Ops own a ring more than they
    // write (overlapped tiling), so their pitch is smaller than the tile and the origin is shifted.
    static constexpr bool PRED_STORE = true;   // branch-free predicated stores (measured per Op: profiles/README.md)
    static constexpr int EXTRA_SMEM = 0;  // bytes of shared memory for the Op's own use (after the ring)
    static constexpr int EXTRA_ARRAYS = 0; // arrays beyond the test's own (fused Ops: the output buffer, slot narrays)
    // transient(a): staged array a is read only at the step it arrives with (point-wise inputs of an Op whose
    // other input is HELD in the ring for HOLD more steps).  Such arrays live in a second, shallower ring of
    // STAGES - HOLD slots (slot = step mod that) instead of occupying every stage: the shared memory saved buys
    // a deeper ring, i.e. more planes in flight.  No extra barriers: a stage's "empty" barrier completes only
    // after the consumers are HOLD steps past it, so the producer is never more than STAGES - HOLD steps ahead
    // of the slowest consumer -- except across an item boundary, where the held stages are handed back early;
    // there the first WARM >= HOLD steps of the next item carry no transient data (see Geo's static_assert).
    static constexpr bool transient(int) { return false; }
    static constexpr int NOUT = 0;
    static constexpr int out_slot(int) { return -1; }
    static constexpr int out_dpl(int) { return 0; }
};

constexpr int round_up_c(int a, int b) { return (a + b - 1) / b * b; }

// Compile-time geometry of an Op.
template <class Op> struct Geo {
    using T = typename Op::real;
    static constexpr int V = 16 / (int)sizeof(T);
    static constexpr int TX = Op::TX;
    static constexpr int TY = Op::TY;
    static constexpr int LX = TX / V;            // threads along x
    static constexpr int NC = Op::NC;            // consumer threads
    static constexpr int NTHREADS = NC + 32 + (Op::NOUT > 0 ? 32 : 0);     // + the producer warp (+ the store warp)
    static constexpr int LY = NC / LX;           // thread rows
    static constexpr int CPT = TY / LY;          // rows ("columns" in z) per thread
    static_assert((NC == 256 || NC == 320 || NC == 384 || NC == 512) && TX % V == 0 && NC % LX == 0 && TY % LY == 0, "bad tile");
    static constexpr int hxp(int a) { return Op::spec(a).hx ? V : 0; }
    static constexpr int bw(int a) { return TX + 2 * hxp(a); }
    static constexpr int bh(int a) { return TY + Op::spec(a).ylo + Op::spec(a).yhi; }
    static constexpr int box_bytes(int a) { return bw(a) * bh(a) * (int)sizeof(T); }
    static constexpr int arr_bytes(int a) { return round_up_c(box_bytes(a), 128); }
    static constexpr bool tr(int a) { return Op::transient(a); }
    static constexpr int arr_off(int a)              // offset inside a stage of the array's own ring
    {
        int o = 0;
        for (int b = 0; b < a; b++)
            if (tr(b) == tr(a)) o += arr_bytes(b);
        return o;
    }
    static constexpr int ring_stage_bytes(bool t)
    {
        int o = 0;
        for (int b = 0; b < Op::NSTAGED; b++)
            if (tr(b) == t) o += arr_bytes(b);
        return o;
    }
    static constexpr int STAGE_BYTES = ring_stage_bytes(false);
    static constexpr int TSTAGE_BYTES = ring_stage_bytes(true);          // 0 for most Ops
    static constexpr int TSLOTS = Op::STAGES - Op::HOLD;                 // slots of the transient ring
    static constexpr int TRING_OFF = Op::STAGES * STAGE_BYTES;
    static_assert(TSTAGE_BYTES == 0 || (Op::HOLD > 0 && Op::WARM >= Op::HOLD), "transient arrays need WARM >= HOLD > 0");
    // address of staged array A for ring stage `st` / transient slot `tst`
    template <int A> B200_DEV static unsigned char* arr_ptr(unsigned char* stages, uint32_t st, uint32_t tst)
    {
        if constexpr (tr(A)) return stages + TRING_OFF + tst * TSTAGE_BYTES + arr_off(A);
        else return stages + st * STAGE_BYTES + arr_off(A);
    }
    static constexpr int NOUT = Op::NOUT;
    static constexpr bool TS = NOUT > 0;                                 // TMA-store output path
    static constexpr int OB = 2;                                         // output staging buffers
    static constexpr int OUT_TILE_BYTES = TX * TY * (int)sizeof(T);
    static constexpr int OUT_BYTES = NOUT * OUT_TILE_BYTES;              // one staging buffer: NOUT tiles
    static constexpr int HDR_BYTES = 256;                                // mbarriers
    static constexpr int RING_BYTES = Op::STAGES * STAGE_BYTES + (TSTAGE_BYTES ? TSLOTS * TSTAGE_BYTES : 0);
    static constexpr int EXTRA_BYTES = round_up_c(Op::EXTRA_SMEM, 128);
    static constexpr int SMEM_BYTES = HDR_BYTES + 128 /*alignment slack*/ + RING_BYTES + EXTRA_BYTES + OB * OUT_BYTES;
    static_assert(SMEM_BYTES <= 232448, "tile does not fit the 227 KB of shared memory");
    static_assert(NOUT <= MAX_OUT && OUT_TILE_BYTES % 128 == 0, "bad output tile");
    // ordinal of output slot SLOT among the Op's TMA-stored outputs
    static constexpr int out_q(int slot)
    {
        for (int q = 0; q < NOUT; q++)
            if (Op::out_slot(q) == slot) return q;
        return -1;
    }
    // does the step of output plane s store anything?  (output q goes to plane s + out_dpl(q))
    static constexpr bool emits(int s, int za, int zb)
    {
        for (int q = 0; q < NOUT; q++)
            if (s + Op::out_dpl(q) >= za && s + Op::out_dpl(q) < zb) return true;
        return false;
    }
    static_assert(Op::STAGES <= MAX_STAGES && Op::NSTAGED <= MAX_STAGED, "too many stages");
    static_assert(Op::STAGES > Op::HOLD, "ring too shallow");
};

// What Op::step() sees.  Everything invariant over an item (tile) or a step is computed once
// there: a store is a row test, one multiply-add for the 32-bit element offset and one 64-bit
// address add off a warp-uniform plane pointer.  PUSH = this launch also stores halo planes into
// a neighbour GPU's memory; single-GPU launches are compiled without that code.
// slot of the step in the transient ring: a member only for the Ops that have transient arrays (empty base otherwise,
// so every other kernel keeps its Ctx layout)
template <bool HAS> struct CtxTransient { static constexpr uint32_t tst = 0; };
template <> struct CtxTransient<true> { uint32_t tst = 0; };

template <class Op, bool PUSH, bool TS = false> struct Ctx : CtxTransient<Geo<Op>::TSTAGE_BYTES != 0> {
    using T = typename Op::real;
    using G = Geo<Op>;
    static constexpr int V = G::V;
    const StreamParams& P;
    unsigned char* stages;      // base of the ring
    unsigned char* extra;       // the Op's own shared memory (Op::EXTRA_SMEM bytes)
    unsigned char* ostage;      // TS: output staging buffer of this step
    int opitch, oshift;         // TS, per item: row pitch of the staging tile and the column it starts at (x-tile 0: TX-V, V)
    uint32_t st;                // ring stage of this step
    int X0, Y0;                 // global x of tile column 0, global y of tile row 0
    int s;                      // output plane of this step
    int rel;                    // s - (first output plane of the item); < 0 during warm-up
    int zb;                     // per item: end of the item's output planes
    int tx, ty;                 // consumer thread coordinates
    int x;                      // per item: global x of this thread's vector
    int xmode;                  // per item: 1 = whole vector inside [xlo,xhi) and 16-byte stores legal,
                                //           2 = some elements inside, 0 = none
    int rows_valid;             // per item: tile rows [0, rows_valid) are inside [ylo, yhi)
    int pvec, pmask;            // per item: store predicates -- whole-vector store / bit v: element v alone (edge vectors)
    unsigned idx0;              // per item: Y0 * nx + x  (element index of tile row 0 within a plane)
    long long poff;             // per step: s * nx * ny  (element offset of the output plane; warp-uniform)

    B200_DEV void begin_item(int X0_, int Y0_)
    {
        X0 = X0_;
        Y0 = Y0_;
        x = X0 + V * tx;
        const bool some = x + V > P.xlo && x < P.xhi;
        const bool all = x >= P.xlo && x + V <= P.xhi;
        xmode = (all && P.vec_ok) ? 1 : (some ? 2 : 0);
        rows_valid = min(G::TY, P.yhi - Y0);
        idx0 = (unsigned)(Y0 * P.nx + x);
        pvec = xmode == 1;
        pmask = 0;
        if (xmode == 2) {
#pragma unroll
            for (int v = 0; v < V; v++)
                if (x + v >= P.xlo && x + v < P.xhi) pmask |= 1 << v;
        }
        if constexpr (TS) {
            oshift = X0 == 0 ? V : 0;
            opitch = G::TX - oshift;
        }
    }

    // Pointer to this thread's 16-byte vector in tile row `row` (tile-local output row, may be
    // negative / >= TY inside the halo) of staged array A, in the stage loaded `back` steps ago.
    template <int A> B200_DEV const T* tile(int row, int back = 0) const
    {
        uint32_t q = st;
        if (back) q = (st + (uint32_t)Op::STAGES - (uint32_t)back) % (uint32_t)Op::STAGES;
        const T* base;
        if constexpr (G::tr(A)) base = reinterpret_cast<const T*>(G::template arr_ptr<A>(stages, q, this->tst));   // back == 0
        else base = reinterpret_cast<const T*>(stages + q * G::STAGE_BYTES + G::arr_off(A));
        return base + (row + Op::spec(A).ylo) * G::bw(A) + G::hxp(A) + V * tx;
    }
    B200_DEV int gx() const { return x; }

    // Read-only pointer into a global array at (x of this thread, tile row, plane s).
    template <int SLOT> B200_DEV const T* gptr(int row) const
    {
        return reinterpret_cast<const T*>(P.arr[SLOT]) + poff + (idx0 + (unsigned)(row * P.nx));
    }
    // true when this thread's whole 16-byte vector at (row, any plane) is inside the array
    B200_DEV bool vec_in_array(int row) const
    {
        return P.vec_ok && x + V <= P.nx && (Y0 + row) >= 0 && (Y0 + row) < P.ny;
    }

    // 16-byte (or element-predicated) store of V values
    B200_DEV void put(T* dst, const T (&val)[V]) const
    {
        if (xmode == 1) {
            VReg<T> r;
#pragma unroll
            for (int v = 0; v < V; v++) r[v] = val[v];
            if constexpr (Op::STREAM_OUT) __stcs(reinterpret_cast<uint4*>(dst), *reinterpret_cast<const uint4*>(r.v));
            else *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(r.v);
        } else {
#pragma unroll
            for (int v = 0; v < V; v++)
                if (x + v >= P.xlo && x + v < P.xhi) dst[v] = val[v];
        }
    }

    // Branch-free store: one predicated 16-byte store (interior vectors) plus V predicated element stores
    // (edge vectors).  Predication instead of branches keeps a step one basic block, so the loads of the next
    // rows are scheduled above the stores of this one.
    B200_DEV static void st_vec_pred(T* dst, const T (&val)[V], int on)
    {
        if constexpr (sizeof(T) == 8) {
            if constexpr (Op::STREAM_OUT)
                asm volatile("{ .reg .pred p; setp.ne.s32 p, %0, 0; @p st.global.cs.v2.f64 [%1], {%2, %3}; }" ::"r"(on), "l"(dst), "d"(val[0]), "d"(val[1]) : "memory");
            else
                asm volatile("{ .reg .pred p; setp.ne.s32 p, %0, 0; @p st.global.v2.f64 [%1], {%2, %3}; }" ::"r"(on), "l"(dst), "d"(val[0]), "d"(val[1]) : "memory");
        } else {
            if constexpr (Op::STREAM_OUT)
                asm volatile("{ .reg .pred p; setp.ne.s32 p, %0, 0; @p st.global.cs.v4.f32 [%1], {%2, %3, %4, %5}; }" ::"r"(on), "l"(dst), "f"(val[0]), "f"(val[1]), "f"(val[2]), "f"(val[3]) : "memory");
            else
                asm volatile("{ .reg .pred p; setp.ne.s32 p, %0, 0; @p st.global.v4.f32 [%1], {%2, %3, %4, %5}; }" ::"r"(on), "l"(dst), "f"(val[0]), "f"(val[1]), "f"(val[2]), "f"(val[3]) : "memory");
        }
    }
    B200_DEV static void st_elem_pred(T* dst, T val, int on)
    {
        if constexpr (sizeof(T) == 8)
            asm volatile("{ .reg .pred p; setp.ne.s32 p, %0, 0; @p st.global.f64 [%1], %2; }" ::"r"(on), "l"(dst), "d"(val) : "memory");
        else
            asm volatile("{ .reg .pred p; setp.ne.s32 p, %0, 0; @p st.global.f32 [%1], %2; }" ::"r"(on), "l"(dst), "f"(val) : "memory");
    }
    B200_DEV void put_pred(T* dst, const T (&val)[V], int row_ok) const
    {
        st_vec_pred(dst, val, pvec & row_ok);
#pragma unroll
        for (int v = 0; v < V; v++) st_elem_pred(dst + v, val[v], (pmask >> v) & row_ok);
    }

    // Store V results of output array SLOT at tile row `row` of the step's output plane (plane s
    // for the 3D tests, the only plane for the 2D tests); interior-predicated.
    // dplane: store into plane s + dplane instead (no halo push for those).
    template <int SLOT> B200_DEV void store(int row, const T (&val)[V], int dplane = 0) const
    {
        const unsigned off = idx0 + (unsigned)(row * P.nx);
        if constexpr (Op::PRED_STORE && !(TS && G::out_q(SLOT) >= 0)) {
            // the common case: branch-free
            put_pred(reinterpret_cast<T*>(P.arr[SLOT]) + (poff + dplane * P.nxny) + off, val, row < rows_valid ? 1 : 0);
        } else {
            if (row < rows_valid && xmode != 0) {
                if constexpr (TS && G::out_q(SLOT) >= 0) {
                    if (xmode == 1) {
                        // whole vector inside the interior: into the staging tile, the store warp sends it with TMA
                        VReg<T> r;
#pragma unroll
                        for (int v = 0; v < V; v++) r[v] = val[v];
                        T* dst = reinterpret_cast<T*>(ostage + G::out_q(SLOT) * G::OUT_TILE_BYTES) + (row * opitch + V * tx - oshift);
                        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(r.v);
                    } else {
                        put(reinterpret_cast<T*>(P.arr[SLOT]) + (poff + dplane * P.nxny) + off, val);
                    }
                } else {
                    put(reinterpret_cast<T*>(P.arr[SLOT]) + (poff + dplane * P.nxny) + off, val);
                }
            }
        }
    }
};

// Ops keep their z queues in register rings indexed by a compile-time phase (no shifting moves):
// the step loop dispatches on phase = step-in-item mod Op::PERIOD.
template <class Op, class C, int PH = 0>
B200_DEV void step_dispatch(Op& op, const C& ctx, typename Op::State& state, int phase)
{
    if constexpr (PH + 1 >= Op::PERIOD) {
        op.template step<PH>(ctx, state);
    } else {
        if (phase == PH) op.template step<PH>(ctx, state);
        else step_dispatch<Op, C, PH + 1>(op, ctx, state, phase);
    }
}

// Producer-side fallback: fill one staged box with bounds-checked scalar loads (zero fill outside).
template <class Op, int A>
B200_DEV void fallback_fill(const StreamParams& P, unsigned char* arr, int X0, int Y0, int plane, int lane)
{
    using G = Geo<Op>;
    using T = typename Op::real;
    constexpr int BW = G::bw(A), BH = G::bh(A);
    T* dst = reinterpret_cast<T*>(arr);
    const T* src = reinterpret_cast<const T*>(P.arr[Op::spec(A).slot]);
    const int gx0 = X0 - G::hxp(A), gy0 = Y0 - Op::spec(A).ylo;
    const bool zin = plane >= 0 && plane < P.ns;
#pragma unroll 4
    for (int e = lane; e < BW * BH; e += 32) {
        const int r = e / BW, c = e - r * BW;
        const int x = gx0 + c, y = gy0 + r;
        T v = T(0);
        if (zin && x >= 0 && x < P.nx && y >= 0 && y < P.ny)
            v = src[((size_t)plane * P.ny + (size_t)y) * P.nx + x];
        dst[e] = v;
    }
}

template <class Op, int A> struct ProducerIssue {
    B200_DEV static void bytes(int s, int za, uint32_t& total)
    {
        if constexpr (A < Op::NSTAGED) {
            if (s + Op::spec(A).lead >= za - Op::spec(A).zlo) total += Geo<Op>::box_bytes(A);
            ProducerIssue<Op, A + 1>::bytes(s, za, total);
        }
    }
    B200_DEV static void tma(const TensorMaps& M, unsigned char* stage, uint64_t* bar, int X0, int Y0, int s, int za)
    {
        if constexpr (A < Op::NSTAGED) {
            if (s + Op::spec(A).lead >= za - Op::spec(A).zlo)
                tma_load_3d_hint(stage + Geo<Op>::arr_off(A), &M.m[A], bar, X0 - Geo<Op>::hxp(A),
                                 Y0 - Op::spec(A).ylo, s + Op::spec(A).lead,
                                 Op::spec(A).dead ? L2_EVICT_FIRST : L2_EVICT_NORMAL);
            ProducerIssue<Op, A + 1>::tma(M, stage, bar, X0, Y0, s, za);
        }
    }
    B200_DEV static void fallback(const StreamParams& P, unsigned char* stage, int X0, int Y0, int s, int za, int lane)
    {
        if constexpr (A < Op::NSTAGED) {
            if (s + Op::spec(A).lead >= za - Op::spec(A).zlo)
                fallback_fill<Op, A>(P, stage + Geo<Op>::arr_off(A), X0, Y0, s + Op::spec(A).lead, lane);
            ProducerIssue<Op, A + 1>::fallback(P, stage, X0, Y0, s, za, lane);
        }
    }
    // Ops with transient arrays: two rings
    B200_DEV static void tma2(const TensorMaps& M, unsigned char* stages, uint32_t st, uint32_t tst, uint64_t* bar, int X0, int Y0, int s, int za)
    {
        if constexpr (A < Op::NSTAGED) {
            if (s + Op::spec(A).lead >= za - Op::spec(A).zlo)
                tma_load_3d_hint(Geo<Op>::template arr_ptr<A>(stages, st, tst), &M.m[A], bar, X0 - Geo<Op>::hxp(A),
                                 Y0 - Op::spec(A).ylo, s + Op::spec(A).lead,
                                 Op::spec(A).dead ? L2_EVICT_FIRST : L2_EVICT_NORMAL);
            ProducerIssue<Op, A + 1>::tma2(M, stages, st, tst, bar, X0, Y0, s, za);
        }
    }
    B200_DEV static void fallback2(const StreamParams& P, unsigned char* stages, uint32_t st, uint32_t tst, int X0, int Y0, int s, int za, int lane)
    {
        if constexpr (A < Op::NSTAGED) {
            if (s + Op::spec(A).lead >= za - Op::spec(A).zlo)
                fallback_fill<Op, A>(P, Geo<Op>::template arr_ptr<A>(stages, st, tst), X0, Y0, s + Op::spec(A).lead, lane);
            ProducerIssue<Op, A + 1>::fallback2(P, stages, st, tst, X0, Y0, s, za, lane);
        }
    }
};

struct ItemCoords { int X0, Y0, za, zb; bool is_end; };

// tile pitch / origin: an Op may declare PX, PY, OX, OY (fused two-sweep Ops); default = the tile itself
template <class Op, class = void> struct TileGrid { static constexpr int PX = Op::TX, PY = Op::TY, OX = 0, OY = 0; };
template <class Op> struct TileGrid<Op, decltype((void)Op::PX)> { static constexpr int PX = Op::PX, PY = Op::PY, OX = Op::OX, OY = Op::OY; };
template <class Op> constexpr int tile_px() { return TileGrid<Op>::PX; }
template <class Op> constexpr int tile_py() { return TileGrid<Op>::PY; }
template <class Op> constexpr int tile_ox() { return TileGrid<Op>::OX; }
template <class Op> constexpr int tile_oy() { return TileGrid<Op>::OY; }

template <class Op, bool PUSH> B200_DEV ItemCoords decode_item(const StreamParams& P, int item)
{
    const int tiles_xy = P.ntx * P.nty;
    ItemCoords c;
    c.is_end = false;
    if constexpr (PUSH) {
        // ends first: position u' in the walk -> unit u of the split dimension
        const int up = item / P.split_stride;
        int w = item - up * P.split_stride;
        int u;
        if (up < P.e_lo) u = up;
        else if (up < P.e_lo + P.e_hi) u = P.split_n - P.e_hi + (up - P.e_lo);
        else {
            const int m = up - P.e_lo - P.e_hi, nm = P.split_n - P.e_lo - P.e_hi;
            u = P.e_lo + (P.reverse ? nm - 1 - m : m);
            if (P.reverse) w = P.split_stride - 1 - w;
        }
        c.is_end = up < P.e_lo + P.e_hi;
        item = u * P.split_stride + w;
    } else {
        if (P.reverse) item = P.nitems - 1 - item;
    }
    const int zc = item / tiles_xy;
    const int t = item - zc * tiles_xy;
    const int tyi = t / P.ntx, txi = t - tyi * P.ntx;
    c.X0 = txi * tile_px<Op>() + tile_ox<Op>();        // multiples of V: vectors stay 16-byte aligned
    c.Y0 = P.ylo + tyi * tile_py<Op>() + tile_oy<Op>();
    c.za = P.z0 + zc * P.zc_len;
    c.zb = min(P.z1, c.za + P.zc_len);
    return c;
}

// Copies what one finished END item contributes to the neighbours' ghosts: for each side, the planes (3D tests) or rows
// (2D tests) of the pushed output array that lie in both the item's box and the side's source range.
template <class Op> B200_DEV void push_item(const StreamParams& P, const ItemCoords& c, int rows_valid, int tid)
{
    using T = typename Op::real;
    using G = Geo<Op>;
    constexpr int V = G::V;
    const T* own = reinterpret_cast<const T*>(P.arr[P.push_slot]);
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
        T* peer = reinterpret_cast<T*>(side ? P.push_hi : P.push_lo);
        if (!peer) continue;
        const int src = side ? P.push_hi_src : P.push_lo_src, cnt = side ? P.push_hi_cnt : P.push_lo_cnt;
        const int dst = side ? P.push_hi_dst : P.push_lo_dst;
        // 3D: planes [p0, p1) x all valid tile rows; 2D: the one plane x rows [r0, r1)
        int p0 = 0, p1 = 1, r0 = 0, r1 = rows_valid;
        if (P.push_dim == 2) { p0 = max(src, c.za); p1 = min(src + cnt, c.zb); }
        else { r0 = max(src - c.Y0, 0); r1 = min(src + cnt - c.Y0, rows_valid); }
        if (p1 <= p0 || r1 <= r0) continue;
        const int nrow = r1 - r0, per_plane = nrow * G::LX;
#pragma unroll 1
        for (int e = tid; e < (p1 - p0) * per_plane; e += G::NC) {
            const int pl = e / per_plane, w = e - pl * per_plane;
            const int row = r0 + w / G::LX, x = c.X0 + V * (w % G::LX);
            if (x + V <= P.xlo || x >= P.xhi) continue;
            const int y = c.Y0 + row, p = p0 + pl;
            const long long from = P.push_dim == 2 ? (long long)p * P.nxny + (long long)y * P.nx + x : (long long)y * P.nx + x;
            const long long to = P.push_dim == 2 ? (long long)(p - src + dst) * P.nxny + (long long)y * P.nx + x
                                                 : (long long)(y - src + dst) * P.nx + x;
            if (P.vec_ok && x >= P.xlo && x + V <= P.xhi) {
                *reinterpret_cast<uint4*>(peer + to) = __ldcg(reinterpret_cast<const uint4*>(own + from));    // L2: written by other warps
            } else {
#pragma unroll
                for (int v = 0; v < V; v++)
                    if (x + v >= P.xlo && x + v < P.xhi) peer[to + v] = __ldcg(own + from + v);
            }
        }
    }
}

// One CTA per SM: 12 consumer warps + the producer warp.  The register file is split over the
// four SM sub-partitions (16384 registers each) and warps are dealt round-robin, so 13 warps
// (4+3+3+3) can have the full 128 registers per thread, while 16+1 warps or 2 CTAs x (8+1) warps
// put 5 warps on one sub-partition and are limited to 96 -- which spills in every double kernel
// (measured: profiles/README.md).  8 + 1 warps (3+2+2+2) or 10 + 1 (3+3+3+2) may use 168: for the Ops that trade warps for
// per-thread reuse (tricubic with two adjacent rows per thread).
template <class Op> struct RegCap { static constexpr int value = Op::NC == 512 ? 96 : Op::NC <= 320 ? 168 : 128; };

template <class Op, bool PUSH, bool TS>
__global__ void __launch_bounds__(Geo<Op>::NTHREADS) __maxnreg__(RegCap<Op>::value)
stream_kernel(const __grid_constant__ StreamParams P, const __grid_constant__ TensorMaps M, const __grid_constant__ OutMaps OM)
{
    using G = Geo<Op>;
    constexpr int S = Op::STAGES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + S;
    uint64_t* ofull = full + 2 * MAX_STAGES;       // [OB] staging buffer written by all consumer warps
    uint64_t* oempty = ofull + G::OB;              // [OB] staging buffer read out by the TMA store
    unsigned char* stages = smem + G::HDR_BYTES;
    unsigned char* extra = stages + G::RING_BYTES;
    unsigned char* ostages = extra + G::EXTRA_BYTES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < S; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], G::NC / 32);
        }
        if constexpr (TS) {
#pragma unroll
            for (int i = 0; i < G::OB; i++) {
                mbar_init(&ofull[i], G::NC / 32);
                mbar_init(&oempty[i], 1);
            }
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == G::NC / 32) {
        // ------------------------------ producer warp ------------------------------
        if (P.use_tma) {
            if (lane != 0) goto finish;
#pragma unroll
            for (int a = 0; a < Op::NSTAGED; a++) tma_prefetch_desc(&M.m[a]);
        }
        uint32_t g = 0;
        bool ordered = false;
        for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
            const ItemCoords c = decode_item<Op, PUSH>(P, item);
            if constexpr (PUSH) {
                // An end item reads ghost planes (the neighbours' previous sweep must have pushed them) and its consumers
                // push into the neighbours' ghost planes (the neighbours must have finished reading the values they replace,
                // which they did in their previous sweep's end items): both are what the neighbours' flags >= wait_value say.
                // The consumers cannot start the item before its first stage is full, i.e. before this wait is over.
                if (c.is_end && !ordered) {
                    if (lane == 0) {
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const unsigned long long* f = P.wait_flag[i];
                            if (f) {
                                unsigned long long cur, t0, now;
                                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                                for (;;) {
                                    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(cur) : "l"(f) : "memory");
                                    if (cur >= P.wait_value) break;
                                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                                    if (now - t0 > 20000000000ull) __trap();      // a neighbour died: fail, do not hang
                                    __nanosleep(100);
                                }
                            }
                        }
                        // the ghost planes were written by the neighbours' generic-proxy stores; this lane reads them through
                        // the async proxy (TMA) next
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                    }
                    __syncwarp(P.use_tma ? 1u : 0xffffffffu);
                    ordered = true;
                }
            }
            for (int s = c.za - Op::WARM; s < c.zb; ++s, ++g) {
                const uint32_t st = g % S, ph = (g / S) & 1u;
                mbar_wait(&empty[st], ph ^ 1u);
                unsigned char* sb = stages + st * G::STAGE_BYTES;
                if (P.use_tma) {
                    uint32_t bytes = 0;
                    ProducerIssue<Op, 0>::bytes(s, c.za, bytes);
                    mbar_arrive_expect_tx(&full[st], bytes);
                    if constexpr (G::TSTAGE_BYTES != 0) ProducerIssue<Op, 0>::tma2(M, stages, st, g % (uint32_t)G::TSLOTS, &full[st], c.X0, c.Y0, s, c.za);
                    else ProducerIssue<Op, 0>::tma(M, sb, &full[st], c.X0, c.Y0, s, c.za);
                } else {
                    if constexpr (G::TSTAGE_BYTES != 0) ProducerIssue<Op, 0>::fallback2(P, stages, st, g % (uint32_t)G::TSLOTS, c.X0, c.Y0, s, c.za, lane);
                    else ProducerIssue<Op, 0>::fallback(P, sb, c.X0, c.Y0, s, c.za, lane);
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[st]);
                }
            }
        }
    } else if (warp == G::NC / 32 + 1) {
        // ------------------------------ store warp (TMA stores) ------------------------------
        if constexpr (TS) {
            if (lane != 0) goto finish;
            uint32_t og = 0;
            for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
                const ItemCoords c = decode_item<Op, PUSH>(P, item);
                for (int s = c.za - Op::WARM; s < c.zb; ++s) {
                    if (!G::emits(s, c.za, c.zb)) continue;
                    const uint32_t ob = og & 1u;
                    mbar_wait(&ofull[ob], (og >> 1) & 1u);
#pragma unroll
                    for (int q = 0; q < G::NOUT; q++) {
                        const int p = s + Op::out_dpl(q);
                        if (p >= c.za && p < c.zb)
                            tma_store_3d(c.X0 == 0 ? &OM.m0[q] : &OM.m[q], ostages + ob * G::OUT_BYTES + q * G::OUT_TILE_BYTES,
                                         c.X0 == 0 ? 0 : c.X0 - G::V, c.Y0 - P.ylo, p);
                    }
                    tma_store_commit();
                    tma_store_wait_read<1>();              // the previous group has been read out of its buffer
                    if (og >= 1u) mbar_arrive(&oempty[ob ^ 1u]);
                    ++og;
                }
            }
            tma_store_wait_all();
        }
    } else {
        // ------------------------------ consumer warps ------------------------------
        Op op(P);
        typename Op::State state;
        Ctx<Op, PUSH, TS> ctx{{}, P, stages, extra, ostages, G::TX, 0, 0u, 0, 0, 0, 0, 0, tid % G::LX, tid / G::LX, 0, 0, 0, 0, 0, 0u, 0ll};
        uint32_t st = 0, ph = 0, rel_st = (uint32_t)(S - Op::HOLD) % S;    // ring stage / parity of this step; stage to hand back
        uint32_t g = 0, og = 0;
        for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
            const ItemCoords c = decode_item<Op, PUSH>(P, item);
            ctx.begin_item(c.X0, c.Y0);
            ctx.zb = c.zb;
            int local = 0, phase = 0;
            for (int s = c.za - Op::WARM; s < c.zb; ++s, ++g, ++local) {
                ctx.st = st;
                if constexpr (G::TSTAGE_BYTES != 0) ctx.tst = g % (uint32_t)G::TSLOTS;
                ctx.s = s;
                ctx.rel = s - c.za;
                ctx.poff = (long long)s * P.nxny;
                op.pre(ctx, state);
                bool emit = false;
                if constexpr (TS) {
                    emit = G::emits(s, c.za, c.zb);
                    if (emit) {
                        // the staging buffer of two emitting steps ago must have left through the TMA store
                        mbar_wait(&oempty[og & 1u], ((og >> 1) & 1u) ^ 1u);
                        ctx.ostage = ostages + (og & 1u) * G::OUT_BYTES;
                    }
                }
                mbar_wait(&full[st], ph);
                step_dispatch<Op, Ctx<Op, PUSH, TS>>(op, ctx, state, phase);
                if constexpr (TS) {
                    if (emit) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&ofull[og & 1u]);
                        ++og;
                    }
                }
                if (local >= Op::HOLD) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[rel_st]);
                }
                if (++st == S) { st = 0; ph ^= 1u; }
                if (++rel_st == S) rel_st = 0;
                if (++phase == Op::PERIOD) phase = 0;
            }
            if constexpr (Op::HOLD > 0) {
                // hand back the stages still held at the end of the item
                const int held = local < Op::HOLD ? local : Op::HOLD;
                __syncwarp();
                if (lane == 0)
                    for (int h = held; h >= 1; h--) mbar_arrive(&empty[(g - (uint32_t)h) % S]);
            }
            if constexpr (PUSH) {
                // the last END item of the grid to complete tells the neighbours that this sweep's halo push is done and
                // that their ghost planes of the previous sweep have been read
                if (c.is_end) {
                    asm volatile("bar.sync 2, %0;" ::"n"(G::NC) : "memory");       // every consumer warp has finished the item
                    // Fused halo push: the part of the planes (3D) / rows (2D) a neighbour needs as ghosts that this item has
                    // just written goes from this GPU's memory (still in L2) straight into the neighbour's ghost planes
                    // through the peer-mapped pointer, over NVLink, while the other CTAs keep sweeping.  Done here, by all
                    // consumer threads, and not inside Op::step(): a push inside the step costs every step registers and
                    // instructions (measured, profiles/r2_push_variants.txt: two inlined step variants +20 % instructions;
                    // one variant with the push behind a branch, inlined or as a cold call, +25-45 % time for the stencils
                    // that sit at their register cap: lapgsrb, tricubic, gaussblur).
                    push_item<Op>(P, c, ctx.rows_valid, tid);
                    asm volatile("bar.sync 2, %0;" ::"n"(G::NC) : "memory");       // ... and has pushed its share
                }
                if (c.is_end && (P.signal_flag[0] || P.signal_flag[1])) {
                    if (tid == 0) {
                        __threadfence_system();
                        const unsigned int prev = atomicAdd(P.done_counter, 1u);
                        if (prev == (unsigned int)P.end_items - 1u) {
                            *P.done_counter = 0u;
                            __threadfence_system();
#pragma unroll
                            for (int i = 0; i < 2; i++)
                                if (P.signal_flag[i])
                                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(P.signal_flag[i]), "l"(P.signal_value) : "memory");
                        }
                    }
                }
            }
        }
    }
finish:
    return;
}

}  // namespace b200