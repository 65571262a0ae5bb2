// b200_launch.cuh -- host-side launch planning for the tile-streaming engine.
//
// Replaces kernelgen_cuda_configure_gird (<test>/cuda/cuda_profiling.cu:36-111 of the
// reference): instead of "128x1x1 blocks, one thread per point" the plan is
//   x tiles of TX, y tiles of TY over the interior, z cut into nzc chunks,
//   grid = min(items, SMs * resident CTAs per SM)  (persistent CTAs),
// with nzc chosen so that the items fill whole rounds of the grid (tail effect) while the
// z-chunks stay long enough to amortise the WARM warm-up planes of the z-march.
#pragma once

#include <mutex>

#include "b200_internal.h"
#include "b200_stream.cuh"

namespace b200 {

template <class Op, bool PUSH, bool TS> struct KernelSetup {
    bool done[16] = {};
    int blocks_per_sm[16] = {};
    std::mutex mu;
    int get(int device, int* bps)
    {
        std::lock_guard<std::mutex> lk(mu);
        if (device < 0 || device >= 16) { set_error("device index %d out of range", device); return B200_ERR_ARG; }
        if (!done[device]) {
            B200_CUDA(cudaFuncSetAttribute(stream_kernel<Op, PUSH, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Geo<Op>::SMEM_BYTES));
            int n = 0;
            B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stream_kernel<Op, PUSH, TS>, Geo<Op>::NTHREADS,
                                                                    Geo<Op>::SMEM_BYTES));
            if (n < 1) { set_error("kernel does not fit on an SM"); return B200_ERR_CUDA; }
            blocks_per_sm[device] = n;
            done[device] = true;
        }
        *bps = blocks_per_sm[device];
        return B200_OK;
    }
};

template <class Op, bool PUSH, bool TS> KernelSetup<Op, PUSH, TS>& kernel_setup()
{
    static KernelSetup<Op, PUSH, TS> s;
    return s;
}

// Pick the number of z-chunks: maximise (grid fill of the last round) x (1 - warm-up overhead).
inline void plan_zchunks(int tiles_xy, int nz, int warm, int grid_cap, int* nzc_out, int* len_out)
{
    int best = 1;
    double best_score = -1.0;
    const int max_chunks = nz;      // chunk length >= 1
    for (int nzc = 1; nzc <= max_chunks; nzc++) {
        const int len = (nz + nzc - 1) / nzc;
        const int real_nzc = (nz + len - 1) / len;
        if (real_nzc != nzc) continue;
        const long long items = (long long)tiles_xy * nzc;
        const long long rounds = (items + grid_cap - 1) / grid_cap;
        const double fill = (double)items / (double)(rounds * grid_cap);
        const double work = (double)len / (double)(len + warm);
        const double score = fill * work;
        if (score > best_score + 1e-9) { best_score = score; best = nzc; }
        if (len <= 4) break;
    }
    *nzc_out = best;
    *len_out = (nz + best - 1) / best;
}

template <class Op, bool PUSH, bool TS>
int launch_variant(const StreamParams& P, const TensorMaps& M, const OutMaps& OM, int device, int num_sms, cudaStream_t stream,
                   int tiles_xy, int nz, StreamParams* planned)
{
    using G = Geo<Op>;
    int bps = 0;
    if (int rc = kernel_setup<Op, PUSH, TS>().get(device, &bps)) return rc;
    StreamParams Q = P;
    const int grid_cap = num_sms * bps;
    plan_zchunks(tiles_xy, nz, Op::WARM, grid_cap, &Q.nzc, &Q.zc_len);
    const long long items = (long long)tiles_xy * Q.nzc;
    if (items > 0x7fffffffLL) { set_error("grid too large"); return B200_ERR_ARG; }
    Q.nitems = (int)items;
    if constexpr (PUSH) {
        // Which units of the split dimension (z-chunks / tile rows) take part in the neighbour ordering: those that read a
        // ghost plane of a staged input or store a plane that is pushed to a neighbour.  They are walked first (decode_item).
        int zlo = 0, zhi = 0, ylo_h = 0, yhi_h = 0;
        for (int a = 0; a < Op::NSTAGED; a++) {
            zlo = Op::spec(a).zlo > zlo ? Op::spec(a).zlo : zlo;
            zhi = Op::spec(a).lead > zhi ? Op::spec(a).lead : zhi;
            ylo_h = Op::spec(a).ylo > ylo_h ? Op::spec(a).ylo : ylo_h;
            yhi_h = Op::spec(a).yhi > yhi_h ? Op::spec(a).yhi : yhi_h;
        }
        const bool planes = Q.push_dim == 2;
        const int n = planes ? Q.ns : Q.ny;                                  // local extent of the split dimension
        const int o0 = planes ? Q.z0 : Q.ylo, o1 = planes ? Q.z1 : Q.yhi;    // output range
        const int len = planes ? Q.zc_len : tile_py<Op>();
        const int rlo = planes ? zlo : ylo_h, rhi = planes ? zhi : yhi_h;    // reach below / above an output unit
        Q.split_n = planes ? Q.nzc : Q.nty;
        Q.split_stride = planes ? tiles_xy : Q.ntx;
        // ghost units: everything outside the owned output range on a side that has a neighbour
        const int glo_end = Q.wait_flag[0] || Q.push_lo ? o0 : 0;            // ghost planes below: [0, o0)
        const int ghi_beg = Q.wait_flag[1] || Q.push_hi ? o1 : n;            // ghost planes above: [o1, n)
        auto is_end_unit = [&](int u) {
            const int a = o0 + u * len, b = (a + len < o1) ? a + len : o1;   // output range of the unit
            if (a - rlo < glo_end || b + rhi > ghi_beg) return true;          // reads a ghost plane
            if (Q.push_lo && a < Q.push_lo_src + Q.push_lo_cnt && b > Q.push_lo_src) return true;
            if (Q.push_hi && a < Q.push_hi_src + Q.push_hi_cnt && b > Q.push_hi_src) return true;
            return false;
        };
        Q.e_lo = 0;
        while (Q.e_lo < Q.split_n && is_end_unit(Q.e_lo)) Q.e_lo++;
        Q.e_hi = 0;
        while (Q.e_lo + Q.e_hi < Q.split_n && is_end_unit(Q.split_n - 1 - Q.e_hi)) Q.e_hi++;
        for (int u = Q.e_lo; u < Q.split_n - Q.e_hi; u++)
            if (is_end_unit(u)) { Q.e_lo = Q.split_n; Q.e_hi = 0; break; }    // cannot happen for a contiguous slab: order everything
        Q.end_items = (Q.e_lo + Q.e_hi) * Q.split_stride;
    }
    const int grid = (int)(items < grid_cap ? items : grid_cap);
    stream_kernel<Op, PUSH, TS><<<grid, G::NTHREADS, G::SMEM_BYTES, stream>>>(Q, M, OM);
    B200_CUDA(cudaGetLastError());
    count_launch();
    (void)planned;
    return B200_OK;
}

template <class Op> int launch_stream(const HostArgs& a)
{
    using T = typename Op::real;
    using G = Geo<Op>;
    const b200_sweep_desc& d = *a.desc;
    const b200_test_info* ti = b200_get_test_info(d.test);

    const bool push = d.push_lo || d.push_hi;

    StreamParams P{};
    TensorMaps M{};
    OutMaps OM{};
    P.nx = d.nx; P.ny = d.ny; P.ns = (ti->ndims == 3) ? d.ns : 1;
    P.nxny = (long long)d.nx * d.ny;
    P.xlo = ti->lo[0]; P.xhi = P.nx - ti->hi[0];
    P.ylo = ti->lo[1]; P.yhi = P.ny - ti->hi[1];
    if (ti->ndims == 3) { P.z0 = ti->lo[2]; P.z1 = P.ns - ti->hi[2]; }
    else                { P.z0 = 0; P.z1 = 1; }
    if (d.out_begin != 0 || d.out_end != 0) {
        int& lo = (ti->ndims == 3) ? P.z0 : P.ylo;
        int& hi = (ti->ndims == 3) ? P.z1 : P.yhi;
        // a slab may start output wherever the stencil's reach (= ghost depth) stays inside the local array;
        // for the whole grid that is the interior, except tricubic2 whose interior is one plane narrower
        const int n_split = (ti->ndims == 3) ? P.ns : P.ny;
        const int rlo = ti->zghost_lo, rhi = n_split - ti->zghost_hi;
        if (d.out_begin < rlo || d.out_end > rhi || d.out_begin > d.out_end) {
            set_error("%s: output range [%d,%d) reaches outside the local array, allowed [%d,%d)", ti->name,
                      d.out_begin, d.out_end, rlo, rhi);
            return B200_ERR_ARG;
        }
        lo = d.out_begin; hi = d.out_end;
    }
    if (P.xhi <= P.xlo || P.yhi <= P.ylo || P.z1 <= P.z0) return B200_OK;   // empty interior: nothing to update
    // in-plane element offsets are 32-bit in the kernels (Ctx::idx0, row * nx): refuse planes they cannot address
    if (P.nxny >= 0x7fffffffLL) { set_error("%s: plane of %lld elements (nx*ny) exceeds the 2^31-1 the kernels index", ti->name, P.nxny); return B200_ERR_ARG; }

    const size_t pitch = (size_t)P.nx * sizeof(T);
    bool aligned = (pitch % 16) == 0;
    for (int q = 0; q < ti->narrays + Op::EXTRA_ARRAYS; q++) {
        if (!a.arrays[q]) { set_error("%s: array slot %d is NULL", ti->name, q); return B200_ERR_ARG; }
        P.arr[q] = a.arrays[q];
        if (((uintptr_t)a.arrays[q]) % 16) aligned = false;
    }
    for (int q = 0; q < B200_MAX_SCALARS; q++) P.sc[q] = d.scalars[q];
    static const bool no_tma = getenv("B200_NO_TMA") != nullptr;
    P.use_tma = aligned && !no_tma;
    P.vec_ok = aligned;

    static const int env_serp = getenv("B200_SERPENTINE") ? atoi(getenv("B200_SERPENTINE")) : 1;
    P.reverse = env_serp ? (d.reverse_order & 1) : 0;
    P.push_slot = -1;
    P.push_dim = ti->ndims == 3 ? 2 : 1;
    if (d.push_lo || d.push_hi) {
        if (ti->exchange_slot < 0) { set_error("%s has no exchanged array", ti->name); return B200_ERR_ARG; }
        // the array being written this sweep is the last slot of the rotation (slot 1 or 2)
        P.push_slot = (ti->rotation == 3) ? 2 : 1;
        P.push_lo = d.push_lo; P.push_lo_src = d.push_lo_src_plane; P.push_lo_dst = d.push_lo_dst_plane; P.push_lo_cnt = d.push_lo_count;
        P.push_hi = d.push_hi; P.push_hi_src = d.push_hi_src_plane; P.push_hi_dst = d.push_hi_dst_plane; P.push_hi_cnt = d.push_hi_count;
        if (((uintptr_t)d.push_lo) % 16 || ((uintptr_t)d.push_hi) % 16) P.vec_ok = 0;
        for (int i = 0; i < 2; i++) {
            P.wait_flag[i] = (const unsigned long long*)d.wait_flag[i];
            P.signal_flag[i] = (unsigned long long*)d.signal_flag[i];
        }
        P.wait_value = d.wait_value;
        P.signal_value = d.signal_value;
        if (int rc = get_done_counter(a.device, &P.done_counter)) return rc;
    }

    P.ntx = (P.nx + tile_px<Op>() - 1) / tile_px<Op>();
    P.nty = (P.yhi - P.ylo + tile_py<Op>() - 1) / tile_py<Op>();

    if (P.use_tma) {
        for (int s = 0; s < Op::NSTAGED; s++) {
            TmaBoxKey key{P.arr[Op::spec(s).slot], P.nx, P.ny, P.ns, (int)sizeof(T), G::bw(s), G::bh(s), 0, 0};
            if (int rc = get_tensor_map(key, &M.m[s])) return rc;
        }
    }
    // TMA-store output path: tensor maps over the vector-reachable interior of every listed output
    static const bool no_ts = getenv("B200_NO_TMA_STORE") != nullptr;
    bool ts = G::TS && P.use_tma && !no_ts && P.nx > 2 * G::V;
    if (ts) {
        for (int q = 0; q < G::NOUT; q++) {
            T* base = reinterpret_cast<T*>(P.arr[Op::out_slot(q)]) + (size_t)P.ylo * P.nx + G::V;
            TmaBoxKey key{base, P.nx - 2 * G::V, P.yhi - P.ylo, P.ns, (int)sizeof(T), Op::TX, Op::TY, P.nx, P.ny};
            if (int rc = get_tensor_map(key, &OM.m[q])) return rc;
            TmaBoxKey key0{base, P.nx - 2 * G::V, P.yhi - P.ylo, P.ns, (int)sizeof(T), Op::TX - G::V, Op::TY, P.nx, P.ny};
            if (int rc = get_tensor_map(key0, &OM.m0[q])) return rc;
        }
    }

    const int tiles_xy = P.ntx * P.nty, nz = P.z1 - P.z0;
    if constexpr (G::TS) {
        if (ts) {
            if (push) return launch_variant<Op, true, true>(P, M, OM, a.device, a.num_sms, a.stream, tiles_xy, nz, nullptr);
            return launch_variant<Op, false, true>(P, M, OM, a.device, a.num_sms, a.stream, tiles_xy, nz, nullptr);
        }
    }
    if (push) return launch_variant<Op, true, false>(P, M, OM, a.device, a.num_sms, a.stream, tiles_xy, nz, nullptr);
    return launch_variant<Op, false, false>(P, M, OM, a.device, a.num_sms, a.stream, tiles_xy, nz, nullptr);
}

template <class Op> int info_stream(KernelInfo* ki, const char* name)
{
    cudaFuncAttributes fa;
    constexpr bool TS = Geo<Op>::TS;
    B200_CUDA(cudaFuncGetAttributes(&fa, stream_kernel<Op, false, TS>));
    ki->regs = fa.numRegs;
    ki->smem_bytes = Geo<Op>::SMEM_BYTES;
    ki->name = name;
    int dev = 0, bps = 0;
    B200_CUDA(cudaGetDevice(&dev));
    if (int rc = kernel_setup<Op, false, TS>().get(dev, &bps)) return rc;
    ki->blocks_per_sm = bps;
    int bps_push = 0;      // also prepares the halo-pushing variant (used by multi-GPU launches)
    if (int rc = kernel_setup<Op, true, TS>().get(dev, &bps_push)) return rc;
    return B200_OK;
}

// ---- small-grid tile policy (float forms of the 3D Ops) --------------------------------------------------
// The decomposition's own ceiling, by the planner's rule (tools/plan_model.py is the same arithmetic in Python):
//   fill  = items / (rounds x CTAs), zwork = len / (len + WARM), yfill = interior rows / (y-tiles x TY).
// With few tiles (512 x 256 x 256: 4 x 6 tiles of 128 x 48) z must be cut into many short chunks, each paying its
// warm-up planes; half-height tiles double the tile count.  Measured for float laplacian (experiment build,
// profiles/r1y_laplacian_float_tile_height.txt): 0.68 -> 0.77 of the roofline at 512 x 256 x 256, 0.76 -> 0.75 at
// 1024 x 1024 x 512 -- so the small form is taken only when the model promises more than SMALL_TILE_GAIN.
// B200_TILE_POLICY = 1 (default; validated on the GPU in round 2: tests/test_gpu_parity.py::test_small_tile_forms_float,
// byte-identical outputs, profiles/r2a_tile_policy.txt: laplacian float 0.689 -> 0.737, lapgsrb float 0.589 -> 0.611 at
// 512 x 256 x 256, nothing lost at 1024 x 1024 x 512) lets the model choose, 0 always takes the default form, 2 always
// takes the small form (parity tests of that form).
template <class Op> double decomposition_score(const b200_test_info* ti, int nx, int ny, int ns, int grid_cap)
{
    const int ylen = ny - ti->lo[1] - ti->hi[1], nz = ns - ti->lo[2] - ti->hi[2];
    if (ylen <= 0 || nz <= 0 || nx <= 0 || grid_cap <= 0) return 0.0;
    const int ntx = (nx + tile_px<Op>() - 1) / tile_px<Op>(), nty = (ylen + tile_py<Op>() - 1) / tile_py<Op>();
    int nzc = 1, len = nz;
    plan_zchunks(ntx * nty, nz, Op::WARM, grid_cap, &nzc, &len);
    const long long items = (long long)ntx * nty * nzc, rounds = (items + grid_cap - 1) / grid_cap;
    const double fill = (double)items / (double)(rounds * grid_cap);
    const double zwork = (double)len / (double)(len + Op::WARM);
    const double yfill = (double)ylen / (double)(nty * tile_py<Op>());
    return fill * zwork * yfill;
}
constexpr double SMALL_TILE_GAIN = 1.05;
inline int tile_policy()
{
    static const int p = getenv("B200_TILE_POLICY") ? atoi(getenv("B200_TILE_POLICY")) : 1;
    return p;
}
template <class Big, class Small> int launch_by_tile_policy(const HostArgs& a)
{
    const int pol = tile_policy();
    if (pol == 2) return launch_stream<Small>(a);
    if (pol == 1) {
        const b200_sweep_desc& d = *a.desc;
        const b200_test_info* ti = b200_get_test_info(d.test);
        if (ti->ndims == 3 &&          // whole grids and z-slabs alike (the extents are those of the local array)
            decomposition_score<Small>(ti, d.nx, d.ny, d.ns, a.num_sms) >
                SMALL_TILE_GAIN * decomposition_score<Big>(ti, d.nx, d.ny, d.ns, a.num_sms))
            return launch_stream<Small>(a);
    }
    return launch_stream<Big>(a);
}
// a 3D Op whose float form has a half-height variant (template <typename T, int TYF>)
#define B200_DEFINE_OP_TILED(name, OpT, TYF_SMALL)                                              \
    int launch_##name(int dtype, const HostArgs& a)                                            \
    {                                                                                          \
        return dtype == B200_F32 ? launch_by_tile_policy<OpT<float>, OpT<float, TYF_SMALL>>(a) \
                                 : launch_stream<OpT<double>>(a);                              \
    }                                                                                          \
    int info_##name(int dtype, KernelInfo* ki)                                                 \
    {                                                                                          \
        return dtype == B200_F32 ? info_stream<OpT<float>>(ki, #name) : info_stream<OpT<double>>(ki, #name); \
    }

#define B200_DEFINE_OP(name, OpT)                                                              \
    int launch_##name(int dtype, const HostArgs& a)                                            \
    {                                                                                          \
        return dtype == B200_F32 ? launch_stream<OpT<float>>(a) : launch_stream<OpT<double>>(a); \
    }                                                                                          \
    int info_##name(int dtype, KernelInfo* ki)                                                 \
    {                                                                                          \
        return dtype == B200_F32 ? info_stream<OpT<float>>(ki, #name) : info_stream<OpT<double>>(ki, #name); \
    }

}  // namespace b200
