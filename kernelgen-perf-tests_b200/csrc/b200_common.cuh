// b200_common.cuh -- PTX wrappers (mbarrier, TMA) and small vector helpers for sm_100a.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

#define B200_DEV __device__ __forceinline__

B200_DEV uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------
B200_DEV void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
B200_DEV void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
B200_DEV void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
B200_DEV void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
B200_DEV bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
B200_DEV void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}

// ---- TMA (cp.async.bulk.tensor), tile mode, 3D ------------------------------------
B200_DEV void tma_prefetch_desc(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
B200_DEV void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)),
          "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// shared -> global tile store (bulk async-group completion); out-of-range parts of the box are not written
B200_DEV void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
B200_DEV void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> B200_DEV void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
B200_DEV void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (TMA) before it is signalled
B200_DEV void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// L2 cache-policy descriptors for the .L2::cache_hint operand (same encodings CUTLASS uses)
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t L2_EVICT_FIRST  = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST   = 0x14F0000000000000ull;

B200_DEV void tma_load_3d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                               uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)),
          "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}

// ---- vector types -----------------------------------------------------------------
template <typename T> struct Vec;            // 16-byte vector of T
template <> struct Vec<float>  { using type = float4;  static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };

// Load N consecutive elements (N*sizeof(T) in {4,8,16} bytes) from an address aligned to that size.
template <int N> B200_DEV void ld_chunk(const float* p, float* out)
{
    if constexpr (N == 4) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else if constexpr (N == 2) {
        const float2 v = *reinterpret_cast<const float2*>(p);
        out[0] = v.x; out[1] = v.y;
    } else {
        static_assert(N == 1, "unsupported chunk");
        out[0] = p[0];
    }
}
template <int N> B200_DEV void ld_chunk(const double* p, double* out)
{
    if constexpr (N == 2) {
        const double2 v = *reinterpret_cast<const double2*>(p);
        out[0] = v.x; out[1] = v.y;
    } else {
        static_assert(N == 1, "unsupported chunk");
        out[0] = p[0];
    }
}

// 16-byte register block of T that can be loaded/stored as one vector.
template <typename T> struct alignas(16) VReg {
    static constexpr int V = 16 / sizeof(T);
    T v[V];
    B200_DEV T& operator[](int i) { return v[i]; }
    B200_DEV const T& operator[](int i) const { return v[i]; }
};

template <typename T> B200_DEV VReg<T> ldv(const T* p)          // p 16-byte aligned (shared or global)
{
    VReg<T> r;
    *reinterpret_cast<uint4*>(r.v) = *reinterpret_cast<const uint4*>(p);
    return r;
}

// Streaming (evict-first) 16-byte global load for data that is read exactly once.
template <typename T> B200_DEV VReg<T> ldv_stream(const T* p)
{
    VReg<T> r;
    uint4 u;
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
    *reinterpret_cast<uint4*>(r.v) = u;
    return r;
}

// x-window around a 16-byte aligned position p: elements p[-L .. V+R), widest aligned loads.
// L, R in {0,1,2}.  w[0] = p[-L].
template <int L, int R, typename T> struct Window {
    static constexpr int V = 16 / sizeof(T);
    T w[L + V + R];
    B200_DEV void load(const T* p)
    {
        if constexpr (L == 1) w[0] = p[-1];
        if constexpr (L == 2) ld_chunk<2>(p - 2, &w[0]);
        // centre: one 16-byte load
        {
            VReg<T> c = ldv(p);
#pragma unroll
            for (int i = 0; i < V; i++) w[L + i] = c[i];
        }
        if constexpr (R == 1) w[L + V] = p[V];
        if constexpr (R == 2) ld_chunk<2>(p + V, &w[L + V]);
    }
    // value at x offset d (relative to output element v): d in [-L, R]
    B200_DEV T at(int v, int d) const { return w[L + v + d]; }
};

}  // namespace b200
