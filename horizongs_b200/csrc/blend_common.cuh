// Building blocks shared by the blend kernels: mbarrier + cp.async.bulk (TMA bulk copy) staging,
// halving warp reduction, vector reductions to global memory.
#pragma once
#include "hgs_common.cuh"
#include "hgs_constants.cuh"

namespace hgs {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier (shared::cta) ------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar`. 16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- vector reductions to global memory (sm_90+) ----------------------------------------------------
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// ---- halving warp reduction --------------------------------------------------------------------------
// Sums N per-lane values over the 32 lanes with ~N + log2(32) shuffles instead of 5*N: at every butterfly
// step lanes pair up, each keeps one half of the value list and sends the other half.  Afterwards every
// lane holds the full 32-lane sum of ONE component; halving_component() tells which.
template <int N, int M>
struct Halving {
    static __device__ __forceinline__ float run(float (&v)[N], int lane) {
        constexpr int H = N / 2, R = N - 2 * H;
        float nv[H + R];
        const bool up = (lane & M) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const float keep = up ? v[2 * i + 1] : v[2 * i];
            const float send = up ? v[2 * i] : v[2 * i + 1];
            nv[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, M);
        }
        if (R) nv[H] = v[N - 1] + __shfl_xor_sync(0xFFFFFFFFu, v[N - 1], M);
        return Halving<H + R, M / 2>::run(nv, lane);
    }
    static __device__ __forceinline__ int comp(const int (&id)[N], int lane) {
        constexpr int H = N / 2, R = N - 2 * H;
        int nid[H + R];
        const bool up = (lane & M) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) nid[i] = up ? id[2 * i + 1] : id[2 * i];
        if (R) nid[H] = id[N - 1];
        return Halving<H + R, M / 2>::comp(nid, lane);
    }
};
template <int N>
struct Halving<N, 0> {
    static __device__ __forceinline__ float run(float (&v)[N], int) {
        static_assert(N == 1, "32 lanes reduce at most 32 components");
        return v[0];
    }
    static __device__ __forceinline__ int comp(const int (&id)[N], int) { return id[0]; }
};
// The same butterfly with the lane's side of every level given as a bit mask (all ones: the lane has bit M set):
// "up ? a : b" becomes one three-input LOP3 on a mask register instead of a predicate that has to be rebuilt from
// the lane id inside the loop (the blend kernels have no predicate registers to spare across their inner loop).
__device__ __forceinline__ float mask_select(float a, float b, unsigned up_mask) {
    return __int_as_float((__float_as_int(a) & up_mask) | (__float_as_int(b) & ~up_mask));
}
struct LaneMasks {
    unsigned m16, m8, m4, m2;
    __device__ __forceinline__ explicit LaneMasks(int lane)
        : m16((lane & 16) ? ~0u : 0u), m8((lane & 8) ? ~0u : 0u), m4((lane & 4) ? ~0u : 0u), m2((lane & 2) ? ~0u : 0u) {
        // opaque to the compiler, which would otherwise turn every mask back into "lane & M" predicates
        asm volatile("" : "+r"(m16), "+r"(m8), "+r"(m4), "+r"(m2));
    }
    template <int M>
    __device__ __forceinline__ unsigned get() const {
        return M == 16 ? m16 : (M == 8 ? m8 : (M == 4 ? m4 : m2));
    }
};
template <int N, int M>
struct HalvingMasked {
    static __device__ __forceinline__ float run(float (&v)[N], const LaneMasks& lm) {
        constexpr int H = N / 2, R = N - 2 * H;
        float nv[H + R];
        const unsigned up = lm.get<M>();
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const float keep = mask_select(v[2 * i + 1], v[2 * i], up);
            const float send = mask_select(v[2 * i], v[2 * i + 1], up);
            nv[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, M);
        }
        if (R) nv[H] = v[N - 1] + __shfl_xor_sync(0xFFFFFFFFu, v[N - 1], M);
        return HalvingMasked<H + R, M / 2>::run(nv, lm);
    }
};
template <int N>
struct HalvingMasked<N, 0> {
    static __device__ __forceinline__ float run(float (&v)[N], const LaneMasks&) {
        static_assert(N == 1, "32 lanes reduce at most 32 components");
        return v[0];
    }
};
template <int M>
struct HalvingMasked<1, M> {   // one component left: plain butterfly, no selection
    static __device__ __forceinline__ float run(float (&v)[1], const LaneMasks& lm) {
        float nv[1] = {v[0] + __shfl_xor_sync(0xFFFFFFFFu, v[0], M)};
        return HalvingMasked<1, M / 2>::run(nv, lm);
    }
};
template <>
struct HalvingMasked<1, 0> {
    static __device__ __forceinline__ float run(float (&v)[1], const LaneMasks&) { return v[0]; }
};
template <int N>
__device__ __forceinline__ float halving_reduce(float (&v)[N], int lane) {
    return Halving<N, 16>::run(v, lane);
}
template <int N>
__device__ __forceinline__ int halving_component(int lane) {
    int id[N];
#pragma unroll
    for (int i = 0; i < N; ++i) id[i] = i;
    return Halving<N, 16>::comp(id, lane);
}


// ---- fast approximate math (MUFU) --------------------------------------------------------------------
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}


// ---- packed per-Gaussian record of the fast 3DGS blend kernels -------------------------------------------
// centre, conic pre-multiplied by -log2(e)/2 so that alpha = o * 2^(A dx^2 + B dx dy + C dy^2) is one MUFU.EX2,
// opacity, cut-off exponent, colour (+ depth), edge-optimum slopes of the per-warp cull test.  Written by
// hgs_blend3d_pack, or directly by the fused projection kernel (project3d.cu).
constexpr float LOG2E = 1.4426950408889634f;
constexpr int REC_BYTES = 64;
struct __align__(16) GRec {
    float x, y, A, B;          // q0
    float C, opac, p2min, _p;  // q1
    float col[4];              // q2
    float kx, ky, _p2, _p3;    // q3
};
static_assert(sizeof(GRec) == REC_BYTES, "record size");
__device__ __forceinline__ GRec make_grec(float mx, float my, float a, float b, float c, float o, const float (&col)[4]) {
    GRec r;
    r.x = mx; r.y = my;
    r.A = -0.5f * LOG2E * a;
    r.B = -LOG2E * b;
    r.C = -0.5f * LOG2E * c;
    r.opac = o;
    // contributes iff o * 2^p2 >= 1/255  <=>  p2 >= -log2(255 o); o <= 0 never contributes
    r.p2min = (o > 0.f) ? -log2f(255.0f * o) : 1.0f;
    r._p = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.col[k] = col[k];
    r.kx = (r.C != 0.f) ? -r.B / (2.f * r.C) : 0.f;
    r.ky = (r.A != 0.f) ? -r.B / (2.f * r.A) : 0.f;
    r._p2 = 0.f; r._p3 = 0.f;
    return r;
}

// ---- tile / sub-tile geometry of the fast blend kernels ----------------------------------------------
// One CTA of 256 threads per 16x16 tile; warp w owns the 8x4 pixel sub-tile ((w & 1) * 8, (w >> 1) * 4).
struct TileGeom {
    int cam, gtile, pi, pj, warp, lane;
    float px, py, X0, X1, Y0, Y1;
    bool inside;
};
__device__ __forceinline__ TileGeom tile_geom(int tile_w, int tile_h, int W, int H) {
    TileGeom g;
    g.cam = blockIdx.z;
    g.gtile = (g.cam * tile_h + blockIdx.y) * tile_w + blockIdx.x;
    g.warp = threadIdx.x >> 5;
    g.lane = threadIdx.x & 31;
    const int sx = blockIdx.x * HGS_TILE_SIZE + (g.warp & 1) * 8, sy = blockIdx.y * HGS_TILE_SIZE + (g.warp >> 1) * 4;
    g.pj = sx + (g.lane & 7);
    g.pi = sy + (g.lane >> 3);
    g.px = (float)g.pj + 0.5f;
    g.py = (float)g.pi + 0.5f;
    g.X0 = (float)sx + 0.5f; g.X1 = (float)sx + 7.5f;
    g.Y0 = (float)sy + 0.5f; g.Y1 = (float)sy + 3.5f;
    g.inside = (g.pi < H && g.pj < W);
    return g;
}


}  // namespace hgs
