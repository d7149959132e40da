// Stage a7: view-dependent colour from spherical harmonics (degree <= 4) and its backward.
//
// Replaces gsplat's spherical_harmonics as reached inside gsplat.rasterization* when
// sh_degree is not None (reference render.py:51,73; SH2 configs).  Basis = Sloan's polynomial
// forms, equal to utils/sh_utils.py:57-112 of the reference on unit vectors; coefficient layout
// [N,K,3] (scene/basic_model.py:369,378).  Restated in oracle/gsplat_oracle.py::spherical_harmonics.
// Fused here: direction = means - campos, normalisation, optional `clamp_min(c + 0.5, 0)`.
//
// One thread per Gaussian, cameras looped in-thread (coefficient gradients are summed over views
// without atomics); only Gaussians that survive culling touch their coefficient rows.
// Roofline: HBM; fwd 12 B per Gaussian + (12 K + 12) B per visible one; bwd 12 K B (zero fill) per Gaussian
// + (24 K + 36) B per visible one.
#include "hgs_common.cuh"
#include "hgs_constants.cuh"
#include "../../include/hgs_raster.h"

namespace {

constexpr int SB = 128;  // threads (= Gaussians) per block

template <int DEG>
struct ShBasis {
    static constexpr int NB = (DEG + 1) * (DEG + 1);
    float b[NB];
};

// basis values (and optionally d/dx, d/dy, d/dz treating x,y,z as free variables)
template <int DEG, bool GRAD>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float* b, float* bx, float* by, float* bz) {
    b[0] = 0.2820947917738781f;
    if (GRAD) { bx[0] = by[0] = bz[0] = 0.f; }
    if (DEG < 1) return;
    b[1] = -0.48860251190292f * y;
    b[2] = 0.48860251190292f * z;
    b[3] = -0.48860251190292f * x;
    if (GRAD) {
        bx[1] = 0.f; by[1] = -0.48860251190292f; bz[1] = 0.f;
        bx[2] = 0.f; by[2] = 0.f; bz[2] = 0.48860251190292f;
        bx[3] = -0.48860251190292f; by[3] = 0.f; bz[3] = 0.f;
    }
    if (DEG < 2) return;
    const float z2 = z * z;
    const float fTmp0B = -1.092548430592079f * z;
    const float fC1 = x * x - y * y;
    const float fS1 = 2.0f * x * y;
    const float pSH6 = 0.9461746957575601f * z2 - 0.3153915652525201f;
    b[4] = 0.5462742152960395f * fS1;
    b[5] = fTmp0B * y;
    b[6] = pSH6;
    b[7] = fTmp0B * x;
    b[8] = 0.5462742152960395f * fC1;
    const float dpSH6 = 2.f * 0.9461746957575601f * z;
    if (GRAD) {
        bx[4] = 0.5462742152960395f * 2.f * y; by[4] = 0.5462742152960395f * 2.f * x; bz[4] = 0.f;
        bx[5] = 0.f; by[5] = fTmp0B; bz[5] = -1.092548430592079f * y;
        bx[6] = 0.f; by[6] = 0.f; bz[6] = dpSH6;
        bx[7] = fTmp0B; by[7] = 0.f; bz[7] = -1.092548430592079f * x;
        bx[8] = 0.5462742152960395f * 2.f * x; by[8] = -0.5462742152960395f * 2.f * y; bz[8] = 0.f;
    }
    if (DEG < 3) return;
    const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    const float fTmp1B = 1.445305721320277f * z;
    const float fC2 = x * fC1 - y * fS1;
    const float fS2 = x * fS1 + y * fC1;
    const float pSH12 = z * (1.865881662950577f * z2 - 1.119528997770346f);
    b[9] = -0.5900435899266435f * fS2;
    b[10] = fTmp1B * fS1;
    b[11] = fTmp0C * y;
    b[12] = pSH12;
    b[13] = fTmp0C * x;
    b[14] = fTmp1B * fC1;
    b[15] = -0.5900435899266435f * fC2;
    const float dTmp0C = -2.f * 2.285228997322329f * z;
    const float dpSH12 = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
    if (GRAD) {
        bx[9] = -0.5900435899266435f * 3.f * fS1; by[9] = -0.5900435899266435f * 3.f * fC1; bz[9] = 0.f;
        bx[10] = fTmp1B * 2.f * y; by[10] = fTmp1B * 2.f * x; bz[10] = 1.445305721320277f * fS1;
        bx[11] = 0.f; by[11] = fTmp0C; bz[11] = dTmp0C * y;
        bx[12] = 0.f; by[12] = 0.f; bz[12] = dpSH12;
        bx[13] = fTmp0C; by[13] = 0.f; bz[13] = dTmp0C * x;
        bx[14] = fTmp1B * 2.f * x; by[14] = -fTmp1B * 2.f * y; bz[14] = 1.445305721320277f * fC1;
        bx[15] = -0.5900435899266435f * 3.f * fC1; by[15] = 0.5900435899266435f * 3.f * fS1; bz[15] = 0.f;
    }
    if (DEG < 4) return;
    const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
    const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
    const float fTmp2B = -1.770130769779931f * z;
    const float fC3 = x * fC2 - y * fS2;
    const float fS3 = x * fS2 + y * fC2;
    const float pSH20 = 1.984313483298443f * z * pSH12 + -1.006230589874905f * pSH6;
    b[16] = 0.6258357354491763f * fS3;
    b[17] = fTmp2B * fS2;
    b[18] = fTmp1C * fS1;
    b[19] = fTmp0D * y;
    b[20] = pSH20;
    b[21] = fTmp0D * x;
    b[22] = fTmp1C * fC1;
    b[23] = fTmp2B * fC2;
    b[24] = 0.6258357354491763f * fC3;
    if (GRAD) {
        const float dTmp0D = -3.f * 4.683325804901025f * z2 + 2.007139630671868f;
        const float dTmp1C = 2.f * 3.31161143515146f * z;
        bx[16] = 0.6258357354491763f * 4.f * fS2; by[16] = 0.6258357354491763f * 4.f * fC2; bz[16] = 0.f;
        bx[17] = fTmp2B * 3.f * fS1; by[17] = fTmp2B * 3.f * fC1; bz[17] = -1.770130769779931f * fS2;
        bx[18] = fTmp1C * 2.f * y; by[18] = fTmp1C * 2.f * x; bz[18] = dTmp1C * fS1;
        bx[19] = 0.f; by[19] = fTmp0D; bz[19] = dTmp0D * y;
        bx[20] = 0.f; by[20] = 0.f;
        bz[20] = 1.984313483298443f * (pSH12 + z * dpSH12) - 1.006230589874905f * dpSH6;
        bx[21] = fTmp0D; by[21] = 0.f; bz[21] = dTmp0D * x;
        bx[22] = fTmp1C * 2.f * x; by[22] = -fTmp1C * 2.f * y; bz[22] = dTmp1C * fC1;
        bx[23] = fTmp2B * 3.f * fC1; by[23] = -fTmp2B * 3.f * fS1; bz[23] = -1.770130769779931f * fC2;
        bx[24] = 0.6258357354491763f * 4.f * fC2; by[24] = -0.6258357354491763f * 4.f * fS2; bz[24] = 0.f;
    }
}

// Direction of Gaussian n seen from camera c (un-normalised)
__device__ __forceinline__ void load_dir(const float* __restrict__ dirs, const float* __restrict__ means,
                                         const float* __restrict__ campos, long long idx, long long n, int c, float& x,
                                         float& y, float& z) {
    if (dirs != nullptr) {
        x = dirs[idx * 3]; y = dirs[idx * 3 + 1]; z = dirs[idx * 3 + 2];
    } else {
        x = means[n * 3] - campos[c * 3];
        y = means[n * 3 + 1] - campos[c * 3 + 1];
        z = means[n * 3 + 2] - campos[c * 3 + 2];
    }
}

// colour of one (camera, Gaussian) pair
template <int DEG>
__device__ __forceinline__ void sh_eval_one(float x, float y, float z, const float* __restrict__ co, int post,
                                            float& r0, float& r1, float& r2) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const float inorm = 1.0f / sqrtf(x * x + y * y + z * z);
    x *= inorm; y *= inorm; z *= inorm;
    float b[NB];
    sh_basis<DEG, false>(x, y, z, b, nullptr, nullptr, nullptr);
    r0 = b[0] * __ldg(co + 0); r1 = b[0] * __ldg(co + 1); r2 = b[0] * __ldg(co + 2);
#pragma unroll
    for (int k = 1; k < NB; ++k) {
        r0 = r0 + b[k] * __ldg(co + k * 3 + 0);
        r1 = r1 + b[k] * __ldg(co + k * 3 + 1);
        r2 = r2 + b[k] * __ldg(co + k * 3 + 2);
    }
    if (post) {
        r0 = fmaxf(r0 + HGS_SH_OFFSET, 0.f);
        r1 = fmaxf(r1 + HGS_SH_OFFSET, 0.f);
        r2 = fmaxf(r2 + HGS_SH_OFFSET, 0.f);
    }
}

// gradient of one (camera, Gaussian) pair: g_co[NB*3] += basis * v, gd = d/d(un-normalised direction)
template <int DEG>
__device__ __forceinline__ void sh_grad_one(float x, float y, float z, const float* __restrict__ co, float v0, float v1,
                                            float v2, bool want_dir, float* g_co, float& gd0, float& gd1, float& gd2) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const float inorm = 1.0f / sqrtf(x * x + y * y + z * z);
    x *= inorm; y *= inorm; z *= inorm;
    float b[NB], bx[NB], by[NB], bz[NB];
    sh_basis<DEG, true>(x, y, z, b, bx, by, bz);
    float vx = 0.f, vy = 0.f, vz = 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        g_co[k * 3 + 0] += b[k] * v0;
        g_co[k * 3 + 1] += b[k] * v1;
        g_co[k * 3 + 2] += b[k] * v2;
        if (want_dir) {
            const float d = co[k * 3] * v0 + co[k * 3 + 1] * v1 + co[k * 3 + 2] * v2;
            vx += bx[k] * d; vy += by[k] * d; vz += bz[k] * d;
        }
    }
    gd0 = gd1 = gd2 = 0.f;
    if (want_dir) {
        const float dd = vx * x + vy * y + vz * z;
        gd0 = (vx - dd * x) * inorm;
        gd1 = (vy - dd * y) * inorm;
        gd2 = (vz - dd * z) * inorm;
    }
}

// Dense variant: one thread per Gaussian.  Only Gaussians that survive culling (radii > 0) touch their
// coefficient row (12 K bytes).  Rows are read/written by their owning thread; every byte of a touched row
// is used, so sector efficiency stays high without shared-memory staging.
template <int DEG>
__global__ void __launch_bounds__(SB) sh_fwd_kernel(const float* __restrict__ dirs, const float* __restrict__ means,
                                                    const float* __restrict__ campos,
                                                    const float* __restrict__ coeffs,
                                                    const int32_t* __restrict__ radii, int C, int N, int K, int post,
                                                    float* __restrict__ colors) {
    const long long n = (long long)blockIdx.x * SB + threadIdx.x;
    if (n >= N) return;
    const float* co = coeffs + n * (long long)(K * 3);
    for (int c = 0; c < C; ++c) {
        const long long idx = (long long)c * N + n;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        if (radii == nullptr || radii[idx] > 0) {
            float x, y, z;
            load_dir(dirs, means, campos, idx, n, c, x, y, z);
            sh_eval_one<DEG>(x, y, z, co, post, r0, r1, r2);
        }
        colors[idx * 3] = r0; colors[idx * 3 + 1] = r1; colors[idx * 3 + 2] = r2;
    }
}

// Work-list variant: one thread per visible (camera, Gaussian) pair; colors is zero-filled by the launcher.
template <int DEG>
__global__ void __launch_bounds__(SB) sh_fwd_vis_kernel(const float* __restrict__ dirs, const float* __restrict__ means,
                                                        const float* __restrict__ campos,
                                                        const float* __restrict__ coeffs,
                                                        const int32_t* __restrict__ vis_ids, long long n_vis, int N,
                                                        int K, int post, float* __restrict__ colors) {
    const long long j = (long long)blockIdx.x * SB + threadIdx.x;
    if (j >= n_vis) return;
    const long long idx = vis_ids[j];
    const int c = (int)(idx / N);
    const long long n = idx - (long long)c * N;
    float x, y, z, r0, r1, r2;
    load_dir(dirs, means, campos, idx, n, c, x, y, z);
    sh_eval_one<DEG>(x, y, z, coeffs + n * (long long)(K * 3), post, r0, r1, r2);
    colors[idx * 3] = r0; colors[idx * 3 + 1] = r1; colors[idx * 3 + 2] = r2;
}

// v_coeffs must be zero-filled by the launcher (rows of culled Gaussians are never touched here).
template <int DEG>
__global__ void __launch_bounds__(SB) sh_bwd_kernel(const float* __restrict__ dirs, const float* __restrict__ means,
                                                    const float* __restrict__ campos,
                                                    const float* __restrict__ coeffs,
                                                    const int32_t* __restrict__ radii,
                                                    const float* __restrict__ colors,
                                                    const float* __restrict__ v_colors, int ld_vc, int C, int N,
                                                    int K, int post, float* __restrict__ v_coeffs,
                                                    float* __restrict__ v_dirs, float* __restrict__ v_means) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const long long n = (long long)blockIdx.x * SB + threadIdx.x;
    if (n >= N) return;
    const bool want_dir = (v_dirs != nullptr) || (v_means != nullptr);
    const float* co = coeffs + n * (long long)(K * 3);
    float g_co[NB * 3];
#pragma unroll
    for (int k = 0; k < NB * 3; ++k) g_co[k] = 0.f;
    float gm0 = 0.f, gm1 = 0.f, gm2 = 0.f;
    bool any = false;
    for (int c = 0; c < C; ++c) {
        const long long idx = (long long)c * N + n;
        float gd0 = 0.f, gd1 = 0.f, gd2 = 0.f;
        if (radii == nullptr || radii[idx] > 0) {
            any = true;
            float v0 = v_colors[idx * ld_vc], v1 = v_colors[idx * ld_vc + 1], v2 = v_colors[idx * ld_vc + 2];
            if (post) {
                if (!(colors[idx * 3] > 0.f)) v0 = 0.f;
                if (!(colors[idx * 3 + 1] > 0.f)) v1 = 0.f;
                if (!(colors[idx * 3 + 2] > 0.f)) v2 = 0.f;
            }
            float x, y, z;
            load_dir(dirs, means, campos, idx, n, c, x, y, z);
            sh_grad_one<DEG>(x, y, z, co, v0, v1, v2, want_dir, g_co, gd0, gd1, gd2);
            gm0 += gd0; gm1 += gd1; gm2 += gd2;
        }
        if (v_dirs != nullptr) { v_dirs[idx * 3] = gd0; v_dirs[idx * 3 + 1] = gd1; v_dirs[idx * 3 + 2] = gd2; }
    }
    if (any) {
        float* out = v_coeffs + n * (long long)(K * 3);
#pragma unroll
        for (int k = 0; k < NB * 3; ++k) out[k] = g_co[k];
    }
    if (v_means != nullptr) { v_means[n * 3] = gm0; v_means[n * 3 + 1] = gm1; v_means[n * 3 + 2] = gm2; }
}

// Fused dense variant used by the rasterization path (radii given): a block owns SB consecutive coefficient
// rows and writes ALL of them with coalesced stores (zeros for culled Gaussians, so no separate memset pass);
// the visible rows of the block are compacted first so that the gradient math runs on densely populated warps
// even when only ~10 % of the Gaussians are visible.  Sums over cameras are taken in-thread (deterministic).
template <int DEG>
__global__ void __launch_bounds__(SB) sh_bwd_fused_kernel(const float* __restrict__ means,
                                                          const float* __restrict__ campos,
                                                          const float* __restrict__ coeffs,
                                                          const int32_t* __restrict__ radii,
                                                          const float* __restrict__ colors,
                                                          const float* __restrict__ v_colors, int ld_vc, int C, int N,
                                                          int K, int post, float* __restrict__ v_coeffs,
                                                          float* __restrict__ v_means) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    constexpr int RL = NB * 3;                 // used floats per row
    extern __shared__ float s_rows[];          // [SB][RL | 1] gradient rows of the visible Gaussians
    __shared__ int s_list[SB];                 // compacted local row indices
    __shared__ int s_slot[SB];                 // local row -> slot in s_rows (-1 = culled)
    __shared__ int s_wcnt[SB / 32];
    constexpr int RS = RL | 1;
    const long long base = (long long)blockIdx.x * SB;
    const long long n = base + threadIdx.x;
    bool vis = false;
    if (n < N)
        for (int c = 0; c < C; ++c) vis |= radii[(long long)c * N + n] > 0;
    // block compaction (order preserving)
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, vis);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int wbase = 0, n_vis = 0;
#pragma unroll
    for (int w = 0; w < SB / 32; ++w) {
        if (w < warp) wbase += s_wcnt[w];
        n_vis += s_wcnt[w];
    }
    const int slot = wbase + __popc(bal & ((1u << lane) - 1u));
    s_slot[threadIdx.x] = vis ? slot : -1;
    if (vis) s_list[slot] = threadIdx.x;
    __syncthreads();
    const int rowlen = K * 3;
    long long rows = N - base;
    if (rows > SB) rows = SB;
    if (n_vis > 0) {
        // coefficient rows of the visible Gaussians -> shared memory, loaded by the whole block (each row is
        // one contiguous 12*NB-byte segment); the slot is overwritten with the gradient row afterwards
        if (v_means != nullptr) {
            const unsigned magic_rl = 0xFFFFFFFFu / (unsigned)RL + 1u;
            for (int i = threadIdx.x; i < n_vis * RL; i += SB) {
                const int j = (int)__umulhi((unsigned)i, magic_rl), cc = i - j * RL;
                s_rows[j * RS + cc] = coeffs[(base + s_list[j]) * (long long)rowlen + cc];
            }
            __syncthreads();
        }
        // thread j < n_vis handles the j-th visible Gaussian of the block
        float gm0 = 0.f, gm1 = 0.f, gm2 = 0.f;
        if (threadIdx.x < n_vis) {
            const int r = s_list[threadIdx.x];
            const long long nn = base + r;
            const float* co = s_rows + threadIdx.x * RS;
            float g_co[RL];
#pragma unroll
            for (int k = 0; k < RL; ++k) g_co[k] = 0.f;
            for (int c = 0; c < C; ++c) {
                const long long idx = (long long)c * N + nn;
                if (radii[idx] <= 0) continue;
                float v0 = v_colors[idx * ld_vc], v1 = v_colors[idx * ld_vc + 1], v2 = v_colors[idx * ld_vc + 2];
                if (post) {
                    if (!(colors[idx * 3] > 0.f)) v0 = 0.f;
                    if (!(colors[idx * 3 + 1] > 0.f)) v1 = 0.f;
                    if (!(colors[idx * 3 + 2] > 0.f)) v2 = 0.f;
                }
                const float x = means[nn * 3] - campos[c * 3], y = means[nn * 3 + 1] - campos[c * 3 + 1],
                            z = means[nn * 3 + 2] - campos[c * 3 + 2];
                float gd0, gd1, gd2;
                sh_grad_one<DEG>(x, y, z, co, v0, v1, v2, v_means != nullptr, g_co, gd0, gd1, gd2);
                gm0 += gd0; gm1 += gd1; gm2 += gd2;
            }
            float* out = s_rows + threadIdx.x * RS;
#pragma unroll
            for (int k = 0; k < RL; ++k) out[k] = g_co[k];
            if (v_means != nullptr) { v_means[nn * 3] = gm0; v_means[nn * 3 + 1] = gm1; v_means[nn * 3 + 2] = gm2; }
        }
        __syncthreads();
    }
    // coalesced write of all rows of the block: 16-byte stores; a float4 whose (at most two) rows are both
    // culled -- the common case -- is a plain zero store
    const int tot = (int)rows * rowlen;                                   // <= 128 * 75
    const unsigned magic = 0xFFFFFFFFu / (unsigned)rowlen + 1u;           // i / rowlen == umulhi(i, magic) here
    float* dst = v_coeffs + base * rowlen;                                // 16-byte aligned: SB * rowlen * 4 % 16 == 0
    const int tot4 = tot >> 2;
    for (int v4 = threadIdx.x; v4 < tot4; v4 += SB) {
        const int i0 = v4 << 2;
        const int r0 = (int)__umulhi((unsigned)i0, magic), r1 = (int)__umulhi((unsigned)(i0 + 3), magic);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s_slot[r0] >= 0 || s_slot[r1] >= 0) {
            float e[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = i0 + k;
                const int r = (int)__umulhi((unsigned)i, magic), c = i - r * rowlen;
                const int sl = s_slot[r];
                e[k] = (sl >= 0 && c < RL) ? s_rows[sl * RS + c] : 0.f;
            }
            o = make_float4(e[0], e[1], e[2], e[3]);
        }
        reinterpret_cast<float4*>(dst)[v4] = o;
    }
    for (int i = (tot4 << 2) + threadIdx.x; i < tot; i += SB) {          // tail of a partial last block
        const int r = (int)__umulhi((unsigned)i, magic), c = i - r * rowlen;
        const int sl = s_slot[r];
        dst[i] = (sl >= 0 && c < RL) ? s_rows[sl * RS + c] : 0.f;
    }
    // culled Gaussians: zero direction gradient
    if (v_means != nullptr && n < N && !vis) { v_means[n * 3] = 0.f; v_means[n * 3 + 1] = 0.f; v_means[n * 3 + 2] = 0.f; }
}

}  // namespace

HGS_API int hgs_sh_fwd(int degree, int K, const float* dirs, const float* means, const float* campos,
                       const float* coeffs, const int32_t* radii, const int32_t* vis_ids, long long n_vis, int C, int N,
                       int post, float* colors, void* stream) {
    if (degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) || C <= 0 || N < 0 || n_vis < 0) return HGS_ERR_INVALID_ARG;
    if (dirs == nullptr && (means == nullptr || campos == nullptr)) return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (vis_ids != nullptr) {
        cudaError_t e = cudaMemsetAsync(colors, 0, (size_t)C * N * 3 * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
        if (n_vis == 0) return 0;
        const int grid = hgs_ceil_div(n_vis, SB);
#define LAUNCH(DEG) sh_fwd_vis_kernel<DEG><<<grid, SB, 0, st>>>(dirs, means, campos, coeffs, vis_ids, n_vis, N, K, post, colors);
        switch (degree) {
            case 0: LAUNCH(0) break;
            case 1: LAUNCH(1) break;
            case 2: LAUNCH(2) break;
            case 3: LAUNCH(3) break;
            default: LAUNCH(4) break;
        }
#undef LAUNCH
        HGS_LAUNCH_CHECK();
        return 0;
    }
    const int grid = hgs_ceil_div(N, SB);
#define LAUNCH(DEG) sh_fwd_kernel<DEG><<<grid, SB, 0, st>>>(dirs, means, campos, coeffs, radii, C, N, K, post, colors);
    switch (degree) {
        case 0: LAUNCH(0) break;
        case 1: LAUNCH(1) break;
        case 2: LAUNCH(2) break;
        case 3: LAUNCH(3) break;
        default: LAUNCH(4) break;
    }
#undef LAUNCH
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_sh_bwd(int degree, int K, const float* dirs, const float* means, const float* campos,
                       const float* coeffs, const int32_t* radii, const int32_t* vis_ids, long long n_vis,
                       const float* colors, const float* v_colors, int ld_v_colors, int C, int N, int post,
                       float* v_coeffs, float* v_dirs, float* v_means, void* stream) {
    if (degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) || C <= 0 || N < 0 || ld_v_colors < 3 || n_vis < 0)
        return HGS_ERR_INVALID_ARG;
    if (dirs == nullptr && (means == nullptr || campos == nullptr)) return HGS_ERR_INVALID_ARG;
    if (post && colors == nullptr) return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    (void)vis_ids; (void)n_vis;
    if (radii != nullptr && dirs == nullptr && v_dirs == nullptr) {
        // rasterization path: fused zero-fill + block-compacted gradient rows, one pass over v_coeffs
        const int grid = hgs_ceil_div(N, SB);
#define LAUNCH(DEG)                                                                                                \
    {                                                                                                              \
        const size_t smem = (size_t)SB * ((((DEG) + 1) * ((DEG) + 1) * 3) | 1) * sizeof(float);                    \
        cudaFuncSetAttribute(sh_bwd_fused_kernel<DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        sh_bwd_fused_kernel<DEG><<<grid, SB, smem, st>>>(means, campos, coeffs, radii, colors, v_colors, ld_v_colors, \
                                                         C, N, K, post, v_coeffs, v_means);                        \
    }
        switch (degree) {
            case 0: LAUNCH(0) break;
            case 1: LAUNCH(1) break;
            case 2: LAUNCH(2) break;
            case 3: LAUNCH(3) break;
            default: LAUNCH(4) break;
        }
#undef LAUNCH
        HGS_LAUNCH_CHECK();
        return 0;
    }
    cudaError_t e = cudaMemsetAsync(v_coeffs, 0, (size_t)N * K * 3 * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    const int grid = hgs_ceil_div(N, SB);
#define LAUNCH(DEG)                                                                                        \
    sh_bwd_kernel<DEG><<<grid, SB, 0, st>>>(dirs, means, campos, coeffs, radii, colors, v_colors, ld_v_colors, C, N, \
                                            K, post, v_coeffs, v_dirs, v_means);
    switch (degree) {
        case 0: LAUNCH(0) break;
        case 1: LAUNCH(1) break;
        case 2: LAUNCH(2) break;
        case 3: LAUNCH(3) break;
        default: LAUNCH(4) break;
    }
#undef LAUNCH
    HGS_LAUNCH_CHECK();
    return 0;
}
