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
#include "sh_math.cuh"
#include "../../include/hgs_raster.h"

namespace {

constexpr int SB = 128;  // threads (= Gaussians) per block

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
                                                        const int32_t* __restrict__ vis_ids, long long n_vis,
                                                        const long long* __restrict__ n_vis_dev, int N,
                                                        int K, int post, float* __restrict__ colors) {
    const long long j = (long long)blockIdx.x * SB + threadIdx.x;
    if (j >= (n_vis_dev != nullptr ? *n_vis_dev : n_vis)) return;   // device-side count: n_vis is only a bound
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

// Work-list variant for one camera (vis_ids = ascending ids of the visible Gaussians, from hgs_isect_prepare):
// AUTONOMOUS WARPS, no CTA barrier.  Round k = visible Gaussians [32k, 32k+32) belongs to one warp: lane j owns
// Gaussian vis_ids[32k + j], so the gradient math runs on dense warps however few Gaussians are visible.  A round
//   1. zero-fills every output row of the id range it covers -- from its first visible id up to the next round's
//      first -- with streaming 16-byte stores (culled Gaussians get their zero rows here, no memset pass),
//   2. stages its 32 coefficient rows in the warp's shared-memory tile, all loads in flight together (coalesced),
//   3. runs sh_grad_one per lane, leaves the coefficient-gradient row in the tile, writes the direction gradient,
//   4. writes the 32 coefficient-gradient rows with coalesced stores.
constexpr int SHW = 8;   // warps per CTA
template <int DEG, bool ZFILL>
__global__ void __launch_bounds__(SHW * 32) sh_bwd_rounds_kernel(const float* __restrict__ means,
                                                                const float* __restrict__ campos,
                                                                const float* __restrict__ coeffs,
                                                                const int32_t* __restrict__ vis_ids, int n_vis,
                                                                const float* __restrict__ colors,
                                                                const float* __restrict__ v_colors, int ld_vc, int N,
                                                                int K, int post, float* __restrict__ v_coeffs,
                                                                float* __restrict__ v_means) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    constexpr int RL = NB * 3;
    constexpr int RS = RL | 1;
    extern __shared__ float s_tiles[];         // [SHW][32][RS]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* rows = s_tiles + warp * 32 * RS;
    const int n_rounds = max(1, (n_vis + 31) >> 5);
    const int rnd = blockIdx.x * SHW + warp;
    if (rnd >= n_rounds) return;
    const int rowlen = K * 3;
    const int j0 = rnd << 5;
    const int nj = min(32, n_vis - j0);        // <= 0 only when nothing is visible
    const bool mine = lane < nj;
    const long long n = mine ? vis_ids[j0 + lane] : 0;
    const long long lo = rnd == 0 ? 0 : vis_ids[j0];
    const long long hi = rnd == n_rounds - 1 ? N : vis_ids[j0 + 32];
    // loads of this lane's Gaussian first, so that they are in flight during the zero fill
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, x = 0.f, y = 0.f, z = 1.f;
    if (mine) {
        v0 = v_colors[n * ld_vc]; v1 = v_colors[n * ld_vc + 1]; v2 = v_colors[n * ld_vc + 2];
        if (post) {
            if (!(colors[n * 3] > 0.f)) v0 = 0.f;
            if (!(colors[n * 3 + 1] > 0.f)) v1 = 0.f;
            if (!(colors[n * 3 + 2] > 0.f)) v2 = 0.f;
        }
        x = means[n * 3] - campos[0]; y = means[n * 3 + 1] - campos[1]; z = means[n * 3 + 2] - campos[2];
    }
    if (ZFILL) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v_means != nullptr)
            for (long long e = lo * 3 + lane; e < hi * 3; e += 32) v_means[e] = 0.f;
        const long long e_lo = lo * rowlen, e_hi = hi * rowlen;
        const long long a_lo = min((e_lo + 3) & ~3ll, e_hi), a_hi = max(e_hi & ~3ll, a_lo);
        for (long long e = e_lo + lane; e < a_lo; e += 32) v_coeffs[e] = 0.f;
        for (long long e = a_lo + 4 * lane; e < a_hi; e += 128) *reinterpret_cast<float4*>(v_coeffs + e) = z4;
        for (long long e = a_hi + lane; e < e_hi; e += 32) v_coeffs[e] = 0.f;
    }
    const unsigned magic_rl = 0xFFFFFFFFu / (unsigned)RL + 1u;
    if (v_means != nullptr && DEG >= 1) {
        float tmp[RL];
#pragma unroll
        for (int k = 0; k < RL; ++k) {
            const int i = lane + 32 * k;
            const int j = (int)__umulhi((unsigned)i, magic_rl), cc = i - j * RL;
            tmp[k] = j < nj ? coeffs[(long long)vis_ids[j0 + j] * rowlen + cc] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < RL; ++k) {
            const int i = lane + 32 * k;
            const int j = (int)__umulhi((unsigned)i, magic_rl), cc = i - j * RL;
            rows[j * RS + cc] = tmp[k];
        }
    }
    __syncwarp();          // tile staged; the zero fill is ordered before the row / direction-gradient stores below
    if (mine) {
        float g_co[RL];
#pragma unroll
        for (int k = 0; k < RL; ++k) g_co[k] = 0.f;
        float gd0, gd1, gd2;
        sh_grad_one<DEG>(x, y, z, rows + lane * RS, v0, v1, v2, v_means != nullptr && DEG >= 1, g_co, gd0, gd1, gd2);
        float* out = rows + lane * RS;
#pragma unroll
        for (int k = 0; k < RL; ++k) out[k] = g_co[k];
        if (v_means != nullptr) { v_means[n * 3] = gd0; v_means[n * 3 + 1] = gd1; v_means[n * 3 + 2] = gd2; }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < RL; ++k) {
        const int i = lane + 32 * k;
        const int j = (int)__umulhi((unsigned)i, magic_rl), cc = i - j * RL;
        if (j < nj) v_coeffs[(long long)vis_ids[j0 + j] * rowlen + cc] = rows[j * RS + cc];
    }
}

}  // namespace

HGS_API int hgs_sh_fwd(int degree, int K, const float* dirs, const float* means, const float* campos,
                       const float* coeffs, const int32_t* radii, const int32_t* vis_ids, long long n_vis,
                       const long long* n_vis_dev, int C, int N, int post, float* colors, void* stream) {
    if (degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) || C <= 0 || N < 0 || n_vis < 0) return HGS_ERR_INVALID_ARG;
    if (dirs == nullptr && (means == nullptr || campos == nullptr)) return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (vis_ids != nullptr) {
        // work-list path: only the rows of the listed (visible) Gaussians are written -- every consumer of the colours
        // inside the rasterization pipeline (record packing, SH backward, fused exchange) goes through the same list
        if (n_vis == 0) return 0;
        const int grid = hgs_ceil_div(n_vis, SB);
#define LAUNCH(DEG) sh_fwd_vis_kernel<DEG><<<grid, SB, 0, st>>>(dirs, means, campos, coeffs, vis_ids, n_vis, n_vis_dev, N, K, post, colors);
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
                       float* v_coeffs, float* v_dirs, float* v_means, int outputs_zeroed, void* stream) {
    if (degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) || C <= 0 || N < 0 || ld_v_colors < 3 || n_vis < 0)
        return HGS_ERR_INVALID_ARG;
    if (dirs == nullptr && (means == nullptr || campos == nullptr)) return HGS_ERR_INVALID_ARG;
    if (post && colors == nullptr) return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (vis_ids != nullptr && C == 1 && dirs == nullptr && v_dirs == nullptr && n_vis < (1ll << 31)) {
        // rasterization path, one camera: dense-warp rounds over the visible work list; the rows of the culled
        // Gaussians are zeroed by memsets at copy-engine speed, the kernel only writes the visible rows
        const int n_rounds = (int)((n_vis + 31) / 32 < 1 ? 1 : (n_vis + 31) / 32);
        const int grid = hgs_ceil_div(n_rounds, SHW);
        if (!outputs_zeroed) {
            cudaError_t e = cudaMemsetAsync(v_coeffs, 0, (size_t)N * K * 3 * sizeof(float), st);
            if (e != cudaSuccess) return (int)e;
            if (v_means != nullptr && (e = cudaMemsetAsync(v_means, 0, (size_t)N * 3 * sizeof(float), st)) != cudaSuccess)
                return (int)e;
        }
#define LAUNCH(DEG)                                                                                                \
    {                                                                                                              \
        const size_t smem = (size_t)SHW * 32 * ((((DEG) + 1) * ((DEG) + 1) * 3) | 1) * sizeof(float);              \
        cudaFuncSetAttribute(sh_bwd_rounds_kernel<DEG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        sh_bwd_rounds_kernel<DEG, false><<<grid, SHW * 32, smem, st>>>(means, campos, coeffs, vis_ids, (int)n_vis,     \
                                                                      colors, v_colors, ld_v_colors, N, K, post,      \
                                                                      v_coeffs, v_means);                            \
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
