// Per-Gaussian backward of one rasterized view in ONE launch: what hgs_blend3d_unpack + hgs_sh_bwd + hgs_project3d_bwd
// do in three passes over the packed gradient rows the blend backward leaves behind (stage a11 -> a7 / a3 backward
// of gsplat.rasterization as called at gaussian_renderer/render.py:40-54).
//
// Work list = ascending ids of the visible Gaussians (hgs_project3d_fwd_bin / hgs_isect_bin_prepare).  AUTONOMOUS WARPS,
// no CTA barrier: round k = visible Gaussians [32k, 32k+32) belongs to one warp, lane j owns Gaussian vis_ids[32k + j].
// A lane reads its 48-byte row of the packed buffer ONCE and from it
//   * copies the view-space mean gradient and the opacity gradient into their dense tensors (autograd hands the first
//     to meta["means2d"].grad, render.py:90-101),
//   * runs the SH backward (coefficient-gradient row through the warp's shared-memory tile so that the 12 K-byte
//     rows leave with coalesced stores; direction gradient kept in registers),
//   * runs the projection backward and writes mean (+ SH direction part), quaternion and scale gradients.
// All dense outputs are zero-filled by the caller; only the rows of visible Gaussians are written here.
// Same device functions and the same -fmad=false as project3d.cu / sh.cu: results are bit-identical to the three
// separate kernels.
//
// Roofline: HBM, gather-bound (64-byte DRAM granules for 12..16-byte parameter rows): per visible Gaussian
// 48 B row + 40 B parameters + 12 K B coefficients + 12 B colours in, 12 K + 40 + 12 B out.
#include "hgs_common.cuh"
#include "hgs_constants.cuh"
#include "project3d_math.cuh"
#include "sh_math.cuh"

namespace {

constexpr int GBW = 8;    // warps per CTA
constexpr int VP = 12;    // floats per row of the packed gradient buffer (blend3d.cu)

template <int DEG>
__global__ void __launch_bounds__(GBW * 32, (DEG <= 2 ? 3 : 2)) gauss_bwd_rounds_kernel(
    const float* __restrict__ vpack, int has_depth, const int32_t* __restrict__ vis_ids, int n_vis,
    const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
    const float* __restrict__ viewmats, const float* __restrict__ Ks, int W, int H, float eps2d, float near_plane,
    float far_plane, const float* __restrict__ campos, const float* __restrict__ coeffs, int K,
    const float* __restrict__ colors, float* __restrict__ v_means2d, float* __restrict__ v_opacities,
    float* __restrict__ v_coeffs, float* __restrict__ v_means, float* __restrict__ v_quats,
    float* __restrict__ v_scales) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    constexpr int RL = NB * 3;
    constexpr int RS = RL | 1;
    extern __shared__ float s_tiles[];         // [GBW][32][RS]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* rows = s_tiles + warp * 32 * RS;
    const int rnd = blockIdx.x * GBW + warp;
    const int j0 = rnd << 5;
    if (j0 >= n_vis) return;
    const int rowlen = K * 3;
    const int nj = min(32, n_vis - j0);
    const bool mine = lane < nj;
    const long long n = mine ? vis_ids[j0 + lane] : 0;
    // this lane's Gaussian: packed gradient row, parameters
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0, qv = make_float4(1.f, 0.f, 0.f, 0.f);
    float px = 0.f, py = 0.f, pz = 1.f, s0 = 1.f, s1 = 1.f, s2 = 1.f;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (mine) {
        const float4* row = reinterpret_cast<const float4*>(vpack + n * VP);
        r0 = row[0]; r1 = row[1]; r2 = row[2];
        px = means[n * 3]; py = means[n * 3 + 1]; pz = means[n * 3 + 2];
        s0 = scales[n * 3]; s1 = scales[n * 3 + 1]; s2 = scales[n * 3 + 2];
        qv = reinterpret_cast<const float4*>(quats)[n];
        v0 = r2.x; v1 = r2.y; v2 = r2.z;
        // the clamp of max(SH + 0.5, 0): no gradient through a clamped channel
        if (!(colors[n * 3] > 0.f)) v0 = 0.f;
        if (!(colors[n * 3 + 1] > 0.f)) v1 = 0.f;
        if (!(colors[n * 3 + 2] > 0.f)) v2 = 0.f;
    }
    // the round's 32 coefficient rows -> shared-memory tile, all loads in flight together (coalesced)
    const unsigned magic_rl = 0xFFFFFFFFu / (unsigned)RL + 1u;
    if (DEG >= 1) {
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
    __syncwarp();
    float gd0 = 0.f, gd1 = 0.f, gd2 = 0.f;
    if (mine) {
        float g_co[RL];
#pragma unroll
        for (int k = 0; k < RL; ++k) g_co[k] = 0.f;
        sh_grad_one<DEG>(px - campos[0], py - campos[1], pz - campos[2], rows + lane * RS, v0, v1, v2, DEG >= 1, g_co, gd0,
                         gd1, gd2);
        float* out = rows + lane * RS;
#pragma unroll
        for (int k = 0; k < RL; ++k) out[k] = g_co[k];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < RL; ++k) {
        const int i = lane + 32 * k;
        const int j = (int)__umulhi((unsigned)i, magic_rl), cc = i - j * RL;
        if (j < nj) v_coeffs[(long long)vis_ids[j0 + j] * rowlen + cc] = rows[j * RS + cc];
    }
    if (!mine) return;
    // dense copies of the two columns autograd consumes as tensors of their own
    reinterpret_cast<float2*>(v_means2d)[n] = make_float2(r0.x, r0.y);
    v_opacities[n] = r1.y;
    // projection backward (row layout: v_means2d 0..1 | v_conics 2..4 | v_opacity 5 | v_colors 8..10 | v_depth 11)
    float g_mean[3] = {0.f, 0.f, 0.f};
    float g_scale[3] = {0.f, 0.f, 0.f};
    float g_quat[4] = {0.f, 0.f, 0.f, 0.f};
    const HgsCam cam = hgs_load_cam(viewmats, Ks, 0);
    Proj3dFwd f;
    if (proj3d_math(cam, px, py, pz, qv.x, qv.y, qv.z, qv.w, s0, s1, s2, (float)W, (float)H, eps2d, near_plane, far_plane,
                    f)) {
        proj3d_bwd_one(cam, f, s0, s1, s2, make_float2(r0.x, r0.y), has_depth ? r2.w : 0.f, r0.z, 0.5f * r0.w, r1.x, g_mean,
                       g_scale, g_quat);
    }
    reinterpret_cast<float4*>(v_quats)[n] = make_float4(g_quat[0], g_quat[1], g_quat[2], g_quat[3]);
    // the order of the sum matches hgs_project3d_bwd with accumulate_means: SH direction part + projection part
    v_means[n * 3] = gd0 + g_mean[0];
    v_means[n * 3 + 1] = gd1 + g_mean[1];
    v_means[n * 3 + 2] = gd2 + g_mean[2];
    v_scales[n * 3] = g_scale[0];
    v_scales[n * 3 + 1] = g_scale[1];
    v_scales[n * 3 + 2] = g_scale[2];
}

}  // namespace

#include "../../include/hgs_raster.h"

HGS_API int hgs_gauss_bwd_fused(const float* vpack, int has_depth, const int32_t* vis_ids, long long n_vis, int N,
                                const float* means, const float* quats, const float* scales, const float* viewmat,
                                const float* Kmat, int width, int height, float eps2d, float near_plane,
                                float far_plane, int sh_degree, int K, const float* campos, const float* coeffs,
                                const float* colors, float* v_means2d, float* v_opacities, float* v_coeffs,
                                float* v_means, float* v_quats, float* v_scales, void* stream) {
    if (N < 0 || n_vis < 0 || n_vis >= (1ll << 31) || width <= 0 || height <= 0 || sh_degree < 0 || sh_degree > 4 ||
        K < (sh_degree + 1) * (sh_degree + 1) || vpack == nullptr || (reinterpret_cast<size_t>(vpack) & 15) ||
        (reinterpret_cast<size_t>(quats) & 15) || (reinterpret_cast<size_t>(v_quats) & 15) ||
        (reinterpret_cast<size_t>(v_means2d) & 7))
        return HGS_ERR_INVALID_ARG;
    if (means == nullptr || quats == nullptr || scales == nullptr || viewmat == nullptr || Kmat == nullptr ||
        campos == nullptr || coeffs == nullptr || colors == nullptr || v_means2d == nullptr || v_opacities == nullptr ||
        v_coeffs == nullptr || v_means == nullptr || v_quats == nullptr || v_scales == nullptr)
        return HGS_ERR_INVALID_ARG;
    if (n_vis == 0) return 0;
    if (vis_ids == nullptr) return HGS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_rounds = (int)((n_vis + 31) / 32);
    const int grid = hgs_ceil_div(n_rounds, GBW);
#define LAUNCH(DEG)                                                                                                   \
    {                                                                                                                 \
        constexpr int RS = (((DEG) + 1) * ((DEG) + 1) * 3) | 1;                                                       \
        const int smem = GBW * 32 * RS * (int)sizeof(float);                                                          \
        cudaError_t e = cudaFuncSetAttribute(gauss_bwd_rounds_kernel<DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             smem);                                                                   \
        if (e != cudaSuccess) return (int)e;                                                                          \
        gauss_bwd_rounds_kernel<DEG><<<grid, GBW * 32, smem, st>>>(                                                   \
            vpack, has_depth, vis_ids, (int)n_vis, means, quats, scales, viewmat, Kmat, width, height, eps2d,         \
            near_plane, far_plane, campos, coeffs, K, colors, v_means2d, v_opacities, v_coeffs, v_means, v_quats,     \
            v_scales);                                                                                                \
    }
    switch (sh_degree) {
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
