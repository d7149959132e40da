// SURVEY.md section 8(f3), first part: the photometric L1 term of the training loss fused into one forward and
// one backward pass over the rendered image.
//
// Replaces the chain of elementwise / reduction kernels PyTorch launches for
//     Ll1 = torch.abs(image - gt_image).mean()                       (utils/loss_utils.py:17-18, train.py:158)
// plus the optional means of the depth channel and of the alpha map (regularisers of the same shape as
// train.py:173,178), and their autograd backward (slice scatter into a zero image, sign, scale: ~12 launches
// and ~8 passes over the 1080p image) by: forward = one read of the render and the ground truth, backward = one
// read and one write of the gradient images the blend backward consumes (already contiguous).
//     L = mean_{p, c < 3} |rc[p, c] - gt[p, c]| + w_depth * mean_p rc[p, 3] + w_alpha * mean_p ra[p]
// The block partial sums are added in a fixed order (deterministic loss value).  Roofline: HBM.
// The SSIM term (loss_utils.py:37-60) is not fused yet.
#include "hgs_common.cuh"
#include "../../include/hgs_raster.h"

namespace {
constexpr int LB = 256;

template <int D>
__global__ void __launch_bounds__(LB) l1_fwd_kernel(const float* __restrict__ rc, const float* __restrict__ ra,
                                                    const float* __restrict__ gt, long long P, float w_depth,
                                                    float w_alpha, float* __restrict__ partials) {
    float acc = 0.f;
    const float inv3p = 1.0f / (3.0f * (float)P), invp = 1.0f / (float)P;
    for (long long p = (long long)blockIdx.x * LB + threadIdx.x; p < P; p += (long long)gridDim.x * LB) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        if (D == 4) {
            const float4 v = reinterpret_cast<const float4*>(rc)[p];
            c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) c[k] = rc[p * 3 + k];
        }
        const float l1 = fabsf(c[0] - gt[p * 3]) + fabsf(c[1] - gt[p * 3 + 1]) + fabsf(c[2] - gt[p * 3 + 2]);
        acc += l1 * inv3p;
        if (D == 4) acc += w_depth * invp * c[3];
        if (ra != nullptr) acc += w_alpha * invp * ra[p];
    }
    __shared__ float s[LB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < LB / 32; ++w) t += s[w];
        partials[blockIdx.x] = t;
    }
}

__global__ void l1_finish_kernel(const float* __restrict__ partials, int n, float* __restrict__ loss) {
    // one warp, fixed order: lane l sums partials l, l + 32, ...; then a butterfly
    float t = 0.f;
    for (int i = threadIdx.x; i < n; i += 32) t += partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
    if (threadIdx.x == 0) loss[0] = t;
}

template <int D>
__global__ void __launch_bounds__(LB) l1_bwd_kernel(const float* __restrict__ rc, const float* __restrict__ gt,
                                                    const float* __restrict__ v_loss, long long P, float w_depth,
                                                    float w_alpha, float* __restrict__ v_rc, float* __restrict__ v_ra) {
    const long long p = (long long)blockIdx.x * LB + threadIdx.x;
    if (p >= P) return;
    const float g = v_loss[0];
    const float g3 = g / (3.0f * (float)P), gp = g / (float)P;
    auto sgn = [](float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); };
    if (D == 4) {
        const float4 v = reinterpret_cast<const float4*>(rc)[p];
        reinterpret_cast<float4*>(v_rc)[p] = make_float4(g3 * sgn(v.x - gt[p * 3]), g3 * sgn(v.y - gt[p * 3 + 1]),
                                                         g3 * sgn(v.z - gt[p * 3 + 2]), w_depth * gp);
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) v_rc[p * 3 + k] = g3 * sgn(rc[p * 3 + k] - gt[p * 3 + k]);
    }
    if (v_ra != nullptr) v_ra[p] = w_alpha * gp;
}
}  // namespace

HGS_API int hgs_l1_loss_partials(void) { return 1184; }

HGS_API int hgs_l1_loss_fwd(const float* render_colors, const float* render_alphas, const float* gt, long long P, int D,
                            float w_depth, float w_alpha, float* partials, float* loss, void* stream) {
    if (P <= 0 || (D != 3 && D != 4) || render_colors == nullptr || gt == nullptr || partials == nullptr || loss == nullptr)
        return HGS_ERR_INVALID_ARG;
    if (D == 4 && (reinterpret_cast<size_t>(render_colors) & 15)) return HGS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)((P + LB - 1) / LB < 1184 ? (P + LB - 1) / LB : 1184);
    if (D == 4) l1_fwd_kernel<4><<<grid, LB, 0, st>>>(render_colors, render_alphas, gt, P, w_depth, w_alpha, partials);
    else l1_fwd_kernel<3><<<grid, LB, 0, st>>>(render_colors, render_alphas, gt, P, w_depth, w_alpha, partials);
    HGS_LAUNCH_CHECK();
    l1_finish_kernel<<<1, 32, 0, st>>>(partials, grid, loss);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_l1_loss_bwd(const float* render_colors, const float* gt, const float* v_loss, long long P, int D,
                            float w_depth, float w_alpha, float* v_render_colors, float* v_render_alphas, void* stream) {
    if (P <= 0 || (D != 3 && D != 4) || render_colors == nullptr || gt == nullptr || v_loss == nullptr ||
        v_render_colors == nullptr)
        return HGS_ERR_INVALID_ARG;
    if (D == 4 && ((reinterpret_cast<size_t>(render_colors) | reinterpret_cast<size_t>(v_render_colors)) & 15))
        return HGS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = hgs_ceil_div(P, LB);
    if (D == 4) l1_bwd_kernel<4><<<grid, LB, 0, st>>>(render_colors, gt, v_loss, P, w_depth, w_alpha, v_render_colors, v_render_alphas);
    else l1_bwd_kernel<3><<<grid, LB, 0, st>>>(render_colors, gt, v_loss, P, w_depth, w_alpha, v_render_colors, v_render_alphas);
    HGS_LAUNCH_CHECK();
    return 0;
}
