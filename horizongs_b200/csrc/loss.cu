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
//
// Second part: the SSIM term (utils/loss_utils.py:20-60, train.py:159).  The reference runs five depthwise 11x11
// convolutions (mu1, mu2, E[x^2], E[y^2], E[xy]) plus ~15 elementwise kernels forward and their autograd backward.
// Here: forward = one kernel per 16x16 tile (halo staged in shared memory, separable 11-tap Gaussian, the five
// moments never leave the SM) that also emits the three per-pixel derivative maps d ssim / d{mu1, E[x^2], E[xy]};
// backward = one kernel that convolves those maps with the same window and assembles
//     d ssim_mean / d x(p) = [ conv(d_mu1)(p) + 2 x(p) conv(d_e1)(p) + y(p) conv(d_e12)(p) ] / (3 P)
// straight into the gradient image (+=).  Zero padding as F.conv2d(padding=5).  Window = the reference's
// float32 gaussian(11, 1.5).  Roofline: HBM / shared memory.
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

namespace {
constexpr int ST = 16;             // tile edge (one thread per pixel)
constexpr int SRAD = 5;            // window radius (11 taps)
constexpr int SHALO = ST + 2 * SRAD;
// torch.Tensor([exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)]) / sum, float32 (loss_utils.py:20-22)
__device__ __constant__ float c_win[11] = {1.028380124e-03f, 7.598758209e-03f, 3.600077331e-02f, 1.093606874e-01f,
                                           2.130055279e-01f, 2.660117149e-01f, 2.130055279e-01f, 1.093606874e-01f,
                                           3.600077331e-02f, 7.598758209e-03f, 1.028380124e-03f};
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;

// grid (tiles_x, tiles_y, cameras); 256 threads; channels 0..2 of channels-last images (img1 has D channels)
__global__ void __launch_bounds__(ST * ST) ssim_fwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                                                          int H, int W, int D, float* __restrict__ dmaps,
                                                          float* __restrict__ partials) {
    __shared__ float s1[SHALO][SHALO + 1], s2[SHALO][SHALO + 1];
    __shared__ float hb[5][SHALO][ST + 1];
    __shared__ float s_red[ST * ST / 32];
    const int tx = threadIdx.x & (ST - 1), ty = threadIdx.x >> 4;
    const int x0 = blockIdx.x * ST - SRAD, y0 = blockIdx.y * ST - SRAD;
    const long long cam_px = (long long)blockIdx.z * H * W;
    const long long P3 = (long long)gridDim.z * H * W * 3;
    const int px = blockIdx.x * ST + tx, py = blockIdx.y * ST + ty;
    const bool inside = px < W && py < H;
    float acc = 0.f;
    for (int ch = 0; ch < 3; ++ch) {
        for (int i = threadIdx.x; i < SHALO * SHALO; i += ST * ST) {
            const int r = i / SHALO, c = i - r * SHALO;
            const int gy = y0 + r, gx = x0 + c;
            float a = 0.f, b = 0.f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const long long p = cam_px + (long long)gy * W + gx;
                a = img1[p * D + ch];
                b = img2[p * 3 + ch];
            }
            s1[r][c] = a;
            s2[r][c] = b;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < SHALO * ST; i += ST * ST) {
            const int r = i / ST, c = i - r * ST;
            float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f, h4 = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                const float w = c_win[k], a = s1[r][c + k], b = s2[r][c + k];
                h0 += w * a; h1 += w * b; h2 += w * a * a; h3 += w * b * b; h4 += w * a * b;
            }
            hb[0][r][c] = h0; hb[1][r][c] = h1; hb[2][r][c] = h2; hb[3][r][c] = h3; hb[4][r][c] = h4;
        }
        __syncthreads();
        float mu1 = 0.f, mu2 = 0.f, e1 = 0.f, e2 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = c_win[k];
            mu1 += w * hb[0][ty + k][tx]; mu2 += w * hb[1][ty + k][tx];
            e1 += w * hb[2][ty + k][tx]; e2 += w * hb[3][ty + k][tx]; e12 += w * hb[4][ty + k][tx];
        }
        if (inside) {
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float sg1 = e1 - mu1_sq, sg2 = e2 - mu2_sq, sg12 = e12 - mu12;
            const float A = 2.f * mu12 + SSIM_C1, B = 2.f * sg12 + SSIM_C2;
            const float Cc = mu1_sq + mu2_sq + SSIM_C1, Dd = sg1 + sg2 + SSIM_C2;
            const float inv_cd = 1.0f / (Cc * Dd);
            const float ssim = A * B * inv_cd;
            acc += ssim;
            // derivatives w.r.t. the three window means that depend on img1: mu1, E[x^2], E[xy]
            const float d_s1 = -ssim / Dd;                       // d ssim / d sigma1^2
            const float d_s12 = 2.f * A * inv_cd;                 // d ssim / d sigma12
            const float d_mu1 = 2.f * mu2 * B * inv_cd - 2.f * mu1 * ssim / Cc - 2.f * mu1 * d_s1 - mu2 * d_s12;
            const long long q = (cam_px + (long long)py * W + px) * 3 + ch;
            dmaps[q] = d_mu1;
            dmaps[P3 + q] = d_s1;
            dmaps[2 * P3 + q] = d_s12;
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < ST * ST / 32; ++w) t += s_red[w];
        partials[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void ssim_finish_kernel(const float* __restrict__ partials, int n, float inv_count, float* __restrict__ out) {
    double t = 0.0;     // one warp, fixed order
    for (int i = threadIdx.x; i < n; i += 32) t += (double)partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
    if (threadIdx.x == 0) out[0] = (float)(t * (double)inv_count);
}

// v_img1[p, ch] += v_ssim * [conv(d_mu1) + 2 x conv(d_e1) + y conv(d_e12)](p) / (3 P)
__global__ void __launch_bounds__(ST * ST) ssim_bwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                                                          const float* __restrict__ dmaps, const float* __restrict__ v_ssim,
                                                          int H, int W, int D, float* __restrict__ v_img1) {
    __shared__ float sm[3][SHALO][SHALO + 1];
    __shared__ float hb[3][SHALO][ST + 1];
    const int tx = threadIdx.x & (ST - 1), ty = threadIdx.x >> 4;
    const int x0 = blockIdx.x * ST - SRAD, y0 = blockIdx.y * ST - SRAD;
    const long long cam_px = (long long)blockIdx.z * H * W;
    const long long P3 = (long long)gridDim.z * H * W * 3;
    const int px = blockIdx.x * ST + tx, py = blockIdx.y * ST + ty;
    const bool inside = px < W && py < H;
    const float scale = v_ssim[0] / (float)P3;
    for (int ch = 0; ch < 3; ++ch) {
        for (int i = threadIdx.x; i < SHALO * SHALO; i += ST * ST) {
            const int r = i / SHALO, c = i - r * SHALO;
            const int gy = y0 + r, gx = x0 + c;
            float a = 0.f, b = 0.f, d = 0.f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const long long q = (cam_px + (long long)gy * W + gx) * 3 + ch;
                a = dmaps[q]; b = dmaps[P3 + q]; d = dmaps[2 * P3 + q];
            }
            sm[0][r][c] = a; sm[1][r][c] = b; sm[2][r][c] = d;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < SHALO * ST; i += ST * ST) {
            const int r = i / ST, c = i - r * ST;
            float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                const float w = c_win[k];
                h0 += w * sm[0][r][c + k]; h1 += w * sm[1][r][c + k]; h2 += w * sm[2][r][c + k];
            }
            hb[0][r][c] = h0; hb[1][r][c] = h1; hb[2][r][c] = h2;
        }
        __syncthreads();
        if (inside) {
            float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                const float w = c_win[k];
                c0 += w * hb[0][ty + k][tx]; c1 += w * hb[1][ty + k][tx]; c2 += w * hb[2][ty + k][tx];
            }
            const long long p = cam_px + (long long)py * W + px;
            const float x = img1[p * D + ch], y = img2[p * 3 + ch];
            v_img1[p * D + ch] += scale * (c0 + 2.f * x * c1 + y * c2);
        }
        __syncthreads();
    }
}
}  // namespace

HGS_API long long hgs_ssim_partials(int C, int H, int W) {
    if (C <= 0 || H <= 0 || W <= 0) return 0;
    return (long long)C * ((H + ST - 1) / ST) * ((W + ST - 1) / ST);
}

HGS_API int hgs_ssim_fwd(const float* render_colors, const float* gt, int C, int H, int W, int D, float* dmaps,
                         float* partials, float* ssim_mean, void* stream) {
    if (C <= 0 || H <= 0 || W <= 0 || D < 3 || render_colors == nullptr || gt == nullptr || dmaps == nullptr ||
        partials == nullptr || ssim_mean == nullptr)
        return HGS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((W + ST - 1) / ST, (H + ST - 1) / ST, C);
    ssim_fwd_kernel<<<grid, ST * ST, 0, st>>>(render_colors, gt, H, W, D, dmaps, partials);
    HGS_LAUNCH_CHECK();
    ssim_finish_kernel<<<1, 32, 0, st>>>(partials, (int)(grid.x * grid.y * grid.z), 1.0f / (3.0f * (float)C * (float)H * (float)W),
                                       ssim_mean);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_ssim_bwd(const float* render_colors, const float* gt, const float* dmaps, const float* v_ssim, int C, int H,
                         int W, int D, float* v_render_colors, void* stream) {
    if (C <= 0 || H <= 0 || W <= 0 || D < 3 || render_colors == nullptr || gt == nullptr || dmaps == nullptr ||
        v_ssim == nullptr || v_render_colors == nullptr)
        return HGS_ERR_INVALID_ARG;
    dim3 grid((W + ST - 1) / ST, (H + ST - 1) / ST, C);
    ssim_bwd_kernel<<<grid, ST * ST, 0, (cudaStream_t)stream>>>(render_colors, gt, dmaps, v_ssim, H, W, D, v_render_colors);
    HGS_LAUNCH_CHECK();
    return 0;
}
