// Stage a12: 2DGS (surfel) alpha-blend rasterization and its backward.
//
// Replaces gsplat's rasterize_to_pixels_2dgs as reached inside gsplat.rasterization_2dgs (reference
// gaussian_renderer/render.py:56-76).  Semantics restated in
// oracle/gsplat_oracle.py::rasterize_to_pixels_2dgs: per pixel (x,y) the ray-splat intersection in the
// surfel's uv frame is s = (h_u x h_v).xy / (h_u x h_v).z with h_u = x*M2 - M0, h_v = y*M2 - M1; the
// weight is min(|s|^2, 2*|mean2d - p|^2) (object-space Gaussian vs. screen-space low-pass filter);
// alpha rules as 3DGS.  Blends colour channels, the camera-space normal, optional distortion
// (last colour channel = depth) and the median depth (last Gaussian blended while T > 0.5).
//
// One CTA per 16x16 tile, one pixel per thread, batches of 128 surfels in shared memory.
// Roofline: FP32 pipe (~52 FLOP + 1 EX2 per pair forward, ~150 backward), not HBM.
#include "hgs_common.cuh"
#include "hgs_constants.cuh"
#include "../../include/hgs_raster.h"

namespace {

constexpr int TS = HGS_TILE_SIZE;
constexpr int BLK = TS * TS;

struct Surfel {       // 16 floats, 64 B
    float x, y, opac, pad;
    float u[3];       // M0
    float v[3];       // M1
    float w[3];       // M2
    float n[3];       // camera-frame normal
};

__device__ __forceinline__ void load_surfel(Surfel& s, int g, const float* __restrict__ means2d,
                                            const float* __restrict__ ray_transforms,
                                            const float* __restrict__ normals, const float* __restrict__ opacities) {
    const float2 xy = reinterpret_cast<const float2*>(means2d)[g];
    s.x = xy.x; s.y = xy.y; s.opac = opacities[g]; s.pad = 0.f;
    const float* rt = ray_transforms + (long long)g * 9;
#pragma unroll
    for (int k = 0; k < 3; ++k) { s.u[k] = rt[k]; s.v[k] = rt[3 + k]; s.w[k] = rt[6 + k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) s.n[k] = normals[(long long)g * 3 + k];
}

struct PairEval {
    float hu[3], hv[3], cr[3];
    float sx, sy, w3, w2, dx, dy, vis, alpha;
    bool valid;
};

__device__ __forceinline__ void eval_pair(const Surfel& s, float px, float py, PairEval& e) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { e.hu[k] = px * s.w[k] - s.u[k]; e.hv[k] = py * s.w[k] - s.v[k]; }
    e.cr[0] = e.hu[1] * e.hv[2] - e.hu[2] * e.hv[1];
    e.cr[1] = e.hu[2] * e.hv[0] - e.hu[0] * e.hv[2];
    e.cr[2] = e.hu[0] * e.hv[1] - e.hu[1] * e.hv[0];
    e.valid = (e.cr[2] != 0.f);
    const float iz = e.valid ? 1.0f / e.cr[2] : 0.f;
    e.sx = e.cr[0] * iz;
    e.sy = e.cr[1] * iz;
    e.w3 = e.sx * e.sx + e.sy * e.sy;
    e.dx = s.x - px;
    e.dy = s.y - py;
    e.w2 = HGS_FILTER_INV_SQUARE_2DGS * (e.dx * e.dx + e.dy * e.dy);
    const float sigma = 0.5f * fminf(e.w3, e.w2);
    e.vis = __expf(-sigma);
    e.alpha = fminf(HGS_ALPHA_MAX, s.opac * e.vis);
    if (sigma < 0.f || e.alpha < HGS_ALPHA_MIN) e.valid = false;
}

template <int D>
__global__ void __launch_bounds__(BLK) blend2d_fwd_kernel(
    const float* __restrict__ means2d, const float* __restrict__ ray_transforms, const float* __restrict__ colors,
    const float* __restrict__ depths, const float* __restrict__ normals, const float* __restrict__ opacities,
    const float* __restrict__ backgrounds, int C, int CH, int W, int H, int tile_w, int tile_h,
    const int32_t* __restrict__ offsets, const int32_t* __restrict__ flatten_ids, int n_isects,
    float* __restrict__ render_colors, float* __restrict__ render_alphas, float* __restrict__ render_normals,
    float* __restrict__ render_distort, float* __restrict__ render_median, int32_t* __restrict__ last_ids,
    int32_t* __restrict__ median_ids) {
    __shared__ Surfel s_g[BLK];
    __shared__ float s_col[BLK * D];
    const int cam = blockIdx.z;
    const int gtile = (cam * tile_h + blockIdx.y) * tile_w + blockIdx.x;
    const int tr = threadIdx.y * TS + threadIdx.x;
    const int pi = blockIdx.y * TS + threadIdx.y;
    const int pj = blockIdx.x * TS + threadIdx.x;
    const float px = (float)pj + 0.5f, py = (float)pi + 0.5f;
    const bool inside = (pi < H && pj < W);
    bool done = !inside;
    const int range_start = offsets[gtile];
    const int range_end = (gtile == C * tile_w * tile_h - 1) ? n_isects : offsets[gtile + 1];
    const int num_batches = (range_end - range_start + BLK - 1) / BLK;

    float T = 1.0f;
    int cur_idx = 0, median_idx = 0;
    float pix[D], nrm[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < D; ++k) pix[k] = 0.f;
    float distort = 0.f, accum_vd = 0.f, median_depth = 0.f;

    for (int b = 0; b < num_batches; ++b) {
        if (__syncthreads_count(done) >= BLK) break;
        const int batch_start = range_start + BLK * b;
        const int idx = batch_start + tr;
        if (idx < range_end) {
            const int g = flatten_ids[idx];
            load_surfel(s_g[tr], g, means2d, ray_transforms, normals, opacities);
#pragma unroll
            for (int k = 0; k < D; ++k) s_col[tr * D + k] = (k < CH) ? colors[(long long)g * CH + k] : depths[g];
        }
        __syncthreads();
        const int batch_size = min(BLK, range_end - batch_start);
        for (int t = 0; t < batch_size && !done; ++t) {
            PairEval e;
            eval_pair(s_g[t], px, py, e);
            if (!e.valid) continue;
            const float next_T = T * (1.0f - e.alpha);
            if (next_T <= HGS_T_EPS) {
                done = true;
                break;
            }
            const float vis = e.alpha * T;
#pragma unroll
            for (int k = 0; k < D; ++k) pix[k] += s_col[t * D + k] * vis;
#pragma unroll
            for (int k = 0; k < 3; ++k) nrm[k] += s_g[t].n[k] * vis;
            const float depth = s_col[t * D + D - 1];
            if (render_distort != nullptr) {
                distort += 2.0f * (vis * depth * (1.0f - T) - vis * accum_vd);
                accum_vd += vis * depth;
            }
            if (T > HGS_MEDIAN_T_2DGS) {
                median_depth = depth;
                median_idx = batch_start + t;
            }
            cur_idx = batch_start + t;
            T = next_T;
        }
    }
    if (inside) {
        const long long pid = ((long long)cam * H + pi) * W + pj;
        render_alphas[pid] = 1.0f - T;
#pragma unroll
        for (int k = 0; k < D; ++k)
            render_colors[pid * D + k] = backgrounds == nullptr ? pix[k] : pix[k] + T * backgrounds[cam * D + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) render_normals[pid * 3 + k] = nrm[k];
        if (render_distort != nullptr) render_distort[pid] = distort;
        render_median[pid] = median_depth;
        last_ids[pid] = cur_idx;
        median_ids[pid] = median_idx;
    }
}

__device__ __forceinline__ float warp_sum2(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

template <int D>
__global__ void __launch_bounds__(BLK) blend2d_bwd_kernel(
    const float* __restrict__ means2d, const float* __restrict__ ray_transforms, const float* __restrict__ colors,
    const float* __restrict__ depths, const float* __restrict__ normals, const float* __restrict__ opacities,
    const float* __restrict__ backgrounds, int C, int CH, int W, int H, int tile_w, int tile_h,
    const int32_t* __restrict__ offsets, const int32_t* __restrict__ flatten_ids, int n_isects,
    const float* __restrict__ render_colors, const float* __restrict__ render_alphas,
    const int32_t* __restrict__ last_ids, const int32_t* __restrict__ median_ids,
    const float* __restrict__ v_render_colors, const float* __restrict__ v_render_alphas,
    const float* __restrict__ v_render_normals, const float* __restrict__ v_render_distort,
    const float* __restrict__ v_render_median, float* __restrict__ v_means2d, float* __restrict__ v_ray_transforms,
    float* __restrict__ v_colors, float* __restrict__ v_depths, float* __restrict__ v_normals,
    float* __restrict__ v_opacities, float* __restrict__ v_densify) {
    __shared__ int s_id[BLK];
    __shared__ Surfel s_g[BLK];
    __shared__ float s_col[BLK * D];
    const int cam = blockIdx.z;
    const int gtile = (cam * tile_h + blockIdx.y) * tile_w + blockIdx.x;
    const int tr = threadIdx.y * TS + threadIdx.x;
    const int lane = tr & 31;
    const int pi = blockIdx.y * TS + threadIdx.y;
    const int pj = blockIdx.x * TS + threadIdx.x;
    const float px = (float)pj + 0.5f, py = (float)pi + 0.5f;
    const bool inside = (pi < H && pj < W);
    const long long pid = ((long long)cam * H + min(pi, H - 1)) * W + min(pj, W - 1);
    const int range_start = offsets[gtile];
    const int range_end = (gtile == C * tile_w * tile_h - 1) ? n_isects : offsets[gtile + 1];
    const int num_batches = (range_end - range_start + BLK - 1) / BLK;
    if (num_batches <= 0) return;

    const float T_final = 1.0f - render_alphas[pid];
    float T = T_final;
    float buffer[D], v_c[D], buffer_n[3] = {0.f, 0.f, 0.f}, v_n[3];
    float bg_dot = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        buffer[k] = 0.f;
        v_c[k] = inside ? v_render_colors[pid * D + k] : 0.f;
        if (backgrounds != nullptr) bg_dot += backgrounds[cam * D + k] * v_c[k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) v_n[k] = (inside && v_render_normals != nullptr) ? v_render_normals[pid * 3 + k] : 0.f;
    const float v_a = inside ? v_render_alphas[pid] : 0.f;
    const float v_dist = (inside && v_render_distort != nullptr) ? v_render_distort[pid] : 0.f;
    const float v_med = (inside && v_render_median != nullptr) ? v_render_median[pid] : 0.f;
    const int med_id = inside ? median_ids[pid] : -1;
    // distortion bookkeeping: totals of the forward pass
    const float accum_w = 1.0f - T_final;
    float accum_d = render_colors[pid * D + D - 1];
    if (backgrounds != nullptr) accum_d -= T_final * backgrounds[cam * D + D - 1];
    float accum_w_buf = accum_w, accum_d_buf = accum_d, distort_buf = 0.f;

    const int bin_final = inside ? last_ids[pid] : 0;
    int warp_bin_final = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xFFFFFFFFu, warp_bin_final, o));

    for (int b = 0; b < num_batches; ++b) {
        __syncthreads();
        const int batch_end = range_end - 1 - BLK * b;
        const int batch_size = min(BLK, batch_end + 1 - range_start);
        const int idx = batch_end - tr;
        if (idx >= range_start) {
            const int g = flatten_ids[idx];
            s_id[tr] = g;
            load_surfel(s_g[tr], g, means2d, ray_transforms, normals, opacities);
#pragma unroll
            for (int k = 0; k < D; ++k) s_col[tr * D + k] = (k < CH) ? colors[(long long)g * CH + k] : depths[g];
        }
        __syncthreads();
        for (int t = max(0, batch_end - warp_bin_final); t < batch_size; ++t) {
            bool valid = inside && (batch_end - t <= bin_final);
            PairEval e;
            e.valid = false;
            if (valid) {
                eval_pair(s_g[t], px, py, e);
                valid = e.valid;
            }
            if (!__any_sync(0xFFFFFFFFu, valid)) continue;
            float v_col[D];
#pragma unroll
            for (int k = 0; k < D; ++k) v_col[k] = 0.f;
            float v_nl[3] = {0.f, 0.f, 0.f};
            float v_u[3] = {0.f, 0.f, 0.f}, v_v[3] = {0.f, 0.f, 0.f}, v_w[3] = {0.f, 0.f, 0.f};
            float v_x = 0.f, v_y = 0.f, v_o = 0.f, v_dx = 0.f, v_dy = 0.f;
            if (valid) {
                const Surfel& s = s_g[t];
                const float ra = 1.0f / (1.0f - e.alpha);
                T *= ra;
                const float fac = e.alpha * T;
                float v_alpha = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    v_col[k] = fac * v_c[k];
                    v_alpha += (s_col[t * D + k] * T - buffer[k] * ra) * v_c[k];
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    v_nl[k] = fac * v_n[k];
                    v_alpha += (s.n[k] * T - buffer_n[k] * ra) * v_n[k];
                }
                v_alpha += T_final * ra * v_a;
                if (backgrounds != nullptr) v_alpha += -T_final * ra * bg_dot;
                const float depth = s_col[t * D + D - 1];
                if (v_render_distort != nullptr) {
                    const float dl_dw =
                        2.0f * (2.0f * (depth * accum_w_buf - accum_d_buf) + (accum_d - depth * accum_w));
                    v_alpha += (dl_dw * T - distort_buf * ra) * v_dist;
                    accum_d_buf -= fac * depth;
                    accum_w_buf -= fac;
                    distort_buf += dl_dw * fac;
                    v_col[D - 1] += 2.0f * fac * (2.0f - 2.0f * T - accum_w + fac) * v_dist;
                }
                if (batch_end - t == med_id) v_col[D - 1] += v_med;

                if (s.opac * e.vis <= HGS_ALPHA_MAX) {
                    const float v_G = s.opac * v_alpha;
                    if (e.w3 <= e.w2) {
                        const float v_sx = v_G * -e.vis * e.sx;
                        const float v_sy = v_G * -e.vis * e.sy;
                        const float iz = 1.0f / e.cr[2];
                        const float vcx = v_sx * iz, vcy = v_sy * iz;
                        const float vcr[3] = {vcx, vcy, -(vcx * e.sx + vcy * e.sy)};
                        // cross = hu x hv : v_hu = hv x v_cross, v_hv = v_cross x hu
                        const float vhu[3] = {e.hv[1] * vcr[2] - e.hv[2] * vcr[1], e.hv[2] * vcr[0] - e.hv[0] * vcr[2],
                                              e.hv[0] * vcr[1] - e.hv[1] * vcr[0]};
                        const float vhv[3] = {vcr[1] * e.hu[2] - vcr[2] * e.hu[1], vcr[2] * e.hu[0] - vcr[0] * e.hu[2],
                                              vcr[0] * e.hu[1] - vcr[1] * e.hu[0]};
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            v_u[k] = -vhu[k];
                            v_v[k] = -vhv[k];
                            v_w[k] = px * vhu[k] + py * vhv[k];
                        }
                        // screen-space positional gradient used for densification
                        v_dx = v_u[2] * s.w[2];
                        v_dy = v_v[2] * s.w[2];
                    } else {
                        v_x = v_G * -e.vis * HGS_FILTER_INV_SQUARE_2DGS * e.dx;
                        v_y = v_G * -e.vis * HGS_FILTER_INV_SQUARE_2DGS * e.dy;
                    }
                    v_o = e.vis * v_alpha;
                }
#pragma unroll
                for (int k = 0; k < D; ++k) buffer[k] += s_col[t * D + k] * fac;
#pragma unroll
                for (int k = 0; k < 3; ++k) buffer_n[k] += s.n[k] * fac;
            }
#pragma unroll
            for (int k = 0; k < D; ++k) v_col[k] = warp_sum2(v_col[k]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                v_nl[k] = warp_sum2(v_nl[k]);
                v_u[k] = warp_sum2(v_u[k]);
                v_v[k] = warp_sum2(v_v[k]);
                v_w[k] = warp_sum2(v_w[k]);
            }
            v_x = warp_sum2(v_x); v_y = warp_sum2(v_y); v_o = warp_sum2(v_o);
            if (v_densify != nullptr) { v_dx = warp_sum2(v_dx); v_dy = warp_sum2(v_dy); }
            if (lane == 0) {
                const int g = s_id[t];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    if (k < CH) atomicAdd(v_colors + (long long)g * CH + k, v_col[k]);
                    else atomicAdd(v_depths + g, v_col[k]);
                }
                float* vr = v_ray_transforms + (long long)g * 9;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    atomicAdd(v_normals + (long long)g * 3 + k, v_nl[k]);
                    atomicAdd(vr + k, v_u[k]);
                    atomicAdd(vr + 3 + k, v_v[k]);
                    atomicAdd(vr + 6 + k, v_w[k]);
                }
                atomicAdd(v_means2d + (long long)g * 2 + 0, v_x);
                atomicAdd(v_means2d + (long long)g * 2 + 1, v_y);
                atomicAdd(v_opacities + g, v_o);
                if (v_densify != nullptr) {
                    atomicAdd(v_densify + (long long)g * 2 + 0, v_dx);
                    atomicAdd(v_densify + (long long)g * 2 + 1, v_dy);
                }
            }
        }
    }
}

}  // namespace

static inline int hgs_count_launch(int e) {
    if (e == 0) __atomic_fetch_add(&g_hgs_launches, 1ull, __ATOMIC_RELAXED);
    return e;
}

#define HGS_DISPATCH_D(D, CALL)              \
    switch (D) {                             \
        case 1: return CALL(1);              \
        case 2: return CALL(2);              \
        case 3: return CALL(3);              \
        case 4: return CALL(4);              \
        case 5: return CALL(5);              \
        case 6: return CALL(6);              \
        case 7: return CALL(7);              \
        case 8: return CALL(8);              \
        default: return HGS_ERR_INVALID_ARG; \
    }

HGS_API int hgs_blend2d_fwd(const float* means2d, const float* ray_transforms, const float* colors,
                            const float* depths, const float* normals, const float* opacities,
                            const float* backgrounds, int C, int N, int CH, int width, int height, int tile_size,
                            const int32_t* isect_offsets, const int32_t* flatten_ids, long long n_isects,
                            float* render_colors, float* render_alphas, float* render_normals, float* render_distort,
                            float* render_median, int32_t* last_ids, int32_t* median_ids, void* stream) {
    (void)N;
    if (tile_size != TS || C <= 0 || width <= 0 || height <= 0 || CH < 0 || n_isects < 0) return HGS_ERR_INVALID_ARG;
    if (n_isects >= (1ll << 31)) return HGS_ERR_TOO_LARGE;
    const int D = CH + (depths != nullptr ? 1 : 0);
    const int tile_w = (width + TS - 1) / TS, tile_h = (height + TS - 1) / TS;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(tile_w, tile_h, C), block(TS, TS);
#define CALL(DD)                                                                                                     \
    (blend2d_fwd_kernel<DD><<<grid, block, 0, st>>>(means2d, ray_transforms, colors, depths, normals, opacities,      \
                                                    backgrounds, C, CH, width, height, tile_w, tile_h, isect_offsets, \
                                                    flatten_ids, (int)n_isects, render_colors, render_alphas,         \
                                                    render_normals, render_distort, render_median, last_ids,          \
                                                    median_ids),                                                      \
     hgs_count_launch((int)cudaGetLastError()))
    HGS_DISPATCH_D(D, CALL)
#undef CALL
}

HGS_API int hgs_blend2d_bwd(const float* means2d, const float* ray_transforms, const float* colors,
                            const float* depths, const float* normals, const float* opacities,
                            const float* backgrounds, int C, int N, int CH, int width, int height, int tile_size,
                            const int32_t* isect_offsets, const int32_t* flatten_ids, long long n_isects,
                            const float* render_colors, const float* render_alphas, const int32_t* last_ids,
                            const int32_t* median_ids, const float* v_render_colors, const float* v_render_alphas,
                            const float* v_render_normals, const float* v_render_distort,
                            const float* v_render_median, float* v_means2d, float* v_ray_transforms, float* v_colors,
                            float* v_depths, float* v_normals, float* v_opacities, float* v_densify, void* stream) {
    (void)N;
    if (tile_size != TS || C <= 0 || width <= 0 || height <= 0 || CH < 0 || n_isects < 0) return HGS_ERR_INVALID_ARG;
    if (n_isects >= (1ll << 31)) return HGS_ERR_TOO_LARGE;
    if (n_isects == 0) return 0;
    const int D = CH + (depths != nullptr ? 1 : 0);
    const int tile_w = (width + TS - 1) / TS, tile_h = (height + TS - 1) / TS;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(tile_w, tile_h, C), block(TS, TS);
#define CALL(DD)                                                                                                     \
    (blend2d_bwd_kernel<DD><<<grid, block, 0, st>>>(                                                                  \
         means2d, ray_transforms, colors, depths, normals, opacities, backgrounds, C, CH, width, height, tile_w,      \
         tile_h, isect_offsets, flatten_ids, (int)n_isects, render_colors, render_alphas, last_ids, median_ids,       \
         v_render_colors, v_render_alphas, v_render_normals, v_render_distort, v_render_median, v_means2d,            \
         v_ray_transforms, v_colors, v_depths, v_normals, v_opacities, v_densify),                                    \
     hgs_count_launch((int)cudaGetLastError()))
    HGS_DISPATCH_D(D, CALL)
#undef CALL
}
