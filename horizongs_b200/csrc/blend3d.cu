// Stage a11: forward alpha-blend rasterization and its backward (3DGS).
//
// Replaces gsplat's rasterize_to_pixels as reached inside gsplat.rasterization (reference
// gaussian_renderer/render.py:40-54).  Semantics restated in oracle/gsplat_oracle.py::rasterize_to_pixels:
// pixel centre (x+.5, y+.5); sigma = .5(a dx^2 + c dy^2) + b dx dy; alpha = min(.999, o exp(-sigma));
// skip if sigma < 0 or alpha < 1/255; stop before the Gaussian that brings T to <= 1e-4.
//
// One CTA per 16x16 tile, one pixel per thread, Gaussians staged in shared memory in batches of 256.
// Roofline: FP32 FMA / MUFU.EX2 / shared-memory broadcast (not HBM): ~26 FLOP + 1 EX2 per (pixel, Gaussian)
// pair forward, ~80 FLOP + 1 EX2 backward plus a warp reduction and L2 atomics per (warp, Gaussian).
#include "hgs_common.cuh"
#include "hgs_constants.cuh"
#include "../../include/hgs_raster.h"

namespace {

constexpr int TS = HGS_TILE_SIZE;
constexpr int BLK = TS * TS;  // 256 threads, also the batch size

template <int D>
__global__ void __launch_bounds__(BLK) blend3d_fwd_kernel(
    const float* __restrict__ means2d, const float* __restrict__ conics, const float* __restrict__ colors,
    const float* __restrict__ depths, const float* __restrict__ opacities, const float* __restrict__ backgrounds,
    int C, int CH, int W, int H, int tile_w, int tile_h, const int32_t* __restrict__ offsets,
    const int32_t* __restrict__ flatten_ids, int n_isects, float* __restrict__ render_colors,
    float* __restrict__ render_alphas, int32_t* __restrict__ last_ids) {
    __shared__ float4 s_xyo[BLK];  // x, y, opacity, -
    __shared__ float4 s_con[BLK];  // a, b, c, -
    __shared__ float s_col[BLK * D];

    const int cam = blockIdx.z;
    const int tile_id = blockIdx.y * tile_w + blockIdx.x;
    const int gtile = cam * tile_w * tile_h + tile_id;
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
    int cur_idx = 0;
    float pix[D];
#pragma unroll
    for (int k = 0; k < D; ++k) pix[k] = 0.f;

    for (int b = 0; b < num_batches; ++b) {
        if (__syncthreads_count(done) >= BLK) break;
        const int batch_start = range_start + BLK * b;
        const int idx = batch_start + tr;
        if (idx < range_end) {
            const int g = flatten_ids[idx];
            const float2 xy = reinterpret_cast<const float2*>(means2d)[g];
            s_xyo[tr] = make_float4(xy.x, xy.y, opacities[g], 0.f);
            s_con[tr] = make_float4(conics[g * 3 + 0], conics[g * 3 + 1], conics[g * 3 + 2], 0.f);
#pragma unroll
            for (int k = 0; k < D; ++k)
                s_col[tr * D + k] = (k < CH) ? colors[(long long)g * CH + k] : depths[g];
        }
        __syncthreads();
        const int batch_size = min(BLK, range_end - batch_start);
        for (int t = 0; t < batch_size && !done; ++t) {
            const float4 xyo = s_xyo[t];
            const float4 con = s_con[t];
            const float dx = xyo.x - px, dy = xyo.y - py;
            const float sigma = 0.5f * (con.x * dx * dx + con.z * dy * dy) + con.y * dx * dy;
            const float alpha = fminf(HGS_ALPHA_MAX, xyo.z * __expf(-sigma));
            if (sigma < 0.f || alpha < HGS_ALPHA_MIN) continue;
            const float next_T = T * (1.0f - alpha);
            if (next_T <= HGS_T_EPS) {
                done = true;
                break;
            }
            const float vis = alpha * T;
#pragma unroll
            for (int k = 0; k < D; ++k) pix[k] += s_col[t * D + k] * vis;
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
        last_ids[pid] = cur_idx;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

template <int D>
__global__ void __launch_bounds__(BLK) blend3d_bwd_kernel(
    const float* __restrict__ means2d, const float* __restrict__ conics, const float* __restrict__ colors,
    const float* __restrict__ depths, const float* __restrict__ opacities, const float* __restrict__ backgrounds,
    int C, int CH, int W, int H, int tile_w, int tile_h, const int32_t* __restrict__ offsets,
    const int32_t* __restrict__ flatten_ids, int n_isects, const float* __restrict__ render_alphas,
    const int32_t* __restrict__ last_ids, const float* __restrict__ v_render_colors,
    const float* __restrict__ v_render_alphas, float* __restrict__ v_means2d, float* __restrict__ v_means2d_abs,
    float* __restrict__ v_conics, float* __restrict__ v_colors, float* __restrict__ v_depths,
    float* __restrict__ v_opacities) {
    __shared__ int s_id[BLK];
    __shared__ float4 s_xyo[BLK];
    __shared__ float4 s_con[BLK];
    __shared__ float s_col[BLK * D];

    const int cam = blockIdx.z;
    const int tile_id = blockIdx.y * tile_w + blockIdx.x;
    const int gtile = cam * tile_w * tile_h + tile_id;
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
    float buffer[D];
    float v_c[D];
    float bg_dot = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        buffer[k] = 0.f;
        v_c[k] = inside ? v_render_colors[pid * D + k] : 0.f;
        if (backgrounds != nullptr) bg_dot += backgrounds[cam * D + k] * v_c[k];
    }
    const float v_a = inside ? v_render_alphas[pid] : 0.f;
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
            const float2 xy = reinterpret_cast<const float2*>(means2d)[g];
            s_xyo[tr] = make_float4(xy.x, xy.y, opacities[g], 0.f);
            s_con[tr] = make_float4(conics[g * 3 + 0], conics[g * 3 + 1], conics[g * 3 + 2], 0.f);
#pragma unroll
            for (int k = 0; k < D; ++k)
                s_col[tr * D + k] = (k < CH) ? colors[(long long)g * CH + k] : depths[g];
        }
        __syncthreads();
        for (int t = max(0, batch_end - warp_bin_final); t < batch_size; ++t) {
            bool valid = inside && (batch_end - t <= bin_final);
            float alpha = 0.f, opac = 0.f, vis = 0.f, dx = 0.f, dy = 0.f;
            float4 con = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const float4 xyo = s_xyo[t];
                con = s_con[t];
                opac = xyo.z;
                dx = xyo.x - px;
                dy = xyo.y - py;
                const float sigma = 0.5f * (con.x * dx * dx + con.z * dy * dy) + con.y * dx * dy;
                vis = __expf(-sigma);
                alpha = fminf(HGS_ALPHA_MAX, opac * vis);
                if (sigma < 0.f || alpha < HGS_ALPHA_MIN) valid = false;
            }
            if (!__any_sync(0xFFFFFFFFu, valid)) continue;
            float v_col[D];
#pragma unroll
            for (int k = 0; k < D; ++k) v_col[k] = 0.f;
            float v_con0 = 0.f, v_con1 = 0.f, v_con2 = 0.f, v_x = 0.f, v_y = 0.f, v_ax = 0.f, v_ay = 0.f, v_o = 0.f;
            if (valid) {
                const float ra = 1.0f / (1.0f - alpha);
                T *= ra;
                const float fac = alpha * T;
                float v_alpha = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    v_col[k] = fac * v_c[k];
                    v_alpha += (s_col[t * D + k] * T - buffer[k] * ra) * v_c[k];
                }
                v_alpha += T_final * ra * v_a;
                if (backgrounds != nullptr) v_alpha += -T_final * ra * bg_dot;
                if (opac * vis <= HGS_ALPHA_MAX) {
                    const float v_sigma = -opac * vis * v_alpha;
                    v_con0 = 0.5f * v_sigma * dx * dx;
                    v_con1 = v_sigma * dx * dy;
                    v_con2 = 0.5f * v_sigma * dy * dy;
                    v_x = v_sigma * (con.x * dx + con.y * dy);
                    v_y = v_sigma * (con.y * dx + con.z * dy);
                    v_ax = fabsf(v_x);
                    v_ay = fabsf(v_y);
                    v_o = vis * v_alpha;
                }
#pragma unroll
                for (int k = 0; k < D; ++k) buffer[k] += s_col[t * D + k] * fac;
            }
#pragma unroll
            for (int k = 0; k < D; ++k) v_col[k] = warp_sum(v_col[k]);
            v_con0 = warp_sum(v_con0); v_con1 = warp_sum(v_con1); v_con2 = warp_sum(v_con2);
            v_x = warp_sum(v_x); v_y = warp_sum(v_y); v_o = warp_sum(v_o);
            if (v_means2d_abs != nullptr) { v_ax = warp_sum(v_ax); v_ay = warp_sum(v_ay); }
            if (lane == 0) {
                const int g = s_id[t];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    if (k < CH) atomicAdd(v_colors + (long long)g * CH + k, v_col[k]);
                    else atomicAdd(v_depths + g, v_col[k]);
                }
                atomicAdd(v_conics + g * 3 + 0, v_con0);
                atomicAdd(v_conics + g * 3 + 1, v_con1);
                atomicAdd(v_conics + g * 3 + 2, v_con2);
                atomicAdd(v_means2d + g * 2 + 0, v_x);
                atomicAdd(v_means2d + g * 2 + 1, v_y);
                if (v_means2d_abs != nullptr) {
                    atomicAdd(v_means2d_abs + g * 2 + 0, v_ax);
                    atomicAdd(v_means2d_abs + g * 2 + 1, v_ay);
                }
                atomicAdd(v_opacities + g, v_o);
            }
        }
    }
}

template <int D>
int launch_fwd(const float* means2d, const float* conics, const float* colors, const float* depths,
               const float* opacities, const float* backgrounds, int C, int CH, int W, int H, int tile_w, int tile_h,
               const int32_t* offsets, const int32_t* flatten_ids, int n_isects, float* render_colors,
               float* render_alphas, int32_t* last_ids, cudaStream_t st) {
    dim3 grid(tile_w, tile_h, C), block(TS, TS);
    blend3d_fwd_kernel<D><<<grid, block, 0, st>>>(means2d, conics, colors, depths, opacities, backgrounds, C, CH, W, H,
                                                   tile_w, tile_h, offsets, flatten_ids, n_isects, render_colors,
                                                   render_alphas, last_ids);
    HGS_LAUNCH_CHECK();
    return 0;
}

template <int D>
int launch_bwd(const float* means2d, const float* conics, const float* colors, const float* depths,
               const float* opacities, const float* backgrounds, int C, int CH, int W, int H, int tile_w, int tile_h,
               const int32_t* offsets, const int32_t* flatten_ids, int n_isects, const float* render_alphas,
               const int32_t* last_ids, const float* v_render_colors, const float* v_render_alphas, float* v_means2d,
               float* v_means2d_abs, float* v_conics, float* v_colors, float* v_depths, float* v_opacities,
               cudaStream_t st) {
    dim3 grid(tile_w, tile_h, C), block(TS, TS);
    blend3d_bwd_kernel<D><<<grid, block, 0, st>>>(means2d, conics, colors, depths, opacities, backgrounds, C, CH, W, H,
                                                   tile_w, tile_h, offsets, flatten_ids, n_isects, render_alphas,
                                                   last_ids, v_render_colors, v_render_alphas, v_means2d,
                                                   v_means2d_abs, v_conics, v_colors, v_depths, v_opacities);
    HGS_LAUNCH_CHECK();
    return 0;
}

}  // namespace

#define HGS_DISPATCH_D(D, CALL)                     \
    switch (D) {                                    \
        case 1: return CALL(1);                     \
        case 2: return CALL(2);                     \
        case 3: return CALL(3);                     \
        case 4: return CALL(4);                     \
        case 5: return CALL(5);                     \
        case 6: return CALL(6);                     \
        case 7: return CALL(7);                     \
        case 8: return CALL(8);                     \
        default: return HGS_ERR_INVALID_ARG;        \
    }

HGS_API int hgs_blend3d_fwd(const float* means2d, const float* conics, const float* colors, const float* depths,
                            const float* opacities, const float* backgrounds, int C, int N, int CH, int width,
                            int height, int tile_size, const int32_t* isect_offsets, const int32_t* flatten_ids,
                            long long n_isects, float* render_colors, float* render_alphas, int32_t* last_ids,
                            void* stream) {
    (void)N;
    if (tile_size != TS || C <= 0 || width <= 0 || height <= 0 || CH < 0 || n_isects < 0) return HGS_ERR_INVALID_ARG;
    if (n_isects >= (1ll << 31)) return HGS_ERR_TOO_LARGE;
    const int D = CH + (depths != nullptr ? 1 : 0);
    const int tile_w = (width + TS - 1) / TS, tile_h = (height + TS - 1) / TS;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DD)                                                                                                    \
    launch_fwd<DD>(means2d, conics, colors, depths, opacities, backgrounds, C, CH, width, height, tile_w, tile_h,    \
                   isect_offsets, flatten_ids, (int)n_isects, render_colors, render_alphas, last_ids, st)
    HGS_DISPATCH_D(D, CALL)
#undef CALL
}

HGS_API int hgs_blend3d_bwd(const float* means2d, const float* conics, const float* colors, const float* depths,
                            const float* opacities, const float* backgrounds, int C, int N, int CH, int width,
                            int height, int tile_size, const int32_t* isect_offsets, const int32_t* flatten_ids,
                            long long n_isects, const float* render_alphas, const int32_t* last_ids,
                            const float* v_render_colors, const float* v_render_alphas, float* v_means2d,
                            float* v_means2d_abs, float* v_conics, float* v_colors, float* v_depths,
                            float* v_opacities, void* stream) {
    (void)N;
    if (tile_size != TS || C <= 0 || width <= 0 || height <= 0 || CH < 0 || n_isects < 0) return HGS_ERR_INVALID_ARG;
    if (n_isects >= (1ll << 31)) return HGS_ERR_TOO_LARGE;
    if (n_isects == 0) return 0;
    const int D = CH + (depths != nullptr ? 1 : 0);
    const int tile_w = (width + TS - 1) / TS, tile_h = (height + TS - 1) / TS;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DD)                                                                                                    \
    launch_bwd<DD>(means2d, conics, colors, depths, opacities, backgrounds, C, CH, width, height, tile_w, tile_h,    \
                   isect_offsets, flatten_ids, (int)n_isects, render_alphas, last_ids, v_render_colors,              \
                   v_render_alphas, v_means2d, v_means2d_abs, v_conics, v_colors, v_depths, v_opacities, st)
    HGS_DISPATCH_D(D, CALL)
#undef CALL
}
