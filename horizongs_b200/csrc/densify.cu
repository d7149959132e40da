// densification statistics fused into one pass over the view-space gradient.
//
// Replaces the chain of boolean-mask scatters in scene/basic_model.py:96-144 (training_statis) for the part
// that consumes the rasterizer's outputs: for every Gaussian visible in the view (radii > 0)
//     grad_accum[n] += || means2d.grad[n] * (W/2, H/2) ||        (basic_model.py:131-136, mean mode)
//     grad_accum[n]  = max(grad_accum[n], ...)                    (:138-139, max mode)
//     denom[n]      += 1                                          (:144)
//     max_radii[n]   = max(max_radii[n], radii[n])                (:139, optional)
// The gradient rows may be strided (slices of the packed blend-gradient buffer are read in place).
// Roofline: HBM; 4 B (radii) per Gaussian + ~24 B per visible one.
#include "hgs_common.cuh"
#include "../../include/hgs_raster.h"

namespace {
__global__ void densify_stats_kernel(const float* __restrict__ v_means2d, int ld, const int32_t* __restrict__ radii,
                                     int C, int N, float half_w, float half_h, int mode_max,
                                     float* __restrict__ grad_accum, float* __restrict__ denom,
                                     float* __restrict__ max_radii) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = 0.f, cnt = 0.f;
    int rmax = 0;
    for (int c = 0; c < C; ++c) {
        const long long idx = (long long)c * N + n;
        const int r = radii[idx];
        if (r <= 0) continue;
        const float gx = v_means2d[idx * ld] * half_w, gy = v_means2d[idx * ld + 1] * half_h;
        const float nrm = sqrtf(gx * gx + gy * gy);
        acc = mode_max ? fmaxf(acc, nrm) : acc + nrm;
        cnt += 1.f;
        rmax = max(rmax, r);
    }
    if (cnt > 0.f) {
        grad_accum[n] = mode_max ? fmaxf(grad_accum[n], acc) : grad_accum[n] + acc;
        denom[n] += cnt;
        if (max_radii != nullptr) max_radii[n] = fmaxf(max_radii[n], (float)rmax);
    }
}
// work-list variant: one thread per visible (camera, Gaussian) pair; several cameras -> atomics
__global__ void densify_stats_vis_kernel(const float* __restrict__ v_means2d, int ld, const int32_t* __restrict__ radii,
                                         const int32_t* __restrict__ vis_ids, long long n_vis, int C, int N,
                                         float half_w, float half_h, int mode_max, float* __restrict__ grad_accum,
                                         float* __restrict__ denom, float* __restrict__ max_radii) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_vis) return;
    const long long idx = vis_ids[j];
    const long long n = idx % N;
    const float gx = v_means2d[idx * ld] * half_w, gy = v_means2d[idx * ld + 1] * half_h;
    const float nrm = sqrtf(gx * gx + gy * gy);
    const float r = (float)radii[idx];
    if (C == 1) {
        grad_accum[n] = mode_max ? fmaxf(grad_accum[n], nrm) : grad_accum[n] + nrm;
        denom[n] += 1.f;
        if (max_radii != nullptr) max_radii[n] = fmaxf(max_radii[n], r);
    } else {
        // non-negative floats order like their bit patterns
        if (mode_max) atomicMax(reinterpret_cast<int*>(grad_accum + n), __float_as_int(nrm));
        else atomicAdd(grad_accum + n, nrm);
        atomicAdd(denom + n, 1.f);
        if (max_radii != nullptr) atomicMax(reinterpret_cast<int*>(max_radii + n), __float_as_int(r));
    }
}
}  // namespace

HGS_API int hgs_densify_stats(const float* v_means2d, int ld_means2d, const int32_t* radii, const int32_t* vis_ids,
                              long long n_vis, int C, int N, int width, int height, int mode_max, float* grad_accum,
                              float* denom, float* max_radii, void* stream) {
    if (C <= 0 || N < 0 || ld_means2d < 2 || width <= 0 || height <= 0 || n_vis < 0) return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    if (vis_ids != nullptr) {
        if (n_vis == 0) return 0;
        densify_stats_vis_kernel<<<hgs_ceil_div(n_vis, 256), 256, 0, (cudaStream_t)stream>>>(
            v_means2d, ld_means2d, radii, vis_ids, n_vis, C, N, 0.5f * (float)width, 0.5f * (float)height, mode_max,
            grad_accum, denom, max_radii);
        HGS_LAUNCH_CHECK();
        return 0;
    }
    densify_stats_kernel<<<hgs_ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(
        v_means2d, ld_means2d, radii, C, N, 0.5f * (float)width, 0.5f * (float)height, mode_max, grad_accum, denom,
        max_radii);
    HGS_LAUNCH_CHECK();
    return 0;
}
