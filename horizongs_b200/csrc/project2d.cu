// Stage a4: fused 2D-Gaussian (surfel) projection and its backward.
//
// Replaces gsplat's fully_fused_projection_2dgs as reached from the reference at
// gaussian_renderer/render.py:171-186 (prefilter_voxel, 2D branch) and inside
// gsplat.rasterization_2dgs (render.py:56-76).  Per surfel: H = [s0*r0, s1*r1, mu] in camera
// space, ray transform M = K H (rows M0, M1, M2), projected centre and extent from M with
// f = (1,1,-1)/(M2.M2*), radius = ceil(3 sqrt(max(1e-4, extent))), view-facing normal.
// Built with -fmad=false; bit-matches oracle/gsplat_oracle.py::_project2d_one.
//
// Roofline: HBM.  fwd 40 B in, 64 B out (+4 B tile count) per surfel; bwd 40+68 B in, 40 B out.
#include "hgs_common.cuh"
#include "hgs_constants.cuh"
#include "project2d_math.cuh"
#include "../../include/hgs_raster.h"

namespace {

constexpr int PB = 256;

__global__ void __launch_bounds__(PB) project2d_fwd_kernel(
    const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
    const float* __restrict__ viewmats, const float* __restrict__ Ks, int N, int W, int H, float near_plane,
    float far_plane, float radius_clip, int tile_size, int tile_w, int tile_h, int32_t* __restrict__ radii,
    float* __restrict__ means2d, float* __restrict__ depths, float* __restrict__ ray_transforms,
    float* __restrict__ normals, int32_t* __restrict__ tiles_per_gauss) {
    __shared__ float s_a[PB * 3];
    __shared__ float s_b[PB * 3];
    const int c = blockIdx.y;
    const long long base = (long long)blockIdx.x * PB;
    const long long n = base + threadIdx.x;
    block_load_rows3<PB>(means, base, N, s_a);
    block_load_rows3<PB>(scales, base, N, s_b);
    __syncthreads();
    const HgsCam cam = hgs_load_cam(viewmats, Ks, c);
    int radius_i = 0, ntiles = 0;
    float o_m2x = 0.f, o_m2y = 0.f, o_depth = 0.f;
    float o_rt[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float o_n[3] = {0.f, 0.f, 0.f};
    if (n < N) {
        const float4 qv = reinterpret_cast<const float4*>(quats)[n];
        Proj2dFwd f;
        bool ok = proj2d_math(cam, s_a[threadIdx.x * 3], s_a[threadIdx.x * 3 + 1], s_a[threadIdx.x * 3 + 2], qv.x, qv.y,
                              qv.z, qv.w, s_b[threadIdx.x * 3], s_b[threadIdx.x * 3 + 1], near_plane, far_plane, f);
        if (ok) {
            const float tmpx = f.f[0] * f.M0[0] * f.M0[0] + f.f[1] * f.M0[1] * f.M0[1] + f.f[2] * f.M0[2] * f.M0[2];
            const float tmpy = f.f[0] * f.M1[0] * f.M1[0] + f.f[1] * f.M1[1] * f.M1[1] + f.f[2] * f.M1[2] * f.M1[2];
            const float hx = f.m2x * f.m2x - tmpx;
            const float hy = f.m2y * f.m2y - tmpy;
            const float radius = ceilf(HGS_RADIUS_SIGMA * sqrtf(fmaxf(fmaxf(hx, hy), HGS_RADIUS_FLOOR_2DGS)));
            bool vis = !(radius <= radius_clip);
            vis = vis && !(f.m2x + radius <= 0.f || f.m2x - radius >= (float)W || f.m2y + radius <= 0.f ||
                           f.m2y - radius >= (float)H);
            if (vis) {
                radius_i = (int)radius;
                o_m2x = f.m2x; o_m2y = f.m2y; o_depth = f.mc[2];
#pragma unroll
                for (int j = 0; j < 3; ++j) { o_rt[j] = f.M0[j]; o_rt[3 + j] = f.M1[j]; o_rt[6 + j] = f.M2[j]; }
                const float dotv = -f.RQ[0][2] * f.mc[0] + -f.RQ[1][2] * f.mc[1] + -f.RQ[2][2] * f.mc[2];
                const float sign = dotv > 0.f ? 1.0f : -1.0f;
#pragma unroll
                for (int i = 0; i < 3; ++i) o_n[i] = f.RQ[i][2] * sign;
                if (tiles_per_gauss != nullptr && radius_i > 0) {
                    int x0, y0, x1, y1;
                    hgs_tile_bbox(o_m2x, o_m2y, (float)radius_i, (float)tile_size, tile_w, tile_h, x0, y0, x1, y1);
                    ntiles = (y1 - y0) * (x1 - x0);
                }
            }
        }
        const long long idx = (long long)c * N + n;
        radii[idx] = radius_i;
        reinterpret_cast<float2*>(means2d)[idx] = make_float2(o_m2x, o_m2y);
        depths[idx] = o_depth;
        if (tiles_per_gauss != nullptr) tiles_per_gauss[idx] = ntiles;
        // 36-byte rows: each thread writes its own row (L2 merges the sectors)
        float* rt = ray_transforms + idx * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) rt[k] = o_rt[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) s_a[threadIdx.x * 3 + k] = o_n[k];
    __syncthreads();
    block_store_rows3<PB>(normals + (long long)c * N * 3, base, N, s_a);
}

// dense variant: one thread per surfel, cameras looped in-thread (deterministic)
__global__ void __launch_bounds__(PB) project2d_bwd_kernel(
    const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
    const float* __restrict__ viewmats, const float* __restrict__ Ks, int C, int N, float near_plane, float far_plane,
    const int32_t* __restrict__ radii, const float* __restrict__ v_means2d, int ld_m2,
    const float* __restrict__ v_depths, int ld_d, const float* __restrict__ v_ray_transforms, int ld_rt,
    const float* __restrict__ v_normals, int ld_n, float* __restrict__ v_means, float* __restrict__ v_quats,
    float* __restrict__ v_scales, int acc) {
    const long long n = (long long)blockIdx.x * PB + threadIdx.x;
    if (n >= N) return;
    bool any = false;
    for (int c = 0; c < C; ++c) any |= radii[(long long)c * N + n] > 0;
    float g_mean[3] = {0.f, 0.f, 0.f}, g_scale[3] = {0.f, 0.f, 0.f}, g_quat[4] = {0.f, 0.f, 0.f, 0.f};
    if (any) {
        const float px = means[n * 3], py = means[n * 3 + 1], pz = means[n * 3 + 2];
        const float s0 = scales[n * 3], s1 = scales[n * 3 + 1];
        const float4 qv = reinterpret_cast<const float4*>(quats)[n];
        for (int c = 0; c < C; ++c) {
            const long long idx = (long long)c * N + n;
            if (radii[idx] <= 0) continue;
            const HgsCam cam = hgs_load_cam(viewmats, Ks, c);
            Proj2dFwd f;
            if (!proj2d_math(cam, px, py, pz, qv.x, qv.y, qv.z, qv.w, s0, s1, near_plane, far_plane, f)) continue;
            proj2d_bwd_one(cam, f, s0, s1, v_means2d, ld_m2, v_depths, ld_d, v_ray_transforms, ld_rt, v_normals, ld_n,
                           idx, g_mean, g_scale, g_quat);
        }
    }
    reinterpret_cast<float4*>(v_quats)[n] = make_float4(g_quat[0], g_quat[1], g_quat[2], g_quat[3]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        v_means[n * 3 + k] = acc ? v_means[n * 3 + k] + g_mean[k] : g_mean[k];
        v_scales[n * 3 + k] = g_scale[k];
    }
}

// work-list variant: one thread per visible (camera, surfel) pair; outputs zero-filled by the launcher
__global__ void __launch_bounds__(PB) project2d_bwd_vis_kernel(
    const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
    const float* __restrict__ viewmats, const float* __restrict__ Ks, int C, int N, float near_plane, float far_plane,
    const int32_t* __restrict__ vis_ids, long long n_vis, const float* __restrict__ v_means2d, int ld_m2,
    const float* __restrict__ v_depths, int ld_d, const float* __restrict__ v_ray_transforms, int ld_rt,
    const float* __restrict__ v_normals, int ld_n, float* __restrict__ v_means, float* __restrict__ v_quats,
    float* __restrict__ v_scales, int acc) {
    const long long j = (long long)blockIdx.x * PB + threadIdx.x;
    if (j >= n_vis) return;
    const long long idx = vis_ids[j];
    const int c = (int)(idx / N);
    const long long n = idx - (long long)c * N;
    const float px = means[n * 3], py = means[n * 3 + 1], pz = means[n * 3 + 2];
    const float s0 = scales[n * 3], s1 = scales[n * 3 + 1];
    const float4 qv = reinterpret_cast<const float4*>(quats)[n];
    float g_mean[3] = {0.f, 0.f, 0.f}, g_scale[3] = {0.f, 0.f, 0.f}, g_quat[4] = {0.f, 0.f, 0.f, 0.f};
    const HgsCam cam = hgs_load_cam(viewmats, Ks, c);
    Proj2dFwd f;
    if (!proj2d_math(cam, px, py, pz, qv.x, qv.y, qv.z, qv.w, s0, s1, near_plane, far_plane, f)) return;
    proj2d_bwd_one(cam, f, s0, s1, v_means2d, ld_m2, v_depths, ld_d, v_ray_transforms, ld_rt, v_normals, ld_n, idx,
                   g_mean, g_scale, g_quat);
    if (C == 1) {
        reinterpret_cast<float4*>(v_quats)[n] = make_float4(g_quat[0], g_quat[1], g_quat[2], g_quat[3]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v_means[n * 3 + k] = acc ? v_means[n * 3 + k] + g_mean[k] : g_mean[k];
            v_scales[n * 3 + k] = g_scale[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(v_quats + n * 4 + k, g_quat[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicAdd(v_means + n * 3 + k, g_mean[k]);
            atomicAdd(v_scales + n * 3 + k, g_scale[k]);
        }
    }
}

}  // namespace

HGS_API int hgs_project2d_fwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                              const float* Ks, int C, int N, int width, int height, float near_plane, float far_plane,
                              float radius_clip, int tile_size, int32_t* radii, float* means2d, float* depths,
                              float* ray_transforms, float* normals, int32_t* tiles_per_gauss, void* stream) {
    if (C <= 0 || N < 0 || width <= 0 || height <= 0 || tile_size <= 0) return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    const int tile_w = (width + tile_size - 1) / tile_size, tile_h = (height + tile_size - 1) / tile_size;
    dim3 grid(hgs_ceil_div(N, PB), C);
    project2d_fwd_kernel<<<grid, PB, 0, (cudaStream_t)stream>>>(means, quats, scales, viewmats, Ks, N, width, height,
                                                                  near_plane, far_plane, radius_clip, tile_size, tile_w,
                                                                  tile_h, radii, means2d, depths, ray_transforms,
                                                                  normals, tiles_per_gauss);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_project2d_bwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                              const float* Ks, int C, int N, int width, int height, float near_plane, float far_plane,
                              const int32_t* radii, const float* v_means2d, int ld_means2d, const float* v_depths,
                              int ld_depths, const float* v_ray_transforms, int ld_ray_transforms,
                              const float* v_normals, int ld_normals, const int32_t* vis_ids, long long n_vis,
                              float* v_means, float* v_quats, float* v_scales, int flags, void* stream) {
    (void)width; (void)height;
    if (C <= 0 || N < 0 || n_vis < 0 || ld_means2d < 2 || ld_depths < 1 || ld_ray_transforms < 9 || ld_normals < 3)
        return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (vis_ids != nullptr) {
        cudaError_t e;
        const int accumulate_means = flags & 1, zeroed = flags & 2;   // see include/hgs_raster.h
        if (!accumulate_means && !zeroed && (e = cudaMemsetAsync(v_means, 0, (size_t)N * 3 * sizeof(float), st)) != cudaSuccess)
            return (int)e;
        if (!zeroed && (e = cudaMemsetAsync(v_quats, 0, (size_t)N * 4 * sizeof(float), st)) != cudaSuccess) return (int)e;
        if (!zeroed && (e = cudaMemsetAsync(v_scales, 0, (size_t)N * 3 * sizeof(float), st)) != cudaSuccess) return (int)e;
        if (n_vis == 0) return 0;
        project2d_bwd_vis_kernel<<<hgs_ceil_div(n_vis, PB), PB, 0, st>>>(
            means, quats, scales, viewmats, Ks, C, N, near_plane, far_plane, vis_ids, n_vis, v_means2d, ld_means2d,
            v_depths, ld_depths, v_ray_transforms, ld_ray_transforms, v_normals, ld_normals, v_means, v_quats, v_scales, accumulate_means);
        HGS_LAUNCH_CHECK();
        return 0;
    }
    project2d_bwd_kernel<<<hgs_ceil_div(N, PB), PB, 0, st>>>(
        means, quats, scales, viewmats, Ks, C, N, near_plane, far_plane, radii, v_means2d, ld_means2d, v_depths,
        ld_depths, v_ray_transforms, ld_ray_transforms, v_normals, ld_normals, v_means, v_quats, v_scales, flags & 1);
    HGS_LAUNCH_CHECK();
    return 0;
}
