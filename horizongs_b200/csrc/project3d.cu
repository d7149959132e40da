// Stage a3: fused 3D-Gaussian projection (world -> camera, quat+scale -> covariance,
// perspective Jacobian, 2D covariance + eps2d, conic, integer radius, culling) with the
// tile-count of stage a8 folded in, and its backward.
//
// Replaces gsplat's fully_fused_projection as reached from the reference at
// gaussian_renderer/render.py:149-165 (prefilter_voxel) and inside gsplat.rasterization
// (render.py:40-54).  Built with -fmad=false: every intermediate is rounded exactly like
// oracle/gsplat_oracle.py::_project3d_one, so radii (integers) are bit-identical to the oracle.
//
// Roofline: HBM.  fwd 40 B in + 28 B out (+4 B tile count) per Gaussian; bwd 40+24(+28) B in, 40 B out.
#include "hgs_common.cuh"
#include "hgs_constants.cuh"
#include "project3d_math.cuh"

namespace {

constexpr int PB = 256;  // threads per block
// Forward kernel in three phases so that the heavy math runs on densely populated warps even when only a
// small, randomly scattered fraction of the Gaussians is on screen:
//   1. every thread: camera-space centre of its Gaussian, near/far test, conservative off-screen test;
//   2. the survivors of the block are compacted and the first n_pass threads do the covariance /
//      Jacobian / conic / radius math, results go to shared memory;
//   3. every thread writes its own output row (zeros for culled Gaussians), coalesced.
__global__ void __launch_bounds__(PB) project3d_fwd_kernel(
    const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
    const float* __restrict__ viewmats, const float* __restrict__ Ks, int N, int W, int H, float eps2d,
    float near_plane, float far_plane, float radius_clip, int tile_size, int tile_w, int tile_h,
    int32_t* __restrict__ radii, float* __restrict__ means2d, float* __restrict__ depths, float* __restrict__ conics,
    float* __restrict__ compensations, int32_t* __restrict__ tiles_per_gauss) {
    __shared__ float s_a[PB * 3];        // means in, conics out
    __shared__ float s_b[PB * 3];        // scales
    __shared__ float s_pc[PB * 3];       // camera-space centres of the survivors
    __shared__ float s_out[PB * 4];      // m2x, m2y, depth, compensation
    __shared__ int s_ri[PB * 2];         // radius, tile count
    __shared__ int s_list[PB];
    __shared__ int s_wcnt[PB / 32];
    __shared__ float s_cam[26];          // viewmat (16), K (9), bound coefficient
    const int c = blockIdx.y;
    const long long base = (long long)blockIdx.x * PB;
    const long long n = base + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    block_load_rows3<PB>(means, base, N, s_a);
    block_load_rows3<PB>(scales, base, N, s_b);
    // defaults: culled
    s_ri[threadIdx.x * 2] = 0; s_ri[threadIdx.x * 2 + 1] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s_out[threadIdx.x * 4 + k] = 0.f;
    // camera: loaded once per block, the derived bound coefficient computed by one thread
    if (threadIdx.x < 16) s_cam[threadIdx.x] = viewmats[c * 16 + threadIdx.x];
    else if (threadIdx.x < 25) s_cam[threadIdx.x] = Ks[c * 9 + threadIdx.x - 16];
    __syncthreads();
    HgsCam cam;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) cam.R[i][j] = s_cam[i * 4 + j];
        cam.t[i] = s_cam[i * 4 + 3];
    }
    cam.fx = s_cam[16]; cam.fy = s_cam[20]; cam.cx = s_cam[18]; cam.cy = s_cam[21];
    if (threadIdx.x == 0) s_cam[25] = proj3d_jf_coeff(cam, (float)W, (float)H);
    __syncthreads();
    const float jf_coeff = s_cam[25];
    const float (*R)[3] = cam.R;

    // ---- phase 1
    bool pass = false;
    if (n < N) {
        const float px = s_a[threadIdx.x * 3 + 0], py = s_a[threadIdx.x * 3 + 1], pz = s_a[threadIdx.x * 3 + 2];
        const float s0 = s_b[threadIdx.x * 3 + 0], s1 = s_b[threadIdx.x * 3 + 1], s2 = s_b[threadIdx.x * 3 + 2];
        const float zc = R[2][0] * px + R[2][1] * py + R[2][2] * pz + cam.t[2];
        if (!(zc < near_plane || zc > far_plane)) {
            const float xc = R[0][0] * px + R[0][1] * py + R[0][2] * pz + cam.t[0];
            const float yc = R[1][0] * px + R[1][1] * py + R[1][2] * pz + cam.t[1];
            pass = !proj3d_surely_offscreen(cam, jf_coeff, xc, yc, zc, fmaxf(fabsf(s0), fmaxf(fabsf(s1), fabsf(s2))),
                                            (float)W, (float)H, eps2d);
            if (pass) { s_pc[threadIdx.x * 3] = xc; s_pc[threadIdx.x * 3 + 1] = yc; s_pc[threadIdx.x * 3 + 2] = zc; }
        }
    }
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, pass);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int wbase = 0, n_pass = 0;
#pragma unroll
    for (int w = 0; w < PB / 32; ++w) {
        if (w < warp) wbase += s_wcnt[w];
        n_pass += s_wcnt[w];
    }
    if (pass) s_list[wbase + __popc(bal & ((1u << lane) - 1u))] = threadIdx.x;
    __syncthreads();

    // ---- phase 2: dense math on the survivors
    float o_ca = 0.f, o_cb = 0.f, o_cc = 0.f;   // conics of the row this thread PROCESSED (written to s_a below)
    int my_row = -1;
    if (threadIdx.x < n_pass) {
        const int r = s_list[threadIdx.x];
        my_row = r;
        const float4 qv = reinterpret_cast<const float4*>(quats)[base + r];
        Proj3dFwd f;
        f.xc = s_pc[r * 3]; f.yc = s_pc[r * 3 + 1]; f.zc = s_pc[r * 3 + 2];
        proj3d_cov_and_project(cam, qv.x, qv.y, qv.z, qv.w, s_b[r * 3], s_b[r * 3 + 1], s_b[r * 3 + 2], (float)W,
                               (float)H, eps2d, f);
        if (f.det > 0.f) {
            const float inv_det = 1.0f / f.det;
            const float b = 0.5f * (f.c00 + f.c11);
            const float v1 = b + sqrtf(fmaxf(b * b - f.det, HGS_EIG_FLOOR));
            const float radius = ceilf(HGS_RADIUS_SIGMA * sqrtf(v1));
            bool vis = !(radius <= radius_clip);
            vis = vis && !(f.m2x + radius <= 0.f || f.m2x - radius >= (float)W || f.m2y + radius <= 0.f ||
                           f.m2y - radius >= (float)H);
            if (vis) {
                const int radius_i = (int)radius;
                int ntiles = 0;
                if (tiles_per_gauss != nullptr && radius_i > 0) {
                    int x0, y0, x1, y1;
                    hgs_tile_bbox(f.m2x, f.m2y, (float)radius_i, (float)tile_size, tile_w, tile_h, x0, y0, x1, y1);
                    ntiles = (y1 - y0) * (x1 - x0);
                }
                s_ri[r * 2] = radius_i; s_ri[r * 2 + 1] = ntiles;
                s_out[r * 4] = f.m2x; s_out[r * 4 + 1] = f.m2y; s_out[r * 4 + 2] = f.zc;
                s_out[r * 4 + 3] = sqrtf(fmaxf(f.det_orig / f.det, 0.f));
                o_ca = f.c11 * inv_det;
                o_cb = -f.c01 * inv_det;
                o_cc = f.c00 * inv_det;
            }
        }
    }
    __syncthreads();   // everyone is done reading means (s_a): reuse it for the conics
    s_a[threadIdx.x * 3 + 0] = 0.f; s_a[threadIdx.x * 3 + 1] = 0.f; s_a[threadIdx.x * 3 + 2] = 0.f;
    __syncthreads();
    if (my_row >= 0) { s_a[my_row * 3 + 0] = o_ca; s_a[my_row * 3 + 1] = o_cb; s_a[my_row * 3 + 2] = o_cc; }
    __syncthreads();

    // ---- phase 3: coalesced rows
    if (n < N) {
        const long long idx = (long long)c * N + n;
        radii[idx] = s_ri[threadIdx.x * 2];
        reinterpret_cast<float2*>(means2d)[idx] = make_float2(s_out[threadIdx.x * 4], s_out[threadIdx.x * 4 + 1]);
        depths[idx] = s_out[threadIdx.x * 4 + 2];
        if (compensations != nullptr) compensations[idx] = s_out[threadIdx.x * 4 + 3];
        if (tiles_per_gauss != nullptr) tiles_per_gauss[idx] = s_ri[threadIdx.x * 2 + 1];
    }
    block_store_rows3<PB>(conics + (long long)c * N * 3, base, N, s_a);
}

#define PROJ3D_BWD_LOAD_AND_RUN(IDX, N_, C_)                                                                         \
    {                                                                                                                \
        const HgsCam cam = hgs_load_cam(viewmats, Ks, (C_));                                                         \
        Proj3dFwd f;                                                                                                 \
        if (proj3d_math(cam, px, py, pz, qv.x, qv.y, qv.z, qv.w, s0, s1, s2, (float)W, (float)H, eps2d, near_plane,  \
                        far_plane, f)) {                                                                             \
            const float2 vm = make_float2(v_means2d[(IDX) * ld_m2], v_means2d[(IDX) * ld_m2 + 1]);                   \
            const float vd = v_depths != nullptr ? v_depths[(IDX) * ld_d] : 0.f;                                     \
            proj3d_bwd_one(cam, f, s0, s1, s2, vm, vd, v_conics[(IDX) * ld_c], 0.5f * v_conics[(IDX) * ld_c + 1],    \
                           v_conics[(IDX) * ld_c + 2], g_mean, g_scale, g_quat);                                     \
        }                                                                                                            \
    }

// Dense variant: one thread per Gaussian, loop over cameras: gradients w.r.t. means/quats/scales are summed
// over the C views without atomics (deterministic).  Gaussians culled in every view only write zeros.
__global__ void __launch_bounds__(PB) project3d_bwd_kernel(
    const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
    const float* __restrict__ viewmats, const float* __restrict__ Ks, int C, int N, int W, int H, float eps2d,
    float near_plane, float far_plane, const int32_t* __restrict__ radii, const float* __restrict__ v_means2d,
    int ld_m2, const float* __restrict__ v_depths, int ld_d, const float* __restrict__ v_conics, int ld_c,
    float* __restrict__ v_means, float* __restrict__ v_quats, float* __restrict__ v_scales, int acc) {
    const long long n = (long long)blockIdx.x * PB + threadIdx.x;
    if (n >= N) return;
    bool any = false;
    for (int c = 0; c < C; ++c) any |= radii[(long long)c * N + n] > 0;
    float px = 0.f, py = 0.f, pz = 0.f, s0 = 1.f, s1 = 1.f, s2 = 1.f;
    float4 qv = make_float4(1.f, 0.f, 0.f, 0.f);
    if (any) {
        px = means[n * 3]; py = means[n * 3 + 1]; pz = means[n * 3 + 2];
        s0 = scales[n * 3]; s1 = scales[n * 3 + 1]; s2 = scales[n * 3 + 2];
        qv = reinterpret_cast<const float4*>(quats)[n];
    }
    float g_mean[3] = {0.f, 0.f, 0.f};
    float g_scale[3] = {0.f, 0.f, 0.f};
    float g_quat[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C && any; ++c) {
        const long long idx = (long long)c * N + n;
        if (radii[idx] <= 0) continue;
        PROJ3D_BWD_LOAD_AND_RUN(idx, N, c)
    }
    reinterpret_cast<float4*>(v_quats)[n] = make_float4(g_quat[0], g_quat[1], g_quat[2], g_quat[3]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        v_means[n * 3 + k] = acc ? v_means[n * 3 + k] + g_mean[k] : g_mean[k];
        v_scales[n * 3 + k] = g_scale[k];
    }
}

// Work-list variant: one thread per VISIBLE (camera, Gaussian) pair (vis_ids = flat indices c*N+n with
// radii > 0, from hgs_isect_prepare), so warps are fully populated even when most Gaussians are culled.
// Outputs are zero-filled by the launcher; with one camera every row is written once (plain stores), with
// several cameras contributions are added atomically.
__global__ void __launch_bounds__(PB) project3d_bwd_vis_kernel(
    const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
    const float* __restrict__ viewmats, const float* __restrict__ Ks, int C, int N, int W, int H, float eps2d,
    float near_plane, float far_plane, const int32_t* __restrict__ vis_ids, long long n_vis,
    const float* __restrict__ v_means2d, int ld_m2, const float* __restrict__ v_depths, int ld_d,
    const float* __restrict__ v_conics, int ld_c, float* __restrict__ v_means, float* __restrict__ v_quats,
    float* __restrict__ v_scales, int acc) {
    const long long j = (long long)blockIdx.x * PB + threadIdx.x;
    if (j >= n_vis) return;
    const long long idx = vis_ids[j];
    const int c = (int)(idx / N);
    const long long n = idx - (long long)c * N;
    const float px = means[n * 3], py = means[n * 3 + 1], pz = means[n * 3 + 2];
    const float s0 = scales[n * 3], s1 = scales[n * 3 + 1], s2 = scales[n * 3 + 2];
    const float4 qv = reinterpret_cast<const float4*>(quats)[n];
    float g_mean[3] = {0.f, 0.f, 0.f};
    float g_scale[3] = {0.f, 0.f, 0.f};
    float g_quat[4] = {0.f, 0.f, 0.f, 0.f};
    PROJ3D_BWD_LOAD_AND_RUN(idx, N, c)
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

// SURVEY.md section 8(f1), first half: LOD level test (scene/lod_model.py:286-290 set_anchor_mask with
// basic_model.py:192-203 map_to_int_level) fused with the anchor prefilter (gaussian_renderer/render.py:120-197:
// project the anchors as Gaussians with scales = scaling[:, :3], keep radii > 0) -- one pass over the anchors
// instead of ~6 elementwise kernels, a boolean gather, a projection launch whose other outputs are thrown away and
// a boolean scatter.  The projection arithmetic is project3d_fwd_kernel's (same device functions, same -fmad=false).
// level_mode: 0 floor, 1 round, 2 ceil.  visible[a] = 1 iff level[a] <= int_level(a) and the anchor's radius > 0.
__global__ void __launch_bounds__(PB) anchor_filter_kernel(
    const float* __restrict__ anchor, const int32_t* __restrict__ level, const float* __restrict__ extra_level,
    const float* __restrict__ scaling, int ld_scaling, const float* __restrict__ rotation,
    const float* __restrict__ cam_center, float resolution_scale, float standard_dist, float inv_log2_fork, int max_level,
    int level_mode, const float* __restrict__ viewmat, const float* __restrict__ Kmat, int N, int W, int H, float eps2d,
    float near_plane, float far_plane, float radius_clip, uint8_t* __restrict__ visible) {
    const long long n = (long long)blockIdx.x * PB + threadIdx.x;
    if (n >= N) return;
    const float px = anchor[n * 3], py = anchor[n * 3 + 1], pz = anchor[n * 3 + 2];
    bool vis = true;
    if (level != nullptr) {
        const float dx = px - cam_center[0], dy = py - cam_center[1], dz = pz - cam_center[2];
        const float dist = sqrtf(dx * dx + dy * dy + dz * dz) * resolution_scale;
        const float pred = log2f(standard_dist / dist) * inv_log2_fork + (extra_level != nullptr ? extra_level[n] : 0.f);
        const float q = level_mode == 0 ? floorf(pred) : (level_mode == 1 ? nearbyintf(pred) : ceilf(pred));
        const int il = (int)fminf(fmaxf(q, 0.f), (float)max_level);
        vis = level[n] <= il;
    }
    if (vis) {
        const HgsCam cam = hgs_load_cam(viewmat, Kmat, 0);
        const float s0 = scaling[n * ld_scaling], s1 = scaling[n * ld_scaling + 1], s2 = scaling[n * ld_scaling + 2];
        const float4 qv = reinterpret_cast<const float4*>(rotation)[n];
        Proj3dFwd f;
        vis = proj3d_math(cam, px, py, pz, qv.x, qv.y, qv.z, qv.w, s0, s1, s2, (float)W, (float)H, eps2d, near_plane,
                          far_plane, f);
        if (vis) {
            vis = false;
            if (f.det > 0.f) {
                const float b = 0.5f * (f.c00 + f.c11);
                const float v1 = b + sqrtf(fmaxf(b * b - f.det, HGS_EIG_FLOOR));
                const float radius = ceilf(HGS_RADIUS_SIGMA * sqrtf(v1));
                vis = !(radius <= radius_clip) &&
                      !(f.m2x + radius <= 0.f || f.m2x - radius >= (float)W || f.m2y + radius <= 0.f ||
                        f.m2y - radius >= (float)H) &&
                      (int)radius > 0;
            }
        }
    }
    visible[n] = vis ? 1 : 0;
}

}  // namespace

#include "../../include/hgs_raster.h"

HGS_API int hgs_project3d_fwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                              const float* Ks, int C, int N, int width, int height, float eps2d, float near_plane,
                              float far_plane, float radius_clip, int tile_size, int32_t* radii, float* means2d,
                              float* depths, float* conics, float* compensations, int32_t* tiles_per_gauss,
                              void* stream) {
    if (C <= 0 || N < 0 || width <= 0 || height <= 0 || tile_size <= 0) return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    const int tile_w = (width + tile_size - 1) / tile_size, tile_h = (height + tile_size - 1) / tile_size;
    dim3 grid(hgs_ceil_div(N, PB), C);
    project3d_fwd_kernel<<<grid, PB, 0, (cudaStream_t)stream>>>(means, quats, scales, viewmats, Ks, N, width, height,
                                                                  eps2d, near_plane, far_plane, radius_clip, tile_size,
                                                                  tile_w, tile_h, radii, means2d, depths, conics,
                                                                  compensations, tiles_per_gauss);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_project3d_bwd(const float* means, const float* quats, const float* scales, const float* viewmats,
                              const float* Ks, int C, int N, int width, int height, float eps2d, float near_plane,
                              float far_plane, const int32_t* radii, const float* v_means2d, int ld_means2d,
                              const float* v_depths, int ld_depths, const float* v_conics, int ld_conics,
                              const int32_t* vis_ids, long long n_vis, float* v_means, float* v_quats,
                              float* v_scales, int flags, void* stream) {
    if (C <= 0 || N < 0 || width <= 0 || height <= 0 || ld_means2d < 2 || ld_conics < 3 || ld_depths < 1 || n_vis < 0)
        return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    if (vis_ids != nullptr) {
        cudaStream_t st = (cudaStream_t)stream;
        cudaError_t e;
        const int accumulate_means = flags & 1, zeroed = flags & 2;   // see include/hgs_raster.h
        if (!accumulate_means && !zeroed && (e = cudaMemsetAsync(v_means, 0, (size_t)N * 3 * sizeof(float), st)) != cudaSuccess)
            return (int)e;
        if (!zeroed && (e = cudaMemsetAsync(v_quats, 0, (size_t)N * 4 * sizeof(float), st)) != cudaSuccess) return (int)e;
        if (!zeroed && (e = cudaMemsetAsync(v_scales, 0, (size_t)N * 3 * sizeof(float), st)) != cudaSuccess) return (int)e;
        if (n_vis == 0) return 0;
        project3d_bwd_vis_kernel<<<hgs_ceil_div(n_vis, PB), PB, 0, st>>>(
            means, quats, scales, viewmats, Ks, C, N, width, height, eps2d, near_plane, far_plane, vis_ids, n_vis,
            v_means2d, ld_means2d, v_depths, ld_depths, v_conics, ld_conics, v_means, v_quats, v_scales, accumulate_means);
        HGS_LAUNCH_CHECK();
        return 0;
    }
    project3d_bwd_kernel<<<hgs_ceil_div(N, PB), PB, 0, (cudaStream_t)stream>>>(
        means, quats, scales, viewmats, Ks, C, N, width, height, eps2d, near_plane, far_plane, radii, v_means2d,
        ld_means2d, v_depths, ld_depths, v_conics, ld_conics, v_means, v_quats, v_scales, flags & 1);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_anchor_filter(const float* anchor, const int32_t* level, const float* extra_level, const float* scaling,
                              int ld_scaling, const float* rotation, const float* cam_center, float resolution_scale,
                              float standard_dist, float fork, int max_level, int level_mode, const float* viewmat,
                              const float* Kmat, int N, int width, int height, float eps2d, float near_plane,
                              float far_plane, float radius_clip, uint8_t* visible, void* stream) {
    if (N < 0 || width <= 0 || height <= 0 || ld_scaling < 3 || anchor == nullptr || scaling == nullptr ||
        rotation == nullptr || viewmat == nullptr || Kmat == nullptr || visible == nullptr || level_mode < 0 ||
        level_mode > 2 || (level != nullptr && (cam_center == nullptr || !(fork > 1.f) || max_level < 0)) ||
        (reinterpret_cast<size_t>(rotation) & 15))
        return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    anchor_filter_kernel<<<hgs_ceil_div(N, PB), PB, 0, (cudaStream_t)stream>>>(
        anchor, level, extra_level, scaling, ld_scaling, rotation, cam_center, resolution_scale, standard_dist,
        level != nullptr ? 1.0f / log2f(fork) : 0.f, max_level, level_mode, viewmat, Kmat, N, width, height, eps2d,
        near_plane, far_plane, radius_clip, visible);
    HGS_LAUNCH_CHECK();
    return 0;
}
