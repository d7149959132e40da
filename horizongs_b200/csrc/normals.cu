// Stage a13 (2DGS post-ops of gsplat.rasterization_2dgs, reference call site gaussian_renderer/render.py:56-76):
//   render_normals (camera frame, out of the blend kernel)  ->  world frame:  n_w = R_c2w n_c
//   render_normals_from_depth = depth_to_normal(depth, c2w, K): back-project every pixel centre with its z-depth,
//       p = t_c2w + depth * R_c2w ((x + .5 - cx) / fx, (y + .5 - cy) / fy, 1), central differences
//       dx = p[y+1][x] - p[y-1][x], dy = p[y][x+1] - p[y][x-1], n = normalize(dx x dy) (F.normalize, eps 1e-12), one-pixel
//       zero border -- oracle/gsplat_oracle.py::depth_to_normal.
// One forward and one backward kernel instead of ~40 eager PyTorch launches (meshgrid, two einsum GEMMs, cross,
// normalize, pad, cuSOLVER inverse); the camera-to-world transform is the closed form of the rigid world-to-camera
// matrix (R^T, -R^T t).  HBM-bound: 32 B read + 24 B written per pixel forward.
#include "hgs_common.cuh"
#include "../../include/hgs_raster.h"

namespace {

constexpr int NB = 256;

struct C2W {
    float R[3][3];   // camera -> world rotation
    float t[3];
    float fx, fy, cx, cy;
};
__device__ __forceinline__ C2W load_c2w(const float* __restrict__ viewmats, const float* __restrict__ Ks, int c) {
    const float* V = viewmats + c * 16;
    const float* K = Ks + c * 9;
    C2W m;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) m.R[i][j] = V[j * 4 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) m.t[i] = -(m.R[i][0] * V[3] + m.R[i][1] * V[7] + m.R[i][2] * V[11]);
    m.fx = K[0]; m.fy = K[4]; m.cx = K[2]; m.cy = K[5];
    return m;
}
// world-space direction of pixel (x, y) per unit z-depth
__device__ __forceinline__ void pixel_dir(const C2W& m, int x, int y, float d[3]) {
    const float u = ((float)x - m.cx + 0.5f) / m.fx, v = ((float)y - m.cy + 0.5f) / m.fy;
#pragma unroll
    for (int i = 0; i < 3; ++i) d[i] = m.R[i][0] * u + m.R[i][1] * v + m.R[i][2];
}
__device__ __forceinline__ void pixel_point(const C2W& m, const float* __restrict__ depth, int ld, long long img, int W,
                                            int x, int y, float p[3]) {
    float d[3];
    pixel_dir(m, x, y, d);
    const float z = depth[(img + (long long)y * W + x) * ld];
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = m.t[i] + z * d[i];
}
__device__ __forceinline__ void cross3(const float a[3], const float b[3], float c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

__global__ void __launch_bounds__(NB) normals_post_fwd_kernel(const float* __restrict__ normals_cam,
                                                              const float* __restrict__ depth, int ld_depth,
                                                              const float* __restrict__ viewmats,
                                                              const float* __restrict__ Ks, int C, int H, int W,
                                                              float* __restrict__ normals_world,
                                                              float* __restrict__ normals_from_depth) {
    const long long pix = (long long)blockIdx.x * NB + threadIdx.x;
    const long long HW = (long long)H * W;
    if (pix >= (long long)C * HW) return;
    const int c = (int)(pix / HW);
    const long long r = pix - (long long)c * HW;
    const int y = (int)(r / W), x = (int)(r - (long long)y * W);
    const C2W m = load_c2w(viewmats, Ks, c);
    if (normals_cam != nullptr) {
        const float n0 = normals_cam[pix * 3], n1 = normals_cam[pix * 3 + 1], n2 = normals_cam[pix * 3 + 2];
#pragma unroll
        for (int i = 0; i < 3; ++i) normals_world[pix * 3 + i] = m.R[i][0] * n0 + m.R[i][1] * n1 + m.R[i][2] * n2;
    }
    if (normals_from_depth != nullptr) {
        float n[3] = {0.f, 0.f, 0.f};
        if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
            const long long img = (long long)c * HW;
            float pu[3], pd[3], pl[3], pr[3], dx[3], dy[3], cr[3];
            pixel_point(m, depth, ld_depth, img, W, x, y + 1, pd);
            pixel_point(m, depth, ld_depth, img, W, x, y - 1, pu);
            pixel_point(m, depth, ld_depth, img, W, x + 1, y, pr);
            pixel_point(m, depth, ld_depth, img, W, x - 1, y, pl);
#pragma unroll
            for (int i = 0; i < 3; ++i) { dx[i] = pd[i] - pu[i]; dy[i] = pr[i] - pl[i]; }
            cross3(dx, dy, cr);
            const float inv = 1.0f / fmaxf(sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]), 1e-12f);
#pragma unroll
            for (int i = 0; i < 3; ++i) n[i] = cr[i] * inv;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) normals_from_depth[pix * 3 + i] = n[i];
    }
}

// gradients of the stencil of interior pixel q = (x, y) w.r.t. its dx and dy
__device__ __forceinline__ bool stencil_grads(const C2W& m, const float* __restrict__ depth, int ld, long long img, int H,
                                              int W, int x, int y, const float* __restrict__ v_nfd, float v_dx[3],
                                              float v_dy[3]) {
    if (!(x >= 1 && x < W - 1 && y >= 1 && y < H - 1)) return false;
    float pu[3], pd[3], pl[3], pr[3], dx[3], dy[3], cr[3];
    pixel_point(m, depth, ld, img, W, x, y + 1, pd);
    pixel_point(m, depth, ld, img, W, x, y - 1, pu);
    pixel_point(m, depth, ld, img, W, x + 1, y, pr);
    pixel_point(m, depth, ld, img, W, x - 1, y, pl);
#pragma unroll
    for (int i = 0; i < 3; ++i) { dx[i] = pd[i] - pu[i]; dy[i] = pr[i] - pl[i]; }
    cross3(dx, dy, cr);
    const float len = sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
    const float* v = v_nfd + (img + (long long)y * W + x) * 3;
    float g[3];
    if (len > 1e-12f) {
        const float inv = 1.0f / len;
        const float n[3] = {cr[0] * inv, cr[1] * inv, cr[2] * inv};
        const float dot = n[0] * v[0] + n[1] * v[1] + n[2] * v[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) g[i] = (v[i] - n[i] * dot) * inv;
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) g[i] = v[i] * 1e12f;      // F.normalize below its eps: x / eps
    }
    cross3(dy, g, v_dx);       // d(dx x dy . g) / d dx = dy x g
    cross3(g, dx, v_dy);       //                 / d dy = g x dx
    return true;
}

__global__ void __launch_bounds__(NB) normals_post_bwd_kernel(const float* __restrict__ depth, int ld_depth,
                                                              const float* __restrict__ viewmats,
                                                              const float* __restrict__ Ks, int C, int H, int W,
                                                              const float* __restrict__ v_normals_world,
                                                              const float* __restrict__ v_nfd,
                                                              float* __restrict__ v_normals_cam,
                                                              float* __restrict__ v_depth, int ld_v_depth) {
    const long long pix = (long long)blockIdx.x * NB + threadIdx.x;
    const long long HW = (long long)H * W;
    if (pix >= (long long)C * HW) return;
    const int c = (int)(pix / HW);
    const long long r = pix - (long long)c * HW;
    const int y = (int)(r / W), x = (int)(r - (long long)y * W);
    const C2W m = load_c2w(viewmats, Ks, c);
    if (v_normals_cam != nullptr) {
        float v[3] = {0.f, 0.f, 0.f};
        if (v_normals_world != nullptr) { v[0] = v_normals_world[pix * 3]; v[1] = v_normals_world[pix * 3 + 1]; v[2] = v_normals_world[pix * 3 + 2]; }
#pragma unroll
        for (int j = 0; j < 3; ++j) v_normals_cam[pix * 3 + j] = m.R[0][j] * v[0] + m.R[1][j] * v[1] + m.R[2][j] * v[2];
    }
    if (v_depth != nullptr) {
        float vp[3] = {0.f, 0.f, 0.f};     // gradient w.r.t. the back-projected point of this pixel
        if (v_nfd != nullptr) {
            const long long img = (long long)c * HW;
            float a[3], b[3];
            // this pixel is the "down" neighbour of (x, y-1), the "up" neighbour of (x, y+1), the "right" neighbour of
            // (x-1, y) and the "left" neighbour of (x+1, y)
            if (stencil_grads(m, depth, ld_depth, img, H, W, x, y - 1, v_nfd, a, b)) { vp[0] += a[0]; vp[1] += a[1]; vp[2] += a[2]; }
            if (stencil_grads(m, depth, ld_depth, img, H, W, x, y + 1, v_nfd, a, b)) { vp[0] -= a[0]; vp[1] -= a[1]; vp[2] -= a[2]; }
            if (stencil_grads(m, depth, ld_depth, img, H, W, x - 1, y, v_nfd, a, b)) { vp[0] += b[0]; vp[1] += b[1]; vp[2] += b[2]; }
            if (stencil_grads(m, depth, ld_depth, img, H, W, x + 1, y, v_nfd, a, b)) { vp[0] -= b[0]; vp[1] -= b[1]; vp[2] -= b[2]; }
        }
        float d[3];
        pixel_dir(m, x, y, d);
        v_depth[pix * ld_v_depth] = vp[0] * d[0] + vp[1] * d[1] + vp[2] * d[2];
    }
}

}  // namespace

HGS_API int hgs_normals_post_fwd(const float* normals_cam, const float* depth, int ld_depth, const float* viewmats,
                                 const float* Ks, int C, int H, int W, float* normals_world, float* normals_from_depth,
                                 void* stream) {
    if (C <= 0 || H <= 0 || W <= 0 || viewmats == nullptr || Ks == nullptr) return HGS_ERR_INVALID_ARG;
    if ((normals_cam == nullptr) != (normals_world == nullptr)) return HGS_ERR_INVALID_ARG;
    if (normals_from_depth != nullptr && (depth == nullptr || ld_depth < 1)) return HGS_ERR_INVALID_ARG;
    const long long n = (long long)C * H * W;
    normals_post_fwd_kernel<<<hgs_ceil_div(n, NB), NB, 0, (cudaStream_t)stream>>>(normals_cam, depth, ld_depth, viewmats, Ks,
                                                                                  C, H, W, normals_world,
                                                                                  normals_from_depth);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_normals_post_bwd(const float* depth, int ld_depth, const float* viewmats, const float* Ks, int C, int H,
                                 int W, const float* v_normals_world, const float* v_normals_from_depth,
                                 float* v_normals_cam, float* v_depth, int ld_v_depth, void* stream) {
    if (C <= 0 || H <= 0 || W <= 0 || viewmats == nullptr || Ks == nullptr) return HGS_ERR_INVALID_ARG;
    if (v_depth != nullptr && (depth == nullptr || ld_depth < 1 || ld_v_depth < 1)) return HGS_ERR_INVALID_ARG;
    const long long n = (long long)C * H * W;
    normals_post_bwd_kernel<<<hgs_ceil_div(n, NB), NB, 0, (cudaStream_t)stream>>>(depth, ld_depth, viewmats, Ks, C, H, W,
                                                                                  v_normals_world, v_normals_from_depth,
                                                                                  v_normals_cam, v_depth, ld_v_depth);
    HGS_LAUNCH_CHECK();
    return 0;
}
