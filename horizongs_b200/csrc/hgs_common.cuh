// Shared device/host helpers for the hgs_raster kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HGS_API extern "C" __attribute__((visibility("default")))

// custom (negative) status codes; positive values are cudaError_t
#define HGS_ERR_INVALID_ARG (-1)
#define HGS_ERR_TOO_LARGE (-2)
#define HGS_ERR_WORKSPACE (-3)

// cumulative number of kernel launches issued by this library (diagnostics only; see hgs_debug_launch_count)
extern unsigned long long g_hgs_launches;

#define HGS_LAUNCH_CHECK()                          \
    do {                                            \
        cudaError_t e__ = cudaGetLastError();       \
        if (e__ != cudaSuccess) return (int)e__;    \
        __atomic_fetch_add(&g_hgs_launches, 1ull, __ATOMIC_RELAXED); \
    } while (0)

static inline int hgs_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- AoS [n,3] <-> per-thread staging through shared memory (coalesced global access) ----
// Block of BLOCK threads owns rows [base, base+BLOCK).  s must hold BLOCK*3 floats.
template <int BLOCK>
__device__ __forceinline__ void block_load_rows3(const float* __restrict__ src, long long base, long long n_rows,
                                                 float* s) {
    const long long lim = n_rows * 3;
    const long long off = base * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int i = threadIdx.x + k * BLOCK;
        long long g = off + i;
        s[i] = (g < lim) ? src[g] : 0.f;
    }
}
template <int BLOCK>
__device__ __forceinline__ void block_store_rows3(float* __restrict__ dst, long long base, long long n_rows,
                                                  const float* s) {
    const long long lim = n_rows * 3;
    const long long off = base * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int i = threadIdx.x + k * BLOCK;
        long long g = off + i;
        if (g < lim) dst[g] = s[i];
    }
}

// camera parameters broadcast to every thread
struct HgsCam {
    float R[3][3];
    float t[3];
    float fx, fy, cx, cy;
};
__device__ __forceinline__ HgsCam hgs_load_cam(const float* __restrict__ viewmats, const float* __restrict__ Ks, int c) {
    HgsCam cam;
    const float* V = viewmats + c * 16;
    const float* K = Ks + c * 9;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) cam.R[i][j] = V[i * 4 + j];
        cam.t[i] = V[i * 4 + 3];
    }
    cam.fx = K[0];
    cam.fy = K[4];
    cam.cx = K[2];
    cam.cy = K[5];
    return cam;
}

// wxyz quaternion -> rotation, normalised with IEEE 1/sqrt (bit-matches the oracle)
__device__ __forceinline__ void hgs_quat_to_rot(float w, float x, float y, float z, float q[3][3], float* inv_norm_out,
                                                float qn[4]) {
    float inv = 1.0f / sqrtf(x * x + y * y + z * z + w * w);
    w *= inv; x *= inv; y *= inv; z *= inv;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    float wx = w * x, wy = w * y, wz = w * z;
    q[0][0] = 1.0f - 2.0f * (y2 + z2); q[0][1] = 2.0f * (xy - wz);        q[0][2] = 2.0f * (xz + wy);
    q[1][0] = 2.0f * (xy + wz);        q[1][1] = 1.0f - 2.0f * (x2 + z2); q[1][2] = 2.0f * (yz - wx);
    q[2][0] = 2.0f * (xz - wy);        q[2][1] = 2.0f * (yz + wx);        q[2][2] = 1.0f - 2.0f * (x2 + y2);
    if (inv_norm_out) *inv_norm_out = inv;
    if (qn) { qn[0] = w; qn[1] = x; qn[2] = y; qn[3] = z; }
}

// gradient of the rotation w.r.t. the raw quaternion, V = dL/dR (row-major), qn = normalised wxyz
__device__ __forceinline__ void hgs_quat_to_rot_vjp(const float qn[4], float inv_norm, const float V[3][3], float vq[4]) {
    const float w = qn[0], x = qn[1], y = qn[2], z = qn[3];
    float vw = 2.f * (z * (V[1][0] - V[0][1]) + y * (V[0][2] - V[2][0]) + x * (V[2][1] - V[1][2]));
    float vx = 2.f * (y * (V[0][1] + V[1][0]) + z * (V[0][2] + V[2][0]) + w * (V[2][1] - V[1][2]) - 2.f * x * (V[1][1] + V[2][2]));
    float vy = 2.f * (x * (V[0][1] + V[1][0]) + w * (V[0][2] - V[2][0]) + z * (V[1][2] + V[2][1]) - 2.f * y * (V[0][0] + V[2][2]));
    float vz = 2.f * (w * (V[1][0] - V[0][1]) + x * (V[0][2] + V[2][0]) + y * (V[1][2] + V[2][1]) - 2.f * z * (V[0][0] + V[1][1]));
    float d = vw * w + vx * x + vy * y + vz * z;
    vq[0] = (vw - d * w) * inv_norm;
    vq[1] = (vx - d * x) * inv_norm;
    vq[2] = (vy - d * y) * inv_norm;
    vq[3] = (vz - d * z) * inv_norm;
}

// tile bounding box of a projected Gaussian (gsplat isect_tiles arithmetic):
// min inclusive, max exclusive, clamped to the tile grid
__device__ __forceinline__ void hgs_tile_bbox(float mx, float my, float radius, float tile_size, int tile_w, int tile_h,
                                              int& x0, int& y0, int& x1, int& y1) {
    float tr = radius / tile_size;
    float tx = mx / tile_size;
    float ty = my / tile_size;
    x0 = (int)fminf(fmaxf(floorf(tx - tr), 0.f), (float)tile_w);
    y0 = (int)fminf(fmaxf(floorf(ty - tr), 0.f), (float)tile_h);
    x1 = (int)fminf(fmaxf(ceilf(tx + tr), 0.f), (float)tile_w);
    y1 = (int)fminf(fmaxf(ceilf(ty + tr), 0.f), (float)tile_h);
}
