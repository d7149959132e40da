// Per-Gaussian projection math of stage a3 shared by the projection kernels (project3d.cu) and the fused
// backward + exchange kernel (exchange_vjp.cu): world -> camera, quat + scale -> covariance, perspective Jacobian,
// blurred 2D covariance (forward), and the analytic VJP of one (camera, Gaussian) pair.  Semantics restated in
// oracle/gsplat_oracle.py::_project3d_one (gsplat fully_fused_projection, reference render.py:149-165).
// Include from translation units built with -fmad=false: every intermediate then rounds like the oracle's.
#pragma once
#include "hgs_common.cuh"
#include "hgs_constants.cuh"

namespace {

struct Proj3dFwd {
    float xc, yc, zc;
    float q[3][3];
    float qn[4];
    float inv_norm;
    float M[3][3];
    float Sc[3][3];
    float rz, rz2, tx, ty;
    bool x_unclamped, y_unclamped;
    float j00, j11, j02, j12;
    float c00, c01, c11;  // blurred 2D covariance
    float det, det_orig;
    float m2x, m2y;
};

__device__ __forceinline__ void proj3d_cov_and_project(const HgsCam& cam, float qw, float qx, float qy, float qz,
                                                       float s0, float s1, float s2, float W, float H, float eps2d,
                                                       Proj3dFwd& o);

// forward math shared by fwd and bwd kernels. returns false when culled by z.
__device__ __forceinline__ bool proj3d_math(const HgsCam& cam, float px, float py, float pz, float qw, float qx, float qy,
                                            float qz, float s0, float s1, float s2, float W, float H, float eps2d,
                                            float near_plane, float far_plane, Proj3dFwd& o) {
    const float (*R)[3] = cam.R;
    o.zc = R[2][0] * px + R[2][1] * py + R[2][2] * pz + cam.t[2];
    if (o.zc < near_plane || o.zc > far_plane) return false;
    o.xc = R[0][0] * px + R[0][1] * py + R[0][2] * pz + cam.t[0];
    o.yc = R[1][0] * px + R[1][1] * py + R[1][2] * pz + cam.t[1];
    proj3d_cov_and_project(cam, qw, qx, qy, qz, s0, s1, s2, W, H, eps2d, o);
    return true;
}

// Conservative early frustum test on the camera-space centre alone: an upper bound of the integer radius
// from ||J||_F^2 * max(scale)^2 (lambda_max of J Sigma J^T <= ||J||_F^2 lambda_max(Sigma)), the tan-fov clamp
// bounds of J, eps2d and the 0.01 eigenvalue floor.  true => the full computation would cull the Gaussian by
// its screen-bounds test, so the covariance math (and the quaternion load) can be skipped with identical output.
// per-camera constant of the bound: fx^2 (1 + Lx^2) + fy^2 (1 + Ly^2), L = largest clamped |x/z|
__device__ __forceinline__ float proj3d_jf_coeff(const HgsCam& cam, float W, float H) {
    const float fx = cam.fx, fy = cam.fy, cx = cam.cx, cy = cam.cy;
    const float tan_fovx = 0.5f * W / fx, tan_fovy = 0.5f * H / fy;
    const float Lx = fmaxf((W - cx) / fx, cx / fx) + HGS_FOV_MARGIN * tan_fovx;
    const float Ly = fmaxf((H - cy) / fy, cy / fy) + HGS_FOV_MARGIN * tan_fovy;
    return fx * fx * (1.0f + Lx * Lx) + fy * fy * (1.0f + Ly * Ly);
}
__device__ __forceinline__ bool proj3d_surely_offscreen(const HgsCam& cam, float jf_coeff, float xc, float yc, float zc,
                                                        float smax, float W, float H, float eps2d) {
    const float fx = cam.fx, fy = cam.fy, cx = cam.cx, cy = cam.cy;
    // approximate reciprocal / square root (MUFU, ~2 ulp): the bound carries a 0.1 % + 1 pixel margin
    const float rz = __fdividef(1.0f, zc);
    const float jf2 = rz * rz * jf_coeff;
    const float v1_bound = jf2 * smax * smax + eps2d + 0.1f + HGS_EIG_FLOOR;
    const float rb = (HGS_RADIUS_SIGMA * (v1_bound * __frsqrt_rn(v1_bound)) + 1.0f) * 1.001f + 0.01f;
    const float m2x = fx * xc * rz + cx, m2y = fy * yc * rz + cy;
    return (m2x + rb < 0.f) || (m2x - rb > W) || (m2y + rb < 0.f) || (m2y - rb > H);
}

__device__ __forceinline__ void proj3d_cov_and_project(const HgsCam& cam, float qw, float qx, float qy, float qz,
                                                       float s0, float s1, float s2, float W, float H, float eps2d,
                                                       Proj3dFwd& o) {
    const float (*R)[3] = cam.R;
    hgs_quat_to_rot(qw, qx, qy, qz, o.q, &o.inv_norm, o.qn);
    const float s[3] = {s0, s1, s2};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o.M[i][j] = o.q[i][j] * s[j];
    float S[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j) {
            S[i][j] = o.M[i][0] * o.M[j][0] + o.M[i][1] * o.M[j][1] + o.M[i][2] * o.M[j][2];
            S[j][i] = S[i][j];
        }
    float A[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) A[i][j] = R[i][0] * S[0][j] + R[i][1] * S[1][j] + R[i][2] * S[2][j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j) {
            o.Sc[i][j] = A[i][0] * R[j][0] + A[i][1] * R[j][1] + A[i][2] * R[j][2];
            o.Sc[j][i] = o.Sc[i][j];
        }

    const float fx = cam.fx, fy = cam.fy, cx = cam.cx, cy = cam.cy;
    float tan_fovx = 0.5f * W / fx;
    float tan_fovy = 0.5f * H / fy;
    float lim_x_pos = (W - cx) / fx + HGS_FOV_MARGIN * tan_fovx;
    float lim_x_neg = cx / fx + HGS_FOV_MARGIN * tan_fovx;
    float lim_y_pos = (H - cy) / fy + HGS_FOV_MARGIN * tan_fovy;
    float lim_y_neg = cy / fy + HGS_FOV_MARGIN * tan_fovy;
    o.rz = 1.0f / o.zc;
    o.rz2 = o.rz * o.rz;
    float xr = o.xc * o.rz, yr = o.yc * o.rz;
    o.x_unclamped = (xr <= lim_x_pos) && (xr >= -lim_x_neg);
    o.y_unclamped = (yr <= lim_y_pos) && (yr >= -lim_y_neg);
    o.tx = o.zc * fminf(lim_x_pos, fmaxf(-lim_x_neg, xr));
    o.ty = o.zc * fminf(lim_y_pos, fmaxf(-lim_y_neg, yr));
    o.j00 = fx * o.rz;
    o.j11 = fy * o.rz;
    o.j02 = -(fx * o.tx * o.rz2);
    o.j12 = -(fy * o.ty * o.rz2);
    float B00 = o.j00 * o.Sc[0][0] + o.j02 * o.Sc[2][0];
    float B01 = o.j00 * o.Sc[0][1] + o.j02 * o.Sc[2][1];
    float B02 = o.j00 * o.Sc[0][2] + o.j02 * o.Sc[2][2];
    float B11 = o.j11 * o.Sc[1][1] + o.j12 * o.Sc[2][1];
    float B12 = o.j11 * o.Sc[1][2] + o.j12 * o.Sc[2][2];
    float c00 = B00 * o.j00 + B02 * o.j02;
    float c01 = B01 * o.j11 + B02 * o.j12;
    float c11 = B11 * o.j11 + B12 * o.j12;
    o.m2x = fx * o.xc * o.rz + cx;
    o.m2y = fy * o.yc * o.rz + cy;
    o.det_orig = c00 * c11 - c01 * c01;
    c00 = c00 + eps2d;
    c11 = c11 + eps2d;
    o.det = c00 * c11 - c01 * c01;
    o.c00 = c00; o.c01 = c01; o.c11 = c11;
}

// gradient of one (camera, Gaussian) pair; accumulates into g_mean / g_scale / g_quat
__device__ __forceinline__ void proj3d_bwd_one(const HgsCam& cam, const Proj3dFwd& f, float s0, float s1, float s2,
                                               float2 vm, float vd, float va, float vb, float vc, float g_mean[3],
                                               float g_scale[3], float g_quat[4]) {
    // conic X = inv(Sigma2'), v_Sigma2 = -X V X
    const float inv_det = 1.0f / f.det;
    const float a = f.c11 * inv_det, b = -f.c01 * inv_det, cc = f.c00 * inv_det;
    const float xv00 = a * va + b * vb, xv01 = a * vb + b * vc;
    const float xv10 = b * va + cc * vb, xv11 = b * vb + cc * vc;
    const float G00 = -(xv00 * a + xv01 * b);
    const float G01 = -(xv00 * b + xv01 * cc);
    const float G11 = -(xv10 * b + xv11 * cc);

    // Sigma2 = J Sc J^T
    const float GJ[2][3] = {{G00 * f.j00, G01 * f.j11, G00 * f.j02 + G01 * f.j12},
                            {G01 * f.j00, G11 * f.j11, G01 * f.j02 + G11 * f.j12}};
    float vSc[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        vSc[0][j] = f.j00 * GJ[0][j];
        vSc[1][j] = f.j11 * GJ[1][j];
        vSc[2][j] = f.j02 * GJ[0][j] + f.j12 * GJ[1][j];
    }
    const float vJ00 = 2.f * (GJ[0][0] * f.Sc[0][0] + GJ[0][1] * f.Sc[1][0] + GJ[0][2] * f.Sc[2][0]);
    const float vJ02 = 2.f * (GJ[0][0] * f.Sc[0][2] + GJ[0][1] * f.Sc[1][2] + GJ[0][2] * f.Sc[2][2]);
    const float vJ11 = 2.f * (GJ[1][0] * f.Sc[0][1] + GJ[1][1] * f.Sc[1][1] + GJ[1][2] * f.Sc[2][1]);
    const float vJ12 = 2.f * (GJ[1][0] * f.Sc[0][2] + GJ[1][1] * f.Sc[1][2] + GJ[1][2] * f.Sc[2][2]);

    const float fx = cam.fx, fy = cam.fy;
    const float rz3 = f.rz2 * f.rz;
    float v_xc = fx * f.rz * vm.x;
    float v_yc = fy * f.rz * vm.y;
    float v_zc = -(fx * f.xc * vm.x + fy * f.yc * vm.y) * f.rz2 + vd;
    v_zc += -fx * f.rz2 * vJ00 - fy * f.rz2 * vJ11;
    if (f.x_unclamped) {
        v_xc += -fx * f.rz2 * vJ02;
        v_zc += 2.f * fx * f.tx * rz3 * vJ02;
    } else {
        v_zc += fx * f.tx * rz3 * vJ02;
    }
    if (f.y_unclamped) {
        v_yc += -fy * f.rz2 * vJ12;
        v_zc += 2.f * fy * f.ty * rz3 * vJ12;
    } else {
        v_zc += fy * f.ty * rz3 * vJ12;
    }

    const float (*R)[3] = cam.R;
    g_mean[0] += R[0][0] * v_xc + R[1][0] * v_yc + R[2][0] * v_zc;
    g_mean[1] += R[0][1] * v_xc + R[1][1] * v_yc + R[2][1] * v_zc;
    g_mean[2] += R[0][2] * v_xc + R[1][2] * v_yc + R[2][2] * v_zc;

    // Sc = R S R^T  ->  vS = R^T vSc R
    float T[3][3], vS[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) T[i][j] = vSc[i][0] * R[0][j] + vSc[i][1] * R[1][j] + vSc[i][2] * R[2][j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) vS[i][j] = R[0][i] * T[0][j] + R[1][i] * T[1][j] + R[2][i] * T[2][j];
    // S = M M^T -> vM = (vS + vS^T) M
    float vM[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            vM[i][j] = (vS[i][0] + vS[0][i]) * f.M[0][j] + (vS[i][1] + vS[1][i]) * f.M[1][j] +
                       (vS[i][2] + vS[2][i]) * f.M[2][j];
    // M = q diag(s)
    const float s[3] = {s0, s1, s2};
    float vq_mat[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        g_scale[j] += f.q[0][j] * vM[0][j] + f.q[1][j] * vM[1][j] + f.q[2][j] * vM[2][j];
#pragma unroll
        for (int i = 0; i < 3; ++i) vq_mat[i][j] = vM[i][j] * s[j];
    }
    float vq[4];
    hgs_quat_to_rot_vjp(f.qn, f.inv_norm, vq_mat, vq);
#pragma unroll
    for (int k = 0; k < 4; ++k) g_quat[k] += vq[k];
}

}  // namespace
