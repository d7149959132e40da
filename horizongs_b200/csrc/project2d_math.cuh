// Per-surfel projection math of stage a4 shared by the projection kernels (project2d.cu) and the fused backward +
// exchange kernel (exchange_vjp.cu): ray transform M = K [R|t] H (forward) and the analytic VJP of one
// (camera, surfel) pair.  Semantics restated in oracle/gsplat_oracle.py::_project2d_one (gsplat 2DGS fork's
// fully_fused_projection_2dgs, reference render.py:171-186).  Forward bit-exactness needs -fmad=false.
#pragma once
#include "hgs_common.cuh"
#include "hgs_constants.cuh"

namespace {

struct Proj2dFwd {
    float mc[3];
    float q[3][3];
    float qn[4];
    float inv_norm;
    float RQ[3][3];
    float M0[3], M1[3], M2[3];
    float dist, f[3];
    float m2x, m2y;
};

__device__ __forceinline__ bool proj2d_math(const HgsCam& cam, float px, float py, float pz, float qw, float qx, float qy,
                                            float qz, float s0, float s1, float near_plane, float far_plane,
                                            Proj2dFwd& o) {
    const float (*R)[3] = cam.R;
    o.mc[2] = R[2][0] * px + R[2][1] * py + R[2][2] * pz + cam.t[2];
    if (o.mc[2] < near_plane || o.mc[2] > far_plane) return false;
    o.mc[0] = R[0][0] * px + R[0][1] * py + R[0][2] * pz + cam.t[0];
    o.mc[1] = R[1][0] * px + R[1][1] * py + R[1][2] * pz + cam.t[1];
    hgs_quat_to_rot(qw, qx, qy, qz, o.q, &o.inv_norm, o.qn);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o.RQ[i][j] = R[i][0] * o.q[0][j] + R[i][1] * o.q[1][j] + R[i][2] * o.q[2][j];
    float WH[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        WH[i][0] = o.RQ[i][0] * s0;
        WH[i][1] = o.RQ[i][1] * s1;
        WH[i][2] = o.mc[i];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        o.M0[j] = cam.fx * WH[0][j] + cam.cx * WH[2][j];
        o.M1[j] = cam.fy * WH[1][j] + cam.cy * WH[2][j];
        o.M2[j] = WH[2][j];
    }
    o.dist = o.M2[0] * o.M2[0] + o.M2[1] * o.M2[1] - o.M2[2] * o.M2[2];
    if (o.dist == 0.f) return false;
    const float invd = 1.0f / o.dist;
    o.f[0] = invd; o.f[1] = invd; o.f[2] = -invd;
    o.m2x = o.f[0] * o.M0[0] * o.M2[0] + o.f[1] * o.M0[1] * o.M2[1] + o.f[2] * o.M0[2] * o.M2[2];
    o.m2y = o.f[0] * o.M1[0] * o.M2[0] + o.f[1] * o.M1[1] * o.M2[1] + o.f[2] * o.M1[2] * o.M2[2];
    return true;
}

// gradient of one (camera, surfel) pair; accumulates into g_mean / g_scale / g_quat
__device__ __forceinline__ void proj2d_bwd_one(const HgsCam& cam, const Proj2dFwd& f, float s0, float s1,
                                               const float* __restrict__ v_means2d, int ld_m2,
                                               const float* __restrict__ v_depths, int ld_d,
                                               const float* __restrict__ v_ray_transforms, int ld_rt,
                                               const float* __restrict__ v_normals, int ld_n, long long idx,
                                               float g_mean[3], float g_scale[3], float g_quat[4]) {
    float vM0[3] = {0.f, 0.f, 0.f}, vM1[3] = {0.f, 0.f, 0.f}, vM2[3] = {0.f, 0.f, 0.f};
    if (v_ray_transforms != nullptr) {
        const float* vr = v_ray_transforms + idx * ld_rt;
#pragma unroll
        for (int j = 0; j < 3; ++j) { vM0[j] = vr[j]; vM1[j] = vr[3 + j]; vM2[j] = vr[6 + j]; }
    }
    if (v_means2d != nullptr) {
        const float vmx = v_means2d[idx * ld_m2], vmy = v_means2d[idx * ld_m2 + 1];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            vM0[j] += vmx * f.f[j] * f.M2[j];
            vM1[j] += vmy * f.f[j] * f.M2[j];
            vM2[j] += vmx * f.f[j] * (f.M0[j] - 2.f * f.M2[j] * f.m2x) + vmy * f.f[j] * (f.M1[j] - 2.f * f.M2[j] * f.m2y);
        }
    }
    float vWH[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        vWH[0][j] = cam.fx * vM0[j];
        vWH[1][j] = cam.fy * vM1[j];
        vWH[2][j] = cam.cx * vM0[j] + cam.cy * vM1[j] + vM2[j];
    }
    float vRQ[3][3];
    float v_mc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        vRQ[i][0] = vWH[i][0] * s0;
        vRQ[i][1] = vWH[i][1] * s1;
        vRQ[i][2] = 0.f;
        v_mc[i] = vWH[i][2];
    }
    g_scale[0] += f.RQ[0][0] * vWH[0][0] + f.RQ[1][0] * vWH[1][0] + f.RQ[2][0] * vWH[2][0];
    g_scale[1] += f.RQ[0][1] * vWH[0][1] + f.RQ[1][1] * vWH[1][1] + f.RQ[2][1] * vWH[2][1];
    if (v_depths != nullptr) v_mc[2] += v_depths[idx * ld_d];
    if (v_normals != nullptr) {
        const float dotv = -f.RQ[0][2] * f.mc[0] + -f.RQ[1][2] * f.mc[1] + -f.RQ[2][2] * f.mc[2];
        const float sign = dotv > 0.f ? 1.0f : -1.0f;
#pragma unroll
        for (int i = 0; i < 3; ++i) vRQ[i][2] = sign * v_normals[idx * ld_n + i];
    }
    const float (*R)[3] = cam.R;
    float vq_mat[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int j = 0; j < 3; ++j) vq_mat[k][j] = R[0][k] * vRQ[0][j] + R[1][k] * vRQ[1][j] + R[2][k] * vRQ[2][j];
        g_mean[k] += R[0][k] * v_mc[0] + R[1][k] * v_mc[1] + R[2][k] * v_mc[2];
    }
    float vq[4];
    hgs_quat_to_rot_vjp(f.qn, f.inv_norm, vq_mat, vq);
#pragma unroll
    for (int k = 0; k < 4; ++k) g_quat[k] += vq[k];
}

}  // namespace
