// Real spherical-harmonics basis (degree <= 4) and the per-(camera, Gaussian) colour evaluation / gradient of
// stage a7, shared by the SH kernels (sh.cu) and the fused backward + exchange kernel (exchange_vjp.cu).
// Basis = Sloan's polynomial forms, equal to utils/sh_utils.py:57-112 of the reference on unit vectors.
// Include from translation units built with -fmad=false.
#pragma once
#include "hgs_common.cuh"
#include "hgs_constants.cuh"

namespace {

template <int DEG>
struct ShBasis {
    static constexpr int NB = (DEG + 1) * (DEG + 1);
    float b[NB];
};

// basis values (and optionally d/dx, d/dy, d/dz treating x,y,z as free variables)
template <int DEG, bool GRAD>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float* b, float* bx, float* by, float* bz) {
    b[0] = 0.2820947917738781f;
    if (GRAD) { bx[0] = by[0] = bz[0] = 0.f; }
    if (DEG < 1) return;
    b[1] = -0.48860251190292f * y;
    b[2] = 0.48860251190292f * z;
    b[3] = -0.48860251190292f * x;
    if (GRAD) {
        bx[1] = 0.f; by[1] = -0.48860251190292f; bz[1] = 0.f;
        bx[2] = 0.f; by[2] = 0.f; bz[2] = 0.48860251190292f;
        bx[3] = -0.48860251190292f; by[3] = 0.f; bz[3] = 0.f;
    }
    if (DEG < 2) return;
    const float z2 = z * z;
    const float fTmp0B = -1.092548430592079f * z;
    const float fC1 = x * x - y * y;
    const float fS1 = 2.0f * x * y;
    const float pSH6 = 0.9461746957575601f * z2 - 0.3153915652525201f;
    b[4] = 0.5462742152960395f * fS1;
    b[5] = fTmp0B * y;
    b[6] = pSH6;
    b[7] = fTmp0B * x;
    b[8] = 0.5462742152960395f * fC1;
    const float dpSH6 = 2.f * 0.9461746957575601f * z;
    if (GRAD) {
        bx[4] = 0.5462742152960395f * 2.f * y; by[4] = 0.5462742152960395f * 2.f * x; bz[4] = 0.f;
        bx[5] = 0.f; by[5] = fTmp0B; bz[5] = -1.092548430592079f * y;
        bx[6] = 0.f; by[6] = 0.f; bz[6] = dpSH6;
        bx[7] = fTmp0B; by[7] = 0.f; bz[7] = -1.092548430592079f * x;
        bx[8] = 0.5462742152960395f * 2.f * x; by[8] = -0.5462742152960395f * 2.f * y; bz[8] = 0.f;
    }
    if (DEG < 3) return;
    const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    const float fTmp1B = 1.445305721320277f * z;
    const float fC2 = x * fC1 - y * fS1;
    const float fS2 = x * fS1 + y * fC1;
    const float pSH12 = z * (1.865881662950577f * z2 - 1.119528997770346f);
    b[9] = -0.5900435899266435f * fS2;
    b[10] = fTmp1B * fS1;
    b[11] = fTmp0C * y;
    b[12] = pSH12;
    b[13] = fTmp0C * x;
    b[14] = fTmp1B * fC1;
    b[15] = -0.5900435899266435f * fC2;
    const float dTmp0C = -2.f * 2.285228997322329f * z;
    const float dpSH12 = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
    if (GRAD) {
        bx[9] = -0.5900435899266435f * 3.f * fS1; by[9] = -0.5900435899266435f * 3.f * fC1; bz[9] = 0.f;
        bx[10] = fTmp1B * 2.f * y; by[10] = fTmp1B * 2.f * x; bz[10] = 1.445305721320277f * fS1;
        bx[11] = 0.f; by[11] = fTmp0C; bz[11] = dTmp0C * y;
        bx[12] = 0.f; by[12] = 0.f; bz[12] = dpSH12;
        bx[13] = fTmp0C; by[13] = 0.f; bz[13] = dTmp0C * x;
        bx[14] = fTmp1B * 2.f * x; by[14] = -fTmp1B * 2.f * y; bz[14] = 1.445305721320277f * fC1;
        bx[15] = -0.5900435899266435f * 3.f * fC1; by[15] = 0.5900435899266435f * 3.f * fS1; bz[15] = 0.f;
    }
    if (DEG < 4) return;
    const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
    const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
    const float fTmp2B = -1.770130769779931f * z;
    const float fC3 = x * fC2 - y * fS2;
    const float fS3 = x * fS2 + y * fC2;
    const float pSH20 = 1.984313483298443f * z * pSH12 + -1.006230589874905f * pSH6;
    b[16] = 0.6258357354491763f * fS3;
    b[17] = fTmp2B * fS2;
    b[18] = fTmp1C * fS1;
    b[19] = fTmp0D * y;
    b[20] = pSH20;
    b[21] = fTmp0D * x;
    b[22] = fTmp1C * fC1;
    b[23] = fTmp2B * fC2;
    b[24] = 0.6258357354491763f * fC3;
    if (GRAD) {
        const float dTmp0D = -3.f * 4.683325804901025f * z2 + 2.007139630671868f;
        const float dTmp1C = 2.f * 3.31161143515146f * z;
        bx[16] = 0.6258357354491763f * 4.f * fS2; by[16] = 0.6258357354491763f * 4.f * fC2; bz[16] = 0.f;
        bx[17] = fTmp2B * 3.f * fS1; by[17] = fTmp2B * 3.f * fC1; bz[17] = -1.770130769779931f * fS2;
        bx[18] = fTmp1C * 2.f * y; by[18] = fTmp1C * 2.f * x; bz[18] = dTmp1C * fS1;
        bx[19] = 0.f; by[19] = fTmp0D; bz[19] = dTmp0D * y;
        bx[20] = 0.f; by[20] = 0.f;
        bz[20] = 1.984313483298443f * (pSH12 + z * dpSH12) - 1.006230589874905f * dpSH6;
        bx[21] = fTmp0D; by[21] = 0.f; bz[21] = dTmp0D * x;
        bx[22] = fTmp1C * 2.f * x; by[22] = -fTmp1C * 2.f * y; bz[22] = dTmp1C * fC1;
        bx[23] = fTmp2B * 3.f * fC1; by[23] = -fTmp2B * 3.f * fS1; bz[23] = -1.770130769779931f * fC2;
        bx[24] = 0.6258357354491763f * 4.f * fC2; by[24] = -0.6258357354491763f * 4.f * fS2; bz[24] = 0.f;
    }
}

// colour of one (camera, Gaussian) pair
template <int DEG>
__device__ __forceinline__ void sh_eval_one(float x, float y, float z, const float* __restrict__ co, int post,
                                            float& r0, float& r1, float& r2) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const float inorm = 1.0f / sqrtf(x * x + y * y + z * z);
    x *= inorm; y *= inorm; z *= inorm;
    float b[NB];
    sh_basis<DEG, false>(x, y, z, b, nullptr, nullptr, nullptr);
    r0 = b[0] * __ldg(co + 0); r1 = b[0] * __ldg(co + 1); r2 = b[0] * __ldg(co + 2);
#pragma unroll
    for (int k = 1; k < NB; ++k) {
        r0 = r0 + b[k] * __ldg(co + k * 3 + 0);
        r1 = r1 + b[k] * __ldg(co + k * 3 + 1);
        r2 = r2 + b[k] * __ldg(co + k * 3 + 2);
    }
    if (post) {
        r0 = fmaxf(r0 + HGS_SH_OFFSET, 0.f);
        r1 = fmaxf(r1 + HGS_SH_OFFSET, 0.f);
        r2 = fmaxf(r2 + HGS_SH_OFFSET, 0.f);
    }
}

// gradient of one (camera, Gaussian) pair: g_co[NB*3] += basis * v, gd = d/d(un-normalised direction)
template <int DEG>
__device__ __forceinline__ void sh_grad_one(float x, float y, float z, const float* __restrict__ co, float v0, float v1,
                                            float v2, bool want_dir, float* g_co, float& gd0, float& gd1, float& gd2) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const float inorm = 1.0f / sqrtf(x * x + y * y + z * z);
    x *= inorm; y *= inorm; z *= inorm;
    float b[NB], bx[NB], by[NB], bz[NB];
    sh_basis<DEG, true>(x, y, z, b, bx, by, bz);
    float vx = 0.f, vy = 0.f, vz = 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        g_co[k * 3 + 0] += b[k] * v0;
        g_co[k * 3 + 1] += b[k] * v1;
        g_co[k * 3 + 2] += b[k] * v2;
        if (want_dir) {
            const float d = co[k * 3] * v0 + co[k * 3 + 1] * v1 + co[k * 3 + 2] * v2;
            vx += bx[k] * d; vy += by[k] * d; vz += bz[k] * d;
        }
    }
    gd0 = gd1 = gd2 = 0.f;
    if (want_dir) {
        const float dd = vx * x + vy * y + vz * z;
        gd0 = (vx - dd * x) * inorm;
        gd1 = (vy - dd * y) * inorm;
        gd2 = (vz - dd * z) * inorm;
    }
}

// direction gradient only (no coefficient gradients): gd = d/d(un-normalised direction) of sum_c v_c * colour_c
template <int DEG>
__device__ __forceinline__ void sh_dirgrad_one(float x, float y, float z, const float* __restrict__ co, float v0, float v1,
                                               float v2, float& gd0, float& gd1, float& gd2) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const float inorm = 1.0f / sqrtf(x * x + y * y + z * z);
    x *= inorm; y *= inorm; z *= inorm;
    float b[NB], bx[NB], by[NB], bz[NB];
    sh_basis<DEG, true>(x, y, z, b, bx, by, bz);
    float vx = 0.f, vy = 0.f, vz = 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        const float d = co[k * 3] * v0 + co[k * 3 + 1] * v1 + co[k * 3 + 2] * v2;
        vx += bx[k] * d; vy += by[k] * d; vz += bz[k] * d;
    }
    const float dd = vx * x + vy * y + vz * z;
    gd0 = (vx - dd * x) * inorm;
    gd1 = (vy - dd * y) * inorm;
    gd2 = (vz - dd * z) * inorm;
}

// basis values of the normalised direction (x, y, z un-normalised on input)
template <int DEG>
__device__ __forceinline__ void sh_basis_of(float x, float y, float z, float* b) {
    const float inorm = 1.0f / sqrtf(x * x + y * y + z * z);
    sh_basis<DEG, false>(x * inorm, y * inorm, z * inorm, b, nullptr, nullptr, nullptr);
}

}  // namespace
