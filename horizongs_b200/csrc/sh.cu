// Stage a7: view-dependent colour from spherical harmonics (degree <= 4) and its backward.
//
// Replaces gsplat's spherical_harmonics as reached inside gsplat.rasterization* when
// sh_degree is not None (reference render.py:51,73; SH2 configs).  Basis = Sloan's polynomial
// forms, equal to utils/sh_utils.py:57-112 of the reference on unit vectors; coefficient layout
// [N,K,3] (scene/basic_model.py:369,378).  Restated in oracle/gsplat_oracle.py::spherical_harmonics.
// Fused here: direction = means - campos, normalisation, optional `clamp_min(c + 0.5, 0)`.
//
// One thread per Gaussian, cameras looped in-thread (coefficient gradients are summed over views
// without atomics).  Coefficient rows are staged through shared memory so global traffic is coalesced.
// Roofline: HBM; fwd (12 K + 12 + 12) B per Gaussian, bwd (24 K + 36) B.
#include "hgs_common.cuh"
#include "hgs_constants.cuh"
#include "../../include/hgs_raster.h"

namespace {

constexpr int SB = 128;  // threads (= Gaussians) per block

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

__device__ __forceinline__ int row_stride(int K) { return (K * 3) | 1; }  // odd -> conflict-free rows

// coalesced copy of rows [base, base+SB) of a [N, K*3] array into padded shared rows (first `used` floats)
__device__ __forceinline__ void stage_rows_in(const float* __restrict__ src, long long base, long long N, int K,
                                              int used, float* s) {
    const int rs = row_stride(K);
    const int rowlen = K * 3;
    long long rows = N - base;
    if (rows > SB) rows = SB;
    if (used == rowlen) {
        const long long tot = rows * rowlen;
        const float* p = src + base * rowlen;
        for (long long i = threadIdx.x; i < tot; i += SB) {
            int r = (int)(i / rowlen), c = (int)(i % rowlen);
            s[r * rs + c] = p[i];
        }
    } else {
        for (long long i = threadIdx.x; i < rows * used; i += SB) {
            int r = (int)(i / used), c = (int)(i % used);
            s[r * rs + c] = src[(base + r) * rowlen + c];
        }
    }
}

template <int DEG>
__global__ void __launch_bounds__(SB) sh_fwd_kernel(const float* __restrict__ dirs, const float* __restrict__ means,
                                                    const float* __restrict__ campos,
                                                    const float* __restrict__ coeffs,
                                                    const int32_t* __restrict__ radii, int C, int N, int K, int post,
                                                    float* __restrict__ colors) {
    extern __shared__ float smem[];
    constexpr int NB = (DEG + 1) * (DEG + 1);
    float* s_co = smem;                          // SB rows of coefficients
    float* s_io = smem + SB * row_stride(K);     // SB*3 staging for means / dirs / colours
    const long long base = (long long)blockIdx.x * SB;
    const long long n = base + threadIdx.x;
    stage_rows_in(coeffs, base, N, K, NB * 3, s_co);
    if (dirs == nullptr) block_load_rows3<SB>(means, base, N, s_io);
    __syncthreads();
    float mx = 0.f, my = 0.f, mz = 0.f;
    if (dirs == nullptr) { mx = s_io[threadIdx.x * 3]; my = s_io[threadIdx.x * 3 + 1]; mz = s_io[threadIdx.x * 3 + 2]; }
    const float* co = s_co + threadIdx.x * row_stride(K);
    for (int c = 0; c < C; ++c) {
        __syncthreads();
        if (dirs != nullptr) {
            block_load_rows3<SB>(dirs + (long long)c * N * 3, base, N, s_io);
            __syncthreads();
        }
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        if (n < N && (radii == nullptr || radii[(long long)c * N + n] > 0)) {
            float x, y, z;
            if (dirs != nullptr) { x = s_io[threadIdx.x * 3]; y = s_io[threadIdx.x * 3 + 1]; z = s_io[threadIdx.x * 3 + 2]; }
            else { x = mx - campos[c * 3]; y = my - campos[c * 3 + 1]; z = mz - campos[c * 3 + 2]; }
            const float inorm = 1.0f / sqrtf(x * x + y * y + z * z);
            x *= inorm; y *= inorm; z *= inorm;
            float b[NB];
            sh_basis<DEG, false>(x, y, z, b, nullptr, nullptr, nullptr);
            r0 = b[0] * co[0]; r1 = b[0] * co[1]; r2 = b[0] * co[2];
#pragma unroll
            for (int k = 1; k < NB; ++k) {
                r0 = r0 + b[k] * co[k * 3 + 0];
                r1 = r1 + b[k] * co[k * 3 + 1];
                r2 = r2 + b[k] * co[k * 3 + 2];
            }
            if (post) {
                r0 = fmaxf(r0 + HGS_SH_OFFSET, 0.f);
                r1 = fmaxf(r1 + HGS_SH_OFFSET, 0.f);
                r2 = fmaxf(r2 + HGS_SH_OFFSET, 0.f);
            }
        }
        __syncthreads();
        s_io[threadIdx.x * 3] = r0; s_io[threadIdx.x * 3 + 1] = r1; s_io[threadIdx.x * 3 + 2] = r2;
        __syncthreads();
        block_store_rows3<SB>(colors + (long long)c * N * 3, base, N, s_io);
    }
}

template <int DEG>
__global__ void __launch_bounds__(SB) sh_bwd_kernel(const float* __restrict__ dirs, const float* __restrict__ means,
                                                    const float* __restrict__ campos,
                                                    const float* __restrict__ coeffs,
                                                    const int32_t* __restrict__ radii,
                                                    const float* __restrict__ colors,
                                                    const float* __restrict__ v_colors, int ld_vc, int C, int N,
                                                    int K, int post, float* __restrict__ v_coeffs,
                                                    float* __restrict__ v_dirs, float* __restrict__ v_means) {
    extern __shared__ float smem[];
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const int rs = row_stride(K);
    float* s_co = smem;                 // coefficients in, coefficient gradients out
    float* s_io = smem + SB * rs;       // SB*3 staging
    float* s_io2 = s_io + SB * 3;       // SB*3 staging (forward colours for the clamp mask)
    const long long base = (long long)blockIdx.x * SB;
    const long long n = base + threadIdx.x;
    const bool want_dir = (v_dirs != nullptr) || (v_means != nullptr);
    stage_rows_in(coeffs, base, N, K, NB * 3, s_co);
    if (dirs == nullptr) block_load_rows3<SB>(means, base, N, s_io);
    __syncthreads();
    float mx = 0.f, my = 0.f, mz = 0.f;
    if (dirs == nullptr) { mx = s_io[threadIdx.x * 3]; my = s_io[threadIdx.x * 3 + 1]; mz = s_io[threadIdx.x * 3 + 2]; }
    float* co = s_co + threadIdx.x * rs;
    float coef[NB * 3];
#pragma unroll
    for (int k = 0; k < NB * 3; ++k) coef[k] = co[k];
    float g_co[NB * 3];
#pragma unroll
    for (int k = 0; k < NB * 3; ++k) g_co[k] = 0.f;
    float gm0 = 0.f, gm1 = 0.f, gm2 = 0.f;

    for (int c = 0; c < C; ++c) {
        __syncthreads();
        if (ld_vc == 3) {
            block_load_rows3<SB>(v_colors + (long long)c * N * 3, base, N, s_io);
        } else if (n < N) {
            const float* vr = v_colors + ((long long)c * N + n) * ld_vc;
            s_io[threadIdx.x * 3] = vr[0]; s_io[threadIdx.x * 3 + 1] = vr[1]; s_io[threadIdx.x * 3 + 2] = vr[2];
        }
        if (post) block_load_rows3<SB>(colors + (long long)c * N * 3, base, N, s_io2);
        __syncthreads();
        float v0 = s_io[threadIdx.x * 3], v1 = s_io[threadIdx.x * 3 + 1], v2 = s_io[threadIdx.x * 3 + 2];
        if (post) {
            if (!(s_io2[threadIdx.x * 3] > 0.f)) v0 = 0.f;
            if (!(s_io2[threadIdx.x * 3 + 1] > 0.f)) v1 = 0.f;
            if (!(s_io2[threadIdx.x * 3 + 2] > 0.f)) v2 = 0.f;
        }
        float x = 0.f, y = 0.f, z = 1.f;
        if (dirs != nullptr) {
            __syncthreads();
            block_load_rows3<SB>(dirs + (long long)c * N * 3, base, N, s_io);
            __syncthreads();
            x = s_io[threadIdx.x * 3]; y = s_io[threadIdx.x * 3 + 1]; z = s_io[threadIdx.x * 3 + 2];
        } else if (n < N) {
            x = mx - campos[c * 3]; y = my - campos[c * 3 + 1]; z = mz - campos[c * 3 + 2];
        }
        float gd0 = 0.f, gd1 = 0.f, gd2 = 0.f;
        if (n < N && (radii == nullptr || radii[(long long)c * N + n] > 0)) {
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
                const float d = coef[k * 3] * v0 + coef[k * 3 + 1] * v1 + coef[k * 3 + 2] * v2;
                vx += bx[k] * d; vy += by[k] * d; vz += bz[k] * d;
            }
            if (want_dir) {
                const float dd = vx * x + vy * y + vz * z;
                gd0 = (vx - dd * x) * inorm;
                gd1 = (vy - dd * y) * inorm;
                gd2 = (vz - dd * z) * inorm;
                gm0 += gd0; gm1 += gd1; gm2 += gd2;
            }
        }
        if (v_dirs != nullptr) {
            __syncthreads();
            s_io[threadIdx.x * 3] = gd0; s_io[threadIdx.x * 3 + 1] = gd1; s_io[threadIdx.x * 3 + 2] = gd2;
            __syncthreads();
            block_store_rows3<SB>(v_dirs + (long long)c * N * 3, base, N, s_io);
        }
    }
    // coefficient gradients: padded rows -> coalesced global rows (unused degrees are zero)
    __syncthreads();
    {
        const int rowlen = K * 3;
        for (int k = NB * 3; k < rowlen; ++k) co[k] = 0.f;
#pragma unroll
        for (int k = 0; k < NB * 3; ++k) co[k] = g_co[k];
        __syncthreads();
        long long rows = N - base;
        if (rows > SB) rows = SB;
        const long long tot = rows * rowlen;
        float* p = v_coeffs + base * rowlen;
        for (long long i = threadIdx.x; i < tot; i += SB) {
            int r = (int)(i / rowlen), cidx = (int)(i % rowlen);
            p[i] = s_co[r * rs + cidx];
        }
    }
    if (v_means != nullptr) {
        __syncthreads();
        s_io[threadIdx.x * 3] = gm0; s_io[threadIdx.x * 3 + 1] = gm1; s_io[threadIdx.x * 3 + 2] = gm2;
        __syncthreads();
        block_store_rows3<SB>(v_means, base, N, s_io);
    }
}

static size_t sh_smem_bytes(int K, bool bwd) { return (size_t)(SB * ((K * 3) | 1) + SB * 3 * (bwd ? 2 : 1)) * sizeof(float); }

}  // namespace

HGS_API int hgs_sh_fwd(int degree, int K, const float* dirs, const float* means, const float* campos,
                       const float* coeffs, const int32_t* radii, int C, int N, int post, float* colors,
                       void* stream) {
    if (degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) || C <= 0 || N < 0) return HGS_ERR_INVALID_ARG;
    if (dirs == nullptr && (means == nullptr || campos == nullptr)) return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sh_smem_bytes(K, false);
    const int grid = hgs_ceil_div(N, SB);
#define LAUNCH(DEG)                                                                                            \
    {                                                                                                          \
        cudaFuncSetAttribute(sh_fwd_kernel<DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        sh_fwd_kernel<DEG><<<grid, SB, smem, st>>>(dirs, means, campos, coeffs, radii, C, N, K, post, colors); \
    }
    switch (degree) {
        case 0: LAUNCH(0) break;
        case 1: LAUNCH(1) break;
        case 2: LAUNCH(2) break;
        case 3: LAUNCH(3) break;
        default: LAUNCH(4) break;
    }
#undef LAUNCH
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_sh_bwd(int degree, int K, const float* dirs, const float* means, const float* campos,
                       const float* coeffs, const int32_t* radii, const float* colors, const float* v_colors,
                       int ld_v_colors, int C, int N, int post, float* v_coeffs, float* v_dirs, float* v_means,
                       void* stream) {
    if (degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) || C <= 0 || N < 0 || ld_v_colors < 3)
        return HGS_ERR_INVALID_ARG;
    if (dirs == nullptr && (means == nullptr || campos == nullptr)) return HGS_ERR_INVALID_ARG;
    if (post && colors == nullptr) return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sh_smem_bytes(K, true);
    const int grid = hgs_ceil_div(N, SB);
#define LAUNCH(DEG)                                                                                                  \
    {                                                                                                                \
        cudaFuncSetAttribute(sh_bwd_kernel<DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
        sh_bwd_kernel<DEG><<<grid, SB, smem, st>>>(dirs, means, campos, coeffs, radii, colors, v_colors,             \
                                                   ld_v_colors, C, N, K, post, v_coeffs, v_dirs, v_means);           \
    }
    switch (degree) {
        case 0: LAUNCH(0) break;
        case 1: LAUNCH(1) break;
        case 2: LAUNCH(2) break;
        case 3: LAUNCH(3) break;
        default: LAUNCH(4) break;
    }
#undef LAUNCH
    HGS_LAUNCH_CHECK();
    return 0;
}
