// Stage a12, fast path: 2DGS (surfel) alpha blending with the same B200 structure as blend3d.cu --
// packed per-surfel records gathered by TMA bulk copies into double-buffered shared memory, per-warp culling,
// halving-butterfly reduction of the 21 per-lane partials, per-warp shared slots, vector reductions to a packed
// per-surfel gradient row.  Semantics identical to the plain kernels in blend2d.cu (and to
// oracle/gsplat_oracle.py::rasterize_to_pixels_2dgs); <= 4 colour channels (every mode the reference uses,
// gaussian_renderer/render.py:56-76).
//
// Cull test: a pixel can only be touched if its ray hits the surfel inside the disc u^2 + v^2 <= 2 tau
// (tau = ln(255 o)) or if it lies within sqrt(tau) pixels of the projected centre (screen-space low-pass term).
// The pixel -> (u, v) map is the homography (u', v', w') = x (M1 x M2) + y (M2 x M0) + (M0 x M1), so the first
// region is the conic u'^2 + v'^2 - 2 tau w'^2 <= 0 in pixel coordinates.  When it is an ellipse the pack
// kernel (in double precision) stores its centre and normalised quadratic form; a warp evaluates the exact
// minimum of that form over its 8x4 pixel rectangle, exactly like the 3DGS kernels -- important for surfels
// seen at grazing angles, whose bounding boxes are huge but whose footprints are thin slivers.  Degenerate or
// unbounded conics are never culled.
#include "blend_common.cuh"
#include "../../include/hgs_raster.h"

namespace {

using namespace hgs;

constexpr int TS = HGS_TILE_SIZE;
constexpr int BLK = TS * TS;
constexpr int REC2_BYTES = 112;
constexpr int SLOT2_BYTES = 112;   // 7 x 16 B: lane-parallel float4 reads are conflict-free
constexpr int FB2 = 64;            // surfels per batch
constexpr int VP2 = 24;            // floats per row of the packed gradient buffer

struct __align__(16) SRec {
    float x, y, opac, _p;   // q0
    float u[3], nx;         // q1: M0, normal.x
    float v[3], ny;         // q2: M1, normal.y
    float w[3], nz;         // q3: M2, normal.z
    float col[4];           // q4
    float ex, ey, QA, QB;   // q5: ellipse centre, normalised form F(d) = QA dx^2 + QB dx dy + QC dy^2 (footprint F <= 1)
    float QC, lp2, kx, ky;  // q6: ..., squared low-pass radius (< 0: never visible), edge-optimum slopes
};
static_assert(sizeof(SRec) == REC2_BYTES, "record size");

__global__ void pack2d_kernel(const float* __restrict__ means2d, const float* __restrict__ ray_transforms,
                              const float* __restrict__ colors, const float* __restrict__ depths,
                              const float* __restrict__ normals, const float* __restrict__ opacities,
                              const int32_t* __restrict__ radii, const int32_t* __restrict__ vis_ids, long long work,
                              int CH, SRec* __restrict__ recs) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= work) return;
    const long long i = vis_ids != nullptr ? (long long)vis_ids[t] : t;
    if (vis_ids == nullptr && radii != nullptr && radii[i] <= 0) return;
    SRec r;
    const float2 m = reinterpret_cast<const float2*>(means2d)[i];
    const float* rt = ray_transforms + i * 9;
    r.x = m.x; r.y = m.y; r.opac = opacities[i]; r._p = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) { r.u[k] = rt[k]; r.v[k] = rt[3 + k]; r.w[k] = rt[6 + k]; }
    r.nx = normals[i * 3]; r.ny = normals[i * 3 + 1]; r.nz = normals[i * 3 + 2];
#pragma unroll
    for (int k = 0; k < 4; ++k) r.col[k] = (k < CH) ? colors[i * CH + k] : ((k == CH && depths != nullptr) ? depths[i] : 0.f);
    // cull data
    float ex = m.x, ey = m.y, QA = 0.f, QB = 0.f, QC = 0.f, lp2 = -1.f;   // QA = QB = QC = 0: never culled
    const float o = r.opac;
    if (o * 255.0f >= 1.0f) {
        const double tau = log(255.0 * (double)o);
        lp2 = (float)tau * 1.02f + 0.05f;
        const double rho2 = 2.0 * tau;
        const double u[3] = {r.u[0], r.u[1], r.u[2]}, v[3] = {r.v[0], r.v[1], r.v[2]}, w[3] = {r.w[0], r.w[1], r.w[2]};
        // (u', v', w') = x a + y b + c
        const double a[3] = {v[1] * w[2] - v[2] * w[1], v[2] * w[0] - v[0] * w[2], v[0] * w[1] - v[1] * w[0]};
        const double b[3] = {w[1] * u[2] - w[2] * u[1], w[2] * u[0] - w[0] * u[2], w[0] * u[1] - w[1] * u[0]};
        const double c[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
        const double sg[3] = {1.0, 1.0, -rho2};
        double qxx = 0, qxy = 0, qyy = 0, qx = 0, qy = 0, q0 = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            qxx += sg[k] * a[k] * a[k]; qxy += sg[k] * a[k] * b[k]; qyy += sg[k] * b[k] * b[k];
            qx += sg[k] * a[k] * c[k];  qy += sg[k] * b[k] * c[k];  q0 += sg[k] * c[k] * c[k];
        }
        const double det2 = qxx * qyy - qxy * qxy;
        if (qxx > 0.0 && qyy > 0.0 && det2 > 1e-9 * qxx * qyy) {
            const double cx = -(qyy * qx - qxy * qy) / det2, cy = -(qxx * qy - qxy * qx) / det2;
            const double kappa = -(q0 + qx * cx + qy * cy);
            if (kappa > 0.0) {
                const double sc = 1.0 / (kappa * 1.02);            // 2 % margin on the footprint
                ex = (float)cx; ey = (float)cy;
                QA = (float)(qxx * sc); QB = (float)(2.0 * qxy * sc); QC = (float)(qyy * sc);
                if (!(QA == QA) || !(QB == QB) || !(QC == QC) || !(ex == ex) || !(ey == ey) || isinf(QA) || isinf(QC)) {
                    ex = m.x; ey = m.y; QA = 0.f; QB = 0.f; QC = 0.f;
                }
            }
        }
    } else {
        QA = 1e30f; QC = 1e30f;                                   // opacity below 1/255: never visible
        ex = -1e30f; ey = -1e30f;
    }
    r.ex = ex; r.ey = ey; r.QA = QA; r.QB = QB;
    r.QC = QC; r.lp2 = lp2;
    r.kx = (QC != 0.f) ? -QB / (2.f * QC) : 0.f;
    r.ky = (QA != 0.f) ? -QB / (2.f * QA) : 0.f;
    float4* dst = reinterpret_cast<float4*>(recs + i);
    const float4* src = reinterpret_cast<const float4*>(&r);
#pragma unroll
    for (int k = 0; k < REC2_BYTES / 16; ++k) dst[k] = src[k];
}

__device__ __forceinline__ const float4* slot2(const unsigned char* stage, int t) {
    return reinterpret_cast<const float4*>(stage + t * SLOT2_BYTES);
}
// keep the surfel if the rectangle of pixel centres [X0,X1]x[Y0,Y1] meets its footprint ellipse
// (minimum of the convex form over the rectangle <= 1) or comes within the low-pass radius of its centre
__device__ __forceinline__ bool cull2_keep(const float4 q0, const float4 q5, const float4 q6, float X0, float X1,
                                           float Y0, float Y1) {
    if (q6.y < 0.f) return false;
    // low-pass disc around the projected centre
    const float ddx = fmaxf(fmaxf(X0 - q0.x, q0.x - X1), 0.f), ddy = fmaxf(fmaxf(Y0 - q0.y, q0.y - Y1), 0.f);
    if (ddx * ddx + ddy * ddy <= q6.y) return true;
    const float u0 = X0 - q5.x, u1 = X1 - q5.x, v0 = Y0 - q5.y, v1 = Y1 - q5.y;
    if (u0 <= 0.f && u1 >= 0.f && v0 <= 0.f && v1 >= 0.f) return true;
    const float A = q5.z, B = q5.w, C = q6.x;
    // value of the form minus a bound of its float32 rounding error (thin, long ellipses cancel heavily)
    auto lower = [&](float u, float v) {
        const float val = (A * u + B * v) * u + C * v * v;
        const float mag = (fabsf(A * u) + fabsf(B * v)) * fabsf(u) + fabsf(C) * v * v;
        return val - 4e-6f * mag;
    };
    float best = lower(u0, fminf(fmaxf(q6.z * u0, v0), v1));
    best = fminf(best, lower(u1, fminf(fmaxf(q6.z * u1, v0), v1)));
    best = fminf(best, lower(fminf(fmaxf(q6.w * v0, u0), u1), v0));
    best = fminf(best, lower(fminf(fmaxf(q6.w * v1, u0), u1), v1));
    return best <= 1.0f;
}

struct Eval2 {
    float hu[3], hv[3], cr[3];
    float sx, sy, w3, w2, dx, dy, vis, alpha, opac;
    bool valid;
};
__device__ __forceinline__ void eval2(const float4* q, float px, float py, Eval2& e) {
    const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
    const float u[3] = {q1.x, q1.y, q1.z}, v[3] = {q2.x, q2.y, q2.z}, w[3] = {q3.x, q3.y, q3.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) { e.hu[k] = px * w[k] - u[k]; e.hv[k] = py * w[k] - v[k]; }
    e.cr[0] = e.hu[1] * e.hv[2] - e.hu[2] * e.hv[1];
    e.cr[1] = e.hu[2] * e.hv[0] - e.hu[0] * e.hv[2];
    e.cr[2] = e.hu[0] * e.hv[1] - e.hu[1] * e.hv[0];
    e.valid = (e.cr[2] != 0.f);
    const float iz = e.valid ? 1.0f / e.cr[2] : 0.f;
    e.sx = e.cr[0] * iz;
    e.sy = e.cr[1] * iz;
    e.w3 = e.sx * e.sx + e.sy * e.sy;
    e.dx = q0.x - px;
    e.dy = q0.y - py;
    e.w2 = HGS_FILTER_INV_SQUARE_2DGS * (e.dx * e.dx + e.dy * e.dy);
    const float sigma = 0.5f * fminf(e.w3, e.w2);
    e.vis = __expf(-sigma);
    e.opac = q0.z;
    e.alpha = fminf(HGS_ALPHA_MAX, e.opac * e.vis);
    if (sigma < 0.f || e.alpha < HGS_ALPHA_MIN) e.valid = false;
}

// =====================================================================================================
// forward
// =====================================================================================================
template <int D, bool NORM_DEPTH, bool DISTORT>
__global__ void __launch_bounds__(BLK) blend2d_fwd_fast_kernel(
    const SRec* __restrict__ recs, const float* __restrict__ backgrounds, int C, int W, int H, int tile_w, int tile_h,
    const int32_t* __restrict__ offsets, const int32_t* __restrict__ flatten_ids, int n_isects,
    float* __restrict__ render_colors, float* __restrict__ render_alphas, float* __restrict__ render_normals,
    float* __restrict__ render_distort, float* __restrict__ render_median, int32_t* __restrict__ last_ids,
    int32_t* __restrict__ median_ids) {
    __shared__ __align__(16) unsigned char s_rec[2][FB2 * SLOT2_BYTES];
    __shared__ __align__(8) uint64_t s_bar[2];
    const TileGeom g = tile_geom(tile_w, tile_h, W, H);
    const int tr = threadIdx.x;
    bool done = !g.inside;
    const int range_start = offsets[g.gtile];
    const int range_end = (g.gtile == C * tile_w * tile_h - 1) ? n_isects : offsets[g.gtile + 1];
    const int nb = (range_end - range_start + FB2 - 1) / FB2;
    if (tr == 0) {
        mbar_init(&s_bar[0], FB2);
        mbar_init(&s_bar[1], FB2);
        mbar_fence_init();
    }
    __syncthreads();
    int g_next = -1;
    if (tr < FB2 && nb > 0) {
        int idx = range_start + tr;
        const int g0 = idx < range_end ? flatten_ids[idx] : -1;
        if (g0 >= 0) {
            mbar_arrive_expect_tx(&s_bar[0], REC2_BYTES);
            bulk_g2s(s_rec[0] + tr * SLOT2_BYTES, recs + g0, REC2_BYTES, &s_bar[0]);
        } else {
            mbar_arrive(&s_bar[0]);
        }
        idx += FB2;
        g_next = idx < range_end ? flatten_ids[idx] : -1;
    }
    float T = 1.0f;
    int cur_idx = 0, median_idx = 0;
    float pix[D], nrm[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < D; ++k) pix[k] = 0.f;
    float distort = 0.f, accum_vd = 0.f, median_depth = 0.f;

    for (int b = 0; b < nb; ++b) {
        const int st = b & 1;
        if (tr < FB2 && b + 1 < nb) {
            const int ns = st ^ 1;
            if (g_next >= 0) {
                mbar_arrive_expect_tx(&s_bar[ns], REC2_BYTES);
                bulk_g2s(s_rec[ns] + tr * SLOT2_BYTES, recs + g_next, REC2_BYTES, &s_bar[ns]);
            } else {
                mbar_arrive(&s_bar[ns]);
            }
            const int idx = range_start + (b + 2) * FB2 + tr;
            g_next = idx < range_end ? flatten_ids[idx] : -1;
        }
        mbar_wait(&s_bar[st], (b >> 1) & 1);
        const int batch_start = range_start + b * FB2;
        const int batch_n = min(FB2, range_end - batch_start);
        const unsigned char* stage = s_rec[st];
        if (!__all_sync(0xFFFFFFFFu, done)) {
            for (int grp = 0; grp * 32 < batch_n; ++grp) {
                const int t = grp * 32 + g.lane;
                bool keep = false;
                if (t < batch_n) {
                    const float4* q = slot2(stage, t);
                    keep = cull2_keep(q[0], q[5], q[6], g.X0, g.X1, g.Y0, g.Y1);
                }
                unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const int tt = grp * 32 + j;
                    const float4* q = slot2(stage, tt);
                    Eval2 e;
                    eval2(q, g.px, g.py, e);
                    if (!done && e.valid) {
                        const float next_T = T * (1.0f - e.alpha);
                        if (next_T <= HGS_T_EPS) {
                            done = true;
                        } else {
                            const float vis = e.alpha * T;
                            const float4 q4 = q[4];
                            const float col[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                            for (int k = 0; k < D; ++k) pix[k] += col[k] * vis;
                            nrm[0] += q[1].w * vis; nrm[1] += q[2].w * vis; nrm[2] += q[3].w * vis;
                            const float depth = col[D - 1];
                            if (DISTORT) {
                                distort += 2.0f * (vis * depth * (1.0f - T) - vis * accum_vd);
                                accum_vd += vis * depth;
                            }
                            if (T > HGS_MEDIAN_T_2DGS) {
                                median_depth = depth;
                                median_idx = batch_start + tt;
                            }
                            cur_idx = batch_start + tt;
                            T = next_T;
                        }
                    }
                }
                if (__all_sync(0xFFFFFFFFu, done)) break;
            }
        }
        if (__syncthreads_count(done) >= BLK) {
            if (b + 1 < nb) mbar_wait(&s_bar[st ^ 1], ((b + 1) >> 1) & 1);
            break;
        }
    }
    if (g.inside) {
        const long long pid = ((long long)g.cam * H + g.pi) * W + g.pj;
        const float alpha_out = 1.0f - T;
        render_alphas[pid] = alpha_out;
        float out[D];
#pragma unroll
        for (int k = 0; k < D; ++k) out[k] = backgrounds == nullptr ? pix[k] : pix[k] + T * backgrounds[g.cam * D + k];
        if (NORM_DEPTH) out[D - 1] = out[D - 1] / fmaxf(alpha_out, HGS_ED_ALPHA_FLOOR);
#pragma unroll
        for (int k = 0; k < D; ++k) render_colors[pid * D + k] = out[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) render_normals[pid * 3 + k] = nrm[k];
        if (DISTORT) render_distort[pid] = distort;
        render_median[pid] = median_depth;
        last_ids[pid] = cur_idx;
        median_ids[pid] = median_idx;
    }
}

// =====================================================================================================
// backward
// =====================================================================================================
constexpr int ACC2_STRIDE = 23;

template <int D>
struct Bwd2Smem {
    unsigned char rec[2][FB2 * SLOT2_BYTES];
    float acc[BLK / 32][FB2 * ACC2_STRIDE];
    int ids[2][FB2];
    unsigned wmask[BLK / 32][FB2 / 32];
    int red[BLK / 32];
    uint64_t bar[2];
};

// per-lane values: [0,1] v_xy (low-pass branch), [2..10] v_M (u,v,w), [11..13] v_normal, [14] v_opacity,
// [15,16] densification gradient, [17..17+D) v_colour
template <int D, bool NORM_DEPTH, bool DISTORT>
__global__ void __launch_bounds__(BLK, 3) blend2d_bwd_fast_kernel(
    const SRec* __restrict__ recs, const float* __restrict__ backgrounds, int C, int W, int H, int tile_w, int tile_h,
    const int32_t* __restrict__ offsets, const int32_t* __restrict__ flatten_ids, int n_isects,
    const float* __restrict__ render_colors, const float* __restrict__ render_alphas,
    const int32_t* __restrict__ last_ids, const int32_t* __restrict__ median_ids,
    const float* __restrict__ v_render_colors, const float* __restrict__ v_render_alphas,
    const float* __restrict__ v_render_normals, const float* __restrict__ v_render_distort,
    const float* __restrict__ v_render_median, float* __restrict__ vpack) {
    constexpr int NV = 17 + D;
    static_assert(NV <= ACC2_STRIDE, "accumulator row too small");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Bwd2Smem<D>& S = *reinterpret_cast<Bwd2Smem<D>*>(smem_raw);
    const TileGeom g = tile_geom(tile_w, tile_h, W, H);
    const int tr = threadIdx.x;
    const int range_start = offsets[g.gtile];
    const int range_end = (g.gtile == C * tile_w * tile_h - 1) ? n_isects : offsets[g.gtile + 1];
    if (range_end <= range_start) return;
    if (tr == 0) {
        mbar_init(&S.bar[0], FB2);
        mbar_init(&S.bar[1], FB2);
        mbar_fence_init();
    }
    const long long pid = ((long long)g.cam * H + min(g.pi, H - 1)) * W + min(g.pj, W - 1);
    const float alpha_out = render_alphas[pid];
    const float T_final = 1.0f - alpha_out;
    float T = T_final;
    float buffer[D], v_c[D], buffer_n[3] = {0.f, 0.f, 0.f}, v_n[3];
    float v_a = g.inside ? v_render_alphas[pid] : 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        buffer[k] = 0.f;
        v_c[k] = g.inside ? v_render_colors[pid * D + k] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) v_n[k] = (g.inside && v_render_normals != nullptr) ? v_render_normals[pid * 3 + k] : 0.f;
    // accumulated (un-normalised) depth of the forward pass, for the distortion terms
    float accum_d = render_colors[pid * D + D - 1];
    if (NORM_DEPTH) {
        const float a_c = fmaxf(alpha_out, HGS_ED_ALPHA_FLOOR);
        const float v_ed = v_c[D - 1];
        v_c[D - 1] = v_ed / a_c;
        if (alpha_out >= HGS_ED_ALPHA_FLOOR && g.inside) v_a += -v_ed * accum_d / a_c;
        accum_d *= a_c;                                           // undo the normalisation
    }
    if (backgrounds != nullptr) accum_d -= T_final * backgrounds[g.cam * D + D - 1];
    float bg_dot = 0.f;
    if (backgrounds != nullptr) {
#pragma unroll
        for (int k = 0; k < D; ++k) bg_dot += backgrounds[g.cam * D + k] * v_c[k];
    }
    const float v_dist = (DISTORT && g.inside) ? v_render_distort[pid] : 0.f;
    const float v_med = (g.inside && v_render_median != nullptr) ? v_render_median[pid] : 0.f;
    const int med_id = g.inside ? median_ids[pid] : -1;
    const float accum_w = 1.0f - T_final;
    float accum_w_buf = accum_w, accum_d_buf = accum_d, distort_buf = 0.f;

    const int bin_final = g.inside ? last_ids[pid] : -1;
    int warp_bin_final = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xFFFFFFFFu, warp_bin_final, o));
    if (g.lane == 0) S.red[g.warp] = warp_bin_final;
    __syncthreads();
    int cta_bin_final = S.red[0];
#pragma unroll
    for (int w = 1; w < BLK / 32; ++w) cta_bin_final = max(cta_bin_final, S.red[w]);
    cta_bin_final = max(cta_bin_final, range_start);
    if (cta_bin_final >= range_end) cta_bin_final = range_end - 1;
    const int top = cta_bin_final;
    const int nb = (top - range_start + 1 + FB2 - 1) / FB2;

    const int my_comp = halving_component<NV>(g.lane);
    const bool is_writer = (g.lane == __ffs(__match_any_sync(0xFFFFFFFFu, my_comp)) - 1);

    int g_next = -1;
    if (tr < FB2) {
        int idx = top - tr;
        const int g0 = idx >= range_start ? flatten_ids[idx] : -1;
        S.ids[0][tr] = g0;
        if (g0 >= 0) {
            mbar_arrive_expect_tx(&S.bar[0], REC2_BYTES);
            bulk_g2s(S.rec[0] + tr * SLOT2_BYTES, recs + g0, REC2_BYTES, &S.bar[0]);
        } else {
            mbar_arrive(&S.bar[0]);
        }
        idx -= FB2;
        g_next = idx >= range_start ? flatten_ids[idx] : -1;
    }

    for (int b = 0; b < nb; ++b) {
        const int st = b & 1;
        if (tr < FB2 && b + 1 < nb) {
            const int ns = st ^ 1;
            S.ids[ns][tr] = g_next;
            if (g_next >= 0) {
                mbar_arrive_expect_tx(&S.bar[ns], REC2_BYTES);
                bulk_g2s(S.rec[ns] + tr * SLOT2_BYTES, recs + g_next, REC2_BYTES, &S.bar[ns]);
            } else {
                mbar_arrive(&S.bar[ns]);
            }
            const int idx = top - (b + 2) * FB2 - tr;
            g_next = idx >= range_start ? flatten_ids[idx] : -1;
        }
        if (g.lane < FB2 / 32) S.wmask[g.warp][g.lane] = 0u;
        mbar_wait(&S.bar[st], (b >> 1) & 1);
        __syncwarp();

        const int batch_end = top - b * FB2;
        const int batch_n = min(FB2, batch_end + 1 - range_start);
        const unsigned char* stage = S.rec[st];
        const int t_first = max(0, batch_end - warp_bin_final);
        for (int grp = t_first >> 5; grp * 32 < batch_n; ++grp) {
            const int t = grp * 32 + g.lane;
            bool keep = false;
            if (t < batch_n && t >= t_first) {
                const float4* q = slot2(stage, t);
                keep = cull2_keep(q[0], q[5], q[6], g.X0, g.X1, g.Y0, g.Y1);
            }
            unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
            unsigned written = 0u;
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const int tt = grp * 32 + j;
                const float4* q = slot2(stage, tt);
                Eval2 e;
                eval2(q, g.px, g.py, e);
                const bool valid = e.valid && (batch_end - tt <= bin_final);
                if (!__any_sync(0xFFFFFFFFu, valid)) continue;
                float val[NV];
#pragma unroll
                for (int k = 0; k < NV; ++k) val[k] = 0.f;
                if (valid) {
                    const float4 q4 = q[4];
                    const float col[4] = {q4.x, q4.y, q4.z, q4.w};
                    const float nn[3] = {q[1].w, q[2].w, q[3].w};
                    const float ra = rcp_approx(1.0f - e.alpha);
                    T *= ra;
                    const float fac = e.alpha * T;
                    float v_alpha = 0.f;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        val[17 + k] = fac * v_c[k];
                        v_alpha += (col[k] * T - buffer[k] * ra) * v_c[k];
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        val[11 + k] = fac * v_n[k];
                        v_alpha += (nn[k] * T - buffer_n[k] * ra) * v_n[k];
                    }
                    v_alpha += T_final * ra * v_a;
                    if (backgrounds != nullptr) v_alpha += -T_final * ra * bg_dot;
                    const float depth = col[D - 1];
                    if (DISTORT) {
                        const float dl_dw =
                            2.0f * (2.0f * (depth * accum_w_buf - accum_d_buf) + (accum_d - depth * accum_w));
                        v_alpha += (dl_dw * T - distort_buf * ra) * v_dist;
                        accum_d_buf -= fac * depth;
                        accum_w_buf -= fac;
                        distort_buf += dl_dw * fac;
                        val[17 + D - 1] += 2.0f * fac * (2.0f - 2.0f * T - accum_w + fac) * v_dist;
                    }
                    if (batch_end - tt == med_id) val[17 + D - 1] += v_med;
                    if (e.opac * e.vis <= HGS_ALPHA_MAX) {
                        const float v_G = e.opac * v_alpha;
                        if (e.w3 <= e.w2) {
                            const float v_sx = v_G * -e.vis * e.sx;
                            const float v_sy = v_G * -e.vis * e.sy;
                            const float iz = 1.0f / e.cr[2];
                            const float vcx = v_sx * iz, vcy = v_sy * iz;
                            const float vcr[3] = {vcx, vcy, -(vcx * e.sx + vcy * e.sy)};
                            const float vhu[3] = {e.hv[1] * vcr[2] - e.hv[2] * vcr[1], e.hv[2] * vcr[0] - e.hv[0] * vcr[2],
                                                  e.hv[0] * vcr[1] - e.hv[1] * vcr[0]};
                            const float vhv[3] = {vcr[1] * e.hu[2] - vcr[2] * e.hu[1], vcr[2] * e.hu[0] - vcr[0] * e.hu[2],
                                                  vcr[0] * e.hu[1] - vcr[1] * e.hu[0]};
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                val[2 + k] = -vhu[k];
                                val[5 + k] = -vhv[k];
                                val[8 + k] = g.px * vhu[k] + g.py * vhv[k];
                            }
                            const float wz = q[3].z;
                            val[15] = val[4] * wz;      // v_u.z * depth
                            val[16] = val[7] * wz;      // v_v.z * depth
                        } else {
                            val[0] = v_G * -e.vis * HGS_FILTER_INV_SQUARE_2DGS * e.dx;
                            val[1] = v_G * -e.vis * HGS_FILTER_INV_SQUARE_2DGS * e.dy;
                        }
                        val[14] = e.vis * v_alpha;
                    }
#pragma unroll
                    for (int k = 0; k < D; ++k) buffer[k] += col[k] * fac;
#pragma unroll
                    for (int k = 0; k < 3; ++k) buffer_n[k] += nn[k] * fac;
                }
                const float r = halving_reduce<NV>(val, g.lane);
                if (is_writer) S.acc[g.warp][tt * ACC2_STRIDE + my_comp] = r;
                written |= 1u << j;
            }
            if (g.lane == 0) S.wmask[g.warp][grp] = written;
        }
        __syncthreads();
        if (tr < batch_n) {
            float sum[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) sum[k] = 0.f;
            bool any = false;
#pragma unroll
            for (int w = 0; w < BLK / 32; ++w) {
                if ((S.wmask[w][tr >> 5] >> (tr & 31)) & 1u) {
                    any = true;
#pragma unroll
                    for (int k = 0; k < NV; ++k) sum[k] += S.acc[w][tr * ACC2_STRIDE + k];
                }
            }
            if (any) {
                // row: [0:2] xy | [2:11] M | [11:14] normal | [14] opacity | [15] - | [16:20] colour | [20:22] densify
                float* row = vpack + (long long)S.ids[st][tr] * VP2;
                red_add_v4(row, sum[0], sum[1], sum[2], sum[3]);
                red_add_v4(row + 4, sum[4], sum[5], sum[6], sum[7]);
                red_add_v4(row + 8, sum[8], sum[9], sum[10], sum[11]);
                red_add_v4(row + 12, sum[12], sum[13], sum[14], 0.f);
                float c4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < D; ++k) c4[k] = sum[17 + k];
                red_add_v4(row + 16, c4[0], c4[1], c4[2], c4[3]);
                red_add_v2(row + 20, sum[15], sum[16]);
            }
        }
        __syncthreads();
    }
}

template <int D, bool NORM, bool DIST>
int launch_fwd2(const SRec* recs, const float* backgrounds, int C, int W, int H, int tile_w, int tile_h,
                const int32_t* offsets, const int32_t* flatten_ids, int n_isects, float* render_colors,
                float* render_alphas, float* render_normals, float* render_distort, float* render_median,
                int32_t* last_ids, int32_t* median_ids, cudaStream_t st) {
    dim3 grid(tile_w, tile_h, C);
    blend2d_fwd_fast_kernel<D, NORM, DIST><<<grid, BLK, 0, st>>>(recs, backgrounds, C, W, H, tile_w, tile_h, offsets,
                                                                 flatten_ids, n_isects, render_colors, render_alphas,
                                                                 render_normals, render_distort, render_median,
                                                                 last_ids, median_ids);
    HGS_LAUNCH_CHECK();
    return 0;
}

template <int D, bool NORM, bool DIST>
int launch_bwd2(const SRec* recs, const float* backgrounds, int C, int W, int H, int tile_w, int tile_h,
                const int32_t* offsets, const int32_t* flatten_ids, int n_isects, const float* render_colors,
                const float* render_alphas, const int32_t* last_ids, const int32_t* median_ids,
                const float* v_render_colors, const float* v_render_alphas, const float* v_render_normals,
                const float* v_render_distort, const float* v_render_median, float* vpack, cudaStream_t st) {
    dim3 grid(tile_w, tile_h, C);
    const int smem = (int)sizeof(Bwd2Smem<D>);
    cudaError_t e = cudaFuncSetAttribute(blend2d_bwd_fast_kernel<D, NORM, DIST>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    blend2d_bwd_fast_kernel<D, NORM, DIST><<<grid, BLK, smem, st>>>(
        recs, backgrounds, C, W, H, tile_w, tile_h, offsets, flatten_ids, n_isects, render_colors, render_alphas,
        last_ids, median_ids, v_render_colors, v_render_alphas, v_render_normals, v_render_distort, v_render_median,
        vpack);
    HGS_LAUNCH_CHECK();
    return 0;
}

}  // namespace

HGS_API size_t hgs_blend2d_pack_bytes(long long CN) { return (size_t)(CN > 0 ? CN : 1) * REC2_BYTES; }

HGS_API int hgs_blend2d_pack(const float* means2d, const float* ray_transforms, const float* colors,
                             const float* depths, const float* normals, const float* opacities,
                             const int32_t* radii, const int32_t* vis_ids, long long n_vis, long long CN, int CH,
                             void* records, void* stream) {
    if (CN < 0 || n_vis < 0 || CH < 0 || CH + (depths != nullptr ? 1 : 0) > 4) return HGS_ERR_INVALID_ARG;
    const long long work = vis_ids != nullptr ? n_vis : CN;
    if (work == 0) return 0;
    pack2d_kernel<<<hgs_ceil_div(work, 256), 256, 0, (cudaStream_t)stream>>>(
        means2d, ray_transforms, colors, depths, normals, opacities, radii, vis_ids, work, CH, (SRec*)records);
    HGS_LAUNCH_CHECK();
    return 0;
}

#define HGS_DISPATCH_FAST2(D, NORM, DIST, CALL)                                          \
    switch ((D) * 4 + ((NORM) ? 2 : 0) + ((DIST) ? 1 : 0)) {                             \
        case 4: return CALL(1, false, false);                                            \
        case 5: return CALL(1, false, true);                                             \
        case 6: return CALL(1, true, false);                                             \
        case 7: return CALL(1, true, true);                                              \
        case 12: return CALL(3, false, false);                                           \
        case 13: return CALL(3, false, true);                                            \
        case 16: return CALL(4, false, false);                                           \
        case 17: return CALL(4, false, true);                                            \
        case 18: return CALL(4, true, false);                                            \
        case 19: return CALL(4, true, true);                                             \
        default: return HGS_ERR_INVALID_ARG;                                             \
    }

HGS_API int hgs_blend2d_fwd_packed(const void* records, const float* backgrounds, int C, int D, int normalize_depth,
                                   int width, int height, int tile_size, const int32_t* isect_offsets,
                                   const int32_t* flatten_ids, long long n_isects, float* render_colors,
                                   float* render_alphas, float* render_normals, float* render_distort,
                                   float* render_median, int32_t* last_ids, int32_t* median_ids, void* stream) {
    if (tile_size != TS || C <= 0 || width <= 0 || height <= 0 || D < 1 || D > 4 || n_isects < 0) return HGS_ERR_INVALID_ARG;
    if (n_isects >= (1ll << 31)) return HGS_ERR_TOO_LARGE;
    const int tile_w = (width + TS - 1) / TS, tile_h = (height + TS - 1) / TS;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DD, NN, XX)                                                                                         \
    launch_fwd2<DD, NN, XX>((const SRec*)records, backgrounds, C, width, height, tile_w, tile_h, isect_offsets,   \
                            flatten_ids, (int)n_isects, render_colors, render_alphas, render_normals,             \
                            render_distort, render_median, last_ids, median_ids, st)
    HGS_DISPATCH_FAST2(D, normalize_depth != 0, render_distort != nullptr, CALL)
#undef CALL
}

HGS_API int hgs_blend2d_bwd_packed(const void* records, const float* backgrounds, int C, int D, int normalize_depth,
                                   int width, int height, int tile_size, const int32_t* isect_offsets,
                                   const int32_t* flatten_ids, long long n_isects, const float* render_colors,
                                   const float* render_alphas, const int32_t* last_ids, const int32_t* median_ids,
                                   const float* v_render_colors, const float* v_render_alphas,
                                   const float* v_render_normals, const float* v_render_distort,
                                   const float* v_render_median, float* vpack, void* stream) {
    if (tile_size != TS || C <= 0 || width <= 0 || height <= 0 || D < 1 || D > 4 || n_isects < 0) return HGS_ERR_INVALID_ARG;
    if (n_isects >= (1ll << 31)) return HGS_ERR_TOO_LARGE;
    if (n_isects == 0) return 0;
    const int tile_w = (width + TS - 1) / TS, tile_h = (height + TS - 1) / TS;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DD, NN, XX)                                                                                         \
    launch_bwd2<DD, NN, XX>((const SRec*)records, backgrounds, C, width, height, tile_w, tile_h, isect_offsets,   \
                            flatten_ids, (int)n_isects, render_colors, render_alphas, last_ids, median_ids,       \
                            v_render_colors, v_render_alphas, v_render_normals, v_render_distort, v_render_median, \
                            vpack, st)
    HGS_DISPATCH_FAST2(D, normalize_depth != 0, v_render_distort != nullptr, CALL)
#undef CALL
}
