// Stage a11: forward alpha-blend rasterization and its backward (3DGS).
//
// Replaces gsplat's rasterize_to_pixels as reached inside gsplat.rasterization (reference
// gaussian_renderer/render.py:40-54).  Semantics restated in oracle/gsplat_oracle.py::rasterize_to_pixels:
// pixel centre (x+.5, y+.5); sigma = .5(a dx^2 + c dy^2) + b dx dy; alpha = min(.999, o exp(-sigma));
// skip if sigma < 0 or alpha < 1/255; stop before the Gaussian that brings T to <= 1e-4.
//
// B200 design (the *_fast kernels, <= 4 render channels -- every mode the reference uses):
//   * pack: one 64-byte record per visible Gaussian (centre, conic pre-multiplied by -log2(e)/2 so that
//     alpha = o * 2^(A dx^2 + B dx dy + C dy^2), opacity, cut-off exponent, colour, edge-optimum slopes).
//   * one CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel sub-tile.  Records of a batch of 128
//     intersections are gathered into shared memory with per-thread TMA bulk copies (cp.async.bulk,
//     mbarrier transaction counting), double buffered, so staging overlaps blending.
//   * per-warp culling: each lane tests one Gaussian of a group of 32 against the warp's sub-tile with an
//     exact concave-maximum-over-rectangle test (plus margin): Gaussians that cannot reach alpha >= 1/255 on
//     any of the warp's 32 pixels are skipped for the whole warp -- result-identical to evaluating them.
//   * backward: per-lane partials are summed over the warp with a halving butterfly (12 shuffles for 10
//     values instead of 50), written to per-warp shared-memory slots, summed over the 8 warps per batch and
//     added to global memory with 3 vector reductions (red.global.add.v4.f32) per (tile, Gaussian).
//   * expected-depth normalisation (render modes ED / RGB+ED) is fused into the epilogue / prologue.
// Roofline: FP32 FMA / MUFU.EX2 / shared-memory pipes (not HBM, no tensor cores: no dense contraction).
// The plain kernels (one pixel per thread, gsplat-style staging) remain for 5..8 channels.
#include "blend_common.cuh"
#include "../../include/hgs_raster.h"

namespace {

using namespace hgs;

constexpr int TS = HGS_TILE_SIZE;
constexpr int BLK = TS * TS;  // 256 threads

// =====================================================================================================
// packed per-Gaussian record
// =====================================================================================================
constexpr float LN2 = 0.6931471805599453f;
constexpr int SLOT_BYTES = 80;   // shared-memory slot (stride 80 B: lane-parallel float4 reads are conflict-free)
constexpr int FB = 128;          // forward batch size (Gaussians staged per mbarrier phase)
constexpr int FB_BWD = 96;       // backward batch size: 4 CTAs/SM fit in shared memory (50 KB each)
constexpr float CULL_MARGIN = 0.02f;  // in log2 units; conservative (float error of the bound is ~1e-5)

__global__ void pack3d_kernel(const float* __restrict__ means2d, const float* __restrict__ conics,
                              const float* __restrict__ colors, const float* __restrict__ depths,
                              const float* __restrict__ opacities, const int32_t* __restrict__ radii,
                              const int32_t* __restrict__ vis_ids, long long CN,
                              const long long* __restrict__ n_dev, int CH, GRec* __restrict__ recs) {
    // CN = number of work items: all C*N Gaussians (vis_ids == NULL) or the visible ones listed in vis_ids
    // (n_dev != NULL: the count lives on the device and CN is only a bound)
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (n_dev != nullptr ? *n_dev : CN)) return;
    const long long i = vis_ids != nullptr ? (long long)vis_ids[t] : t;
    if (vis_ids == nullptr && radii != nullptr && radii[i] <= 0) return;
    const float2 m = reinterpret_cast<const float2*>(means2d)[i];
    const float a = conics[i * 3 + 0], b = conics[i * 3 + 1], c = conics[i * 3 + 2];
    const float o = opacities[i];
    float col[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) col[k] = (k < CH) ? colors[i * CH + k] : ((k == CH && depths != nullptr) ? depths[i] : 0.f);
    const GRec r = make_grec(m.x, m.y, a, b, c, o, col);
    float4* dst = reinterpret_cast<float4*>(recs + i);
    const float4* src = reinterpret_cast<const float4*>(&r);
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[k] = src[k];
}

// can this Gaussian reach alpha >= 1/255 (and sigma >= 0) anywhere on the pixel-centre rectangle
// [X0,X1] x [Y0,Y1]?  Maximum of the concave quadratic p2(u,v) = A u^2 + B u v + C v^2 over the rectangle.
__device__ __forceinline__ bool cull_keep(const float4 q0, const float4 q1, const float4 q3, float X0, float X1,
                                          float Y0, float Y1) {
    const float u0 = X0 - q0.x, u1 = X1 - q0.x, v0 = Y0 - q0.y, v1 = Y1 - q0.y;
    const float A = q0.z, B = q0.w, C = q1.x;
    const float thr = q1.z - CULL_MARGIN;
    if (u0 <= 0.f && u1 >= 0.f && v0 <= 0.f && v1 >= 0.f) return thr <= 0.f;
    // value of the exponent plus a bound of its float32 rounding error (long thin Gaussians cancel heavily)
    auto upper = [&](float u, float v) {
        const float val = (A * u + B * v) * u + C * v * v;
        const float mag = (fabsf(A * u) + fabsf(B * v)) * fabsf(u) + fabsf(C) * v * v;
        return val + 4e-6f * mag;
    };
    float best = upper(u0, fminf(fmaxf(q3.x * u0, v0), v1));
    best = fmaxf(best, upper(u1, fminf(fmaxf(q3.x * u1, v0), v1)));
    best = fmaxf(best, upper(fminf(fmaxf(q3.y * v0, u0), u1), v0));
    best = fmaxf(best, upper(fminf(fmaxf(q3.y * v1, u0), u1), v1));
    return best >= thr;
}

__device__ __forceinline__ const float4* slot_q(const unsigned char* stage, int t) {
    return reinterpret_cast<const float4*>(stage + t * SLOT_BYTES);
}

// =====================================================================================================
// fast forward
// =====================================================================================================
template <int D, bool NORM_DEPTH>
__global__ void __launch_bounds__(BLK, 5) blend3d_fwd_fast_kernel(
    const GRec* __restrict__ recs, const float* __restrict__ backgrounds, int C, int W, int H, int tile_w, int tile_h,
    const int32_t* __restrict__ offsets, const int32_t* __restrict__ flatten_ids, int n_isects,
    float* __restrict__ render_colors, float* __restrict__ render_alphas, int32_t* __restrict__ last_ids) {
    __shared__ __align__(16) unsigned char s_rec[2][FB * SLOT_BYTES];
    __shared__ __align__(8) uint64_t s_bar[2];
    const TileGeom g = tile_geom(tile_w, tile_h, W, H);
    const int tr = threadIdx.x;
    bool done = !g.inside;

    const int range_start = offsets[g.gtile];
    const int range_end = (g.gtile == C * tile_w * tile_h - 1) ? n_isects : offsets[g.gtile + 1];
    const int nb = (range_end - range_start + FB - 1) / FB;

    if (tr == 0) {
        mbar_init(&s_bar[0], FB);
        mbar_init(&s_bar[1], FB);
        mbar_fence_init();
    }
    __syncthreads();

    // loader threads (tr < FB): one record per thread and batch
    int g_next = -1;
    if (tr < FB && nb > 0) {
        int idx = range_start + tr;
        const int g0 = idx < range_end ? flatten_ids[idx] : -1;
        if (g0 >= 0) {
            mbar_arrive_expect_tx(&s_bar[0], REC_BYTES);
            bulk_g2s(s_rec[0] + tr * SLOT_BYTES, recs + g0, REC_BYTES, &s_bar[0]);
        } else {
            mbar_arrive(&s_bar[0]);
        }
        idx += FB;
        g_next = idx < range_end ? flatten_ids[idx] : -1;
    }

    float T = 1.0f;
    int cur_idx = 0;
    float pix[D];
#pragma unroll
    for (int k = 0; k < D; ++k) pix[k] = 0.f;
    // a finished pixel (outside the image, or transmittance exhausted) is one whose x coordinate is NaN: no separate
    // flag to test in the inner loop
    bool finished = done;
    float px = done ? __int_as_float(0x7fc00000) : g.px;
    __shared__ unsigned short s_keep[BLK / 32][32];   // slots of the records that survive the warp's cull, in order
    unsigned short* keep_list = s_keep[g.warp];

    for (int b = 0; b < nb; ++b) {
        const int st = b & 1;
        if (tr < FB && b + 1 < nb) {
            const int ns = st ^ 1;
            if (g_next >= 0) {
                mbar_arrive_expect_tx(&s_bar[ns], REC_BYTES);
                bulk_g2s(s_rec[ns] + tr * SLOT_BYTES, recs + g_next, REC_BYTES, &s_bar[ns]);
            } else {
                mbar_arrive(&s_bar[ns]);
            }
            const int idx = range_start + (b + 2) * FB + tr;
            g_next = idx < range_end ? flatten_ids[idx] : -1;
        }
        mbar_wait(&s_bar[st], (b >> 1) & 1);

        const int batch_start = range_start + b * FB;
        const int batch_n = min(FB, range_end - batch_start);
        const unsigned char* stage = s_rec[st];
        if (!__all_sync(0xFFFFFFFFu, finished)) {
            for (int grp = 0; grp * 32 < batch_n; ++grp) {
                const int t = grp * 32 + g.lane;
                bool keep = false;
                if (t < batch_n) {
                    const float4* q = slot_q(stage, t);
                    keep = cull_keep(q[0], q[1], q[3], g.X0, g.X1, g.Y0, g.Y1);
                }
                const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
                if (m == 0u) continue;
                if (keep) keep_list[__popc(m & ((1u << g.lane) - 1u))] = (unsigned short)t;
                __syncwarp();
                const int n_keep = __popc(m);
                // a finished pixel is one whose x coordinate is NaN: the exponent is NaN, "p2 <= 0" fails, and the
                // alpha threshold stays an immediate (no per-iteration copies of a loop-carried threshold)
                for (int i = 0; i < n_keep; ++i) {
                    const int tt = keep_list[i];
                    const float4* q = slot_q(stage, tt);
                    const float4 q0 = q[0];
                    const float2 q1 = *reinterpret_cast<const float2*>(q + 1);
                    const float dx = q0.x - px, dy = q0.y - g.py;
                    const float p2 = (q0.z * dx + q0.w * dy) * dx + (q1.x * dy) * dy;
                    const float alpha = fminf(HGS_ALPHA_MAX, q1.y * ex2_approx(p2));
                    if (p2 <= 0.f && alpha >= HGS_ALPHA_MIN) {
                        const float next_T = T * (1.0f - alpha);
                        // if (next_T <= T_EPS) px = NaN, as one predicated move
                        asm("{\n\t.reg .pred ps;\n\t"
                            "setp.le.f32 ps, %1, %2;\n\t"
                            "@ps mov.b32 %0, 0x7fc00000;\n\t}"
                            : "+f"(px)
                            : "f"(next_T), "f"(HGS_T_EPS));
                        if (!(next_T <= HGS_T_EPS)) {
                            const float vis = alpha * T;
                            const float4 q2 = q[2];
                            pix[0] += q2.x * vis;
                            if (D > 1) pix[1] += q2.y * vis;
                            if (D > 2) pix[2] += q2.z * vis;
                            if (D > 3) pix[3] += q2.w * vis;
                            cur_idx = batch_start + tt;
                            T = next_T;
                        }
                    }
                }
                finished = (px != px);
                __syncwarp();
                if (__all_sync(0xFFFFFFFFu, finished)) break;
            }
        }
        if (__syncthreads_count(finished) >= BLK) {
            if (b + 1 < nb) mbar_wait(&s_bar[st ^ 1], ((b + 1) >> 1) & 1);  // drain the in-flight prefetch
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
        if (D == 4) {
            reinterpret_cast<float4*>(render_colors)[pid] = make_float4(out[0], out[1], out[2], out[3]);
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) render_colors[pid * D + k] = out[k];
        }
        last_ids[pid] = cur_idx;
    }
}

// =====================================================================================================
// pair statistics (measurement aid, never on the timed path): P_eval = (pixel, Gaussian) pairs a per-pixel
// front-to-back loop visits before the pixel stops, P_blend = pairs that pass the alpha test and are blended
// =====================================================================================================
__global__ void __launch_bounds__(BLK) blend3d_stats_kernel(const GRec* __restrict__ recs, int C, int W, int H,
                                                            int tile_w, int tile_h,
                                                            const int32_t* __restrict__ offsets,
                                                            const int32_t* __restrict__ flatten_ids, int n_isects,
                                                            unsigned long long* __restrict__ counters) {
    const TileGeom g = tile_geom(tile_w, tile_h, W, H);
    const int range_start = offsets[g.gtile];
    const int range_end = (g.gtile == C * tile_w * tile_h - 1) ? n_isects : offsets[g.gtile + 1];
    unsigned long long n_eval = 0, n_blend = 0;
    if (g.inside) {
        float T = 1.0f;
        for (int idx = range_start; idx < range_end; ++idx) {
            const float4* q = reinterpret_cast<const float4*>(recs + flatten_ids[idx]);
            const float4 q0 = q[0], q1 = q[1];
            ++n_eval;
            const float dx = q0.x - g.px, dy = q0.y - g.py;
            const float p2 = (q0.z * dx + q0.w * dy) * dx + (q1.x * dy) * dy;
            const float alpha = fminf(HGS_ALPHA_MAX, q1.y * ex2_approx(p2));
            if (p2 > 0.f || alpha < HGS_ALPHA_MIN) continue;
            const float next_T = T * (1.0f - alpha);
            if (next_T <= HGS_T_EPS) break;
            ++n_blend;
            T = next_T;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_eval += __shfl_xor_sync(0xFFFFFFFFu, n_eval, o);
        n_blend += __shfl_xor_sync(0xFFFFFFFFu, n_blend, o);
    }
    // culling statistics (whole tile range, no early termination): (warp, Gaussian) pairs surviving the 8x4 cull,
    // and the iteration counts two independent half-warps would need with 4x4 / 8x2 half rectangles
    unsigned long long n_full = 0, n_h44 = 0, n_h82 = 0;
    for (int base = range_start; base < range_end; base += 32) {
        const int idx = base + g.lane;
        bool k0 = false, ka = false, kb = false, kc = false, kd = false;
        if (idx < range_end) {
            const float4* q = reinterpret_cast<const float4*>(recs + flatten_ids[idx]);
            const float4 q0 = q[0], q1 = q[1], q3 = q[3];
            k0 = cull_keep(q0, q1, q3, g.X0, g.X1, g.Y0, g.Y1);
            ka = cull_keep(q0, q1, q3, g.X0, g.X0 + 3.f, g.Y0, g.Y1);
            kb = cull_keep(q0, q1, q3, g.X0 + 4.f, g.X1, g.Y0, g.Y1);
            kc = cull_keep(q0, q1, q3, g.X0, g.X1, g.Y0, g.Y0 + 1.f);
            kd = cull_keep(q0, q1, q3, g.X0, g.X1, g.Y0 + 2.f, g.Y1);
        }
        n_full += __popc(__ballot_sync(0xFFFFFFFFu, k0));
        n_h44 += max(__popc(__ballot_sync(0xFFFFFFFFu, ka)), __popc(__ballot_sync(0xFFFFFFFFu, kb)));
        n_h82 += max(__popc(__ballot_sync(0xFFFFFFFFu, kc)), __popc(__ballot_sync(0xFFFFFFFFu, kd)));
    }
    if (g.lane == 0) {
        atomicAdd(&counters[0], n_eval);
        atomicAdd(&counters[1], n_blend);
        atomicAdd(&counters[2], n_full);
        atomicAdd(&counters[3], n_h44);
        atomicAdd(&counters[4], n_h82);
    }
}

// =====================================================================================================
// fast backward
// =====================================================================================================
constexpr int ACC_STRIDE = SLOT_BYTES / 8;  // accumulator row of a slot: 10 floats at HALF the slot's byte offset
constexpr int VP = 12;          // floats per row of the packed gradient buffer

template <int D, int FBB>
struct BwdSmem {
    unsigned char rec[2][FBB * SLOT_BYTES];
    float acc[BLK / 32][FBB * ACC_STRIDE];
    int ids[2][FBB];
    unsigned wmask[BLK / 32][FBB / 32];
    unsigned short keep[BLK / 32][32];   // slots that survive the warp's cull, in order
    int red[BLK / 32];
    uint64_t bar[2];
};

template <int D, bool NORM_DEPTH, int FBB>
__global__ void __launch_bounds__(BLK, 4) blend3d_bwd_fast_kernel(
    const GRec* __restrict__ recs, const float* __restrict__ backgrounds, int C, int W, int H, int tile_w, int tile_h,
    const int32_t* __restrict__ offsets, const int32_t* __restrict__ flatten_ids, int n_isects,
    const float* __restrict__ render_colors, const float* __restrict__ render_alphas,
    const int32_t* __restrict__ last_ids, const float* __restrict__ v_render_colors,
    const float* __restrict__ v_render_alphas, float* __restrict__ vpack) {
    constexpr int NV = 6 + D;
    static_assert(NV <= ACC_STRIDE, "accumulator row too small");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem<D, FBB>& S = *reinterpret_cast<BwdSmem<D, FBB>*>(smem_raw);
    const TileGeom g = tile_geom(tile_w, tile_h, W, H);
    const int tr = threadIdx.x;

    const int range_start = offsets[g.gtile];
    const int range_end = (g.gtile == C * tile_w * tile_h - 1) ? n_isects : offsets[g.gtile + 1];
    if (range_end <= range_start) return;

    if (tr == 0) {
        mbar_init(&S.bar[0], FBB);
        mbar_init(&S.bar[1], FBB);
        mbar_fence_init();
    }

    // per-pixel state
    const long long pid = ((long long)g.cam * H + min(g.pi, H - 1)) * W + min(g.pj, W - 1);
    const float alpha_out = render_alphas[pid];
    const float T_final = 1.0f - alpha_out;
    float T = T_final;
    float v_c[D];
    float v_a = g.inside ? v_render_alphas[pid] : 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) v_c[k] = g.inside ? v_render_colors[pid * D + k] : 0.f;
    if (NORM_DEPTH) {
        // out_d = acc_d / max(alpha, eps): fold the quotient rule into v_c[D-1] and v_a
        const float a_c = fmaxf(alpha_out, HGS_ED_ALPHA_FLOOR);
        const float v_ed = v_c[D - 1];
        v_c[D - 1] = v_ed / a_c;
        if (alpha_out >= HGS_ED_ALPHA_FLOOR && g.inside) v_a += -v_ed * render_colors[pid * D + D - 1] / a_c;
    }
    float bg_dot = 0.f;
    if (backgrounds != nullptr) {
#pragma unroll
        for (int k = 0; k < D; ++k) bg_dot += backgrounds[g.cam * D + k] * v_c[k];
    }
    // d(out)/d(alpha_i) = T_i * (c_i . v_c) + (T_final (v_a - bg.v_c) - S_i) / (1 - alpha_i), where
    // S_i = sum over the Gaussians behind i of w_j (c_j . v_c) is carried as ONE scalar (s_behind)
    const float c0 = T_final * (v_a - bg_dot);
    float s_behind = 0.f;
    const int bin_final = g.inside ? last_ids[pid] : -1;
    int warp_bin_final = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xFFFFFFFFu, warp_bin_final, o));
    if (g.lane == 0) S.red[g.warp] = warp_bin_final;
    __syncthreads();
    int cta_bin_final = S.red[0];
#pragma unroll
    for (int w = 1; w < BLK / 32; ++w) cta_bin_final = max(cta_bin_final, S.red[w]);
    // gsplat semantics: last_ids defaults to 0, so index 0 is always replayed by the tile that owns it
    cta_bin_final = max(cta_bin_final, range_start);
    if (cta_bin_final >= range_end) cta_bin_final = range_end - 1;

    // batches run back to front over [range_start, cta_bin_final]
    const int top = cta_bin_final;
    const int nb = (top - range_start + 1 + FBB - 1) / FBB;

    // every lane ends the butterfly with the full sum of ONE component (halving_component) and stores it itself
    const int my_comp = halving_component<NV>(g.lane);
    unsigned char* const acc_lane = reinterpret_cast<unsigned char*>(&S.acc[g.warp][my_comp]);
    const LaneMasks lm(g.lane);
    // colour-gradient pairs in the lane's keep / send order of the first butterfly level
    const bool up16 = (g.lane & 16) != 0;
    float vK[D / 2 + 1], vS[D / 2 + 1];
#pragma unroll
    for (int k = 0; k < D / 2; ++k) {
        vK[k] = up16 ? v_c[2 * k + 1] : v_c[2 * k];
        vS[k] = up16 ? v_c[2 * k] : v_c[2 * k + 1];
    }

    int g_next = -1;
    if (tr < FBB) {
        int idx = top - tr;
        const int g0 = idx >= range_start ? flatten_ids[idx] : -1;
        S.ids[0][tr] = g0;
        if (g0 >= 0) {
            mbar_arrive_expect_tx(&S.bar[0], REC_BYTES);
            bulk_g2s(S.rec[0] + tr * SLOT_BYTES, recs + g0, REC_BYTES, &S.bar[0]);
        } else {
            mbar_arrive(&S.bar[0]);
        }
        idx -= FBB;
        g_next = idx >= range_start ? flatten_ids[idx] : -1;
    }

    for (int b = 0; b < nb; ++b) {
        const int st = b & 1;
        if (tr < FBB && b + 1 < nb) {
            const int ns = st ^ 1;
            S.ids[ns][tr] = g_next;
            if (g_next >= 0) {
                mbar_arrive_expect_tx(&S.bar[ns], REC_BYTES);
                bulk_g2s(S.rec[ns] + tr * SLOT_BYTES, recs + g_next, REC_BYTES, &S.bar[ns]);
            } else {
                mbar_arrive(&S.bar[ns]);
            }
            const int idx = top - (b + 2) * FBB - tr;
            g_next = idx >= range_start ? flatten_ids[idx] : -1;
        }
        if (g.lane < FBB / 32) S.wmask[g.warp][g.lane] = 0u;
        mbar_wait(&S.bar[st], (b >> 1) & 1);
        __syncwarp();

        const int batch_end = top - b * FBB;                       // element t <-> index batch_end - t
        const int batch_n = min(FBB, batch_end + 1 - range_start);
        const unsigned char* stage = S.rec[st];
        const int t_first = max(0, batch_end - warp_bin_final);
        // this pixel replays slots t >= batch_end - bin_final (as a slot byte offset, clamped against overflow)
        const int off_min = max(-1, min(batch_end - bin_final, FBB)) * SLOT_BYTES;
        for (int grp = t_first >> 5; grp * 32 < batch_n; ++grp) {
            const int t = grp * 32 + g.lane;
            bool keep = false;
            if (t < batch_n && t >= t_first) {
                const float4* q = slot_q(stage, t);
                keep = cull_keep(q[0], q[1], q[3], g.X0, g.X1, g.Y0, g.Y1);
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
            // lean loop: no divergent branch around the gradient math (lanes that do not blend select zeros);
            // the first butterfly level is fed with values already in the lane's keep / send order (u, w = the
            // lane's own / its partner's coordinate); lane-side selections are LOP3s on mask registers; the
            // survivor list holds slot byte offsets (= 2 x the accumulator row offset), bit 15 set once a survivor
            // turned out to blend nowhere in the warp
            int pos = 0;
            if (m != 0u) {
                pos = __popc(m & ((1u << g.lane) - 1u));
                if (keep) S.keep[g.warp][pos] = (unsigned short)(t * SLOT_BYTES);
                __syncwarp();
            }
            unsigned short* const kl = S.keep[g.warp];
            const int n_keep = __popc(m);
            for (int i = 0; i < n_keep; ++i) {
                const int off = kl[i];
                const float4* q = reinterpret_cast<const float4*>(stage + off);
                const float4 q0 = q[0];
                const float2 q1 = *reinterpret_cast<const float2*>(q + 1);
                const float dx = q0.x - g.px, dy = q0.y - g.py;
                const float p2 = (q0.z * dx + q0.w * dy) * dx + (q1.x * dy) * dy;
                const float vis = ex2_approx(p2);
                const float opac = q1.y;
                const float av = opac * vis;
                const float alpha = fminf(HGS_ALPHA_MAX, av);
                const bool valid = (off >= off_min) && p2 <= 0.f && alpha >= HGS_ALPHA_MIN;
                if (!__any_sync(0xFFFFFFFFu, valid)) {
                    kl[i] = (unsigned short)(off | 0x8000);   // every lane stores the same value: no divergence
                    continue;
                }
                const float4 q2 = q[2];
                const float col[4] = {q2.x, q2.y, q2.z, q2.w};
                const float ra = rcp_approx(1.0f - alpha);  // 1 - alpha in [1e-3, 1]: 1-ulp reciprocal
                // if (valid) T *= ra, as ONE predicated multiply (the compiler prefers multiply + select)
                asm("{\n\t.reg .pred pv;\n\t"
                    "setp.ne.b32 pv, %1, 0;\n\t"
                    "@pv mul.f32 %0, %0, %2;\n\t}"
                    : "+f"(T)
                    : "r"((int)valid), "f"(ra));
                const float fac = (valid ? alpha : 0.f) * T;
                float cdot = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) cdot += col[k] * v_c[k];
                const float v_alpha = T * cdot + ra * (c0 - s_behind);
                s_behind += fac * cdot;
                // moments of v_sigma over the pixels; the flush turns them into v_means2d / v_conics
                float v_o;  // = (valid && av <= ALPHA_MAX) ? vis * v_alpha : 0, one compare chained on `valid`
                asm("{\n\t.reg .pred pv, pq;\n\t"
                    "setp.ne.b32 pv, %1, 0;\n\t"
                    "setp.le.and.f32 pq, %2, %3, pv;\n\t"
                    "selp.f32 %0, %4, 0f00000000, pq;\n\t}"
                    : "=f"(v_o)
                    : "r"((int)valid), "f"(av), "f"(HGS_ALPHA_MAX), "f"(vis * v_alpha));
                const float v_sigma = -opac * v_o;
                const float u = mask_select(dy, dx, lm.m16), w = mask_select(dx, dy, lm.m16);
                constexpr int H = NV / 2, R = NV - 2 * H;
                float nv[H + R];
                const float k0 = v_sigma * u, s0 = v_sigma * w;           // components 0, 1: Mx, My
                const float k1 = k0 * u, s1 = s0 * w;                     // components 2, 3: Mxx, Myy
                const float mxy = k0 * w;
                const float k2 = mask_select(v_o, mxy, lm.m16);           // components 4, 5: Mxy, v_opacity
                const float s2 = mask_select(mxy, v_o, lm.m16);
                nv[0] = k0 + __shfl_xor_sync(0xFFFFFFFFu, s0, 16);
                nv[1] = k1 + __shfl_xor_sync(0xFFFFFFFFu, s1, 16);
                nv[2] = k2 + __shfl_xor_sync(0xFFFFFFFFu, s2, 16);
#pragma unroll
                for (int k = 0; k < D / 2; ++k)                           // components 6 + 2k, 7 + 2k
                    nv[3 + k] = fac * vK[k] + __shfl_xor_sync(0xFFFFFFFFu, fac * vS[k], 16);
                if (R) {
                    const float last = fac * v_c[D - 1];
                    nv[H] = last + __shfl_xor_sync(0xFFFFFFFFu, last, 16);
                }
                const float r = HalvingMasked<H + R, 8>::run(nv, lm);
                // every lane holds the full sum of ITS component (lanes sharing a component hold identical
                // bits: the last butterfly levels are commutative adds), so all 32 lanes store: no predicate
                *reinterpret_cast<float*>(acc_lane + (off >> 1)) = r;
            }
            __syncwarp();
            const unsigned wm = __ballot_sync(0xFFFFFFFFu, keep && (S.keep[g.warp][pos] & 0x8000u) == 0u);
            if (g.lane == 0) S.wmask[g.warp][grp] = wm;
        }
        __syncthreads();
        // flush: thread t sums the 8 warps' slots of Gaussian t and adds them to global memory
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
                    for (int k = 0; k < NV; ++k) sum[k] += S.acc[w][tr * ACC_STRIDE + k];
                }
            }
            if (any) {
                // moments -> gradients: v_xy = (a Mx + b My, b Mx + c My) with (a, b, c) = -ln2 (2A, B, 2C);
                // v_conic = (Mxx / 2, Mxy, Myy / 2)
                const float4* q = slot_q(stage, tr);
                const float4 q0 = q[0], q1 = q[1];
                const float Mx = sum[0], My = sum[1];
                sum[0] = -LN2 * (2.f * q0.z * Mx + q0.w * My);
                sum[1] = -LN2 * (q0.w * Mx + 2.f * q1.x * My);
                const float Myy = sum[3];  // component order Mxx, Myy, Mxy
                sum[2] *= 0.5f;
                sum[3] = sum[4];
                sum[4] = 0.5f * Myy;
                float* row = vpack + (long long)S.ids[st][tr] * VP;
                red_add_v4(row, sum[0], sum[1], sum[2], sum[3]);
                red_add_v2(row + 4, sum[4], sum[5]);
                if (D == 4) red_add_v4(row + 8, sum[6], sum[7], sum[8], sum[9]);
                else if (D == 3) { red_add_v2(row + 8, sum[6], sum[7]); atomicAdd(row + 10, sum[8]); }
                else if (D == 2) red_add_v2(row + 8, sum[6], sum[7]);
                else atomicAdd(row + 8, sum[6]);
            }
        }
        __syncthreads();
    }
}

// =====================================================================================================
// plain kernels (any channel count up to 8)
// =====================================================================================================
template <int D>
__global__ void __launch_bounds__(BLK) blend3d_fwd_kernel(
    const float* __restrict__ means2d, const float* __restrict__ conics, const float* __restrict__ colors,
    const float* __restrict__ depths, const float* __restrict__ opacities, const float* __restrict__ backgrounds,
    int C, int CH, int W, int H, int tile_w, int tile_h, const int32_t* __restrict__ offsets,
    const int32_t* __restrict__ flatten_ids, int n_isects, float* __restrict__ render_colors,
    float* __restrict__ render_alphas, int32_t* __restrict__ last_ids) {
    __shared__ float4 s_xyo[BLK];  // x, y, opacity, -
    __shared__ float4 s_con[BLK];  // a, b, c, -
    __shared__ float s_col[BLK * D];

    const int cam = blockIdx.z;
    const int tile_id = blockIdx.y * tile_w + blockIdx.x;
    const int gtile = cam * tile_w * tile_h + tile_id;
    const int tr = threadIdx.x;
    const int pi = blockIdx.y * TS + (tr >> 4);
    const int pj = blockIdx.x * TS + (tr & 15);
    const float px = (float)pj + 0.5f, py = (float)pi + 0.5f;
    const bool inside = (pi < H && pj < W);
    bool done = !inside;

    const int range_start = offsets[gtile];
    const int range_end = (gtile == C * tile_w * tile_h - 1) ? n_isects : offsets[gtile + 1];
    const int num_batches = (range_end - range_start + BLK - 1) / BLK;

    float T = 1.0f;
    int cur_idx = 0;
    float pix[D];
#pragma unroll
    for (int k = 0; k < D; ++k) pix[k] = 0.f;

    for (int b = 0; b < num_batches; ++b) {
        if (__syncthreads_count(done) >= BLK) break;
        const int batch_start = range_start + BLK * b;
        const int idx = batch_start + tr;
        if (idx < range_end) {
            const int g = flatten_ids[idx];
            const float2 xy = reinterpret_cast<const float2*>(means2d)[g];
            s_xyo[tr] = make_float4(xy.x, xy.y, opacities[g], 0.f);
            s_con[tr] = make_float4(conics[(long long)g * 3 + 0], conics[(long long)g * 3 + 1], conics[(long long)g * 3 + 2], 0.f);
#pragma unroll
            for (int k = 0; k < D; ++k)
                s_col[tr * D + k] = (k < CH) ? colors[(long long)g * CH + k] : depths[g];
        }
        __syncthreads();
        const int batch_size = min(BLK, range_end - batch_start);
        for (int t = 0; t < batch_size && !done; ++t) {
            const float4 xyo = s_xyo[t];
            const float4 con = s_con[t];
            const float dx = xyo.x - px, dy = xyo.y - py;
            const float sigma = 0.5f * (con.x * dx * dx + con.z * dy * dy) + con.y * dx * dy;
            const float alpha = fminf(HGS_ALPHA_MAX, xyo.z * __expf(-sigma));
            if (sigma < 0.f || alpha < HGS_ALPHA_MIN) continue;
            const float next_T = T * (1.0f - alpha);
            if (next_T <= HGS_T_EPS) {
                done = true;
                break;
            }
            const float vis = alpha * T;
#pragma unroll
            for (int k = 0; k < D; ++k) pix[k] += s_col[t * D + k] * vis;
            cur_idx = batch_start + t;
            T = next_T;
        }
    }
    if (inside) {
        const long long pid = ((long long)cam * H + pi) * W + pj;
        render_alphas[pid] = 1.0f - T;
#pragma unroll
        for (int k = 0; k < D; ++k)
            render_colors[pid * D + k] = backgrounds == nullptr ? pix[k] : pix[k] + T * backgrounds[cam * D + k];
        last_ids[pid] = cur_idx;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

template <int D>
__global__ void __launch_bounds__(BLK) blend3d_bwd_kernel(
    const float* __restrict__ means2d, const float* __restrict__ conics, const float* __restrict__ colors,
    const float* __restrict__ depths, const float* __restrict__ opacities, const float* __restrict__ backgrounds,
    int C, int CH, int W, int H, int tile_w, int tile_h, const int32_t* __restrict__ offsets,
    const int32_t* __restrict__ flatten_ids, int n_isects, const float* __restrict__ render_alphas,
    const int32_t* __restrict__ last_ids, const float* __restrict__ v_render_colors,
    const float* __restrict__ v_render_alphas, float* __restrict__ v_means2d, float* __restrict__ v_means2d_abs,
    float* __restrict__ v_conics, float* __restrict__ v_colors, float* __restrict__ v_depths,
    float* __restrict__ v_opacities) {
    __shared__ int s_id[BLK];
    __shared__ float4 s_xyo[BLK];
    __shared__ float4 s_con[BLK];
    __shared__ float s_col[BLK * D];

    const int cam = blockIdx.z;
    const int tile_id = blockIdx.y * tile_w + blockIdx.x;
    const int gtile = cam * tile_w * tile_h + tile_id;
    const int tr = threadIdx.x;
    const int lane = tr & 31;
    const int pi = blockIdx.y * TS + (tr >> 4);
    const int pj = blockIdx.x * TS + (tr & 15);
    const float px = (float)pj + 0.5f, py = (float)pi + 0.5f;
    const bool inside = (pi < H && pj < W);
    const long long pid = ((long long)cam * H + min(pi, H - 1)) * W + min(pj, W - 1);

    const int range_start = offsets[gtile];
    const int range_end = (gtile == C * tile_w * tile_h - 1) ? n_isects : offsets[gtile + 1];
    const int num_batches = (range_end - range_start + BLK - 1) / BLK;
    if (num_batches <= 0) return;

    const float T_final = 1.0f - render_alphas[pid];
    float T = T_final;
    float buffer[D];
    float v_c[D];
    float bg_dot = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        buffer[k] = 0.f;
        v_c[k] = inside ? v_render_colors[pid * D + k] : 0.f;
        if (backgrounds != nullptr) bg_dot += backgrounds[cam * D + k] * v_c[k];
    }
    const float v_a = inside ? v_render_alphas[pid] : 0.f;
    const int bin_final = inside ? last_ids[pid] : 0;
    int warp_bin_final = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xFFFFFFFFu, warp_bin_final, o));

    for (int b = 0; b < num_batches; ++b) {
        __syncthreads();
        const int batch_end = range_end - 1 - BLK * b;
        const int batch_size = min(BLK, batch_end + 1 - range_start);
        const int idx = batch_end - tr;
        if (idx >= range_start) {
            const int g = flatten_ids[idx];
            s_id[tr] = g;
            const float2 xy = reinterpret_cast<const float2*>(means2d)[g];
            s_xyo[tr] = make_float4(xy.x, xy.y, opacities[g], 0.f);
            s_con[tr] = make_float4(conics[(long long)g * 3 + 0], conics[(long long)g * 3 + 1], conics[(long long)g * 3 + 2], 0.f);
#pragma unroll
            for (int k = 0; k < D; ++k)
                s_col[tr * D + k] = (k < CH) ? colors[(long long)g * CH + k] : depths[g];
        }
        __syncthreads();
        for (int t = max(0, batch_end - warp_bin_final); t < batch_size; ++t) {
            bool valid = inside && (batch_end - t <= bin_final);
            float alpha = 0.f, opac = 0.f, vis = 0.f, dx = 0.f, dy = 0.f;
            float4 con = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const float4 xyo = s_xyo[t];
                con = s_con[t];
                opac = xyo.z;
                dx = xyo.x - px;
                dy = xyo.y - py;
                const float sigma = 0.5f * (con.x * dx * dx + con.z * dy * dy) + con.y * dx * dy;
                vis = __expf(-sigma);
                alpha = fminf(HGS_ALPHA_MAX, opac * vis);
                if (sigma < 0.f || alpha < HGS_ALPHA_MIN) valid = false;
            }
            if (!__any_sync(0xFFFFFFFFu, valid)) continue;
            float v_col[D];
#pragma unroll
            for (int k = 0; k < D; ++k) v_col[k] = 0.f;
            float v_con0 = 0.f, v_con1 = 0.f, v_con2 = 0.f, v_x = 0.f, v_y = 0.f, v_ax = 0.f, v_ay = 0.f, v_o = 0.f;
            if (valid) {
                const float ra = 1.0f / (1.0f - alpha);
                T *= ra;
                const float fac = alpha * T;
                float v_alpha = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    v_col[k] = fac * v_c[k];
                    v_alpha += (s_col[t * D + k] * T - buffer[k] * ra) * v_c[k];
                }
                v_alpha += T_final * ra * v_a;
                if (backgrounds != nullptr) v_alpha += -T_final * ra * bg_dot;
                if (opac * vis <= HGS_ALPHA_MAX) {
                    const float v_sigma = -opac * vis * v_alpha;
                    v_con0 = 0.5f * v_sigma * dx * dx;
                    v_con1 = v_sigma * dx * dy;
                    v_con2 = 0.5f * v_sigma * dy * dy;
                    v_x = v_sigma * (con.x * dx + con.y * dy);
                    v_y = v_sigma * (con.y * dx + con.z * dy);
                    v_ax = fabsf(v_x);
                    v_ay = fabsf(v_y);
                    v_o = vis * v_alpha;
                }
#pragma unroll
                for (int k = 0; k < D; ++k) buffer[k] += s_col[t * D + k] * fac;
            }
#pragma unroll
            for (int k = 0; k < D; ++k) v_col[k] = warp_sum(v_col[k]);
            v_con0 = warp_sum(v_con0); v_con1 = warp_sum(v_con1); v_con2 = warp_sum(v_con2);
            v_x = warp_sum(v_x); v_y = warp_sum(v_y); v_o = warp_sum(v_o);
            if (v_means2d_abs != nullptr) { v_ax = warp_sum(v_ax); v_ay = warp_sum(v_ay); }
            if (lane == 0) {
                const int g = s_id[t];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    if (k < CH) atomicAdd(v_colors + (long long)g * CH + k, v_col[k]);
                    else atomicAdd(v_depths + g, v_col[k]);
                }
                atomicAdd(v_conics + (long long)g * 3 + 0, v_con0);
                atomicAdd(v_conics + (long long)g * 3 + 1, v_con1);
                atomicAdd(v_conics + (long long)g * 3 + 2, v_con2);
                atomicAdd(v_means2d + (long long)g * 2 + 0, v_x);
                atomicAdd(v_means2d + (long long)g * 2 + 1, v_y);
                if (v_means2d_abs != nullptr) {
                    atomicAdd(v_means2d_abs + (long long)g * 2 + 0, v_ax);
                    atomicAdd(v_means2d_abs + (long long)g * 2 + 1, v_ay);
                }
                atomicAdd(v_opacities + g, v_o);
            }
        }
    }
}

template <int D>
int launch_fwd(const float* means2d, const float* conics, const float* colors, const float* depths,
               const float* opacities, const float* backgrounds, int C, int CH, int W, int H, int tile_w, int tile_h,
               const int32_t* offsets, const int32_t* flatten_ids, int n_isects, float* render_colors,
               float* render_alphas, int32_t* last_ids, cudaStream_t st) {
    dim3 grid(tile_w, tile_h, C);
    blend3d_fwd_kernel<D><<<grid, BLK, 0, st>>>(means2d, conics, colors, depths, opacities, backgrounds, C, CH, W, H,
                                                 tile_w, tile_h, offsets, flatten_ids, n_isects, render_colors,
                                                 render_alphas, last_ids);
    HGS_LAUNCH_CHECK();
    return 0;
}

template <int D>
int launch_bwd(const float* means2d, const float* conics, const float* colors, const float* depths,
               const float* opacities, const float* backgrounds, int C, int CH, int W, int H, int tile_w, int tile_h,
               const int32_t* offsets, const int32_t* flatten_ids, int n_isects, const float* render_alphas,
               const int32_t* last_ids, const float* v_render_colors, const float* v_render_alphas, float* v_means2d,
               float* v_means2d_abs, float* v_conics, float* v_colors, float* v_depths, float* v_opacities,
               cudaStream_t st) {
    dim3 grid(tile_w, tile_h, C);
    blend3d_bwd_kernel<D><<<grid, BLK, 0, st>>>(means2d, conics, colors, depths, opacities, backgrounds, C, CH, W, H,
                                                 tile_w, tile_h, offsets, flatten_ids, n_isects, render_alphas,
                                                 last_ids, v_render_colors, v_render_alphas, v_means2d, v_means2d_abs,
                                                 v_conics, v_colors, v_depths, v_opacities);
    HGS_LAUNCH_CHECK();
    return 0;
}

template <int D, bool NORM>
int launch_fwd_fast(const GRec* recs, const float* backgrounds, int C, int W, int H, int tile_w, int tile_h,
                    const int32_t* offsets, const int32_t* flatten_ids, int n_isects, float* render_colors,
                    float* render_alphas, int32_t* last_ids, cudaStream_t st) {
    dim3 grid(tile_w, tile_h, C);
    blend3d_fwd_fast_kernel<D, NORM><<<grid, BLK, 0, st>>>(recs, backgrounds, C, W, H, tile_w, tile_h, offsets,
                                                           flatten_ids, n_isects, render_colors, render_alphas,
                                                           last_ids);
    HGS_LAUNCH_CHECK();
    return 0;
}

template <int D, bool NORM>
int launch_bwd_fast(const GRec* recs, const float* backgrounds, int C, int W, int H, int tile_w, int tile_h,
                    const int32_t* offsets, const int32_t* flatten_ids, int n_isects, const float* render_colors,
                    const float* render_alphas, const int32_t* last_ids, const float* v_render_colors,
                    const float* v_render_alphas, float* vpack, cudaStream_t st) {
    dim3 grid(tile_w, tile_h, C);
    const int smem = (int)sizeof(BwdSmem<D, FB_BWD>);
    auto* kern = blend3d_bwd_fast_kernel<D, NORM, FB_BWD>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, BLK, smem, st>>>(recs, backgrounds, C, W, H, tile_w, tile_h, offsets, flatten_ids, n_isects,
                                  render_colors, render_alphas, last_ids, v_render_colors, v_render_alphas, vpack);
    HGS_LAUNCH_CHECK();
    return 0;
}

// The view-space mean gradient and the opacity gradient are consumed as dense tensors by autograd (retain_grad() of
// meta["means2d"] clones the gradient; the opacity gradient is accumulated into a leaf): reading them as strided
// views of the 48-byte vpack rows touches every sector of vpack (288 MB at 6 M Gaussians).  This copies the two
// columns of the VISIBLE rows into dense, zero-filled tensors instead (zero fill by the launcher).
__global__ void unpack_vpack_kernel(const float* __restrict__ vpack, const int32_t* __restrict__ vis_ids, long long n_vis,
                                    float* __restrict__ v_means2d, float* __restrict__ v_opacities) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_vis) return;
    const long long i = vis_ids[t];
    const float4 q0 = reinterpret_cast<const float4*>(vpack + i * VP)[0];
    const float2 q1 = reinterpret_cast<const float2*>(vpack + i * VP + 4)[0];
    reinterpret_cast<float2*>(v_means2d)[i] = make_float2(q0.x, q0.y);
    v_opacities[i] = q1.y;
}

// rows[ids[j]] := 0 (16-byte stores; R4 float4 per row)
__global__ void zero_rows_kernel(float4* __restrict__ rows, int R4, const int32_t* __restrict__ ids, long long n_ids) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ids * R4) return;
    const long long j = t / R4;
    rows[(long long)ids[j] * R4 + (t - j * R4)] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace

#define HGS_DISPATCH_D(D, CALL)                     \
    switch (D) {                                    \
        case 1: return CALL(1);                     \
        case 2: return CALL(2);                     \
        case 3: return CALL(3);                     \
        case 4: return CALL(4);                     \
        case 5: return CALL(5);                     \
        case 6: return CALL(6);                     \
        case 7: return CALL(7);                     \
        case 8: return CALL(8);                     \
        default: return HGS_ERR_INVALID_ARG;        \
    }

HGS_API int hgs_blend3d_fwd(const float* means2d, const float* conics, const float* colors, const float* depths,
                            const float* opacities, const float* backgrounds, int C, int N, int CH, int width,
                            int height, int tile_size, const int32_t* isect_offsets, const int32_t* flatten_ids,
                            long long n_isects, float* render_colors, float* render_alphas, int32_t* last_ids,
                            void* stream) {
    (void)N;
    if (tile_size != TS || C <= 0 || width <= 0 || height <= 0 || CH < 0 || n_isects < 0) return HGS_ERR_INVALID_ARG;
    if (n_isects >= (1ll << 31)) return HGS_ERR_TOO_LARGE;
    const int D = CH + (depths != nullptr ? 1 : 0);
    const int tile_w = (width + TS - 1) / TS, tile_h = (height + TS - 1) / TS;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DD)                                                                                                    \
    launch_fwd<DD>(means2d, conics, colors, depths, opacities, backgrounds, C, CH, width, height, tile_w, tile_h,    \
                   isect_offsets, flatten_ids, (int)n_isects, render_colors, render_alphas, last_ids, st)
    HGS_DISPATCH_D(D, CALL)
#undef CALL
}

HGS_API int hgs_blend3d_bwd(const float* means2d, const float* conics, const float* colors, const float* depths,
                            const float* opacities, const float* backgrounds, int C, int N, int CH, int width,
                            int height, int tile_size, const int32_t* isect_offsets, const int32_t* flatten_ids,
                            long long n_isects, const float* render_alphas, const int32_t* last_ids,
                            const float* v_render_colors, const float* v_render_alphas, float* v_means2d,
                            float* v_means2d_abs, float* v_conics, float* v_colors, float* v_depths,
                            float* v_opacities, void* stream) {
    (void)N;
    if (tile_size != TS || C <= 0 || width <= 0 || height <= 0 || CH < 0 || n_isects < 0) return HGS_ERR_INVALID_ARG;
    if (n_isects >= (1ll << 31)) return HGS_ERR_TOO_LARGE;
    if (n_isects == 0) return 0;
    const int D = CH + (depths != nullptr ? 1 : 0);
    const int tile_w = (width + TS - 1) / TS, tile_h = (height + TS - 1) / TS;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DD)                                                                                                    \
    launch_bwd<DD>(means2d, conics, colors, depths, opacities, backgrounds, C, CH, width, height, tile_w, tile_h,    \
                   isect_offsets, flatten_ids, (int)n_isects, render_alphas, last_ids, v_render_colors,              \
                   v_render_alphas, v_means2d, v_means2d_abs, v_conics, v_colors, v_depths, v_opacities, st)
    HGS_DISPATCH_D(D, CALL)
#undef CALL
}

// ---- fast path --------------------------------------------------------------------------------------
HGS_API size_t hgs_blend3d_pack_bytes(long long CN) { return (size_t)(CN > 0 ? CN : 1) * REC_BYTES; }

HGS_API int hgs_blend3d_pack(const float* means2d, const float* conics, const float* colors, const float* depths,
                             const float* opacities, const int32_t* radii, const int32_t* vis_ids, long long n_vis,
                             const long long* n_vis_dev, long long CN, int CH, void* records, void* stream) {
    if (CN < 0 || n_vis < 0 || CH < 0 || CH + (depths != nullptr ? 1 : 0) > 4) return HGS_ERR_INVALID_ARG;
    const long long work = vis_ids != nullptr ? n_vis : CN;
    if (work == 0) return 0;
    pack3d_kernel<<<hgs_ceil_div(work, 256), 256, 0, (cudaStream_t)stream>>>(
        means2d, conics, colors, depths, opacities, radii, vis_ids, work, vis_ids != nullptr ? n_vis_dev : nullptr, CH,
        (GRec*)records);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_blend3d_stats(const void* records, int C, int width, int height, int tile_size,
                              const int32_t* isect_offsets, const int32_t* flatten_ids, long long n_isects,
                              unsigned long long* counters, void* stream) {
    if (tile_size != TS || C <= 0 || width <= 0 || height <= 0 || n_isects < 0 || n_isects >= (1ll << 31))
        return HGS_ERR_INVALID_ARG;
    const int tile_w = (width + TS - 1) / TS, tile_h = (height + TS - 1) / TS;
    dim3 grid(tile_w, tile_h, C);
    blend3d_stats_kernel<<<grid, BLK, 0, (cudaStream_t)stream>>>((const GRec*)records, C, width, height, tile_w, tile_h,
                                                                 isect_offsets, flatten_ids, (int)n_isects, counters);
    HGS_LAUNCH_CHECK();
    return 0;
}

#define HGS_DISPATCH_FAST(D, NORM, CALL)                         \
    switch ((D) * 2 + ((NORM) ? 1 : 0)) {                        \
        case 2: return CALL(1, false);                           \
        case 3: return CALL(1, true);                            \
        case 4: return CALL(2, false);                           \
        case 5: return CALL(2, true);                            \
        case 6: return CALL(3, false);                           \
        case 7: return CALL(3, true);                            \
        case 8: return CALL(4, false);                           \
        case 9: return CALL(4, true);                            \
        default: return HGS_ERR_INVALID_ARG;                     \
    }

HGS_API int hgs_blend3d_fwd_packed(const void* records, const float* backgrounds, int C, int D, int normalize_depth,
                                   int width, int height, int tile_size, const int32_t* isect_offsets,
                                   const int32_t* flatten_ids, long long n_isects, float* render_colors,
                                   float* render_alphas, int32_t* last_ids, void* stream) {
    if (tile_size != TS || C <= 0 || width <= 0 || height <= 0 || D < 1 || D > 4 || n_isects < 0) return HGS_ERR_INVALID_ARG;
    if (n_isects >= (1ll << 31)) return HGS_ERR_TOO_LARGE;
    const int tile_w = (width + TS - 1) / TS, tile_h = (height + TS - 1) / TS;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DD, NN)                                                                                               \
    launch_fwd_fast<DD, NN>((const GRec*)records, backgrounds, C, width, height, tile_w, tile_h, isect_offsets,     \
                            flatten_ids, (int)n_isects, render_colors, render_alphas, last_ids, st)
    HGS_DISPATCH_FAST(D, normalize_depth != 0, CALL)
#undef CALL
}

HGS_API int hgs_blend3d_bwd_packed(const void* records, const float* backgrounds, int C, int D, int normalize_depth,
                                   int width, int height, int tile_size, const int32_t* isect_offsets,
                                   const int32_t* flatten_ids, long long n_isects, const float* render_colors,
                                   const float* render_alphas, const int32_t* last_ids, const float* v_render_colors,
                                   const float* v_render_alphas, float* vpack, void* stream) {
    if (tile_size != TS || C <= 0 || width <= 0 || height <= 0 || D < 1 || D > 4 || n_isects < 0) return HGS_ERR_INVALID_ARG;
    if (n_isects >= (1ll << 31)) return HGS_ERR_TOO_LARGE;
    if (n_isects == 0) return 0;
    const int tile_w = (width + TS - 1) / TS, tile_h = (height + TS - 1) / TS;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DD, NN)                                                                                               \
    launch_bwd_fast<DD, NN>((const GRec*)records, backgrounds, C, width, height, tile_w, tile_h, isect_offsets,     \
                            flatten_ids, (int)n_isects, render_colors, render_alphas, last_ids, v_render_colors,    \
                            v_render_alphas, vpack, st)
    HGS_DISPATCH_FAST(D, normalize_depth != 0, CALL)
#undef CALL
}

HGS_API int hgs_blend3d_unpack(const float* vpack, const int32_t* vis_ids, long long n_vis, long long CN, float* v_means2d,
                               float* v_opacities, int outputs_zeroed, void* stream) {
    if (vpack == nullptr || v_means2d == nullptr || v_opacities == nullptr || n_vis < 0 || CN < 0 ||
        (n_vis > 0 && vis_ids == nullptr))
        return HGS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (!outputs_zeroed) {
        if ((e = cudaMemsetAsync(v_means2d, 0, (size_t)CN * 2 * sizeof(float), st)) != cudaSuccess) return (int)e;
        if ((e = cudaMemsetAsync(v_opacities, 0, (size_t)CN * sizeof(float), st)) != cudaSuccess) return (int)e;
    }
    if (n_vis == 0) return 0;
    unpack_vpack_kernel<<<hgs_ceil_div(n_vis, 256), 256, 0, st>>>(vpack, vis_ids, n_vis, v_means2d, v_opacities);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_zero_rows(float* rows, int row_floats, const int32_t* ids, long long n_ids, void* stream) {
    if (rows == nullptr || row_floats <= 0 || row_floats % 4 != 0 || n_ids < 0 || (reinterpret_cast<size_t>(rows) & 15))
        return HGS_ERR_INVALID_ARG;
    if (n_ids == 0) return 0;
    if (ids == nullptr) return HGS_ERR_INVALID_ARG;
    const int R4 = row_floats / 4;
    zero_rows_kernel<<<hgs_ceil_div(n_ids * R4, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(rows), R4,
                                                                                      ids, n_ids);
    HGS_LAUNCH_CHECK();
    return 0;
}
