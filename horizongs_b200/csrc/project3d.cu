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
#include "bin_common.cuh"
#include "sh_math.cuh"
#include "blend_common.cuh"

namespace {

constexpr int PB = 256;  // threads per block
constexpr int CH = hgs_bin::CP_TILE;   // Gaussians per CTA of the forward kernel (4 per thread)
constexpr int PER = CH / PB;

// Forward kernel.  One CTA owns CH = 1024 consecutive Gaussians of one camera, so that the heavy math runs on
// densely populated warps even when only a small, randomly scattered fraction of the Gaussians is on screen:
//   1. every thread, 4 consecutive rows (16-byte loads / stores): camera-space centre, near/far test, conservative
//      off-screen test; zeros are written to every output row; the survivors' row numbers are compacted, in order,
//      into shared memory;
//   2. the survivors, one per thread: quaternion load, covariance / Jacobian / conic / radius math, exact screen
//      test, tile count; a visible Gaussian overwrites its output rows;
//   3. (BIN) ordering, fused (what hgs_isect_bin_prepare's first kernel does from tiles_per_gauss): the visible
//      Gaussians with tiles are compacted in ascending flat-index order across the CTAs (decoupled look-back;
//      CTAs take their chunk by ticket so that every predecessor is running), their records {flat index, depth
//      bits, tile box} are written in that order, and the super-tile histogram is updated.
struct BinArgs {
    hgs_bin::BinGeom G;
    unsigned long long* flags;
    uint32_t *ticket, *super_count;
    int32_t* visible_ids;
    hgs_bin::VisRec* vrec;
    long long* counts_dev;
};

//   (SHADE) shading, fused (what hgs_sh_fwd + hgs_blend3d_pack do from the arrays this kernel writes): in phase 2 a
//      visible Gaussian also evaluates its view-dependent colour (SHADE = SH degree 0..4; -1: colours given) and
//      writes its 64-byte blend record -- centre, conic, depth are still in registers, so the five 64-byte-granule
//      gathers of the pack kernel and the second read of the centre disappear.
struct ShadeArgs {
    const float* feats;       // SH coefficients [N,K,3] (SHADE >= 0) or colours [N,3] (SHADE == -1)
    const float* opacities;   // [N]
    const float* campos;      // [C,3]
    int K, depth_channel;
    float* colors;            // [C,N,3] out (SHADE >= 0): the clamped SH colour, visible rows only
    hgs::GRec* recs;          // [C*N] out
};
constexpr int SHADE_NONE = -2, SHADE_RGB = -1;

template <bool BIN, int SHADE>
__global__ void __launch_bounds__(PB, 6) project3d_fwd_kernel(
    const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
    const float* __restrict__ viewmats, const float* __restrict__ Ks, int N, int nblk_cam, int W, int H, float eps2d,
    float near_plane, float far_plane, float radius_clip, int tile_size, int tile_w, int tile_h,
    int32_t* __restrict__ radii, float* __restrict__ means2d, float* __restrict__ depths, float* __restrict__ conics,
    float* __restrict__ compensations, int32_t* __restrict__ tiles_per_gauss, BinArgs B, ShadeArgs S) {
    __shared__ unsigned short s_list[CH];      // local rows of the phase-1 survivors, ascending
    __shared__ int s_seg[PER * (PB / 32) + 1];
    __shared__ float s_cam[26];                // viewmat (16), K (9), bound coefficient
    __shared__ uint4 s_rec[BIN ? CH : 1];      // BIN: {flat index (~0: none), depth bits, box lo, box hi} per survivor
    __shared__ uint32_t s_bid, s_nbig;
    __shared__ unsigned long long s_excl, s_isect;
    __shared__ unsigned short s_big[BIN ? CH : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (BIN) {
        if (threadIdx.x == 0) {
            s_bid = atomicAdd(B.ticket, 1u);   // chunks are taken in start order (look-back needs the predecessors running)
            s_nbig = 0;
            s_isect = 0;
        }
        __syncthreads();
    }
    const uint32_t bid = BIN ? s_bid : (uint32_t)(blockIdx.y * nblk_cam + blockIdx.x);
    const int c = (int)(bid / (uint32_t)nblk_cam);
    const long long base = (long long)(bid - (uint32_t)c * nblk_cam) * CH;
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

    // ---- phase 1: rows base + 4 * threadIdx.x + k (16-byte loads and stores when the rows are aligned)
    bool pass[PER];
    {
        const long long n0 = base + (long long)threadIdx.x * PER;
        const long long idx0 = (long long)c * N + n0;
        const bool full = n0 + PER <= N;
        const bool vec = full && (idx0 & 3) == 0;           // n0 is a multiple of 4; idx0 too unless C > 1 and N % 4 != 0
        float m[PER * 3], sc[PER * 3];
        if (full) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float4 a = reinterpret_cast<const float4*>(means + n0 * 3)[j];
                const float4 q = reinterpret_cast<const float4*>(scales + n0 * 3)[j];
                m[4 * j] = a.x; m[4 * j + 1] = a.y; m[4 * j + 2] = a.z; m[4 * j + 3] = a.w;
                sc[4 * j] = q.x; sc[4 * j + 1] = q.y; sc[4 * j + 2] = q.z; sc[4 * j + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < PER * 3; ++e) {
                const bool in = n0 * 3 + e < (long long)N * 3;
                m[e] = in ? means[n0 * 3 + e] : 0.f;
                sc[e] = in ? scales[n0 * 3 + e] : 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            pass[k] = false;
            if (n0 + k < N) {
                const float px = m[3 * k], py = m[3 * k + 1], pz = m[3 * k + 2];
                // zc decides the near / far cull: exactly the oracle's expression
                const float zc = R[2][0] * px + R[2][1] * py + R[2][2] * pz + cam.t[2];
                if (!(zc < near_plane || zc > far_plane)) {
                    // the off-screen test is conservative (0.1 % + 1 px margin): fused arithmetic is fine here
                    const float xc = __fmaf_rn(R[0][0], px, __fmaf_rn(R[0][1], py, __fmaf_rn(R[0][2], pz, cam.t[0])));
                    const float yc = __fmaf_rn(R[1][0], px, __fmaf_rn(R[1][1], py, __fmaf_rn(R[1][2], pz, cam.t[1])));
                    pass[k] = !proj3d_surely_offscreen(cam, jf_coeff, xc, yc, zc,
                                                       fmaxf(fabsf(sc[3 * k]), fmaxf(fabsf(sc[3 * k + 1]), fabsf(sc[3 * k + 2]))),
                                                       (float)W, (float)H, eps2d);
                }
            }
        }
        // defaults: culled
        if (vec) {
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<int4*>(radii + idx0) = make_int4(0, 0, 0, 0);
            reinterpret_cast<float4*>(means2d + idx0 * 2)[0] = z4;
            reinterpret_cast<float4*>(means2d + idx0 * 2)[1] = z4;
            *reinterpret_cast<float4*>(depths + idx0) = z4;
#pragma unroll
            for (int j = 0; j < 3; ++j) reinterpret_cast<float4*>(conics + idx0 * 3)[j] = z4;
            if (compensations != nullptr) *reinterpret_cast<float4*>(compensations + idx0) = z4;
            if (tiles_per_gauss != nullptr) *reinterpret_cast<int4*>(tiles_per_gauss + idx0) = make_int4(0, 0, 0, 0);
        } else {
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                if (n0 + k >= N) break;
                const long long idx = idx0 + k;
                radii[idx] = 0;
                means2d[idx * 2] = 0.f; means2d[idx * 2 + 1] = 0.f;
                depths[idx] = 0.f;
                conics[idx * 3] = 0.f; conics[idx * 3 + 1] = 0.f; conics[idx * 3 + 2] = 0.f;
                if (compensations != nullptr) compensations[idx] = 0.f;
                if (tiles_per_gauss != nullptr) tiles_per_gauss[idx] = 0;
            }
        }
    }
    {
        // ordered compaction of the survivors' local rows: exclusive scan of the per-thread counts over the CTA
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) cnt += pass[k] ? 1 : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_seg[warp] = incl;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < PB / 32; ++w) {
            const int x = s_seg[w];
            if (w < warp) wbase += x;
            total += x;
        }
        int pos = wbase + incl - cnt;
#pragma unroll
        for (int k = 0; k < PER; ++k)
            if (pass[k]) s_list[pos++] = (unsigned short)(threadIdx.x * PER + k);
        __syncthreads();
        if (threadIdx.x == 0) s_seg[32] = total;
        __syncthreads();
    }
    const int n_pass = s_seg[32];

    // ---- phase 2: dense math on the survivors
    uint32_t n_vis_t = 0;       // BIN: visible Gaussians with tiles found by this thread
    for (int i = threadIdx.x; i < n_pass; i += PB) {
        const int r = s_list[i];
        const long long n = base + r;
        const long long idx = (long long)c * N + n;
        const float4 qv = reinterpret_cast<const float4*>(quats)[n];
        const float px = means[n * 3], py = means[n * 3 + 1], pz = means[n * 3 + 2];
        Proj3dFwd f;
        f.zc = R[2][0] * px + R[2][1] * py + R[2][2] * pz + cam.t[2];
        f.xc = R[0][0] * px + R[0][1] * py + R[0][2] * pz + cam.t[0];
        f.yc = R[1][0] * px + R[1][1] * py + R[1][2] * pz + cam.t[1];
        proj3d_cov_and_project(cam, qv.x, qv.y, qv.z, qv.w, scales[n * 3], scales[n * 3 + 1], scales[n * 3 + 2], (float)W,
                               (float)H, eps2d, f);
        uint4 rec = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
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
                int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
                if (tiles_per_gauss != nullptr && radius_i > 0) {
                    hgs_tile_bbox(f.m2x, f.m2y, (float)radius_i, (float)tile_size, tile_w, tile_h, x0, y0, x1, y1);
                    ntiles = (y1 - y0) * (x1 - x0);
                }
                radii[idx] = radius_i;
                reinterpret_cast<float2*>(means2d)[idx] = make_float2(f.m2x, f.m2y);
                depths[idx] = f.zc;
                conics[idx * 3] = f.c11 * inv_det;
                conics[idx * 3 + 1] = -f.c01 * inv_det;
                conics[idx * 3 + 2] = f.c00 * inv_det;
                if (compensations != nullptr) compensations[idx] = sqrtf(fmaxf(f.det_orig / f.det, 0.f));
                if (tiles_per_gauss != nullptr) tiles_per_gauss[idx] = ntiles;
                if (SHADE != SHADE_NONE) {
                    float col[4];
                    if (SHADE == SHADE_RGB) {
                        col[0] = S.feats[n * 3]; col[1] = S.feats[n * 3 + 1]; col[2] = S.feats[n * 3 + 2];
                    } else {
                        constexpr int DEG = SHADE < 0 ? 0 : SHADE;
                        sh_eval_one<DEG>(px - S.campos[c * 3], py - S.campos[c * 3 + 1], pz - S.campos[c * 3 + 2],
                                         S.feats + n * (long long)(S.K * 3), 1, col[0], col[1], col[2]);
                        S.colors[idx * 3] = col[0]; S.colors[idx * 3 + 1] = col[1]; S.colors[idx * 3 + 2] = col[2];
                    }
                    col[3] = S.depth_channel ? f.zc : 0.f;
                    const hgs::GRec gr = hgs::make_grec(f.m2x, f.m2y, f.c11 * inv_det, -f.c01 * inv_det, f.c00 * inv_det,
                                              S.opacities[n], col);
                    float4* dst = reinterpret_cast<float4*>(S.recs + idx);
                    const float4* src = reinterpret_cast<const float4*>(&gr);
#pragma unroll
                    for (int k = 0; k < 4; ++k) dst[k] = src[k];
                }
                if (BIN && ntiles > 0) {
                    rec = make_uint4((uint32_t)idx, __float_as_uint(f.zc), (uint32_t)x0 | ((uint32_t)y0 << 16),
                                     (uint32_t)x1 | ((uint32_t)y1 << 16));
                    ++n_vis_t;
                }
            }
        }
        if (BIN) s_rec[i] = rec;
    }
    if (!BIN) return;

    // ---- phase 3: ordered compaction of the visible Gaussians, records, super-tile histogram
    __syncthreads();
    // position of survivor slot i among the CTA's visible ones: rounds of PB slots, (round, warp) segments in order
    const int n_rounds = (n_pass + PB - 1) / PB;        // <= PER
    for (int k = 0; k < PER; ++k) {
        const int i = k * PB + threadIdx.x;
        const bool v = k < n_rounds && i < n_pass && s_rec[i].x != 0xFFFFFFFFu;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, v);
        if (lane == 0) s_seg[k * (PB / 32) + warp] = __popc(bal);
    }
    __syncthreads();
    if (warp == 0) {
        const int v = s_seg[lane];
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        s_seg[lane] = incl - v;
        if (lane == 31) {
            s_seg[32] = incl;
            volatile unsigned long long* vf = B.flags;      // publish the block total at once
            vf[bid] = (bid == 0 ? LB_PREFIX : LB_AGG) | (unsigned long long)incl;
        }
    }
    __syncthreads();
    const uint32_t tot = (uint32_t)s_seg[32];
    // histogram first (it does not need the prefix of the earlier blocks), tile total
    unsigned long long isects = 0;
    for (int k = 0; k < n_rounds; ++k) {
        const int i = k * PB + threadIdx.x;
        if (i >= n_pass) break;
        const uint4 rec = s_rec[i];
        if (rec.x == 0xFFFFFFFFu) continue;
        const int x0 = (int)(rec.z & 0xFFFFu), y0 = (int)(rec.z >> 16), x1 = (int)(rec.w & 0xFFFFu), y1 = (int)(rec.w >> 16);
        isects += (unsigned long long)((x1 - x0) * (y1 - y0));
        if (hgs_bin::super_area(x0, y0, x1, y1) > hgs_bin::BIG_AREA) s_big[atomicAdd(&s_nbig, 1u)] = (unsigned short)i;
        else hgs_bin::super_hist_add(B.G, B.super_count, rec.x, x0, y0, x1, y1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) isects += __shfl_xor_sync(0xFFFFFFFFu, isects, o);
    if (lane == 0 && isects) atomicAdd(&s_isect, isects);
    __syncthreads();
    const uint32_t nbig = s_nbig;
    for (uint32_t b = 0; b < nbig; ++b) {
        const uint4 rec = s_rec[s_big[b]];
        hgs_bin::super_hist_add_cta(B.G, B.super_count, rec.x, (int)(rec.z & 0xFFFFu), (int)(rec.z >> 16),
                                    (int)(rec.w & 0xFFFFu), (int)(rec.w >> 16));
    }
    if (warp == 0) {
        const unsigned long long excl = hgs_bin::lookback_exclusive(B.flags, bid, tot);
        if (lane == 0) {
            s_excl = excl;
            if (bid == gridDim.x - 1) B.counts_dev[0] = (long long)(excl + tot);
            if (s_isect) atomicAdd(reinterpret_cast<unsigned long long*>(B.counts_dev + 1), s_isect);
        }
    }
    __syncthreads();
    const long long out0 = (long long)s_excl;
    for (int k = 0; k < n_rounds; ++k) {
        const int i = k * PB + threadIdx.x;
        const uint4 rec = i < n_pass ? s_rec[i] : make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
        const bool v = rec.x != 0xFFFFFFFFu;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, v);
        if (v) {
            const long long p = out0 + s_seg[k * (PB / 32) + warp] + __popc(bal & ((1u << lane) - 1u));
            B.visible_ids[p] = (int32_t)rec.x;
            hgs_bin::VisRec r;
            r.g = rec.x; r.depth_bits = rec.y; r.xy0 = rec.z; r.xy1 = rec.w;
            B.vrec[p] = r;
        }
    }
}
static_assert(PER * (PB / 32) == 32, "one warp scans the (round, warp) segment counts");

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
    const int nblk_cam = hgs_ceil_div(N, CH);
    dim3 grid(nblk_cam, C);
    project3d_fwd_kernel<false, SHADE_NONE><<<grid, PB, 0, (cudaStream_t)stream>>>(
        means, quats, scales, viewmats, Ks, N, nblk_cam, width, height, eps2d, near_plane, far_plane, radius_clip, tile_size,
        tile_w, tile_h, radii, means2d, depths, conics, compensations, tiles_per_gauss, BinArgs{}, ShadeArgs{});
    HGS_LAUNCH_CHECK();
    return 0;
}

// projection + first kernel of the ordering stage (hgs_isect_bin_prepare's compaction and histogram) in one launch;
// temp as for hgs_isect_bin_prepare; the caller follows with hgs_isect_bin_scan.
// shade: -2 none; -1 colours [N,3] given in feats; 0..4 SH degree with coefficients feats [N,K,3], view direction
// means - campos[c], colour = max(SH + 0.5, 0) written to colors_out [C,N,3] (visible rows only).  With shade != -2 the
// 64-byte blend records of the visible Gaussians (hgs_blend3d_pack's, bit for bit) are written to records.
HGS_API int hgs_project3d_fwd_bin(const float* means, const float* quats, const float* scales, const float* viewmats,
                                  const float* Ks, int C, int N, int width, int height, float eps2d, float near_plane,
                                  float far_plane, float radius_clip, int tile_size, int32_t* radii, float* means2d,
                                  float* depths, float* conics, float* compensations, int32_t* tiles_per_gauss,
                                  int32_t* visible_ids, long long* counts_dev, void* temp, size_t temp_bytes,
                                  int shade, const float* feats, int K, const float* opacities, const float* campos,
                                  int depth_channel, float* colors_out, void* records, void* stream) {
    if (C <= 0 || N < 0 || width <= 0 || height <= 0 || tile_size <= 0 || tiles_per_gauss == nullptr ||
        visible_ids == nullptr || counts_dev == nullptr || shade < SHADE_NONE || shade > 4)
        return HGS_ERR_INVALID_ARG;
    if (shade != SHADE_NONE && (feats == nullptr || opacities == nullptr || records == nullptr ||
                                (reinterpret_cast<size_t>(records) & 15) ||
                                (shade >= 0 && (campos == nullptr || colors_out == nullptr || K < (shade + 1) * (shade + 1)))))
        return HGS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int tile_w = (width + tile_size - 1) / tile_size, tile_h = (height + tile_size - 1) / tile_size;
    const long long CN = (long long)C * N, tt = (long long)C * tile_w * tile_h;
    if (CN >= (1ll << hgs_bin::ID_BITS) || tt >= (1ll << 29) || tile_w >= 65536 || tile_h >= 65536) return HGS_ERR_TOO_LARGE;
    cudaError_t e = cudaMemsetAsync(counts_dev, 0, 3 * sizeof(long long), st);
    if (e != cudaSuccess) return (int)e;
    if (N == 0) return 0;
    BinArgs B;
    B.G = hgs_bin::make_geom(N, tile_size, tile_w, tile_h);
    const hgs_bin::BinTemp T = hgs_bin::bin_temp(temp, CN, (long long)C * B.G.stw * B.G.sth, tt);
    if (temp_bytes < T.bytes) return HGS_ERR_WORKSPACE;
    if ((e = cudaMemsetAsync(temp, 0, T.zero_bytes, st)) != cudaSuccess) return (int)e;
    B.flags = T.flags; B.ticket = T.ticket; B.super_count = T.super_count;
    B.visible_ids = visible_ids; B.vrec = T.vrec; B.counts_dev = counts_dev;
    const int nblk_cam = hgs_ceil_div(N, CH);
    ShadeArgs S;
    S.feats = feats; S.opacities = opacities; S.campos = campos; S.K = K; S.depth_channel = depth_channel;
    S.colors = colors_out; S.recs = reinterpret_cast<hgs::GRec*>(records);
#define LAUNCH(SH)                                                                                                    \
    project3d_fwd_kernel<true, SH><<<nblk_cam * C, PB, 0, st>>>(                                                      \
        means, quats, scales, viewmats, Ks, N, nblk_cam, width, height, eps2d, near_plane, far_plane, radius_clip,    \
        tile_size, tile_w, tile_h, radii, means2d, depths, conics, compensations, tiles_per_gauss, B, S);
    switch (shade) {
        case SHADE_NONE: LAUNCH(SHADE_NONE) break;
        case SHADE_RGB: LAUNCH(SHADE_RGB) break;
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
