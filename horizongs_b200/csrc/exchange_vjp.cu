// Stage e, fused form: the backward of the per-Gaussian stages (a7 spherical harmonics,
// a3 projection) FUSED with the exchange of the view-sharded gradients over NVLink peer memory.
//
// A view's gradient w.r.t. the 38 floats of a Gaussian has 27 floats of SH-coefficient gradient, and that part is
// RANK ONE: v_coeffs[k][c] = basis_k(direction) * v_colour[c], where the direction is (mean - camera position)
// and every rank holds the means.  So the ranks do not exchange 38 gradients per visible Gaussian (160 B,
// csrc/exchange.cu) but a 64-byte record:
//   push   : one thread per visible Gaussian runs the camera-specific backward -- projection VJP (means, quats,
//            scales), SH direction gradient (into means), densification norm -- from the row the blend backward
//            left in `vpack`, and the record {v_means 3, v_opacity, v_quats 4, v_scales 3, norm, v_colour 3} is
//            staged in shared memory and stored with TMA bulk copies into a mailbox slot in EVERY peer's memory
//            (release flag; exchange_common.cuh).  The heavy math runs once per (Gaussian, view) pair, on the
//            rank that rendered the view, on dense warps.
//   reduce : on every rank, one lane per (Gaussian, view) pair expands the rank-one SH part (basis of that
//            view's direction x v_colour) and the pairs of a Gaussian are summed IN RANK ORDER; the dense
//            gradient tensors are written once (zeros for untouched rows) and the densification statistics
//            (scene/basic_model.py:131-144) are updated in the same pass.
// This replaces, per step and rank: sh_bwd + project3d_bwd + densify_stats + the gradient all-reduce, with
// 2.5x less NVLink traffic than the sparse all-reduce and 25x less than the dense one.
// All replicas add the same numbers in the same order: bit-identical gradients on every rank.
// (Built with fused multiply-add: the backward needs no bit-equality with the oracle; the single-GPU backward
// kernels use the same device functions with -fmad=false.)
#include <stdlib.h>

#include "exchange_common.cuh"
#include "hgs_constants.cuh"
#include "project2d_math.cuh"
#include "project3d_math.cuh"
#include "sh_math.cuh"

namespace {

constexpr int VJ_ROW = 16;        // floats per record: v_means 3, v_opacity | v_quats 4 | v_scales 3, norm | v_colour 3, -
constexpr int VJ_R4 = VJ_ROW / 4;
constexpr int VJ_CAM = 28;        // viewmat 16, K 9, camera position 3 (slot header, after the row count)
constexpr int VJ_PUSH = 256;      // records computed and staged per iteration of a push CTA (one per thread)

__device__ __forceinline__ HgsCam cam_from(const float* c) {
    HgsCam cam;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) cam.R[i][j] = c[i * 4 + j];
        cam.t[i] = c[i * 4 + 3];
    }
    cam.fx = c[16 + 0];
    cam.fy = c[16 + 4];
    cam.cx = c[16 + 2];
    cam.cy = c[16 + 5];
    return cam;
}

// The push kernel: thread t of a CTA computes the record of one visible Gaussian (the camera-specific backward),
// the CTA stages 256 records in shared memory and stores them with one TMA bulk copy per peer (double buffered).
template <int DEG>
__global__ void __launch_bounds__(VJ_PUSH) vjp_push_kernel(
    ExPeers P, ExLayout L, const float4* __restrict__ vpack, const float* __restrict__ colors_fwd,
    const float* __restrict__ viewmat, const float* __restrict__ Kmat, const float* __restrict__ campos,
    const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
    const float* __restrict__ coeffs, int K, float Wf, float Hf, float eps2d, float near_plane, float far_plane,
    const int32_t* __restrict__ ids, int n_rows, int parity, unsigned long long flag_value) {
    __shared__ __align__(128) float4 stage[2][VJ_PUSH * VJ_R4];
    const size_t soff = slot_offset(L, parity, P.rank);
    write_block_entries(P, L, soff, ids, n_rows);
    if (blockIdx.x == 0 && threadIdx.x <= VJ_CAM) {
        const int t = threadIdx.x;
        if (P.mc != nullptr) {
            unsigned char* hdr = P.mc + soff;
            if (t == 0) {
                mc_st_u64(hdr, (unsigned long long)(long long)n_rows);
            } else {
                const int k = t - 1;
                mc_st_f32(reinterpret_cast<float*>(hdr + 8) + k, k < 16 ? viewmat[k] : (k < 25 ? Kmat[k - 16] : campos[k - 25]));
            }
        }
        for (int q = 0; q < (P.mc != nullptr ? 0 : P.world); ++q) {
            unsigned char* hdr = P.base[q] + soff;
            if (t == 0) {
                *reinterpret_cast<long long*>(hdr) = n_rows;
            } else {
                const int k = t - 1;
                reinterpret_cast<float*>(hdr + 8)[k] = k < 16 ? viewmat[k] : (k < 25 ? Kmat[k - 16] : campos[k - 25]);
            }
        }
    }
    float camf[VJ_CAM];
#pragma unroll
    for (int k = 0; k < VJ_CAM; ++k) camf[k] = k < 16 ? viewmat[k] : (k < 25 ? Kmat[k - 16] : campos[k - 25]);
    const HgsCam cam = cam_from(camf);
    const int rowlen = K * 3;
    const int n_chunks = (n_rows + VJ_PUSH - 1) / VJ_PUSH;
    int it = 0;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
        const int st = it & 1;
        const int r0 = chunk * VJ_PUSH;
        const int nr = min(VJ_PUSH, n_rows - r0);
        if (P.mc == nullptr) {
            if (threadIdx.x < P.world) bulk_wait_read<1>();   // stage `st` was last read by the stores of iteration it - 2
            __syncthreads();
        }
        if (threadIdx.x < nr) {
            const long long n = ids[r0 + threadIdx.x];
            const float4 q0 = vpack[n * 3], q1 = vpack[n * 3 + 1];     // v_means2d, v_conics a b | c, v_opacity
            float4 q2 = vpack[n * 3 + 2];                              // v_colour, v_depth
            if (colors_fwd != nullptr) {
                // gsplat's clamp_min(colour + 0.5, 0): no gradient through a clamped channel
                if (!(colors_fwd[n * 3 + 0] > 0.f)) q2.x = 0.f;
                if (!(colors_fwd[n * 3 + 1] > 0.f)) q2.y = 0.f;
                if (!(colors_fwd[n * 3 + 2] > 0.f)) q2.z = 0.f;
            }
            const float px = means[n * 3], py = means[n * 3 + 1], pz = means[n * 3 + 2];
            const float s0 = scales[n * 3], s1 = scales[n * 3 + 1], s2 = scales[n * 3 + 2];
            const float4 qv = reinterpret_cast<const float4*>(quats)[n];
            float g_mean[3] = {0.f, 0.f, 0.f}, g_scale[3] = {0.f, 0.f, 0.f}, g_quat[4] = {0.f, 0.f, 0.f, 0.f};
            if (DEG >= 1) {
                float gd0, gd1, gd2;
                sh_dirgrad_one<(DEG >= 1 ? DEG : 1)>(px - camf[25], py - camf[26], pz - camf[27], coeffs + n * rowlen, q2.x,
                                                     q2.y, q2.z, gd0, gd1, gd2);
                g_mean[0] = gd0; g_mean[1] = gd1; g_mean[2] = gd2;
            }
            Proj3dFwd f;
            if (proj3d_math(cam, px, py, pz, qv.x, qv.y, qv.z, qv.w, s0, s1, s2, Wf, Hf, eps2d, near_plane, far_plane, f))
                proj3d_bwd_one(cam, f, s0, s1, s2, make_float2(q0.x, q0.y), q2.w, q0.z, 0.5f * q0.w, q1.x, g_mean, g_scale,
                               g_quat);
            const float gx = q0.x * (0.5f * Wf), gy = q0.y * (0.5f * Hf);
            const float4 r0v = make_float4(g_mean[0], g_mean[1], g_mean[2], q1.y);
            const float4 r1v = make_float4(g_quat[0], g_quat[1], g_quat[2], g_quat[3]);
            const float4 r2v = make_float4(g_scale[0], g_scale[1], g_scale[2], sqrtf(gx * gx + gy * gy));
            const float4 r3v = make_float4(q2.x, q2.y, q2.z, 0.f);
            if (P.mc != nullptr) {
                // one multicast store per 16 bytes: the switch replicates the record into every rank's mailbox
                float4* dst = reinterpret_cast<float4*>(P.mc + soff + L.rows_off) + (size_t)(r0 + threadIdx.x) * VJ_R4;
                mc_st_v4(dst, r0v); mc_st_v4(dst + 1, r1v); mc_st_v4(dst + 2, r2v); mc_st_v4(dst + 3, r3v);
            } else {
                float4* rec = stage[st] + threadIdx.x * VJ_R4;
                rec[0] = r0v; rec[1] = r1v; rec[2] = r2v; rec[3] = r3v;
            }
        }
        if (P.mc != nullptr) continue;
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x < P.world) {
            bulk_s2g(P.base[threadIdx.x] + soff + L.rows_off + (size_t)r0 * VJ_ROW * 4, stage[st],
                     (unsigned)(nr * VJ_ROW * 4));
            bulk_commit();
        }
    }
    publish_push(P, flag_value);
}

// The 2DGS push kernel: the same record, computed from the 24-float row the surfel blend backward leaves in its
// vpack ([0:2] v_means2d, [2:11] v_ray_transforms, [11:14] v_normals, [14] v_opacity, [16:16+3] v_colour,
// [19] v_depth when the depth channel is rendered, [20:22] densification gradient; see hgs_blend2d_bwd_packed)
// with the surfel projection VJP (project2d_math.cuh).  The densification norm uses v_means2d + the densification
// gradient, as meta["means2d"].grad does on one GPU (rendering.py _DensifyProbe / _DensifyInject).
constexpr int VJ_ROW2D = 24;
template <int DEG>
__global__ void __launch_bounds__(VJ_PUSH) vjp_push2d_kernel(
    ExPeers P, ExLayout L, const float* __restrict__ vpack, int has_depth, const float* __restrict__ colors_fwd,
    const float* __restrict__ viewmat, const float* __restrict__ Kmat, const float* __restrict__ campos,
    const float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
    const float* __restrict__ coeffs, int K, float Wf, float Hf, float near_plane, float far_plane,
    const int32_t* __restrict__ ids, int n_rows, int parity, unsigned long long flag_value) {
    __shared__ __align__(128) float4 stage[2][VJ_PUSH * VJ_R4];
    const size_t soff = slot_offset(L, parity, P.rank);
    write_block_entries(P, L, soff, ids, n_rows);
    if (blockIdx.x == 0 && threadIdx.x <= VJ_CAM) {
        const int t = threadIdx.x;
        if (P.mc != nullptr) {
            unsigned char* hdr = P.mc + soff;
            if (t == 0) {
                mc_st_u64(hdr, (unsigned long long)(long long)n_rows);
            } else {
                const int k = t - 1;
                mc_st_f32(reinterpret_cast<float*>(hdr + 8) + k, k < 16 ? viewmat[k] : (k < 25 ? Kmat[k - 16] : campos[k - 25]));
            }
        }
        for (int q = 0; q < (P.mc != nullptr ? 0 : P.world); ++q) {
            unsigned char* hdr = P.base[q] + soff;
            if (t == 0) {
                *reinterpret_cast<long long*>(hdr) = n_rows;
            } else {
                const int k = t - 1;
                reinterpret_cast<float*>(hdr + 8)[k] = k < 16 ? viewmat[k] : (k < 25 ? Kmat[k - 16] : campos[k - 25]);
            }
        }
    }
    float camf[VJ_CAM];
#pragma unroll
    for (int k = 0; k < VJ_CAM; ++k) camf[k] = k < 16 ? viewmat[k] : (k < 25 ? Kmat[k - 16] : campos[k - 25]);
    const HgsCam cam = cam_from(camf);
    const int rowlen = K * 3;
    const int n_chunks = (n_rows + VJ_PUSH - 1) / VJ_PUSH;
    int it = 0;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
        const int st = it & 1;
        const int r0 = chunk * VJ_PUSH;
        const int nr = min(VJ_PUSH, n_rows - r0);
        if (P.mc == nullptr) {
            if (threadIdx.x < P.world) bulk_wait_read<1>();
            __syncthreads();
        }
        if (threadIdx.x < nr) {
            const long long n = ids[r0 + threadIdx.x];
            const float* row = vpack + n * VJ_ROW2D;
            float v0 = row[16], v1 = row[17], v2 = row[18];
            if (colors_fwd != nullptr) {
                if (!(colors_fwd[n * 3 + 0] > 0.f)) v0 = 0.f;
                if (!(colors_fwd[n * 3 + 1] > 0.f)) v1 = 0.f;
                if (!(colors_fwd[n * 3 + 2] > 0.f)) v2 = 0.f;
            }
            const float px = means[n * 3], py = means[n * 3 + 1], pz = means[n * 3 + 2];
            const float s0 = scales[n * 3], s1 = scales[n * 3 + 1];
            const float4 qv = reinterpret_cast<const float4*>(quats)[n];
            float g_mean[3] = {0.f, 0.f, 0.f}, g_scale[3] = {0.f, 0.f, 0.f}, g_quat[4] = {0.f, 0.f, 0.f, 0.f};
            if (DEG >= 1) {
                float gd0, gd1, gd2;
                sh_dirgrad_one<(DEG >= 1 ? DEG : 1)>(px - camf[25], py - camf[26], pz - camf[27], coeffs + n * rowlen, v0, v1,
                                                     v2, gd0, gd1, gd2);
                g_mean[0] = gd0; g_mean[1] = gd1; g_mean[2] = gd2;
            }
            Proj2dFwd f;
            if (proj2d_math(cam, px, py, pz, qv.x, qv.y, qv.z, qv.w, s0, s1, near_plane, far_plane, f))
                proj2d_bwd_one(cam, f, s0, s1, vpack, VJ_ROW2D, has_depth ? vpack + 19 : nullptr, VJ_ROW2D, vpack + 2, VJ_ROW2D,
                               vpack + 11, VJ_ROW2D, n, g_mean, g_scale, g_quat);
            const float gx = (row[0] + row[20]) * (0.5f * Wf), gy = (row[1] + row[21]) * (0.5f * Hf);
            const float4 r0v = make_float4(g_mean[0], g_mean[1], g_mean[2], row[14]);
            const float4 r1v = make_float4(g_quat[0], g_quat[1], g_quat[2], g_quat[3]);
            const float4 r2v = make_float4(g_scale[0], g_scale[1], g_scale[2], sqrtf(gx * gx + gy * gy));
            const float4 r3v = make_float4(v0, v1, v2, 0.f);
            if (P.mc != nullptr) {
                float4* dst = reinterpret_cast<float4*>(P.mc + soff + L.rows_off) + (size_t)(r0 + threadIdx.x) * VJ_R4;
                mc_st_v4(dst, r0v); mc_st_v4(dst + 1, r1v); mc_st_v4(dst + 2, r2v); mc_st_v4(dst + 3, r3v);
            } else {
                float4* rec = stage[st] + threadIdx.x * VJ_R4;
                rec[0] = r0v; rec[1] = r1v; rec[2] = r2v; rec[3] = r3v;
            }
        }
        if (P.mc != nullptr) continue;
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x < P.world) {
            bulk_s2g(P.base[threadIdx.x] + soff + L.rows_off + (size_t)r0 * VJ_ROW * 4, stage[st],
                     (unsigned)(nr * VJ_ROW * 4));
            bulk_commit();
        }
    }
    publish_push(P, flag_value);
}

// The reduce kernel.  One CTA of 8 warps owns RANGE consecutive Gaussian ids (RANGE / 256 block entries per
// source).  Set-up (one round trip, a few barriers): the sources' presence bitmaps of the range, their union, the
// ordered list of touched ids and the prefix sum of their (Gaussian, source) PAIR counts go to shared memory.
// The pairs are the unit of work: one lane = one pair, so a Gaussian seen by three views costs three lanes, not
// three serial iterations of one lane.  The touched list is cut into rounds of whole Gaussians holding at most 32
// pairs (a new round starts whenever the running pair count passes a multiple of T = 33 - world), round k
// belongs to warp k % 8, and from there the warps are AUTONOMOUS (no CTA barrier).  A round
//   1. every lane loads its pair's 64-byte record and its Gaussian's mean; meanwhile the id range the round
//      covers is zero-filled with streaming 16-byte stores,
//   2. every lane expands the rank-one SH part of its pair: basis(mean - that view's camera position) x v_colour,
//   3. the pairs of one Gaussian are summed IN RANK ORDER by the lane of its first pair (through a
//      shared-memory tile; single-source Gaussians -- the majority -- skip it), which writes the per-Gaussian
//      rows and leaves the coefficient-gradient row in the tile,
//   4. writes the round's coefficient-gradient rows, coalesced.  (The id range the round covers -- from its first
//      touched id up to the next round's first -- was zeroed with streaming 16-byte stores in step 1, so the
//      untouched ids in between end up as zero rows.)
// DEG >= 0: SH coefficients [N,K,3] of degree DEG (the `post` clamp mask was applied by the pusher);
// DEG == -1: plain colours [N,3] (K == 1)
constexpr int VJ_WARPS = 8;
constexpr int VJ_RANGE = 2048;                 // Gaussian ids per CTA
constexpr int VJ_NW = VJ_RANGE / 32;           // bitmap words per CTA
constexpr int VJ_OUT = 13;                     // per-pair outputs besides the coefficient gradients
struct VjSmem {      // followed by unsigned bits[world][NW], int first[world][NW], then the per-warp tiles
    float cam[EX_MAX_W][VJ_CAM];
    unsigned uni[VJ_NW];                // union of the sources' bitmaps
    int upre[VJ_NW];                    // touched ids before word w
    int pstart[VJ_RANGE + 4];           // pairs before touched Gaussian j (pstart[n_touched] = all pairs)
    unsigned short list[VJ_RANGE];      // touched local ids, ascending
    unsigned short round_first[VJ_RANGE + 32];   // first touched Gaussian of round r
    int scan[VJ_WARPS];
    int n_touched, n_rounds;
    int pad[2];
};
static_assert(sizeof(VjSmem) % 16 == 0, "tile alignment");

template <int DEG>
__global__ void __launch_bounds__(VJ_WARPS * 32, (DEG <= 2 ? 3 : 2)) vjp_reduce_kernel(
    ExLayout L, const unsigned char* mailbox, int parity, const int* __restrict__ status,
    const float* __restrict__ means, int K, float* __restrict__ v_means, float* __restrict__ v_quats,
    float* __restrict__ v_scales, float* __restrict__ v_opac, float* __restrict__ v_coeffs,
    float* __restrict__ grad_accum, float* __restrict__ denom) {
    constexpr int NB = DEG >= 0 ? (DEG + 1) * (DEG + 1) : 1;
    constexpr int RL = NB * 3;
    constexpr int RS = RL | 1;
    constexpr int PS = (RL + VJ_OUT) | 1;     // stride of a pair's partial-result row
    constexpr int NW = VJ_NW, RANGE = VJ_RANGE;
    constexpr int TB = VJ_WARPS * 32;
    constexpr int PER = RANGE / TB;           // touched-list entries per thread in the pair-count scan
    extern __shared__ __align__(128) unsigned char ex_smem[];
    VjSmem& S = *reinterpret_cast<VjSmem*>(ex_smem);
    const int world = L.world;
    unsigned* s_bits = reinterpret_cast<unsigned*>(ex_smem + sizeof(VjSmem));              // [world][NW]
    int* s_first = reinterpret_cast<int*>(s_bits + world * NW);                            // [world][NW]
    float* s_tiles = reinterpret_cast<float*>(s_first + world * NW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // a peer never arrived, or a contribution overflowed its slot: behave as if nobody had pushed anything (all
    // gradients of this step become zero, the statistics stay) -- the host raises when it reads the status
    const bool dead = __ldcg(status) != 0;
    const long long id0 = (long long)blockIdx.x * RANGE;
    const int ent0 = blockIdx.x * (RANGE / EX_IDS);
    // ---- set-up: bitmaps (and first-record indices) of the range from every source, cameras
    for (int i = tid; i < world * NW; i += TB) {
        const int s = i / NW, w = i - s * NW;
        const int ent = ent0 + (w >> 3), ww = w & 7;
        unsigned bits = 0u;
        int first = 0;
        if (ent < L.n_blocks && !dead) {
            const unsigned* e = reinterpret_cast<const unsigned*>(mailbox + slot_offset(L, parity, s) + L.ent_off) +
                                (size_t)ent * EX_ENTRY_WORDS;
            first = (int)__ldcg(e);
            for (int k = 0; k < ww; ++k) first += __popc(__ldcg(e + 1 + k));
            bits = __ldcg(e + 1 + ww);
        }
        s_bits[i] = bits;
        s_first[i] = first;
    }
    for (int i = tid; i < world * VJ_CAM; i += TB) {
        const int s = i / VJ_CAM, k = i - s * VJ_CAM;
        S.cam[s][k] = __ldcg(reinterpret_cast<const float*>(mailbox + slot_offset(L, parity, s) + 8) + k);
    }
    __syncthreads();
    if (tid < NW) {
        unsigned u = 0u;
        for (int s = 0; s < world; ++s) u |= s_bits[s * NW + tid];
        S.uni[tid] = u;
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive prefix of the words' popcounts (each lane scans NW / 32 consecutive words)
        constexpr int WPL = (NW + 31) / 32;
        int local[WPL], sum = 0;
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
            const int w = lane * WPL + k;
            local[k] = sum;
            sum += w < NW ? __popc(S.uni[w]) : 0;
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        const int excl = incl - sum;
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
            const int w = lane * WPL + k;
            if (w < NW) S.upre[w] = excl + local[k];
        }
        if (lane == 31) S.n_touched = incl;
    }
    __syncthreads();
    for (int w = tid; w < NW; w += TB) {
        unsigned u = S.uni[w];
        int j = S.upre[w];
        while (u) {
            const int b = __ffs(u) - 1;
            u &= u - 1;
            S.list[j++] = (unsigned short)(w * 32 + b);
        }
    }
    __syncthreads();
    const int n_touched = S.n_touched;
    auto present = [&](int l) -> unsigned {
        unsigned m = 0u;
        for (int s = 0; s < world; ++s) m |= ((s_bits[s * NW + (l >> 5)] >> (l & 31)) & 1u) << s;
        return m;
    };
    {
        // pair counts of the touched Gaussians and their exclusive prefix (thread t: entries [t * PER, t * PER + PER))
        int cnt[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int j = tid * PER + k;
            cnt[k] = j < n_touched ? __popc(present(S.list[j])) : 0;
            sum += cnt[k];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) S.scan[warp] = incl;
        __syncthreads();
        int base = incl - sum;
        for (int w = 0; w < warp; ++w) base += S.scan[w];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int j = tid * PER + k;
            if (j <= n_touched) S.pstart[j] = base;       // pstart[n_touched] = all pairs
            base += cnt[k];
        }
        if (tid == TB - 1 && n_touched == RANGE) S.pstart[RANGE] = base;
    }
    __syncthreads();
    // rounds: whole Gaussians, at most 32 pairs: a new round starts where floor(pstart / T) changes
    const int T = 33 - world;
    for (int j = tid; j < n_touched; j += TB) {
        const int r = S.pstart[j] / T;
        if (j == 0 || r != S.pstart[j - 1] / T) S.round_first[r] = (unsigned short)j;
    }
    if (tid == 0) S.n_rounds = n_touched > 0 ? S.pstart[n_touched - 1] / T + 1 : 0;
    __syncthreads();
    // ---- autonomous warps from here on
    const int n_rounds = S.n_rounds;
    const int rowlen = K * 3;
    const long long id_end = min(id0 + RANGE, L.n_ids);
    // one tile per warp, used as [32][RS] coefficient rows (steps 1-2, 4) and as [32][PS] partial results (step 3)
    float* rows = s_tiles + warp * (32 * PS);
    float* part = rows;
    for (int rnd = warp; rnd < max(n_rounds, 1); rnd += VJ_WARPS) {
        const int j0 = n_rounds > 0 ? S.round_first[rnd] : 0;
        const int j1 = n_rounds > 0 ? (rnd + 1 < n_rounds ? S.round_first[rnd + 1] : n_touched) : 0;
        const int nG = j1 - j0;                                // Gaussians of the round (<= 32)
        const int p0 = n_rounds > 0 ? S.pstart[j0] : 0;
        const int nP = n_rounds > 0 ? S.pstart[j1] - p0 : 0;   // pairs of the round (<= 32)
        // lane -> its pair: Gaussian g (index in the round) and source s
        int g = 0;
        {
            // the lane of Gaussian gg's first pair marks a segment head; g = number of heads at or before the lane - 1
            unsigned heads = 0u;
            const int my_off = lane < nG ? S.pstart[j0 + lane] - p0 : 32;
            for (int gg = 0; gg < nG; ++gg) heads |= 1u << __shfl_sync(0xFFFFFFFFu, my_off, gg);
            g = __popc(heads & (0xFFFFFFFFu >> (31 - lane))) - 1;
        }
        const bool mine = lane < nP;
        const int l = mine ? S.list[j0 + g] : 0;
        const long long n = id0 + l;
        const int seg_off = mine ? S.pstart[j0 + g] - p0 : 0;       // lane of the Gaussian's first pair
        const int seg_len = mine ? S.pstart[j0 + g + 1] - S.pstart[j0 + g] : 0;
        int src = 0;
        if (mine) {
            unsigned m = present(l);
            for (int k = lane - seg_off; k > 0; --k) m &= m - 1;    // (lane - seg_off)-th source that saw it
            src = __ffs(m) - 1;
        }
        // (1) loads: the pair's record and the Gaussian's mean
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = q0, q3 = q0;
        float px = 0.f, py = 0.f, pz = 1.f;
        if (mine) {
            const int rec_i = s_first[src * NW + (l >> 5)] + __popc(s_bits[src * NW + (l >> 5)] & ((1u << (l & 31)) - 1u));
            const float4* rec = reinterpret_cast<const float4*>(mailbox + slot_offset(L, parity, src) + L.rows_off) +
                                (size_t)rec_i * VJ_R4;
            q0 = __ldcg(rec); q1 = __ldcg(rec + 1); q2 = __ldcg(rec + 2); q3 = __ldcg(rec + 3);
            if (DEG >= 1) { px = means[n * 3]; py = means[n * 3 + 1]; pz = means[n * 3 + 2]; }
        }
        // the id range this round covers: [its first touched id (or the CTA's first id), the next round's first).
        // All of its output rows are zeroed first (plain streaming stores, issued while the loads above are in
        // flight); the Gaussians of the round overwrite theirs in steps 3 and 4.
        const long long lo = rnd == 0 ? id0 : id0 + S.list[j0];
        const long long hi = rnd + 1 >= n_rounds ? id_end : id0 + S.list[j1];
        {
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (long long gi = lo + lane; gi < hi; gi += 32) {
                reinterpret_cast<float4*>(v_quats)[gi] = z4;
                v_opac[gi] = 0.f;
            }
            for (long long e = lo * 3 + lane; e < hi * 3; e += 32) {
                v_means[e] = 0.f;
                v_scales[e] = 0.f;
            }
            const long long e_lo = lo * rowlen, e_hi = hi * rowlen;
            const long long a_lo = min((e_lo + 3) & ~3ll, e_hi), a_hi = max(e_hi & ~3ll, a_lo);
            for (long long e = e_lo + lane; e < a_lo; e += 32) v_coeffs[e] = 0.f;
            for (long long e = a_lo + 4 * lane; e < a_hi; e += 128) *reinterpret_cast<float4*>(v_coeffs + e) = z4;
            for (long long e = a_hi + lane; e < e_hi; e += 32) v_coeffs[e] = 0.f;
        }
        // (2) the pair's contribution: the rank-one SH part expanded, the rest as the pusher computed it
        float g_co[RL];
        float o[VJ_OUT];     // v_means 0..2 | v_quats 3..6 | v_scales 7..9 | opacity 10 | norm 11 | count 12
#pragma unroll
        for (int k = 0; k < RL; ++k) g_co[k] = 0.f;
#pragma unroll
        for (int k = 0; k < VJ_OUT; ++k) o[k] = 0.f;
        if (mine) {
            if (DEG >= 0) {
                float b[NB];
                if (DEG >= 1) sh_basis_of<(DEG >= 0 ? DEG : 0)>(px - S.cam[src][25], py - S.cam[src][26], pz - S.cam[src][27], b);
                else b[0] = 0.2820947917738781f;
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                    g_co[k * 3] = b[k] * q3.x; g_co[k * 3 + 1] = b[k] * q3.y; g_co[k * 3 + 2] = b[k] * q3.z;
                }
            } else {
                g_co[0] = q3.x; g_co[1] = q3.y; g_co[2] = q3.z;
            }
            o[0] = q0.x; o[1] = q0.y; o[2] = q0.z; o[10] = q0.w;
            o[3] = q1.x; o[4] = q1.y; o[5] = q1.z; o[6] = q1.w;
            o[7] = q2.x; o[8] = q2.y; o[9] = q2.z; o[11] = q2.w;
            o[12] = 1.f;
        }
        // (3) sum a Gaussian's pairs in rank order at its first pair's lane
        const bool head = mine && lane == seg_off;
        if (mine && !head) {
            float* pr = part + lane * PS;
#pragma unroll
            for (int k = 0; k < RL; ++k) pr[k] = g_co[k];
#pragma unroll
            for (int k = 0; k < VJ_OUT; ++k) pr[RL + k] = o[k];
        }
        __syncwarp();
        if (head) {
            for (int t = 1; t < seg_len; ++t) {
                const float* pr = part + (lane + t) * PS;
#pragma unroll
                for (int k = 0; k < RL; ++k) g_co[k] += pr[k];
#pragma unroll
                for (int k = 0; k < VJ_OUT; ++k) o[k] += pr[RL + k];
            }
        }
        __syncwarp();      // the partial results are consumed: the tile becomes the rows tile again
        if (head) {
            float* out = rows + g * RS;
#pragma unroll
            for (int k = 0; k < RL; ++k) out[k] = g_co[k];
            v_means[n * 3] = o[0];
            v_means[n * 3 + 1] = o[1];
            v_means[n * 3 + 2] = o[2];
            reinterpret_cast<float4*>(v_quats)[n] = make_float4(o[3], o[4], o[5], o[6]);
            v_scales[n * 3] = o[7];
            v_scales[n * 3 + 1] = o[8];
            v_scales[n * 3 + 2] = o[9];
            v_opac[n] = o[10];
            if (grad_accum != nullptr) grad_accum[n] += o[11];
            if (denom != nullptr) denom[n] += o[12];
        }
        __syncwarp();
        // (4) the round's coefficient-gradient rows, coalesced: consecutive lanes write consecutive floats of a row
        {
            const unsigned magic_rl = 0xFFFFFFFFu / (unsigned)RL + 1u;
#pragma unroll
            for (int k = 0; k < RL; ++k) {
                const int i = lane + 32 * k;
                const int j = (int)__umulhi((unsigned)i, magic_rl), cc = i - j * RL;
                if (j < nG) v_coeffs[(id0 + S.list[j0 + j]) * (long long)rowlen + cc] = rows[j * RS + cc];
            }
        }
        __syncwarp();      // the tiles are reused by this warp's next round
    }
}

template <int DEG>
int launch_reduce(const ExLayout& L, const unsigned char* mailbox, int parity, const int* status, const float* means,
                  int K, float* v_means, float* v_quats, float* v_scales, float* v_opac, float* v_coeffs,
                  float* grad_accum, float* denom, cudaStream_t st) {
    constexpr int NB = DEG >= 0 ? (DEG + 1) * (DEG + 1) : 1;
    constexpr int RL = NB * 3;
    const int smem = (int)(sizeof(VjSmem) + (size_t)L.world * VJ_NW * 8 +
                           (size_t)VJ_WARPS * 32 * ((RL + VJ_OUT) | 1) * sizeof(float));
    cudaError_t e = cudaFuncSetAttribute(vjp_reduce_kernel<DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    const int grid = (int)((L.n_ids + VJ_RANGE - 1) / VJ_RANGE);
    vjp_reduce_kernel<DEG><<<grid, VJ_WARPS * 32, smem, st>>>(L, mailbox, parity, status, means, K, v_means, v_quats,
                                                             v_scales, v_opac, v_coeffs, grad_accum, denom);
    HGS_LAUNCH_CHECK();
    return 0;
}

template <int DEG>
int launch_push(const ExPeers& P, const ExLayout& L, const float* vpack, const float* colors_fwd, const float* viewmat,
                const float* Kmat, const float* campos, const float* means, const float* quats, const float* scales,
                const float* coeffs, int K, int width, int height, float eps2d, float near_plane, float far_plane,
                const int32_t* ids, int n_rows, int parity, unsigned long long flag_value, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int chunks = (n_rows + VJ_PUSH - 1) / VJ_PUSH;
    const int grid = chunks < 1 ? 1 : (chunks > sms * 4 ? sms * 4 : chunks);
    vjp_push_kernel<DEG><<<grid, VJ_PUSH, 0, st>>>(P, L, reinterpret_cast<const float4*>(vpack), colors_fwd, viewmat, Kmat,
                                                   campos, means, quats, scales, coeffs, K, (float)width, (float)height,
                                                   eps2d, near_plane, far_plane, ids, n_rows, parity, flag_value);
    HGS_LAUNCH_CHECK();
    return 0;
}

template <int DEG>
int launch_push2d(const ExPeers& P, const ExLayout& L, const float* vpack, int has_depth, const float* colors_fwd,
                  const float* viewmat, const float* Kmat, const float* campos, const float* means, const float* quats,
                  const float* scales, const float* coeffs, int K, int width, int height, float near_plane, float far_plane,
                  const int32_t* ids, int n_rows, int parity, unsigned long long flag_value, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int chunks = (n_rows + VJ_PUSH - 1) / VJ_PUSH;
    const int grid = chunks < 1 ? 1 : (chunks > sms * 4 ? sms * 4 : chunks);
    vjp_push2d_kernel<DEG><<<grid, VJ_PUSH, 0, st>>>(P, L, vpack, has_depth, colors_fwd, viewmat, Kmat, campos, means, quats,
                                                     scales, coeffs, K, (float)width, (float)height, near_plane, far_plane, ids,
                                                     n_rows, parity, flag_value);
    HGS_LAUNCH_CHECK();
    return 0;
}

}  // namespace

HGS_API size_t hgs_exchange_vjp_mailbox_bytes(int world, long long n_ids, long long cap_rows) {
    if (ex_bad_geometry(world, 0, n_ids, cap_rows)) return 0;
    return EX_CTRL_BYTES + (size_t)2 * world * make_layout(world, n_ids, cap_rows, VJ_ROW).slot_bytes;
}

HGS_API int hgs_exchange_vjp_push(int sh_degree, int K, const float* vpack, const float* colors_fwd, const float* viewmat,
                                  const float* Kmat, const float* campos, const float* means, const float* quats,
                                  const float* scales, const float* coeffs, int width, int height, float eps2d,
                                  float near_plane, float far_plane, long long n_ids, const int32_t* ids,
                                  long long n_rows, long long cap_rows, void* const* mailboxes_host, void* multicast_base,
                                  int world, int rank, unsigned long long step, void* stream) {
    if (ex_bad_geometry(world, rank, n_ids, cap_rows) || n_rows < 0 || vpack == nullptr || viewmat == nullptr ||
        Kmat == nullptr || campos == nullptr || means == nullptr || quats == nullptr || scales == nullptr ||
        (reinterpret_cast<size_t>(vpack) & 15) || (reinterpret_cast<size_t>(quats) & 15) || sh_degree < -1 ||
        sh_degree > 4 || width <= 0 || height <= 0)
        return HGS_ERR_INVALID_ARG;
    if (sh_degree >= 0 ? K < (sh_degree + 1) * (sh_degree + 1) : K != 1) return HGS_ERR_INVALID_ARG;
    if (sh_degree >= 1 && coeffs == nullptr) return HGS_ERR_INVALID_ARG;
    if (n_rows > 0 && ids == nullptr) return HGS_ERR_INVALID_ARG;
    if (n_rows > cap_rows) n_rows = -1;   // overflow: push no records and a negative row count (see exchange_wait_kernel)
    ExPeers P;
    if (int e = ex_fill_peers(P, mailboxes_host, world, rank, multicast_base)) return e;
    const ExLayout L = make_layout(world, n_ids, cap_rows, VJ_ROW);
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DEG)                                                                                                       \
    launch_push<DEG>(P, L, vpack, colors_fwd, viewmat, Kmat, campos, means, quats, scales, coeffs, K, width, height, eps2d, \
                     near_plane, far_plane, ids, (int)n_rows, (int)(step & 1ull), step + 1ull, st)
    switch (sh_degree) {
        case 1: return CALL(1);
        case 2: return CALL(2);
        case 3: return CALL(3);
        case 4: return CALL(4);
        default: return CALL(0);     // degree 0 and plain colours: no direction gradient
    }
#undef CALL
}

HGS_API int hgs_exchange_vjp_reduce(int sh_degree, int K, const float* means, long long n_ids, long long cap_rows,
                                    const void* mailbox, int world, int rank, unsigned long long step, float* v_means,
                                    float* v_quats, float* v_scales, float* v_opacities, float* v_coeffs,
                                    float* grad_accum, float* denom, int* status_dev, void* stream) {
    if (ex_bad_geometry(world, rank, n_ids, cap_rows) || mailbox == nullptr || status_dev == nullptr || sh_degree < -1 ||
        sh_degree > 4)
        return HGS_ERR_INVALID_ARG;
    if (sh_degree >= 0 ? K < (sh_degree + 1) * (sh_degree + 1) : K != 1) return HGS_ERR_INVALID_ARG;
    if (means == nullptr || v_means == nullptr || v_quats == nullptr || v_scales == nullptr || v_opacities == nullptr ||
        v_coeffs == nullptr)
        return HGS_ERR_INVALID_ARG;
    if ((reinterpret_cast<size_t>(v_quats) & 15) || (reinterpret_cast<size_t>(v_coeffs) & 15)) return HGS_ERR_INVALID_ARG;
    const ExLayout L = make_layout(world, n_ids, cap_rows, VJ_ROW);
    cudaStream_t st = (cudaStream_t)stream;
    exchange_wait_kernel<<<1, 32, 0, st>>>((const unsigned char*)mailbox, L, (int)(step & 1ull), step + 1ull, status_dev);
    HGS_LAUNCH_CHECK();
    const unsigned char* mb = (const unsigned char*)mailbox;
    const int parity = (int)(step & 1ull);
#define CALL(DEG)                                                                                                      \
    launch_reduce<DEG>(L, mb, parity, status_dev, means, K, v_means, v_quats, v_scales, v_opacities, v_coeffs, grad_accum, \
                       denom, st)
    switch (sh_degree) {
        case -1: return CALL(-1);
        case 0: return CALL(0);
        case 1: return CALL(1);
        case 2: return CALL(2);
        case 3: return CALL(3);
        default: return CALL(4);
    }
#undef CALL
}

HGS_API int hgs_exchange_vjp_push_2dgs(int sh_degree, int K, const float* vpack24, int has_depth, const float* colors_fwd,
                                       const float* viewmat, const float* Kmat, const float* campos, const float* means,
                                       const float* quats, const float* scales, const float* coeffs, int width, int height,
                                       float near_plane, float far_plane, long long n_ids, const int32_t* ids,
                                       long long n_rows, long long cap_rows, void* const* mailboxes_host,
                                       void* multicast_base, int world, int rank, unsigned long long step, void* stream) {
    if (ex_bad_geometry(world, rank, n_ids, cap_rows) || n_rows < 0 || vpack24 == nullptr || viewmat == nullptr ||
        Kmat == nullptr || campos == nullptr || means == nullptr || quats == nullptr || scales == nullptr ||
        (reinterpret_cast<size_t>(quats) & 15) || sh_degree < -1 || sh_degree > 4 || width <= 0 || height <= 0)
        return HGS_ERR_INVALID_ARG;
    if (sh_degree >= 0 ? K < (sh_degree + 1) * (sh_degree + 1) : K != 1) return HGS_ERR_INVALID_ARG;
    if (sh_degree >= 1 && coeffs == nullptr) return HGS_ERR_INVALID_ARG;
    if (n_rows > 0 && ids == nullptr) return HGS_ERR_INVALID_ARG;
    if (n_rows > cap_rows) n_rows = -1;   // overflow: push no records and a negative row count (see exchange_wait_kernel)
    ExPeers P;
    if (int e = ex_fill_peers(P, mailboxes_host, world, rank, multicast_base)) return e;
    const ExLayout L = make_layout(world, n_ids, cap_rows, VJ_ROW);
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(DEG)                                                                                                       \
    launch_push2d<DEG>(P, L, vpack24, has_depth, colors_fwd, viewmat, Kmat, campos, means, quats, scales, coeffs, K, width, \
                       height, near_plane, far_plane, ids, (int)n_rows, (int)(step & 1ull), step + 1ull, st)
    switch (sh_degree) {
        case 1: return CALL(1);
        case 2: return CALL(2);
        case 3: return CALL(3);
        case 4: return CALL(4);
        default: return CALL(0);
    }
#undef CALL
}
