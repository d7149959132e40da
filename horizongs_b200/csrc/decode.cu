// the anchor -> neural-Gaussian decode of the LOD model, fused.
//
// Replaces scene/basic_model.py:297-371 generate_neural_gaussians (view_dim 3, appearance_dim 0, colour_dim 3,
// smooth_complement == 1, basic_model.py:43-44) -- in PyTorch ~60 launches: three MLPs
//     Linear(F + 3, F) -> ReLU -> Linear(F, {k, 7k, 3k})   (+ Tanh / - / optional Sigmoid; scene/lod_model.py:67-84)
// on cat(anchor_feat, unit view direction), the opacity > 0 mask, a [V k, 22] concatenation, a boolean gather and
// the post-processing (scale = grid_scale[3:6] * sigmoid, rot = normalize, xyz = anchor + offset * grid_scale[0:3])
// -- by three kernels with the MLP weights resident in shared memory:
//   count   : one warp per visible anchor evaluates the opacity MLP, stores the k opacities, the mask bits and the
//             number of kept offsets (the caller scans the counts: positions of the compacted rows);
//   forward : one warp per visible anchor evaluates the cov and colour MLPs (lane j = hidden unit j, then
//             lane o = output o) and writes the kept Gaussians straight into the rasterizer's input tensors
//             (means [M,3], colors [M,3], opacities [M], scales [M,3], quats [M,4]) at their compacted rows;
//   backward: one warp per visible anchor recomputes the activations, turns the gradients of its rows into output
//             gradients, back-propagates through the three MLPs and writes the gradients w.r.t. anchor,
//             anchor_feat, offset and grid scaling; the weight / bias gradients are accumulated in shared memory
//             per CTA and added to global memory once per CTA.
// F (feat_dim) = 32 = one lane per hidden unit; k (n_offsets) <= 16.  Roofline: FP32 FMA / shared memory.
#include "hgs_common.cuh"
#include "../../include/hgs_raster.h"

namespace {

constexpr int DF = 32;                  // feat_dim = hidden width
constexpr int DINMAX = DF + 3;          // MLP input: features (+ unit view direction when view_dim == 3)
constexpr int DKMAX = 16;               // max offsets per anchor
constexpr int DCMAX = 48;               // max colour floats per offset (3 = RGB, 3 (d + 1)^2 = SH of degree d <= 3)
constexpr int DWARPS = 8;
constexpr int DSMEM_MAX = 227 * 1024;

// shape of one decode call: k offsets, cd colour floats per offset, vd = 0 | 3 view-direction inputs
struct DShape {
    int k, cd, vd;
    __host__ __device__ int tot() const { return (8 + cd) * k; }     // outputs: k opacity + 7k cov + cd k colour
    __host__ __device__ int din() const { return DF + vd; }
};

struct DecodeMlp {
    const float* W1[3];   // [F, F+3]   0 = opacity, 1 = cov, 2 = colour
    const float* b1[3];   // [F]
    const float* W2[3];   // [out, F]
    const float* b2[3];   // [out]
};
struct DecodeMlpGrad {
    float* W1[3];
    float* b1[3];
    float* W2[3];
    float* b2[3];
};

// shared-memory image of the weights: W1[3][F][F+3] | b1[3][F] | W2T[F][TOT] | W2[TOT][F] | b2[TOT]
struct SmemW {
    float* W1;
    float* b1;
    float* W2T;
    float* W2;
    float* b2;
};
__host__ __device__ inline int smem_w_floats(int tot, int din) { return 3 * DF * din + 3 * DF + 2 * DF * tot + tot; }
__device__ __forceinline__ SmemW carve(float* base, int tot, int din) {
    SmemW s;
    s.W1 = base;
    s.b1 = s.W1 + 3 * DF * din;
    s.W2T = s.b1 + 3 * DF;
    s.W2 = s.W2T + DF * tot;
    s.b2 = s.W2 + tot * DF;
    return s;
}
// output o of the concatenated output vector belongs to MLP m with local index o - first(m)
__device__ __forceinline__ int mlp_of(int o, int k) { return o < k ? 0 : (o < 8 * k ? 1 : 2); }
__device__ __forceinline__ int first_of(int m, int k) { return m == 0 ? 0 : (m == 1 ? k : 8 * k); }

__device__ void load_weights(const DecodeMlp& P, const SmemW& S, const DShape D, bool need_w2_rowmajor) {
    const int tot = D.tot(), k = D.k, DIN = D.din();
    for (int i = threadIdx.x; i < 3 * DF * DIN; i += blockDim.x) S.W1[i] = P.W1[i / (DF * DIN)][i % (DF * DIN)];
    for (int i = threadIdx.x; i < 3 * DF; i += blockDim.x) S.b1[i] = P.b1[i / DF][i % DF];
    for (int i = threadIdx.x; i < tot * DF; i += blockDim.x) {
        const int o = i / DF, j = i - o * DF;
        const int m = mlp_of(o, k);
        const float w = P.W2[m][(o - first_of(m, k)) * DF + j];
        S.W2T[j * tot + o] = w;
        if (need_w2_rowmajor) S.W2[i] = w;
    }
    for (int o = threadIdx.x; o < tot; o += blockDim.x) {
        const int m = mlp_of(o, k);
        S.b2[o] = P.b2[m][o - first_of(m, k)];
    }
    __syncthreads();
}

// hidden layer of MLP m for this warp's anchor: lane j returns the pre-activation of hidden unit j
__device__ __forceinline__ float hidden_pre(const SmemW& S, int m, float x_lane, float d0, float d1, float d2, int lane,
                                            int vd) {
    const float* w = S.W1 + (m * DF + lane) * (DF + vd);
    float acc = S.b1[m * DF + lane];
#pragma unroll
    for (int i = 0; i < DF; ++i) acc += w[i] * __shfl_sync(0xFFFFFFFFu, x_lane, i);
    if (vd == 3) acc += w[DF] * d0 + w[DF + 1] * d1 + w[DF + 2] * d2;
    return acc;
}
// outputs [o_begin, o_end) of the concatenated output vector -> s_out (pre-activation); s_h = [3][F] hidden activations
__device__ __forceinline__ void outputs(const SmemW& S, const float* s_h, int k, int tot, int o_begin, int o_end,
                                        float* s_out, int lane) {
    for (int o = o_begin + lane; o < o_end; o += 32) {
        const float* h = s_h + mlp_of(o, k) * DF;
        float acc = S.b2[o];
#pragma unroll
        for (int j = 0; j < DF; ++j) acc += S.W2T[j * tot + o] * h[j];
        s_out[o] = acc;
    }
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

struct WarpAnchor {
    float x_lane, d0, d1, d2, inv_dist;   // feature of this lane, unit view direction
    float ax, ay, az;
};
__device__ __forceinline__ WarpAnchor load_anchor(const float* __restrict__ anchor, const float* __restrict__ feat,
                                                  const float* __restrict__ cam, long long a, int lane) {
    WarpAnchor w;
    w.x_lane = feat[a * DF + lane];
    w.ax = anchor[a * 3]; w.ay = anchor[a * 3 + 1]; w.az = anchor[a * 3 + 2];
    const float vx = w.ax - cam[0], vy = w.ay - cam[1], vz = w.az - cam[2];
    w.inv_dist = 1.0f / sqrtf(vx * vx + vy * vy + vz * vz);
    w.d0 = vx * w.inv_dist; w.d1 = vy * w.inv_dist; w.d2 = vz * w.inv_dist;
    return w;
}

// ---- count: opacity MLP only ------------------------------------------------------------------------------
__global__ void __launch_bounds__(DWARPS * 32) decode_count_kernel(DecodeMlp P, const float* __restrict__ anchor,
                                                                  const float* __restrict__ feat,
                                                                  const float* __restrict__ cam,
                                                                  const long long* __restrict__ vis, long long V, DShape D,
                                                                  float* __restrict__ opac_all,
                                                                  int32_t* __restrict__ bits, int32_t* __restrict__ cnt) {
    extern __shared__ float sm[];
    const int tot = D.tot(), k = D.k;
    const SmemW S = carve(sm, tot, D.din());
    load_weights(P, S, D, false);
    float* s_h = sm + smem_w_floats(tot, D.din()) + (threadIdx.x >> 5) * (3 * DF + tot);
    float* s_out = s_h + 3 * DF;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long v = (long long)blockIdx.x * DWARPS + warp; v < V; v += (long long)gridDim.x * DWARPS) {
        const long long a = vis[v];
        const WarpAnchor w = load_anchor(anchor, feat, cam, a, lane);
        s_h[lane] = fmaxf(hidden_pre(S, 0, w.x_lane, w.d0, w.d1, w.d2, lane, D.vd), 0.f);
        __syncwarp();
        outputs(S, s_h, k, tot, 0, k, s_out, lane);
        __syncwarp();
        float op = 0.f;
        if (lane < k) {
            op = tanhf(s_out[lane]);
            opac_all[v * k + lane] = op;
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, lane < k && op > 0.f);
        if (lane == 0) { bits[v] = (int32_t)m; cnt[v] = __popc(m); }
        __syncwarp();
    }
}

// ---- forward ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DWARPS * 32) decode_fwd_kernel(
    DecodeMlp P, const float* __restrict__ anchor, const float* __restrict__ feat, const float* __restrict__ offset,
    const float* __restrict__ scaling, const float* __restrict__ cam, const long long* __restrict__ vis, long long V,
    DShape D, int color_sigmoid, const float* __restrict__ opac_all, const int32_t* __restrict__ bits,
    const long long* __restrict__ row0, float* __restrict__ xyz, float* __restrict__ color, float* __restrict__ opacity,
    float* __restrict__ scales, float* __restrict__ quats) {
    extern __shared__ float sm[];
    const int tot = D.tot(), k = D.k, cd = D.cd;
    const SmemW S = carve(sm, tot, D.din());
    load_weights(P, S, D, false);
    float* s_h = sm + smem_w_floats(tot, D.din()) + (threadIdx.x >> 5) * (3 * DF + tot);
    float* s_out = s_h + 3 * DF;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long v = (long long)blockIdx.x * DWARPS + warp; v < V; v += (long long)gridDim.x * DWARPS) {
        const unsigned m = (unsigned)bits[v];
        if (m == 0u) continue;          // warp-uniform
        const long long a = vis[v];
        const WarpAnchor w = load_anchor(anchor, feat, cam, a, lane);
        s_h[DF + lane] = fmaxf(hidden_pre(S, 1, w.x_lane, w.d0, w.d1, w.d2, lane, D.vd), 0.f);
        s_h[2 * DF + lane] = fmaxf(hidden_pre(S, 2, w.x_lane, w.d0, w.d1, w.d2, lane, D.vd), 0.f);
        __syncwarp();
        outputs(S, s_h, k, tot, k, tot, s_out, lane);
        __syncwarp();
        if (lane < k && ((m >> lane) & 1u)) {
            const long long r = row0[v] + __popc(m & ((1u << lane) - 1u));
            const float* sr = s_out + k + 7 * lane;
            const float* co = s_out + 8 * k + cd * lane;
            const float* gs = scaling + a * 6;
            const float* of = offset + (a * k + lane) * 3;
            xyz[r * 3] = w.ax + of[0] * gs[0];
            xyz[r * 3 + 1] = w.ay + of[1] * gs[1];
            xyz[r * 3 + 2] = w.az + of[2] * gs[2];
            for (int c = 0; c < cd; ++c) color[r * cd + c] = color_sigmoid ? sigmoidf_(co[c]) : co[c];
#pragma unroll
            for (int c = 0; c < 3; ++c) scales[r * 3 + c] = gs[3 + c] * sigmoidf_(sr[c]);
            const float n = fmaxf(sqrtf(sr[3] * sr[3] + sr[4] * sr[4] + sr[5] * sr[5] + sr[6] * sr[6]), 1e-12f);
            reinterpret_cast<float4*>(quats)[r] = make_float4(sr[3] / n, sr[4] / n, sr[5] / n, sr[6] / n);
            opacity[r] = opac_all[v * k + lane];
        }
        __syncwarp();
    }
}

// ---- backward -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DWARPS * 32) decode_bwd_kernel(
    DecodeMlp P, DecodeMlpGrad G, const float* __restrict__ anchor, const float* __restrict__ feat,
    const float* __restrict__ offset, const float* __restrict__ scaling, const float* __restrict__ cam,
    const long long* __restrict__ vis, long long V, DShape D, int color_sigmoid, const float* __restrict__ opac_all,
    const int32_t* __restrict__ bits, const long long* __restrict__ row0, const float* __restrict__ v_xyz,
    const float* __restrict__ v_color, const float* __restrict__ v_opacity, const float* __restrict__ v_scales,
    const float* __restrict__ v_quats, float* __restrict__ g_anchor, float* __restrict__ g_feat,
    float* __restrict__ g_offset, float* __restrict__ g_scaling) {
    extern __shared__ float sm[];
    const int tot = D.tot(), k = D.k, cd = D.cd, DIN = D.din();
    const SmemW S = carve(sm, tot, DIN);
    load_weights(P, S, D, true);
    // CTA-wide gradient accumulators: dW1[3][F][F+vd] | db1[3][F] | dW2[TOT][F] | db2[TOT]
    float* a_W1 = sm + smem_w_floats(tot, DIN);
    float* a_b1 = a_W1 + 3 * DF * DIN;
    float* a_W2 = a_b1 + 3 * DF;
    float* a_b2 = a_W2 + tot * DF;
    const int n_acc = 3 * DF * DIN + 3 * DF + tot * DF + tot;
    for (int i = threadIdx.x; i < n_acc; i += blockDim.x) a_W1[i] = 0.f;
    __syncthreads();
    // per-warp scratch: h[3][F] | pre-mask [3][F] as floats | out[TOT] | gout[TOT] | x[F+3] | dpre[3][F]
    float* s_h = a_W1 + n_acc + (threadIdx.x >> 5) * (3 * DF + 2 * tot + DINMAX + 3 * DF + 8);
    float* s_out = s_h + 3 * DF;
    float* s_gout = s_out + tot;
    float* s_x = s_gout + tot;
    float* s_dpre = s_x + DINMAX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long v = (long long)blockIdx.x * DWARPS + warp; v < V; v += (long long)gridDim.x * DWARPS) {
        const unsigned m = (unsigned)bits[v];
        if (m == 0u) continue;          // no kept offset: no gradient flows through this anchor
        const long long a = vis[v];
        const WarpAnchor w = load_anchor(anchor, feat, cam, a, lane);
        float pre[3];
#pragma unroll
        for (int mm = 0; mm < 3; ++mm) {
            pre[mm] = hidden_pre(S, mm, w.x_lane, w.d0, w.d1, w.d2, lane, D.vd);
            s_h[mm * DF + lane] = fmaxf(pre[mm], 0.f);
        }
        s_x[lane] = w.x_lane;
        if (lane < 3) s_x[DF + lane] = lane == 0 ? w.d0 : (lane == 1 ? w.d1 : w.d2);
        for (int o = lane; o < tot; o += 32) s_gout[o] = 0.f;
        __syncwarp();
        outputs(S, s_h, k, tot, k, tot, s_out, lane);
        __syncwarp();
        // output gradients of the kept offsets, and the direct paths (anchor, offset, grid scaling)
        float ga0 = 0.f, ga1 = 0.f, ga2 = 0.f, gs[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (lane < k && ((m >> lane) & 1u)) {
            const long long r = row0[v] + __popc(m & ((1u << lane) - 1u));
            const float* sr = s_out + k + 7 * lane;
            const float* co = s_out + 8 * k + cd * lane;
            const float* gsc = scaling + a * 6;
            const float* of = offset + (a * k + lane) * 3;
            const float op = opac_all[v * k + lane];
            s_gout[lane] = v_opacity[r] * (1.0f - op * op);                       // tanh'
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float vx = v_xyz[r * 3 + c];
                if (c == 0) ga0 = vx; else if (c == 1) ga1 = vx; else ga2 = vx;
                g_offset[(a * k + lane) * 3 + c] = vx * gsc[c];
                gs[c] = vx * of[c];
                const float sg = sigmoidf_(sr[c]);
                const float vs = v_scales[r * 3 + c];
                gs[3 + c] = vs * sg;
                s_gout[k + 7 * lane + c] = vs * gsc[3 + c] * sg * (1.0f - sg);
            }
            for (int c = 0; c < cd; ++c) {
                const float vc = v_color[r * cd + c];
                if (color_sigmoid) {
                    const float sc = sigmoidf_(co[c]);
                    s_gout[8 * k + cd * lane + c] = vc * sc * (1.0f - sc);
                } else {
                    s_gout[8 * k + cd * lane + c] = vc;
                }
            }
            // rot = q / max(|q|, eps): v_q -> (v - (v . qn) qn) / |q|
            const float n = fmaxf(sqrtf(sr[3] * sr[3] + sr[4] * sr[4] + sr[5] * sr[5] + sr[6] * sr[6]), 1e-12f);
            const float4 vq = reinterpret_cast<const float4*>(v_quats)[r];
            const float qn[4] = {sr[3] / n, sr[4] / n, sr[5] / n, sr[6] / n};
            const float vv[4] = {vq.x, vq.y, vq.z, vq.w};
            const float dot = vv[0] * qn[0] + vv[1] * qn[1] + vv[2] * qn[2] + vv[3] * qn[3];
#pragma unroll
            for (int q = 0; q < 4; ++q) s_gout[k + 7 * lane + 3 + q] = (vv[q] - dot * qn[q]) / n;
        } else if (lane < k) {
#pragma unroll
            for (int c = 0; c < 3; ++c) g_offset[(a * k + lane) * 3 + c] = 0.f;
        }
        // warp sums of the direct paths
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ga0 += __shfl_xor_sync(0xFFFFFFFFu, ga0, o);
            ga1 += __shfl_xor_sync(0xFFFFFFFFu, ga1, o);
            ga2 += __shfl_xor_sync(0xFFFFFFFFu, ga2, o);
#pragma unroll
            for (int c = 0; c < 6; ++c) gs[c] += __shfl_xor_sync(0xFFFFFFFFu, gs[c], o);
        }
        if (lane == 0) { g_scaling[a * 6] = gs[0]; g_scaling[a * 6 + 1] = gs[1]; g_scaling[a * 6 + 2] = gs[2];
                         g_scaling[a * 6 + 3] = gs[3]; g_scaling[a * 6 + 4] = gs[4]; g_scaling[a * 6 + 5] = gs[5]; }
        __syncwarp();
        // second layer: dW2[o][j] += gout[o] h[j], db2[o] += gout[o], dh[j] = sum_o W2[o][j] gout[o]
        float dh[3] = {0.f, 0.f, 0.f};
        for (int o = 0; o < tot; ++o) {
            const float g = s_gout[o];
            if (g == 0.f) continue;     // warp-uniform (shared-memory broadcast)
            const int mm = mlp_of(o, k);
            atomicAdd(&a_W2[o * DF + lane], g * s_h[mm * DF + lane]);
            const float c = S.W2[o * DF + lane] * g;
            if (mm == 0) dh[0] += c; else if (mm == 1) dh[1] += c; else dh[2] += c;
        }
        for (int o = lane; o < tot; o += 32) {
            const float g = s_gout[o];
            if (g != 0.f) atomicAdd(&a_b2[o], g);
        }
        // first layer: dpre = dh * relu'; dW1[j][i] += dpre[j] x[i]; db1[j] += dpre[j]; dx[i] = sum_j W1[j][i] dpre[j]
#pragma unroll
        for (int mm = 0; mm < 3; ++mm) {
            const float dp = pre[mm] > 0.f ? dh[mm] : 0.f;
            s_dpre[mm * DF + lane] = dp;
            if (dp != 0.f) {
                float* dw = a_W1 + (mm * DF + lane) * DIN;
                for (int i = 0; i < DIN; ++i) atomicAdd(&dw[i], dp * s_x[i]);
                atomicAdd(&a_b1[mm * DF + lane], dp);
            }
        }
        __syncwarp();
        float dx = 0.f, dd = 0.f;          // lane i: d/d feat[i]; lanes 0..2 also d/d dir[lane]
        for (int mm = 0; mm < 3; ++mm) {
#pragma unroll
            for (int j = 0; j < DF; ++j) {
                const float dp = s_dpre[mm * DF + j];
                dx += S.W1[(mm * DF + j) * DIN + lane] * dp;
                if (lane < D.vd) dd += S.W1[(mm * DF + j) * DIN + DF + lane] * dp;
            }
        }
        g_feat[a * DF + lane] = dx;
        // unit direction d = v / |v|: dL/dv = (dd - (dd . d) d) / |v|; v = anchor - cam
        const float dd0 = __shfl_sync(0xFFFFFFFFu, dd, 0), dd1 = __shfl_sync(0xFFFFFFFFu, dd, 1),
                    dd2 = __shfl_sync(0xFFFFFFFFu, dd, 2);
        if (lane == 0) {
            const float dot = dd0 * w.d0 + dd1 * w.d1 + dd2 * w.d2;
            g_anchor[a * 3] = ga0 + (dd0 - dot * w.d0) * w.inv_dist;
            g_anchor[a * 3 + 1] = ga1 + (dd1 - dot * w.d1) * w.inv_dist;
            g_anchor[a * 3 + 2] = ga2 + (dd2 - dot * w.d2) * w.inv_dist;
        }
        __syncwarp();
    }
    __syncthreads();
    // one global add per accumulator and CTA
    for (int i = threadIdx.x; i < 3 * DF * DIN; i += blockDim.x)
        if (a_W1[i] != 0.f) atomicAdd(&G.W1[i / (DF * DIN)][i % (DF * DIN)], a_W1[i]);
    for (int i = threadIdx.x; i < 3 * DF; i += blockDim.x)
        if (a_b1[i] != 0.f) atomicAdd(&G.b1[i / DF][i % DF], a_b1[i]);
    for (int i = threadIdx.x; i < tot * DF; i += blockDim.x) {
        const int o = i / DF, j = i - o * DF, mm = mlp_of(o, k);
        if (a_W2[i] != 0.f) atomicAdd(&G.W2[mm][(o - first_of(mm, k)) * DF + j], a_W2[i]);
    }
    for (int o = threadIdx.x; o < tot; o += blockDim.x) {
        const int mm = mlp_of(o, k);
        if (a_b2[o] != 0.f) atomicAdd(&G.b2[mm][o - first_of(mm, k)], a_b2[o]);
    }
}

int fill_mlp(DecodeMlp& P, const float* const* w) {
    for (int m = 0; m < 3; ++m) {
        P.W1[m] = w[m * 4]; P.b1[m] = w[m * 4 + 1]; P.W2[m] = w[m * 4 + 2]; P.b2[m] = w[m * 4 + 3];
        if (P.W1[m] == nullptr || P.b1[m] == nullptr || P.W2[m] == nullptr || P.b2[m] == nullptr) return HGS_ERR_INVALID_ARG;
    }
    return 0;
}
int decode_grid(long long V) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (V + DWARPS - 1) / DWARPS;
    return (int)(want < 1 ? 1 : (want > 2ll * sms ? 2ll * sms : want));
}

}  // namespace

static bool bad_shape(int feat_dim, int k, int view_dim, int color_dim) {
    return feat_dim != DF || k < 1 || k > DKMAX || (view_dim != 0 && view_dim != 3) || color_dim < 3 || color_dim > DCMAX ||
           color_dim % 3 != 0;
}

HGS_API int hgs_decode_count(const float* const* mlp_host, const float* anchor, const float* feat, const float* cam_center,
                             const long long* vis, long long V, int feat_dim, int k, int view_dim, int color_dim,
                             float* opac_all, int32_t* bits, int32_t* cnt, void* stream) {
    DecodeMlp P;
    if (mlp_host == nullptr || fill_mlp(P, mlp_host) || bad_shape(feat_dim, k, view_dim, color_dim) || V < 0)
        return HGS_ERR_INVALID_ARG;
    if (V == 0) return 0;
    const DShape D{k, color_dim, view_dim};
    const int smem = (smem_w_floats(D.tot(), D.din()) + DWARPS * (3 * DF + D.tot())) * (int)sizeof(float);
    if (smem > DSMEM_MAX) return HGS_ERR_TOO_LARGE;
    cudaError_t e = cudaFuncSetAttribute(decode_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    decode_count_kernel<<<decode_grid(V), DWARPS * 32, smem, (cudaStream_t)stream>>>(P, anchor, feat, cam_center, vis, V, D,
                                                                                      opac_all, bits, cnt);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_decode_fwd(const float* const* mlp_host, const float* anchor, const float* feat, const float* offset,
                           const float* scaling, const float* cam_center, const long long* vis, long long V, int feat_dim,
                           int k, int view_dim, int color_dim, int color_sigmoid, const float* opac_all,
                           const int32_t* bits, const long long* row0, float* xyz, float* color, float* opacity,
                           float* scales, float* quats, void* stream) {
    DecodeMlp P;
    if (mlp_host == nullptr || fill_mlp(P, mlp_host) || bad_shape(feat_dim, k, view_dim, color_dim) || V < 0)
        return HGS_ERR_INVALID_ARG;
    if (V == 0) return 0;
    if (reinterpret_cast<size_t>(quats) & 15) return HGS_ERR_INVALID_ARG;
    const DShape D{k, color_dim, view_dim};
    const int smem = (smem_w_floats(D.tot(), D.din()) + DWARPS * (3 * DF + D.tot())) * (int)sizeof(float);
    if (smem > DSMEM_MAX) return HGS_ERR_TOO_LARGE;
    cudaError_t e = cudaFuncSetAttribute(decode_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    decode_fwd_kernel<<<decode_grid(V), DWARPS * 32, smem, (cudaStream_t)stream>>>(
        P, anchor, feat, offset, scaling, cam_center, vis, V, D, color_sigmoid, opac_all, bits, row0, xyz, color, opacity,
        scales, quats);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_decode_bwd(const float* const* mlp_host, float* const* mlp_grad_host, const float* anchor, const float* feat,
                           const float* offset, const float* scaling, const float* cam_center, const long long* vis,
                           long long V, int feat_dim, int k, int view_dim, int color_dim, int color_sigmoid,
                           const float* opac_all, const int32_t* bits, const long long* row0, const float* v_xyz,
                           const float* v_color, const float* v_opacity, const float* v_scales, const float* v_quats,
                           float* g_anchor, float* g_feat, float* g_offset, float* g_scaling, void* stream) {
    DecodeMlp P;
    if (mlp_host == nullptr || mlp_grad_host == nullptr || fill_mlp(P, mlp_host) ||
        bad_shape(feat_dim, k, view_dim, color_dim) || V < 0)
        return HGS_ERR_INVALID_ARG;
    DecodeMlpGrad G;
    for (int m = 0; m < 3; ++m) {
        G.W1[m] = mlp_grad_host[m * 4]; G.b1[m] = mlp_grad_host[m * 4 + 1];
        G.W2[m] = mlp_grad_host[m * 4 + 2]; G.b2[m] = mlp_grad_host[m * 4 + 3];
        if (G.W1[m] == nullptr || G.b1[m] == nullptr || G.W2[m] == nullptr || G.b2[m] == nullptr) return HGS_ERR_INVALID_ARG;
    }
    if (V == 0) return 0;
    if (reinterpret_cast<size_t>(v_quats) & 15) return HGS_ERR_INVALID_ARG;
    const DShape D{k, color_dim, view_dim};
    const int tot = D.tot(), din = D.din();
    const int n_acc = 3 * DF * din + 3 * DF + tot * DF + tot;
    const int smem = (smem_w_floats(tot, din) + n_acc + DWARPS * (3 * DF + 2 * tot + DINMAX + 3 * DF + 8)) * (int)sizeof(float);
    if (smem > DSMEM_MAX) return HGS_ERR_TOO_LARGE;
    cudaError_t e = cudaFuncSetAttribute(decode_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    decode_bwd_kernel<<<decode_grid(V), DWARPS * 32, smem, (cudaStream_t)stream>>>(
        P, G, anchor, feat, offset, scaling, cam_center, vis, V, D, color_sigmoid, opac_all, bits, row0, v_xyz, v_color,
        v_opacity, v_scales, v_quats, g_anchor, g_feat, g_offset, g_scaling);
    HGS_LAUNCH_CHECK();
    return 0;
}
