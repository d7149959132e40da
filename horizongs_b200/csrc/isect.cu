// Stages a8-a10: tile intersection, tile|depth ordering, per-tile ranges.  Integer work, bit-exact
// against oracle/gsplat_oracle.py::isect_tiles / isect_offset_encode (= gsplat isect_tiles(sort=True)
// + isect_offset_encode as reached inside gsplat.rasterization*, reference render.py:40,62).
//
// B200-first ordering.  gsplat emits I (key,value) pairs Gaussian-major and runs a generic 64-bit LSD radix
// sort over 32 + tile_bits + cam_bits key bits (6 passes x 24 B r/w per pair through HBM).  The same order is
// produced here by BINNING, with every pair crossing memory twice and the sort itself on the SM:
//   1. bin_count:   one pass over tiles_per_gauss: ordered compaction of the Gaussians that touch a tile
//                   (single-pass scan, decoupled look-back) + histogram of the (camera, tile) bins;
//   2. tile_scan:   exclusive scan of the histogram = the per-tile ranges (isect_offsets, the a10 output);
//   3. bin_scatter: every visible Gaussian drops (depth bits << 32 | flat index) into its tiles' ranges
//                   (slot by atomic cursor: arrival order is arbitrary);
//   4. tile_sort:   one CTA per (camera, tile) sorts its range by the 64-bit (depth, flat index) key -- a
//                   bitonic network held in registers (blocked, 1..16 keys per thread), exchanged by warp
//                   shuffles and, for the widest strides only, through shared memory -- and writes
//                   isect_ids / flatten_ids.  Ranges longer than 4096 are sorted in 4096-key chunks and merged
//                   by the same network with the widest strides through (L2-resident) global memory.
// Within a tile the keys (depth bits, flat index) are unique, so the sorted order does not depend on the
// arrival order: it is the stable ascending sort of gsplat's cam|tile|depth keys over the Gaussian-major
// emission (ties on depth are broken by ascending flat index in both).
// Roofline: HBM for steps 1-3 (16 B per Gaussian + 8 B per pair), shuffle / issue for step 4.
#include "hgs_common.cuh"
#include "hgs_constants.cuh"
#include "bin_common.cuh"
#include "../../include/hgs_raster.h"

namespace {

using namespace hgs_bin;

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int SCAN_THREADS = 1024;

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
static int n_bits_of(long long n) {  // floor(log2(n)) + 1 for n >= 1
    int b = 0;
    while (n > 0) { ++b; n >>= 1; }
    return b;
}

// block-wide exclusive scan of one value per thread (RS_THREADS threads); returns the exclusive prefix and
// the block total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp /*[RS_WARPS]*/, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
        const uint32_t x = s_warp[w];
        if (w < warp) wbase += x;
        tot += x;
    }
    __syncthreads();
    total = tot;
    return wbase + incl - v;
}

// ---------------------------------------------------------------------------------------------
// large exclusive scan (i32 in, i32 out, i64 total): reduce / scan-of-sums / scan.
// n is read from device memory (n_dev) and bounded by cap (grid size).
// ---------------------------------------------------------------------------------------------
constexpr int LS_THREADS = 256;
constexpr int LS_ITEMS = 16;
constexpr int LS_TILE = LS_THREADS * LS_ITEMS;

__device__ __forceinline__ long long block_reduce_ll(long long v, long long* s_tmp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = v;
    __syncthreads();
    long long r = 0;
    for (int w = 0; w < LS_THREADS / 32; ++w) r += s_tmp[w];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(LS_THREADS) ls_reduce_kernel(const int32_t* __restrict__ in,
                                                               const long long* __restrict__ n_dev,
                                                               long long* __restrict__ block_sums) {
    __shared__ long long s_tmp[LS_THREADS / 32];
    const long long n = *n_dev;
    const long long base = (long long)blockIdx.x * LS_TILE;
    long long sum = 0;
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        long long idx = base + i * LS_THREADS + threadIdx.x;
        if (idx < n) sum += in[idx];
    }
    long long tot = block_reduce_ll(sum, s_tmp);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) ls_scan_sums_kernel(long long* __restrict__ block_sums, int nb,
                                                                    long long* __restrict__ total) {
    __shared__ long long s_sum[SCAN_THREADS];
    const int per = (nb + SCAN_THREADS - 1) / SCAN_THREADS;
    const int b = threadIdx.x * per;
    int e = b + per;
    if (e > nb) e = nb;
    long long sum = 0;
    for (int i = b; i < e; ++i) sum += block_sums[i];
    s_sum[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < SCAN_THREADS; off <<= 1) {
        long long v = threadIdx.x >= off ? s_sum[threadIdx.x - off] : 0;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    long long run = s_sum[threadIdx.x] - sum;
    for (int i = b; i < e; ++i) {
        long long v = block_sums[i];
        block_sums[i] = run;
        run += v;
    }
    if (threadIdx.x == SCAN_THREADS - 1) total[0] = s_sum[SCAN_THREADS - 1];
}

__global__ void __launch_bounds__(LS_THREADS) ls_scan_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                             const long long* __restrict__ n_dev,
                                                             const long long* __restrict__ block_sums) {
    // thread t owns LS_ITEMS consecutive items (blocked arrangement)
    __shared__ long long s_warp[LS_THREADS / 32];
    const long long n = *n_dev;
    if ((long long)blockIdx.x * LS_TILE >= n) return;
    const long long base = (long long)blockIdx.x * LS_TILE + (long long)threadIdx.x * LS_ITEMS;
    int32_t v[LS_ITEMS];
    long long sum = 0;
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        long long idx = base + i;
        v[i] = idx < n ? in[idx] : 0;
        sum += v[i];
    }
    long long incl = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    long long wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += s_warp[w];
    long long run = block_sums[blockIdx.x] + wbase + incl - sum;
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        long long idx = base + i;
        if (idx < n) out[idx] = (int32_t)run;
        run += v[i];
    }
}

// exclusive scan of in[0..*n_dev) (n <= cap); total -> total_dev[0].  temp: ceil(cap/LS_TILE) long longs
static int large_scan(const int32_t* in, int32_t* out, long long* total_dev, const long long* n_dev, long long cap,
                      void* temp, size_t temp_bytes, cudaStream_t st) {
    if (cap <= 0) return (int)cudaMemsetAsync(total_dev, 0, sizeof(long long), st);
    const int nb = hgs_ceil_div(cap, LS_TILE);
    if (temp_bytes < (size_t)nb * sizeof(long long)) return HGS_ERR_WORKSPACE;
    long long* sums = (long long*)temp;
    ls_reduce_kernel<<<nb, LS_THREADS, 0, st>>>(in, n_dev, sums);
    HGS_LAUNCH_CHECK();
    ls_scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(sums, nb, total_dev);
    HGS_LAUNCH_CHECK();
    ls_scan_kernel<<<nb, LS_THREADS, 0, st>>>(in, out, n_dev, sums);
    HGS_LAUNCH_CHECK();
    return 0;
}

__global__ void set_ll_kernel(long long* p, long long v) { *p = v; }

// ---------------------------------------------------------------------------------------------
// a8 kernels
// ---------------------------------------------------------------------------------------------
__global__ void isect_count_kernel(const float* __restrict__ means2d, const int32_t* __restrict__ radii, long long CN,
                                   int tile_size, int tile_w, int tile_h, int32_t* __restrict__ tiles_per_gauss) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CN) return;
    int r = radii[i];
    int cnt = 0;
    if (r > 0) {
        float2 m = reinterpret_cast<const float2*>(means2d)[i];
        int x0, y0, x1, y1;
        hgs_tile_bbox(m.x, m.y, (float)r, (float)tile_size, tile_w, tile_h, x0, y0, x1, y1);
        cnt = (y1 - y0) * (x1 - x0);
    }
    tiles_per_gauss[i] = cnt;
}

// Gaussian-major emission of full 64-bit keys (gsplat sort=False layout)
__global__ void isect_emit_kernel(const float* __restrict__ means2d, const int32_t* __restrict__ radii,
                                  const float* __restrict__ depths, const int32_t* __restrict__ cum, long long CN,
                                  int N, int tile_size, int tile_w, int tile_h, int tile_bits,
                                  long long* __restrict__ isect_ids, int32_t* __restrict__ flatten_ids) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CN) return;
    int r = radii[i];
    if (r <= 0) return;
    float2 m = reinterpret_cast<const float2*>(means2d)[i];
    int x0, y0, x1, y1;
    hgs_tile_bbox(m.x, m.y, (float)r, (float)tile_size, tile_w, tile_h, x0, y0, x1, y1);
    const long long cam_enc = (i / N) << (32 + tile_bits);
    const long long depth_enc = (long long)__float_as_int(depths[i]);
    long long cur = cum[i];
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
            long long tile_id = (long long)y * tile_w + x;
            isect_ids[cur] = cam_enc | (tile_id << 32) | depth_enc;
            flatten_ids[cur] = (int32_t)i;
            ++cur;
        }
}

// =============================================================================================
// sorted path by binning (see the header comment)
// =============================================================================================
constexpr int SORT_CHUNK = 2048;                 // keys one CTA sorts in registers (8 per thread)

// ---- phase 1a: ordered compaction of the Gaussians with tiles + super-tile histogram --------------------
__global__ void __launch_bounds__(RS_THREADS) bin_count_kernel(
    const float* __restrict__ means2d, const int32_t* __restrict__ radii, const float* __restrict__ depths,
    const int32_t* __restrict__ tiles_per_gauss, long long CN, BinGeom G, int nblk,
    unsigned long long* __restrict__ flags, uint32_t* __restrict__ ticket, uint32_t* __restrict__ super_count,
    int32_t* __restrict__ visible_ids, VisRec* __restrict__ vrec, long long* __restrict__ counts_dev) {
    __shared__ uint32_t s_warp[RS_WARPS];
    __shared__ uint32_t s_bid, s_nbig;
    __shared__ unsigned long long s_excl, s_isect;
    __shared__ uint32_t s_list[CP_TILE];
    __shared__ uint2 s_box[CP_TILE];
    __shared__ uint32_t s_big[CP_TILE];
    if (threadIdx.x == 0) {
        s_bid = atomicAdd(ticket, 1u);   // logical block id = start order: every predecessor is already running
        s_nbig = 0;
        s_isect = 0;
    }
    __syncthreads();
    const uint32_t bid = s_bid;
    constexpr int PER = CP_TILE / RS_THREADS;  // 4 consecutive elements per thread (order preserving)
    const long long first = (long long)bid * CP_TILE + (long long)threadIdx.x * PER;
    int tcount[PER];
    if (first + PER <= CN && (reinterpret_cast<uintptr_t>(tiles_per_gauss) & 15) == 0) {
        const int4 q = *reinterpret_cast<const int4*>(tiles_per_gauss + first);
        tcount[0] = q.x; tcount[1] = q.y; tcount[2] = q.z; tcount[3] = q.w;
    } else {
#pragma unroll
        for (int i = 0; i < PER; ++i) tcount[i] = first + i < CN ? tiles_per_gauss[first + i] : 0;
    }
    uint32_t cnt = 0;
    unsigned long long isects = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        cnt += tcount[i] > 0 ? 1u : 0u;
        isects += tcount[i] > 0 ? (unsigned long long)tcount[i] : 0ull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) isects += __shfl_xor_sync(0xFFFFFFFFu, isects, o);
    if ((threadIdx.x & 31) == 0 && isects) atomicAdd(&s_isect, isects);
    uint32_t tot;
    uint32_t pos = block_excl_scan(cnt, s_warp, tot);
#pragma unroll
    for (int i = 0; i < PER; ++i)
        if (tcount[i] > 0) s_list[pos++] = (uint32_t)(first + i);
    // publish the block total at once; the prefix of the earlier blocks is only needed for the final writes
    if (threadIdx.x == 0) {
        volatile unsigned long long* vf = flags;
        vf[bid] = (bid == 0 ? LB_PREFIX : LB_AGG) | (unsigned long long)tot;
    }
    __syncthreads();
    // the block's visible Gaussians: tile boxes, histogram of the super-tiles they touch
    const int n_super = G.stw * G.sth;
    for (uint32_t i = threadIdx.x; i < tot; i += RS_THREADS) {
        const uint32_t g = s_list[i];
        int x0, y0, x1, y1;
        tile_box(G, means2d, radii, g, x0, y0, x1, y1);
        s_box[i] = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)x1 | ((uint32_t)y1 << 16));
        const int sx0 = x0 / ST, sy0 = y0 / ST, sx1 = (x1 - 1) / ST, sy1 = (y1 - 1) / ST;
        if ((sx1 - sx0 + 1) * (sy1 - sy0 + 1) > BIG_AREA) {
            s_big[atomicAdd(&s_nbig, 1u)] = i;
            continue;
        }
        uint32_t* srow = super_count + (((long long)(g / (uint32_t)G.N) * n_super) * SUB + (g % SUB)) * PAD;
        for (int y = sy0; y <= sy1; ++y)
            for (int x = sx0; x <= sx1; ++x) atomicAdd(srow + (long long)(y * G.stw + x) * (SUB * PAD), 1u);
    }
    __syncthreads();
    const uint32_t nbig = s_nbig;
    for (uint32_t b = 0; b < nbig; ++b) {
        const uint32_t i = s_big[b], g = s_list[i];
        const uint2 bx = s_box[i];
        const int sx0 = (int)(bx.x & 0xFFFFu) / ST, sy0 = (int)(bx.x >> 16) / ST;
        const int sw = ((int)(bx.y & 0xFFFFu) - 1) / ST - sx0 + 1, sarea = sw * (((int)(bx.y >> 16) - 1) / ST - sy0 + 1);
        uint32_t* srow = super_count + (((long long)(g / (uint32_t)G.N) * n_super) * SUB + (g % SUB)) * PAD;
        for (int k = threadIdx.x; k < sarea; k += RS_THREADS)
            atomicAdd(srow + (long long)((sy0 + k / sw) * G.stw + sx0 + k % sw) * (SUB * PAD), 1u);
    }
    if (threadIdx.x < 32) {
        // decoupled look-back (warp 0): exclusive prefix of the block totals of all earlier blocks
        const int lane = threadIdx.x;
        volatile unsigned long long* vf = flags;
        unsigned long long excl = 0;
        if (bid > 0) {
            long long j = (long long)bid - 1;
            while (true) {
                const long long idx = j - lane;
                unsigned long long v;
                do {
                    v = idx >= 0 ? vf[idx] : LB_PREFIX;
                } while (__any_sync(0xFFFFFFFFu, (v >> 62) == 0));
                const unsigned m = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
                const int stop = m ? __ffs(m) - 1 : 31;      // nearest predecessor that already knows its prefix
                unsigned long long c = lane <= stop ? (v & LB_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
                excl += c;
                if (m) break;
                j -= 32;
            }
            if (lane == 0) vf[bid] = LB_PREFIX | (excl + tot);
        }
        if (lane == 0) {
            s_excl = excl;
            if ((int)bid == nblk - 1) counts_dev[0] = (long long)(excl + tot);
            if (s_isect) atomicAdd(reinterpret_cast<unsigned long long*>(counts_dev + 1), s_isect);
        }
    }
    __syncthreads();
    const long long out0 = (long long)s_excl;
    for (uint32_t i = threadIdx.x; i < tot; i += RS_THREADS) {
        const uint32_t g = s_list[i];
        visible_ids[out0 + i] = (int32_t)g;
        const uint2 bx = s_box[i];
        VisRec r;
        r.g = g; r.depth_bits = __float_as_uint(depths[g]); r.xy0 = bx.x; r.xy1 = bx.y;
        vrec[out0 + i] = r;
    }
}


// ---- phase 1b: exclusive scan of the (padded) sub-bin counters, one CTA: soff[i], total -> counts_dev[2] ----
__global__ void __launch_bounds__(SCAN_THREADS) hist_scan_kernel(const uint32_t* __restrict__ hist_padded, int n,
                                                                 int32_t* __restrict__ soff,
                                                                 long long* __restrict__ counts_dev) {
    __shared__ long long s_warp[SCAN_THREADS / 32];
    // gather the counters into soff (independent strided loads, many in flight), then scan in place
#pragma unroll 8
    for (int i = threadIdx.x; i < n; i += SCAN_THREADS) soff[i] = (int32_t)__ldcg(hist_padded + (long long)i * PAD);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
    const int b = min(threadIdx.x * per, n), e = min(b + per, n);
    long long sum = 0;
    for (int i = b; i < e; ++i) sum += (uint32_t)soff[i];
    long long incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const long long w = s_warp[lane];
        long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= o) wi += t;
        }
        s_warp[lane] = wi - w;      // exclusive prefix of the warp totals
        if (lane == 31) counts_dev[2] = wi;
    }
    __syncthreads();
    long long run = s_warp[warp] + incl - sum;
    for (int i = b; i < e; ++i) {
        const uint32_t c = (uint32_t)soff[i];
        soff[i] = (int32_t)run;     // totals >= 2^31 are rejected by the caller before the ranges are used
        run += c;
    }
}
static_assert(SCAN_THREADS == 1024, "hist_scan_kernel scans 32 warp totals with one warp");

// ---- phase 2a: every visible Gaussian drops one key per super-tile it touches into that super-tile's range:
// key = depth bits << 32 | flat index << 4 | mask of the super-tile's ST x ST tiles it touches (bit ly * ST + lx)
__device__ __forceinline__ unsigned long long make_key(const VisRec& r, int sx, int sy) {
    const int x0 = (int)(r.xy0 & 0xFFFFu), y0 = (int)(r.xy0 >> 16), x1 = (int)(r.xy1 & 0xFFFFu), y1 = (int)(r.xy1 >> 16);
    const int lx0 = max(x0 - sx * ST, 0), lx1 = min(x1 - sx * ST, ST);
    const int ly0 = max(y0 - sy * ST, 0), ly1 = min(y1 - sy * ST, ST);
    const uint32_t cols = ((1u << lx1) - 1u) & ~((1u << lx0) - 1u);
    uint32_t mask = 0;
    for (int ly = ly0; ly < ly1; ++ly) mask |= cols << (ly * ST);
    return ((unsigned long long)r.depth_bits << 32) | ((unsigned long long)r.g << (32 - ID_BITS)) | mask;
}

__global__ void __launch_bounds__(RS_THREADS) bin_scatter_kernel(
    const VisRec* __restrict__ vrec, const long long* __restrict__ counts_dev, BinGeom G,
    const int32_t* __restrict__ soff, uint32_t* __restrict__ cursor, unsigned long long* __restrict__ bucket,
    long long cap_isects, long long cap_super) {
    __shared__ uint32_t s_nbig;
    __shared__ uint32_t s_big[RS_THREADS];
    const long long n_vis = counts_dev[0];
    if (counts_dev[1] > cap_isects || counts_dev[2] > cap_super) return;   // outputs too small: the caller re-runs
    if ((long long)blockIdx.x * RS_THREADS >= n_vis) return;
    if (threadIdx.x == 0) s_nbig = 0;
    __syncthreads();
    const long long j = (long long)blockIdx.x * RS_THREADS + threadIdx.x;
    const int n_super = G.stw * G.sth;
    if (j < n_vis) {
        const VisRec r = vrec[j];
        const int sx0 = (int)(r.xy0 & 0xFFFFu) / ST, sy0 = (int)(r.xy0 >> 16) / ST;
        const int sx1 = ((int)(r.xy1 & 0xFFFFu) - 1) / ST, sy1 = ((int)(r.xy1 >> 16) - 1) / ST;
        if ((sx1 - sx0 + 1) * (sy1 - sy0 + 1) > BIG_AREA) {
            s_big[atomicAdd(&s_nbig, 1u)] = threadIdx.x;
        } else {
            const long long base = ((long long)(r.g / (uint32_t)G.N) * n_super) * SUB + (r.g % SUB);
            for (int y = sy0; y <= sy1; ++y)
                for (int x = sx0; x <= sx1; ++x) {
                    const long long t = base + (long long)(y * G.stw + x) * SUB;
                    const uint32_t slot = atomicAdd(&cursor[t * PAD], 1u);
                    bucket[(long long)soff[t] + slot] = make_key(r, x, y);
                }
        }
    }
    __syncthreads();
    const uint32_t nbig = s_nbig;
    for (uint32_t b = 0; b < nbig; ++b) {
        const VisRec r = vrec[(long long)blockIdx.x * RS_THREADS + s_big[b]];
        const int sx0 = (int)(r.xy0 & 0xFFFFu) / ST, sy0 = (int)(r.xy0 >> 16) / ST;
        const int sw = ((int)(r.xy1 & 0xFFFFu) - 1) / ST - sx0 + 1, sarea = sw * (((int)(r.xy1 >> 16) - 1) / ST - sy0 + 1);
        const long long base = ((long long)(r.g / (uint32_t)G.N) * n_super) * SUB + (r.g % SUB);
        for (int k = threadIdx.x; k < sarea; k += RS_THREADS) {
            const int x = sx0 + k % sw, y = sy0 + k / sw;
            const long long t = base + (long long)(y * G.stw + x) * SUB;
            const uint32_t slot = atomicAdd(&cursor[t * PAD], 1u);
            bucket[(long long)soff[t] + slot] = make_key(r, x, y);
        }
    }
}

// ---- phase 2b: per-super-tile sort ------------------------------------------------------------------
// Ascending bitonic network in its uniform-direction form (first stage of every merge pairs i with
// i ^ (k - 1), the others i with i ^ j; every exchange puts the minimum at the lower index, so +inf padding at
// the top never moves).  Blocked layout: thread t holds keys [t * IPT, (t + 1) * IPT).
// The network is written once for two key types: the 64-bit keys, and 32-bit keys (see sort_small32).
typedef unsigned long long u64;
__device__ __forceinline__ u64 shfl_xor_key(u64 v, int m) {
    const unsigned lo = __shfl_xor_sync(0xFFFFFFFFu, (unsigned)v, m);
    const unsigned hi = __shfl_xor_sync(0xFFFFFFFFu, (unsigned)(v >> 32), m);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t shfl_xor_key(uint32_t v, int m) { return __shfl_xor_sync(0xFFFFFFFFu, v, m); }
template <typename K>
__device__ __forceinline__ void cmp_exch(K& a, K& b) {
    const K lo = a < b ? a : b, hi = a < b ? b : a;
    a = lo;
    b = hi;
}
template <typename K>
__device__ __forceinline__ K pick(bool lower, K mine, K other) {
    const K lo = mine < other ? mine : other, hi = mine < other ? other : mine;
    return lower ? lo : hi;
}

// stages j = j_hi, j_hi / 2, ..., 1 (plain i ^ j exchanges); s_buf: RS_THREADS * IPT keys
// `active`: the warp holds at least one real key (a warp of +inf padding never changes: it only takes part in the
// shared-memory stages, where its partners read it)
template <int IPT, typename K>
__device__ __forceinline__ void merge_xor_stages(K (&v)[IPT], int j_hi, K* s_buf, bool active = true) {
    const int t = threadIdx.x;
    int j = j_hi;
    for (; j >= 32 * IPT; j >>= 1) {          // partner in another warp: through shared memory ([a][t] layout)
        const int m = j / IPT;
#pragma unroll
        for (int a = 0; a < IPT; ++a) s_buf[a * RS_THREADS + t] = v[a];
        __syncthreads();
        const bool lower = (t & m) == 0;
#pragma unroll
        for (int a = 0; a < IPT; ++a) v[a] = pick(lower, v[a], s_buf[a * RS_THREADS + (t ^ m)]);
        __syncthreads();
    }
    if (!active) return;
    for (; j >= IPT; j >>= 1) {               // partner in this warp: shuffles
        const int m = j / IPT;
        const bool lower = (t & m) == 0;
#pragma unroll
        for (int a = 0; a < IPT; ++a) v[a] = pick(lower, v[a], shfl_xor_key(v[a], m));
    }
#pragma unroll
    for (int jj = IPT / 2; jj > 0; jj >>= 1) {  // partner in this thread
        if (jj <= j) {
#pragma unroll
            for (int a = 0; a < IPT; ++a)
                if ((a & jj) == 0) cmp_exch(v[a], v[a | jj]);
        }
    }
}

// full sort of the CTA's RS_THREADS * IPT keys; k_max = smallest power of two >= the number of real keys
// (the +inf padding above it is already in place)
template <int IPT, typename K>
__device__ __forceinline__ void cta_sort(K (&v)[IPT], int k_max, K* s_buf, bool active = true) {
    const int t = threadIdx.x;
#pragma unroll
    for (int k = 2; k <= IPT; k <<= 1) {      // merges inside the thread
        if (!active) break;
#pragma unroll
        for (int a = 0; a < IPT; ++a)
            if ((a & (k >> 1)) == 0) cmp_exch(v[a], v[a ^ (k - 1)]);
#pragma unroll
        for (int jj = k >> 2; jj > 0; jj >>= 1) {
#pragma unroll
            for (int a = 0; a < IPT; ++a)
                if ((a & jj) == 0) cmp_exch(v[a], v[a | jj]);
        }
    }
    for (int k = 2 * IPT; k <= k_max; k <<= 1) {
        // mirror stage: key (t, a) meets key (t ^ (k / IPT - 1), IPT - 1 - a)
        const int m = k / IPT - 1;
        const bool lower = (t & (k / IPT / 2)) == 0;
        K o[IPT];
        if (m >= 32) {
#pragma unroll
            for (int a = 0; a < IPT; ++a) s_buf[a * RS_THREADS + t] = v[a];
            __syncthreads();
#pragma unroll
            for (int a = 0; a < IPT; ++a) o[a] = s_buf[(IPT - 1 - a) * RS_THREADS + (t ^ m)];
            __syncthreads();
#pragma unroll
            for (int a = 0; a < IPT; ++a) v[a] = pick(lower, v[a], o[a]);
        } else if (active) {
#pragma unroll
            for (int a = 0; a < IPT; ++a) o[a] = shfl_xor_key(v[IPT - 1 - a], m);
#pragma unroll
            for (int a = 0; a < IPT; ++a) v[a] = pick(lower, v[a], o[a]);
        }
        merge_xor_stages<IPT, K>(v, k >> 2, s_buf, active);
    }
}

// keys [0, n) of buf -> registers (blocked), coalesced through shared memory; padding = +inf
template <int IPT>
__device__ __forceinline__ void load_blocked(unsigned long long (&v)[IPT], const unsigned long long* buf, int n,
                                             unsigned long long* s_buf) {
    const int t = threadIdx.x;
#pragma unroll
    for (int a = 0; a < IPT; ++a) {
        const int i = a * RS_THREADS + t;
        s_buf[i] = i < n ? buf[i] : KEY_INF;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < IPT; ++a) v[a] = s_buf[t * IPT + a];
    __syncthreads();
}
template <int IPT>
__device__ __forceinline__ void stage_blocked(const unsigned long long (&v)[IPT], unsigned long long* s_buf) {
#pragma unroll
    for (int a = 0; a < IPT; ++a) s_buf[threadIdx.x * IPT + a] = v[a];
    __syncthreads();
}

// sort n <= RS_THREADS * IPT keys of `bucket`; the sorted keys end up in s_buf[0, n)
template <int IPT>
__device__ __forceinline__ void sort_small(const unsigned long long* bucket, int n, unsigned long long* s_buf) {
    unsigned long long v[IPT];
    load_blocked<IPT>(v, bucket, n, s_buf);
    int k_max = 2;
    while (k_max < n) k_max <<= 1;
    cta_sort<IPT, u64>(v, k_max, s_buf, (threadIdx.x & ~31) * IPT < n);
    stage_blocked<IPT>(v, s_buf);
}

// The same with a 32-bit network (a third of the work per stage).  Within a super-tile the depth bits span a small
// range, so the keys are first ordered by key32 = ((depth bits - min) >> s) << B | arrival index -- B = bits of the
// padded length, s = whatever makes the depth part fit -- which is unique and sorts them by depth up to the dropped
// bits; a few odd-even exchange passes on the full 64-bit keys then put neighbours that agree in the kept depth bits
// (and exact depth ties, which the arrival index ordered arbitrarily) into (depth, flat index) order.  A pathological
// range (more than MAX_FIX rounds needed) falls back to the 64-bit network.  The sorted keys end up in s_buf[0, n).
constexpr int MAX_FIX = 24;
template <int IPT>
__device__ __forceinline__ void sort_small32(const u64* bucket, int n, u64* s_buf, uint32_t* s_k32, uint32_t* s_red) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    uint32_t dmin = 0xFFFFFFFFu, dmax = 0u;
    for (int i = t; i < n; i += RS_THREADS) {
        const u64 key = bucket[i];
        s_buf[i] = key;
        const uint32_t d = (uint32_t)(key >> 32);
        dmin = min(dmin, d);
        dmax = max(dmax, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmin = min(dmin, __shfl_xor_sync(0xFFFFFFFFu, dmin, o));
        dmax = max(dmax, __shfl_xor_sync(0xFFFFFFFFu, dmax, o));
    }
    if (lane == 0) { s_red[warp] = dmin; s_red[RS_WARPS + warp] = dmax; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) { dmin = min(dmin, s_red[w]); dmax = max(dmax, s_red[RS_WARPS + w]); }
    int k_max = 2;
    while (k_max < n) k_max <<= 1;
    const int B = 31 - __clz(k_max);                       // k_max = 2^B >= n: arrival indices fit in B bits
    const int range_bits = 32 - __clz(dmax - dmin);         // 0 when all depths are equal
    const int sh = max(0, range_bits - (32 - B));
    uint32_t v[IPT];
#pragma unroll
    for (int a = 0; a < IPT; ++a) {
        const int i = t * IPT + a;
        v[a] = i < n ? ((((uint32_t)(s_buf[i] >> 32) - dmin) >> sh) << B) | (uint32_t)i : 0xFFFFFFFFu;
    }
    cta_sort<IPT, uint32_t>(v, k_max, s_k32, (t & ~31) * IPT < n);
    // bring the 64-bit keys into that order (gather into registers, then store blocked)
    u64 w64[IPT];
#pragma unroll
    for (int a = 0; a < IPT; ++a) w64[a] = t * IPT + a < n ? s_buf[v[a] & ((1u << B) - 1u)] : KEY_INF;
    __syncthreads();
#pragma unroll
    for (int a = 0; a < IPT; ++a)
        if (t * IPT + a < n) s_buf[t * IPT + a] = w64[a];
    __syncthreads();
    // odd-even exchange passes on the full keys until nothing moves
    int round = 0;
    for (; round < MAX_FIX; ++round) {
        bool moved = false;
        for (int j = 2 * t; j + 1 < n; j += 2 * RS_THREADS) {
            const u64 a = s_buf[j], b = s_buf[j + 1];
            if (b < a) { s_buf[j] = b; s_buf[j + 1] = a; moved = true; }
        }
        __syncthreads();
        for (int j = 2 * t + 1; j + 1 < n; j += 2 * RS_THREADS) {
            const u64 a = s_buf[j], b = s_buf[j + 1];
            if (b < a) { s_buf[j] = b; s_buf[j + 1] = a; moved = true; }
        }
        if (!__syncthreads_or(moved)) break;
    }
    if (round == MAX_FIX) {
        // long runs of equal (kept) depth bits: finish with the 64-bit network
        u64 k64[IPT];
#pragma unroll
        for (int a = 0; a < IPT; ++a) k64[a] = t * IPT + a < n ? s_buf[t * IPT + a] : KEY_INF;
        __syncthreads();
        cta_sort<IPT, u64>(k64, k_max, s_buf, (t & ~31) * IPT < n);
        stage_blocked<IPT>(k64, s_buf);
    }
}

// one global-memory exchange stage of the big merge: pairs (i, p) with bit `h` of i clear,
// p = mirror ? i ^ (2h - 1) : i | h
__device__ __forceinline__ void global_stage(unsigned long long* buf, int n, int h, bool mirror) {
    for (int q = threadIdx.x;; q += RS_THREADS) {
        const int i = ((q & ~(h - 1)) << 1) | (q & (h - 1));
        if (i >= n) break;
        const int p = mirror ? (i ^ (2 * h - 1)) : (i | h);
        if (p < n) {
            unsigned long long a = buf[i], b = buf[p];
            if (b < a) {
                buf[i] = b;
                buf[p] = a;
            }
        }
    }
    __syncthreads();
}

// ranges longer than SORT_CHUNK: chunk sorts (32-bit network, as above) + merges whose wide strides go through
// (L2-resident) global memory; sorted in place
__device__ void sort_big(unsigned long long* bucket, int n, unsigned long long* s_buf, uint32_t* s_k32, uint32_t* s_red) {
    constexpr int IPT = SORT_CHUNK / RS_THREADS;
    const int n_chunks = (n + SORT_CHUNK - 1) / SORT_CHUNK;
    unsigned long long v[IPT];
    for (int c = 0; c < n_chunks; ++c) {
        unsigned long long* cb = bucket + (long long)c * SORT_CHUNK;
        const int cn = min(SORT_CHUNK, n - c * SORT_CHUNK);
        sort_small32<IPT>(cb, cn, s_buf, s_k32, s_red);
        __syncthreads();
        for (int i = threadIdx.x; i < cn; i += RS_THREADS) cb[i] = s_buf[i];
        __syncthreads();
    }
    int P = SORT_CHUNK;
    while (P < n) P <<= 1;
    for (int k = 2 * SORT_CHUNK; k <= P; k <<= 1) {
        global_stage(bucket, n, k >> 1, true);
        for (int j = k >> 2; j >= SORT_CHUNK; j >>= 1) global_stage(bucket, n, j, false);
        for (int c = 0; c < n_chunks; ++c) {
            unsigned long long* cb = bucket + (long long)c * SORT_CHUNK;
            const int cn = min(SORT_CHUNK, n - c * SORT_CHUNK);
            load_blocked<IPT>(v, cb, cn, s_buf);
            merge_xor_stages<IPT, u64>(v, SORT_CHUNK >> 1, s_buf);
            stage_blocked<IPT>(v, s_buf);
            for (int i = threadIdx.x; i < cn; i += RS_THREADS) cb[i] = s_buf[i];
            __syncthreads();
        }
    }
}

// One CTA per (camera, super-tile): sorts the super-tile's keys by (depth bits, flat index) in place and counts
// the keys of each of its tiles (= the tile histogram) from the masks the keys carry.
__global__ void __launch_bounds__(RS_THREADS) super_sort_kernel(
    BinGeom G, int total_super, const int32_t* __restrict__ soff, const long long* __restrict__ counts_dev,
    unsigned long long* bucket, uint32_t* __restrict__ tile_count, long long cap_isects, long long cap_super) {
    if (counts_dev[1] > cap_isects || counts_dev[2] > cap_super) return;
    __shared__ __align__(16) unsigned long long s_buf[SORT_CHUNK];
    __shared__ uint32_t s_k32[SORT_CHUNK];
    __shared__ uint32_t s_red[2 * RS_WARPS];
    __shared__ uint32_t s_cnt[ST2];
    const int s = blockIdx.x;
    const int n_super = G.stw * G.sth, n_tiles = G.tile_w * G.tile_h;
    const int cam = s / n_super, sy = (s % n_super) / G.stw, sx = s % G.stw;
    const long long begin = soff[(long long)s * SUB];
    const long long end = s + 1 < total_super ? (long long)soff[(long long)(s + 1) * SUB] : counts_dev[2];
    const int n = (int)(end - begin);
    if (threadIdx.x < ST2) s_cnt[threadIdx.x] = 0;
    uint32_t cnt[ST2];
#pragma unroll
    for (int k = 0; k < ST2; ++k) cnt[k] = 0;
    if (n > 0) {
        unsigned long long* bk = bucket + begin;
        const bool big = n > SORT_CHUNK;
        if (n <= RS_THREADS) sort_small32<1>(bk, n, s_buf, s_k32, s_red);
        else if (n <= 2 * RS_THREADS) sort_small32<2>(bk, n, s_buf, s_k32, s_red);
        else if (n <= 4 * RS_THREADS) sort_small32<4>(bk, n, s_buf, s_k32, s_red);
        else if (!big) sort_small32<8>(bk, n, s_buf, s_k32, s_red);
        else sort_big(bk, n, s_buf, s_k32, s_red);
        for (int i = threadIdx.x; i < n; i += RS_THREADS) {
            const unsigned long long key = big ? bk[i] : s_buf[i];
            if (!big) bk[i] = key;
            const uint32_t mask = (uint32_t)key & ((1u << ST2) - 1u);
#pragma unroll
            for (int k = 0; k < ST2; ++k) cnt[k] += (mask >> k) & 1u;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ST2; ++k) {
        uint32_t c = cnt[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt[k], c);
    }
    __syncthreads();
    if (threadIdx.x < ST2) {
        const int tx = sx * ST + (threadIdx.x % ST), ty = sy * ST + (threadIdx.x / ST);
        if (tx < G.tile_w && ty < G.tile_h) tile_count[(long long)cam * n_tiles + ty * G.tile_w + tx] = s_cnt[threadIdx.x];
    }
}
static_assert(ST2 <= 8 && ST2 <= RS_WARPS, "one warp per tile of the super-tile, masks are 8 bits");
static_assert(SORT_CHUNK == 8 * RS_THREADS, "largest in-register sort class");

__device__ __forceinline__ long long block_sum_u32(const uint32_t* __restrict__ a, long long lo, long long hi,
                                                   long long* s_red) {
    long long v = 0;
    for (long long i = lo + threadIdx.x; i < hi; i += RS_THREADS) v += a[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    long long r = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) r += s_red[w];
    return r;
}

// One CTA per (camera, super-tile): the start of each of its tiles' ranges is the sum of the tile histogram before
// that tile (summed here, no scan launch; also written out as isect_offsets), and every tile's range is the
// ORDER-PRESERVING selection of the super-tile's sorted keys by the tile's mask bit: warp w serves tile w.
__global__ void __launch_bounds__(RS_THREADS) super_expand_kernel(
    BinGeom G, int total_super, int tile_bits, const int32_t* __restrict__ soff, const long long* __restrict__ counts_dev,
    const unsigned long long* __restrict__ bucket, const uint32_t* __restrict__ tile_count,
    int32_t* __restrict__ offsets, long long* __restrict__ isect_ids, int32_t* __restrict__ flatten_ids,
    long long cap_isects, long long cap_super) {
    if (counts_dev[1] > cap_isects || counts_dev[2] > cap_super) return;
    constexpr int CH = 1024;
    __shared__ __align__(16) unsigned long long s_key[CH];
    __shared__ long long s_red[RS_WARPS];
    const int s = blockIdx.x;
    const int n_super = G.stw * G.sth, n_tiles = G.tile_w * G.tile_h;
    const int cam = s / n_super, sy = (s % n_super) / G.stw, sx = s % G.stw;
    // tile rows ty0 (and ty0 + 1): first tile index of the super-tile in each
    const int tx0 = sx * ST, ty0 = sy * ST;
    const long long A = (long long)cam * n_tiles + (long long)ty0 * G.tile_w + tx0;
    long long row_start[ST];
    row_start[0] = block_sum_u32(tile_count, 0, A, s_red);
#pragma unroll
    for (int r = 1; r < ST; ++r)
        row_start[r] = ty0 + r < G.tile_h ? row_start[r - 1] + block_sum_u32(tile_count, A + (long long)(r - 1) * G.tile_w,
                                                                             A + (long long)r * G.tile_w, s_red)
                                          : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = warp % ST, ly = warp / ST;
    const bool live = warp < ST2 && tx0 + lx < G.tile_w && ty0 + ly < G.tile_h;
    long long wpos = 0;
    unsigned long long tkey_hi = 0;
    if (live) {
        const long long t_idx = A + (long long)ly * G.tile_w + lx;
        wpos = row_start[ly];
        for (int x = 0; x < lx; ++x) wpos += tile_count[t_idx - lx + x];
        if (lane == 0) offsets[t_idx] = (int32_t)wpos;
        tkey_hi = (((unsigned long long)cam << tile_bits) | (unsigned long long)((ty0 + ly) * G.tile_w + tx0 + lx)) << 32;
    }
    const long long begin = soff[(long long)s * SUB];
    const long long end = s + 1 < total_super ? (long long)soff[(long long)(s + 1) * SUB] : counts_dev[2];
    const int n = (int)(end - begin);
    const unsigned lt = (1u << lane) - 1u;
    for (int c0 = 0; c0 < n; c0 += CH) {
        const int cn = min(CH, n - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < cn; i += RS_THREADS) s_key[i] = bucket[begin + c0 + i];
        __syncthreads();
        if (!live) continue;
        for (int i0 = 0; i0 < cn; i0 += 32) {
            const int i = i0 + lane;
            const unsigned long long key = i < cn ? s_key[i] : 0ull;
            const bool in = ((uint32_t)key >> warp) & 1u;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, in);
            if (in) {
                const long long p = wpos + __popc(m & lt);
                isect_ids[p] = (long long)(tkey_hi | (key >> 32));
                flatten_ids[p] = (int32_t)(((uint32_t)key) >> (32 - ID_BITS));
            }
            wpos += __popc(m);
        }
    }
}

__global__ void offset_encode_kernel(const long long* __restrict__ isect_ids, long long n_isects, int n_tiles,
                                     int tile_bits, int total_tiles, int32_t* __restrict__ offsets) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_isects) return;
    const long long tmask = (1ll << tile_bits) - 1;
    const long long hi = isect_ids[i] >> 32;
    const int id_curr = (int)((hi >> tile_bits) * n_tiles + (hi & tmask));
    if (i == 0)
        for (int t = 0; t <= id_curr; ++t) offsets[t] = 0;
    if (i == n_isects - 1)
        for (int t = id_curr + 1; t < total_tiles; ++t) offsets[t] = (int32_t)n_isects;
    if (i > 0) {
        const long long hp = isect_ids[i - 1] >> 32;
        const int id_prev = (int)((hp >> tile_bits) * n_tiles + (hp & tmask));
        for (int t = id_prev + 1; t <= id_curr; ++t) offsets[t] = (int32_t)i;
    }
}

}  // namespace

HGS_API int hgs_isect_count(const float* means2d, const int32_t* radii, long long CN, int tile_size, int tile_w,
                            int tile_h, int32_t* tiles_per_gauss, void* stream) {
    if (CN < 0 || tile_size <= 0 || tile_w <= 0 || tile_h <= 0) return HGS_ERR_INVALID_ARG;
    if (CN == 0) return 0;
    isect_count_kernel<<<hgs_ceil_div(CN, 256), 256, 0, (cudaStream_t)stream>>>(means2d, radii, CN, tile_size, tile_w,
                                                                                 tile_h, tiles_per_gauss);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API size_t hgs_scan_temp_bytes(long long n) {
    return align_up((size_t)(hgs_ceil_div(n > 0 ? n : 1, LS_TILE)) * sizeof(long long)) + 256;
}

HGS_API int hgs_exclusive_scan_i32(const int32_t* in, int32_t* out, long long* total_dev, long long n, void* temp,
                                   size_t temp_bytes, void* stream) {
    if (n < 0) return HGS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (temp_bytes < hgs_scan_temp_bytes(n)) return HGS_ERR_WORKSPACE;
    // the element count lives in the last 256 bytes of temp
    long long* n_dev = (long long*)((char*)temp + hgs_scan_temp_bytes(n) - 256);
    set_ll_kernel<<<1, 1, 0, st>>>(n_dev, n);
    HGS_LAUNCH_CHECK();
    return large_scan(in, out, total_dev, n_dev, n, temp, temp_bytes - 256, st);
}

HGS_API int hgs_isect_emit(const float* means2d, const int32_t* radii, const float* depths, const int32_t* cum, int C,
                           int N, int tile_size, int tile_w, int tile_h, long long* isect_ids, int32_t* flatten_ids,
                           void* stream) {
    if (C <= 0 || N < 0 || tile_size <= 0) return HGS_ERR_INVALID_ARG;
    const long long CN = (long long)C * N;
    if (CN == 0) return 0;
    const int tile_bits = n_bits_of((long long)tile_w * tile_h);
    isect_emit_kernel<<<hgs_ceil_div(CN, 256), 256, 0, (cudaStream_t)stream>>>(
        means2d, radii, depths, cum, CN, N, tile_size, tile_w, tile_h, tile_bits, isect_ids, flatten_ids);
    HGS_LAUNCH_CHECK();
    return 0;
}

// temp layout of the sorted path (zero-filled by phase 1 up to zero_bytes; shared by both phases):
// look-back flags [ceil(CN / CP_TILE)] u64 | tile histogram [C*T] u32 | super-tile sub-bin histogram [S*SUB] u32 |
// sub-bin cursors [S*SUB] u32 | ticket || sub-bin offsets [S*SUB] i32 | visible-Gaussian records [CN] x 16 B
HGS_API size_t hgs_isect_bin_temp_bytes(long long CN, int C, int tile_w, int tile_h) {
    const BinGeom G = make_geom(1, 16, tile_w > 0 ? tile_w : 1, tile_h > 0 ? tile_h : 1);
    return bin_temp(nullptr, CN, (long long)C * G.stw * G.sth, (long long)C * tile_w * tile_h).bytes;
}

HGS_API size_t hgs_isect_bin_bucket_bytes(long long n_super_isects) {
    // one 8-byte sort key per (Gaussian, super-tile) pair
    return align_up((size_t)(n_super_isects > 0 ? n_super_isects : 1) * sizeof(unsigned long long));
}

HGS_API int hgs_isect_bin_prepare(const float* means2d, const int32_t* radii, const float* depths,
                                  const int32_t* tiles_per_gauss, int C, int N, int tile_size, int tile_w, int tile_h, int32_t* visible_ids,
                                  long long* counts_dev, void* temp, size_t temp_bytes, void* stream) {
    if (C <= 0 || N < 0 || tile_size <= 0 || tile_w <= 0 || tile_h <= 0) return HGS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const long long CN = (long long)C * N;
    const long long tt_ll = (long long)C * tile_w * tile_h;
    // flat indices travel in ID_BITS bits of the sort keys
    if (CN >= (1ll << ID_BITS) || tt_ll >= (1ll << 29) || tile_w >= 65536 || tile_h >= 65536) return HGS_ERR_TOO_LARGE;
    if (n_bits_of((long long)tile_w * tile_h) + n_bits_of(C) > 32) return HGS_ERR_TOO_LARGE;
    if (CN == 0) return (int)cudaMemsetAsync(counts_dev, 0, 3 * sizeof(long long), st);
    const BinGeom G = make_geom(N, tile_size, tile_w, tile_h);
    const int total_super = C * G.stw * G.sth;
    const BinTemp T = bin_temp(temp, CN, total_super, tt_ll);
    if (temp_bytes < T.bytes) return HGS_ERR_WORKSPACE;
    const int nblk = hgs_ceil_div(CN, CP_TILE);
    cudaError_t e = cudaMemsetAsync(temp, 0, T.zero_bytes, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(counts_dev, 0, 3 * sizeof(long long), st);
    if (e != cudaSuccess) return (int)e;
    bin_count_kernel<<<nblk, RS_THREADS, 0, st>>>(means2d, radii, depths, tiles_per_gauss, CN, G, nblk, T.flags, T.ticket,
                                                  T.super_count, visible_ids, T.vrec, counts_dev);
    HGS_LAUNCH_CHECK();
    hist_scan_kernel<<<1, SCAN_THREADS, 0, st>>>(T.super_count, total_super * SUB, T.soff, counts_dev);
    HGS_LAUNCH_CHECK();
    return 0;
}

// second launch of phase 1 alone (after hgs_project3d_fwd_bin, which fuses the first one into the projection)
HGS_API int hgs_isect_bin_scan(int C, int N, int tile_size, int tile_w, int tile_h, long long* counts_dev, void* temp,
                               size_t temp_bytes, void* stream) {
    if (C <= 0 || N < 0 || tile_size <= 0 || tile_w <= 0 || tile_h <= 0) return HGS_ERR_INVALID_ARG;
    if (N == 0) return 0;
    const long long CN = (long long)C * N;
    const BinGeom G = make_geom(N, tile_size, tile_w, tile_h);
    const int total_super = C * G.stw * G.sth;
    const BinTemp T = bin_temp(temp, CN, total_super, (long long)C * tile_w * tile_h);
    if (temp_bytes < T.bytes) return HGS_ERR_WORKSPACE;
    hist_scan_kernel<<<1, SCAN_THREADS, 0, (cudaStream_t)stream>>>(T.super_count, total_super * SUB, T.soff, counts_dev);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_isect_bin_sorted(const long long* counts_dev, int C, int N,
                                 long long n_visible_bound, long long n_isects, long long n_super_isects, int tile_size,
                                 int tile_w, int tile_h, int32_t* isect_offsets, long long* isect_ids,
                                 int32_t* flatten_ids, void* temp, size_t temp_bytes, void* bucket, size_t bucket_bytes,
                                 void* stream) {
    if (C <= 0 || N < 0 || n_isects < 0 || n_visible_bound < 0 || n_super_isects < 0 || tile_size <= 0 || tile_w <= 0 ||
        tile_h <= 0)
        return HGS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const long long CN = (long long)C * N;
    const int n_tiles = tile_w * tile_h;
    const long long tt_ll = (long long)C * n_tiles;
    if (n_isects >= (1ll << 31) || tt_ll >= (1ll << 29)) return HGS_ERR_TOO_LARGE;
    if (n_isects == 0 || n_visible_bound == 0)
        return (int)cudaMemsetAsync(isect_offsets, 0, (size_t)tt_ll * sizeof(int32_t), st);
    const BinGeom G = make_geom(N, tile_size, tile_w, tile_h);
    const int total_super = C * G.stw * G.sth;
    const BinTemp T = bin_temp(temp, CN, total_super, tt_ll);
    if (temp_bytes < T.bytes) return HGS_ERR_WORKSPACE;
    if (bucket_bytes < hgs_isect_bin_bucket_bytes(n_super_isects)) return HGS_ERR_WORKSPACE;
    unsigned long long* keys = (unsigned long long*)bucket;
    // n_isects / n_super_isects are CAPACITIES of the outputs / the bucket (>= the exact counts when the caller has read
    // them; a guess when it has not): every kernel returns at once when counts_dev says they do not fit, leaving temp
    // untouched, and the caller -- who reads counts_dev afterwards -- calls again with larger buffers
    const long long cap_super = (long long)(bucket_bytes / sizeof(unsigned long long));
    bin_scatter_kernel<<<hgs_ceil_div(n_visible_bound, RS_THREADS), RS_THREADS, 0, st>>>(
        T.vrec, counts_dev, G, T.soff, T.cursor, keys, n_isects, cap_super);
    HGS_LAUNCH_CHECK();
    super_sort_kernel<<<total_super, RS_THREADS, 0, st>>>(G, total_super, T.soff, counts_dev, keys, T.tile_count, n_isects,
                                                          cap_super);
    HGS_LAUNCH_CHECK();
    super_expand_kernel<<<total_super, RS_THREADS, 0, st>>>(G, total_super, n_bits_of(n_tiles), T.soff, counts_dev, keys,
                                                            T.tile_count, isect_offsets, isect_ids, flatten_ids, n_isects,
                                                            cap_super);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_isect_offset_encode(const long long* isect_ids, long long n_isects, int C, int tile_w, int tile_h,
                                    int32_t* isect_offsets, void* stream) {
    if (C <= 0 || tile_w <= 0 || tile_h <= 0 || n_isects < 0) return HGS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = tile_w * tile_h;
    const int total_tiles = C * n_tiles;
    if (n_isects == 0) return (int)cudaMemsetAsync(isect_offsets, 0, (size_t)total_tiles * sizeof(int32_t), st);
    offset_encode_kernel<<<hgs_ceil_div(n_isects, 256), 256, 0, st>>>(isect_ids, n_isects, n_tiles,
                                                                      n_bits_of(n_tiles), total_tiles, isect_offsets);
    HGS_LAUNCH_CHECK();
    return 0;
}
