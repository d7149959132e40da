// Stages a8-a10: tile intersection, tile|depth ordering, per-tile ranges.  Integer work, bit-exact
// against oracle/gsplat_oracle.py::isect_tiles / isect_offset_encode (= gsplat isect_tiles(sort=True)
// + isect_offset_encode as reached inside gsplat.rasterization*, reference render.py:40,62).
//
// B200-first ordering.  gsplat emits I (key,value) pairs Gaussian-major and runs a generic
// 64-bit LSD radix sort over 32 + tile_bits + cam_bits key bits (6 passes x 24 B r/w per pair).
// The same order is produced here with far less HBM traffic:
//   1. compact the Gaussians that touch at least one tile (typically 10-20 % of a large scene) and
//      depth-order them once (u32 key = depth bits, 4 stable 8-bit passes, plus a camera pass when C > 1);
//   2. emit the pairs in that order, one thread per intersection (fully coalesced writes): pairs are already
//      depth-ordered, ties in flat-index order;
//   3. stable-partition the I pairs by (cam, tile) only: ceil((tile_bits+cam_bits)/8) passes of 8 B pairs.
// A stable sort on (cam, tile) of a sequence ordered by (depth, flat index) is exactly the stable sort
// on the full cam|tile|depth key of the Gaussian-major sequence: within one tile a Gaussian appears once,
// so ties on the full key are ties on depth between different Gaussians and both orders break them by
// ascending flat index.
//
// Radix pass = 3 launches (chunk histogram, per-digit row scan, chunk scatter); no inter-block waiting.
// Element counts that are only known on the device (visible Gaussians) are read from device memory by
// over-provisioned grids, so phase 1 needs no host round trip.  The scatter re-orders each 2048-pair tile in
// shared memory so that global writes are contiguous runs per digit.
// Roofline: HBM.  Per pass 4 B (hist) + 8 B + 8 B per pair.
#include "hgs_common.cuh"
#include "hgs_constants.cuh"
#include "../../include/hgs_raster.h"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048 pairs per block iteration
constexpr int RADIX = 256;
constexpr int RS_MAX_CHUNKS = 148 * 4;
constexpr int SCAN_THREADS = 1024;
constexpr int CP_TILE = 1024;  // elements per block in the compaction kernels

struct DigitSpec {
    int shift;
    uint32_t mask;
    int from_val_div;  // 0: digit from key; >0: digit from (val / from_val_div)
};
__device__ __forceinline__ uint32_t digit_of(const DigitSpec& ds, uint32_t key, uint32_t val) {
    uint32_t src = ds.from_val_div > 0 ? (val / (uint32_t)ds.from_val_div) : key;
    return (src >> ds.shift) & ds.mask;
}

// lanes of the warp holding the same 8-bit digit (fixed cost: 8 ballots; __match_any_sync degrades to one
// iteration per distinct value, i.e. 32 iterations on the low tile-id byte)
__device__ __forceinline__ unsigned peers_of(uint32_t d) {
    unsigned peers = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const unsigned b = __ballot_sync(0xFFFFFFFFu, (d >> k) & 1u);
        peers &= ((d >> k) & 1u) ? b : ~b;
    }
    return peers;
}

// chunk c of a pass over n elements with gridDim.x chunks: [begin, end)
__device__ __forceinline__ void chunk_range(long long n, long long& begin, long long& end) {
    const long long tiles = (n + RS_TILE - 1) / RS_TILE;
    const long long tpc = (tiles + gridDim.x - 1) / gridDim.x;
    begin = (long long)blockIdx.x * tpc * RS_TILE;
    end = begin + tpc * RS_TILE;
    if (end > n) end = n;
}
static int n_chunks_for(long long cap) {
    long long tiles = (cap + RS_TILE - 1) / RS_TILE;
    if (tiles < 1) tiles = 1;
    return (int)(tiles < RS_MAX_CHUNKS ? tiles : RS_MAX_CHUNKS);
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

__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const uint32_t* __restrict__ keys,
                                                                const uint32_t* __restrict__ vals,
                                                                const long long* __restrict__ n_dev, DigitSpec ds,
                                                                uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[RADIX];
    for (int i = threadIdx.x; i < RADIX; i += RS_THREADS) s_hist[i] = 0;
    __syncthreads();
    long long begin, end;
    chunk_range(*n_dev, begin, end);
    for (long long base = begin; base < end; base += RS_TILE) {
        // 8 independent loads in flight per thread, then warp-aggregated shared-memory increments
        uint32_t d[RS_ITEMS];
        bool valid[RS_ITEMS];
#pragma unroll
        for (int k = 0; k < RS_ITEMS; ++k) {
            const long long i = base + k * RS_THREADS + threadIdx.x;
            valid[k] = i < end;
            d[k] = valid[k] ? digit_of(ds, keys[i], ds.from_val_div > 0 ? vals[i] : 0u) : 0u;
        }
#pragma unroll
        for (int k = 0; k < RS_ITEMS; ++k) {
            const unsigned vmask = __ballot_sync(0xFFFFFFFFu, valid[k]);   // invalid lanes leave the peer sets
            const unsigned m = peers_of(d[k]) & vmask;
            const int leader = __ffs(m) - 1;
            if (valid[k] && (int)(threadIdx.x & 31) == leader) atomicAdd(&s_hist[d[k]], (uint32_t)__popc(m));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX; i += RS_THREADS) hist[(long long)i * gridDim.x + blockIdx.x] = s_hist[i];
}

// hist is [RADIX][n_chunks]: block d turns row d into its exclusive scan (over chunks) and writes the row total
__global__ void __launch_bounds__(RS_THREADS) radix_rowscan_kernel(uint32_t* __restrict__ hist, int n_chunks,
                                                                   uint32_t* __restrict__ rowsum) {
    __shared__ uint32_t s_warp[RS_WARPS];
    uint32_t* row = hist + (long long)blockIdx.x * n_chunks;
    uint32_t carry = 0;
    for (int base = 0; base < n_chunks; base += RS_THREADS) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n_chunks ? row[i] : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(v, s_warp, tot);
        if (i < n_chunks) row[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) rowsum[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(RS_THREADS, 4) radix_scatter_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, const long long* __restrict__ n_dev, DigitSpec ds,
    const uint32_t* __restrict__ hist_scanned, const uint32_t* __restrict__ rowsum) {
    __shared__ uint32_t s_base[RADIX];    // running global position of the next element of digit d (this chunk)
    __shared__ uint32_t s_start[RADIX];   // first local sorted index of digit d inside the current tile
    __shared__ uint32_t s_delta[RADIX];   // global position - local sorted index for digit d (mod 2^32)
    __shared__ uint32_t s_whist[RS_WARPS][RADIX];
    __shared__ uint32_t s_warp[RS_WARPS];
    __shared__ uint32_t s_key[RS_TILE];
    __shared__ uint32_t s_val[RS_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    {
        // global base of digit d = sum of the totals of smaller digits + this chunk's offset inside digit d
        uint32_t tot;
        const uint32_t digit_base = block_excl_scan(rowsum[threadIdx.x], s_warp, tot);
        s_base[threadIdx.x] = digit_base + hist_scanned[(long long)threadIdx.x * gridDim.x + blockIdx.x];
    }
    long long begin, end;
    chunk_range(*n_dev, begin, end);

    for (long long tile = begin; tile < end; tile += RS_TILE) {
        uint32_t key[RS_ITEMS], val[RS_ITEMS], dig[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            long long idx = tile + warp * (RS_ITEMS * 32) + i * 32 + lane;
            const bool valid = idx < end;
            key[i] = valid ? keys_in[idx] : 0xFFFFFFFFu;
            val[i] = valid ? vals_in[idx] : 0u;
            // padding sorts last inside the tile (largest digit, highest index) and is never written out
            dig[i] = valid ? digit_of(ds, key[i], val[i]) : (uint32_t)(RADIX - 1);
        }
#pragma unroll
        for (int i = 0; i < RADIX / 32; ++i) s_whist[warp][i * 32 + lane] = 0;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const unsigned m = peers_of(dig[i]);
            rank[i] = s_whist[warp][dig[i]] + (uint32_t)__popc(m & lt_mask);
            __syncwarp();
            if (lane == __ffs(m) - 1) s_whist[warp][dig[i]] += (uint32_t)__popc(m);
            __syncwarp();
        }
        __syncthreads();
        uint32_t cnt_d = 0;
        {
            // thread d: exclusive prefix over warps inside digit d, and the tile's count of digit d
            const int d = threadIdx.x;  // RS_THREADS == RADIX
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) {
                uint32_t t = s_whist[w][d];
                s_whist[w][d] = cnt_d;
                cnt_d += t;
            }
        }
        uint32_t tot;
        const uint32_t start_d = block_excl_scan(cnt_d, s_warp, tot);  // contains __syncthreads
        s_start[threadIdx.x] = start_d;
        s_delta[threadIdx.x] = s_base[threadIdx.x] - start_d;
        s_base[threadIdx.x] += cnt_d;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const uint32_t q = s_start[dig[i]] + s_whist[warp][dig[i]] + rank[i];
            s_key[q] = key[i];
            s_val[q] = val[i];
        }
        __syncthreads();
        const long long rem = end - tile;
        const int n_valid = (int)(rem < RS_TILE ? rem : RS_TILE);
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const int q = i * RS_THREADS + threadIdx.x;
            if (q < n_valid) {
                const uint32_t k = s_key[q], v = s_val[q];
                const uint32_t pos = s_delta[digit_of(ds, k, v)] + (uint32_t)q;
                keys_out[pos] = k;
                vals_out[pos] = v;
            }
        }
        __syncthreads();
    }
}
static_assert(RS_THREADS == RADIX, "one thread per digit in the warp-prefix step");

// one stable LSD pass: (keys_in, vals_in) -> (keys_out, vals_out); n lives on the device, cap bounds it
static int radix_pass(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                      const long long* n_dev, long long cap, DigitSpec ds, uint32_t* hist, cudaStream_t st) {
    const int nc = n_chunks_for(cap);
    radix_hist_kernel<<<nc, RS_THREADS, 0, st>>>(keys_in, vals_in, n_dev, ds, hist);
    HGS_LAUNCH_CHECK();
    uint32_t* rowsum = hist + (size_t)RADIX * RS_MAX_CHUNKS;
    radix_rowscan_kernel<<<RADIX, RS_THREADS, 0, st>>>(hist, nc, rowsum);
    HGS_LAUNCH_CHECK();
    radix_scatter_kernel<<<nc, RS_THREADS, 0, st>>>(keys_in, vals_in, keys_out, vals_out, n_dev, ds, hist, rowsum);
    HGS_LAUNCH_CHECK();
    return 0;
}
constexpr size_t HIST_BYTES = ((size_t)RADIX * RS_MAX_CHUNKS + RADIX) * sizeof(uint32_t);  // + row totals

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
static int n_bits_of(long long n) {  // floor(log2(n)) + 1 for n >= 1
    int b = 0;
    while (n > 0) { ++b; n >>= 1; }
    return b;
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

// ---- compaction of the Gaussians that touch at least one tile (order-preserving) ---------------------
__global__ void __launch_bounds__(RS_THREADS) vis_count_kernel(const int32_t* __restrict__ tiles_per_gauss,
                                                               long long CN, long long* __restrict__ blk_count) {
    __shared__ long long s_tmp[RS_THREADS / 32];
    const long long base = (long long)blockIdx.x * CP_TILE;
    long long c = 0;
#pragma unroll
    for (int i = 0; i < CP_TILE / RS_THREADS; ++i) {
        const long long idx = base + i * RS_THREADS + threadIdx.x;
        if (idx < CN && tiles_per_gauss[idx] > 0) ++c;
    }
    const long long tot = block_reduce_ll(c, s_tmp);
    if (threadIdx.x == 0) blk_count[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(RS_THREADS) compact_kernel(const float* __restrict__ depths,
                                                             const int32_t* __restrict__ tiles_per_gauss, long long CN,
                                                             const long long* __restrict__ blk_base,
                                                             uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                             int32_t* __restrict__ visible_ids) {
    __shared__ uint32_t s_warp[RS_WARPS];
    constexpr int PER = CP_TILE / RS_THREADS;  // 4 consecutive elements per thread (order preserving)
    const long long first = (long long)blockIdx.x * CP_TILE + (long long)threadIdx.x * PER;
    bool vis[PER];
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const long long idx = first + i;
        vis[i] = idx < CN && tiles_per_gauss[idx] > 0;
        cnt += vis[i] ? 1u : 0u;
    }
    uint32_t tot;
    uint32_t pos = block_excl_scan(cnt, s_warp, tot);
    const long long out0 = blk_base[blockIdx.x];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        if (vis[i]) {
            const long long idx = first + i;
            keys[out0 + pos] = (uint32_t)__float_as_int(depths[idx]);
            vals[out0 + pos] = (uint32_t)idx;
            if (visible_ids != nullptr) visible_ids[out0 + pos] = (int32_t)idx;
            ++pos;
        }
    }
}

__global__ void gather_counts_kernel(const uint32_t* __restrict__ vals, const int32_t* __restrict__ tiles_per_gauss,
                                     const long long* __restrict__ n_dev, int32_t* __restrict__ order,
                                     int32_t* __restrict__ cnt_sorted) {
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *n_dev) return;
    uint32_t g = vals[j];
    order[j] = (int32_t)g;
    cnt_sorted[j] = tiles_per_gauss[g];
}

// ---- depth-ordered emission, one thread per intersection ----------------------------------------------
// Block b owns emissions [b*RS_TILE, (b+1)*RS_TILE): it locates the (at most RS_TILE) sorted Gaussians that
// produce them, stages their tile boxes in shared memory and lets every thread binary-search its Gaussian.
__global__ void __launch_bounds__(RS_THREADS) emit_sorted_kernel(
    const float* __restrict__ means2d, const int32_t* __restrict__ radii, const int32_t* __restrict__ order,
    const int32_t* __restrict__ cum_sorted, long long n_vis, long long n_isects, int N, int tile_size, int tile_w,
    int tile_h, int tile_bits, uint32_t* __restrict__ tkeys, uint32_t* __restrict__ vals) {
    __shared__ int s_cum[RS_TILE];
    __shared__ uint32_t s_g[RS_TILE];
    __shared__ uint32_t s_xy[RS_TILE];   // y0 << 16 | x0
    __shared__ int s_w[RS_TILE];
    __shared__ long long s_j[2];
    const long long e0 = (long long)blockIdx.x * RS_TILE;
    long long e1 = e0 + RS_TILE;
    if (e1 > n_isects) e1 = n_isects;
    if (threadIdx.x < 64) {
        // warp 0 / warp 1: last j with cum_sorted[j] <= target, 32-ary search (cum_sorted[0] == 0 <= target)
        const int lane = threadIdx.x & 31;
        const long long target = threadIdx.x < 32 ? e0 : e1 - 1;
        long long lo = 0, hi = n_vis;                       // answer in [lo, hi)
        while (hi - lo > 32) {
            const long long step = (hi - lo + 31) >> 5;
            const long long idx = lo + lane * step;
            const bool le = idx < hi && (long long)cum_sorted[idx] <= target;
            const int p = __popc(__ballot_sync(0xFFFFFFFFu, le));   // >= 1: probes are sorted, lane 0 is true
            const long long nlo = lo + (long long)(p - 1) * step;
            const long long nhi = lo + (long long)p * step;
            lo = nlo;
            hi = nhi < hi ? nhi : hi;
        }
        const long long idx = lo + lane;
        const bool le = idx < hi && (long long)cum_sorted[idx] <= target;
        const int p = __popc(__ballot_sync(0xFFFFFFFFu, le));
        if (lane == 0) s_j[threadIdx.x >> 5] = lo + p - 1;
    }
    __syncthreads();
    const long long j0 = s_j[0];
    const int nj = (int)(s_j[1] - j0 + 1);   // <= RS_TILE: every compacted Gaussian emits at least once
    for (int j = threadIdx.x; j < nj; j += RS_THREADS) {
        const uint32_t g = (uint32_t)order[j0 + j];
        const float2 m = reinterpret_cast<const float2*>(means2d)[g];
        int x0, y0, x1, y1;
        hgs_tile_bbox(m.x, m.y, (float)radii[g], (float)tile_size, tile_w, tile_h, x0, y0, x1, y1);
        s_cum[j] = cum_sorted[j0 + j];
        s_g[j] = g;
        s_xy[j] = ((uint32_t)y0 << 16) | (uint32_t)x0;
        s_w[j] = x1 - x0;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const long long e = e0 + i * RS_THREADS + threadIdx.x;
        if (e < e1) {
            int lo = 0, hi = nj - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if ((long long)s_cum[mid] <= e) lo = mid; else hi = mid - 1;
            }
            const int k = (int)(e - s_cum[lo]);
            const int w = s_w[lo];
            const uint32_t xy = s_xy[lo];
            const int y = (int)(xy >> 16) + k / w, x = (int)(xy & 0xFFFFu) + k % w;
            const uint32_t g = s_g[lo];
            tkeys[e] = ((g / (uint32_t)N) << tile_bits) | (uint32_t)(y * tile_w + x);
            vals[e] = g;
        }
    }
}

__global__ void finalize_sorted_kernel(const uint32_t* __restrict__ tkeys, const uint32_t* __restrict__ vals,
                                       const float* __restrict__ depths, long long n_isects, int n_tiles,
                                       int tile_bits, int total_tiles, long long* __restrict__ isect_ids,
                                       int32_t* __restrict__ offsets) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_isects) return;
    const uint32_t tk = tkeys[i];
    const uint32_t g = vals[i];
    isect_ids[i] = ((long long)tk << 32) | (long long)__float_as_int(depths[g]);
    const uint32_t tmask = (1u << tile_bits) - 1u;
    const int id_curr = (int)(tk >> tile_bits) * n_tiles + (int)(tk & tmask);
    if (i == 0)
        for (int t = 0; t <= id_curr; ++t) offsets[t] = 0;
    if (i == n_isects - 1)
        for (int t = id_curr + 1; t < total_tiles; ++t) offsets[t] = (int32_t)n_isects;
    if (i > 0) {
        const uint32_t tp = tkeys[i - 1];
        const int id_prev = (int)(tp >> tile_bits) * n_tiles + (int)(tp & tmask);
        for (int t = id_prev + 1; t <= id_curr; ++t) offsets[t] = (int32_t)i;
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

// temp layout for prepare: keysA, valsA, keysB, valsB (CN u32 each), cnt_sorted (CN i32), hist,
// block counts/bases (ceil(CN/CP_TILE) long longs), scan temp
HGS_API size_t hgs_isect_prepare_temp_bytes(long long CN) {
    const size_t a = align_up((size_t)(CN > 0 ? CN : 1) * 4);
    const size_t nb = (size_t)hgs_ceil_div(CN > 0 ? CN : 1, CP_TILE);
    return 5 * a + align_up(HIST_BYTES) + align_up(nb * sizeof(long long)) + hgs_scan_temp_bytes(CN);
}

HGS_API int hgs_isect_prepare(const float* depths, const int32_t* tiles_per_gauss, int C, int N, int32_t* order,
                              int32_t* cum_sorted, int32_t* visible_ids, long long* counts_dev, void* temp,
                              size_t temp_bytes, void* stream) {
    if (C <= 0 || N < 0) return HGS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const long long CN = (long long)C * N;
    if (CN >= (1ll << 31)) return HGS_ERR_TOO_LARGE;
    if (CN == 0) return (int)cudaMemsetAsync(counts_dev, 0, 2 * sizeof(long long), st);
    if (temp_bytes < hgs_isect_prepare_temp_bytes(CN)) return HGS_ERR_WORKSPACE;
    const size_t a = align_up((size_t)CN * 4);
    const int nblk = hgs_ceil_div(CN, CP_TILE);
    char* p = (char*)temp;
    uint32_t* kA = (uint32_t*)p; p += a;
    uint32_t* vA = (uint32_t*)p; p += a;
    uint32_t* kB = (uint32_t*)p; p += a;
    uint32_t* vB = (uint32_t*)p; p += a;
    int32_t* cnt_sorted = (int32_t*)p; p += a;
    uint32_t* hist = (uint32_t*)p; p += align_up(HIST_BYTES);
    long long* blk = (long long*)p; p += align_up((size_t)nblk * sizeof(long long));
    void* scan_temp = p;
    long long* n_vis_dev = counts_dev;      // counts_dev[0] = visible Gaussians, counts_dev[1] = intersections

    // 1. order-preserving compaction of the Gaussians with at least one tile
    vis_count_kernel<<<nblk, RS_THREADS, 0, st>>>(tiles_per_gauss, CN, blk);
    HGS_LAUNCH_CHECK();
    ls_scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(blk, nblk, n_vis_dev);
    HGS_LAUNCH_CHECK();
    compact_kernel<<<nblk, RS_THREADS, 0, st>>>(depths, tiles_per_gauss, CN, blk, kA, vA, visible_ids);
    HGS_LAUNCH_CHECK();
    // 2. stable LSD sort of the depth bits (then of the camera index when C > 1)
    uint32_t *ki = kA, *vi = vA, *ko = kB, *vo = vB;
    for (int pass = 0; pass < 4; ++pass) {
        DigitSpec ds{pass * 8, 0xFFu, 0};
        int rc = radix_pass(ki, vi, ko, vo, n_vis_dev, CN, ds, hist, st);
        if (rc) return rc;
        uint32_t* t;
        t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    if (C > 1) {
        const int cam_bits = n_bits_of(C);
        for (int shift = 0; shift < cam_bits; shift += 8) {
            int nb = cam_bits - shift < 8 ? cam_bits - shift : 8;
            DigitSpec ds{shift, (1u << nb) - 1u, N};
            int rc = radix_pass(ki, vi, ko, vo, n_vis_dev, CN, ds, hist, st);
            if (rc) return rc;
            uint32_t* t;
            t = ki; ki = ko; ko = t;
            t = vi; vi = vo; vo = t;
        }
    }
    // 3. per-Gaussian tile counts in sorted order and their exclusive scan
    gather_counts_kernel<<<hgs_ceil_div(CN, 256), 256, 0, st>>>(vi, tiles_per_gauss, n_vis_dev, order, cnt_sorted);
    HGS_LAUNCH_CHECK();
    return large_scan(cnt_sorted, cum_sorted, counts_dev + 1, n_vis_dev, CN, scan_temp, hgs_scan_temp_bytes(CN) - 256,
                      st);
}

// temp layout for sorted: tkeyA, tkeyB, valsT (I u32 each), hist, element count
HGS_API size_t hgs_isect_sorted_temp_bytes(long long CN, long long n_isects) {
    (void)CN;
    size_t a = align_up((size_t)(n_isects > 0 ? n_isects : 1) * 4);
    return 3 * a + align_up(HIST_BYTES) + 256;
}

HGS_API int hgs_isect_sorted(const float* means2d, const int32_t* radii, const float* depths, const int32_t* order,
                             const int32_t* cum_sorted, int C, int N, long long n_visible, long long n_isects,
                             int tile_size, int tile_w, int tile_h, long long* isect_ids, int32_t* flatten_ids,
                             int32_t* isect_offsets, void* temp, size_t temp_bytes, void* stream) {
    if (C <= 0 || N < 0 || n_isects < 0 || n_visible < 0 || tile_size <= 0 || tile_w <= 0 || tile_h <= 0)
        return HGS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const long long CN = (long long)C * N;
    const int n_tiles = tile_w * tile_h;
    const long long total_tiles_ll = (long long)C * n_tiles;
    if (n_isects >= (1ll << 31) || total_tiles_ll >= (1ll << 31) || tile_w >= 65536 || tile_h >= 65536)
        return HGS_ERR_TOO_LARGE;
    const int total_tiles = (int)total_tiles_ll;
    if (n_isects == 0) return (int)cudaMemsetAsync(isect_offsets, 0, (size_t)total_tiles * sizeof(int32_t), st);
    if (n_visible == 0) return HGS_ERR_INVALID_ARG;
    const int tile_bits = n_bits_of(n_tiles);
    const int cam_bits = n_bits_of(C);
    if (tile_bits + cam_bits > 32) return HGS_ERR_TOO_LARGE;
    if (temp_bytes < hgs_isect_sorted_temp_bytes(CN, n_isects)) return HGS_ERR_WORKSPACE;
    const size_t a = align_up((size_t)n_isects * 4);
    char* p = (char*)temp;
    uint32_t* kA = (uint32_t*)p; p += a;
    uint32_t* kB = (uint32_t*)p; p += a;
    uint32_t* vT = (uint32_t*)p; p += a;
    uint32_t* hist = (uint32_t*)p; p += align_up(HIST_BYTES);
    long long* n_dev = (long long*)p;
    set_ll_kernel<<<1, 1, 0, st>>>(n_dev, n_isects);
    HGS_LAUNCH_CHECK();

    // when C == 1 the camera bit is always 0: sort tile bits only
    const int key_bits = (C > 1) ? tile_bits + cam_bits : tile_bits;
    const int n_pass = (key_bits + 7) / 8;
    // choose ping-pong start so the last pass lands the values in flatten_ids
    uint32_t* vF = (uint32_t*)flatten_ids;
    uint32_t *ki = kA, *ko = kB;
    uint32_t* vi = (n_pass % 2 == 0) ? vF : vT;
    uint32_t* vo = (n_pass % 2 == 0) ? vT : vF;

    emit_sorted_kernel<<<hgs_ceil_div(n_isects, RS_TILE), RS_THREADS, 0, st>>>(
        means2d, radii, order, cum_sorted, n_visible, n_isects, N, tile_size, tile_w, tile_h, tile_bits, ki, vi);
    HGS_LAUNCH_CHECK();
    for (int pass = 0; pass < n_pass; ++pass) {
        int shift = pass * 8;
        int nb = key_bits - shift < 8 ? key_bits - shift : 8;
        DigitSpec ds{shift, (1u << nb) - 1u, 0};
        int rc = radix_pass(ki, vi, ko, vo, n_dev, n_isects, ds, hist, st);
        if (rc) return rc;
        uint32_t* t;
        t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    // now (ki, vi) hold the result and vi == flatten_ids (values are already in their output buffer)
    finalize_sorted_kernel<<<hgs_ceil_div(n_isects, 256), 256, 0, st>>>(ki, vi, depths, n_isects, n_tiles, tile_bits,
                                                                        total_tiles, isect_ids, isect_offsets);
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
