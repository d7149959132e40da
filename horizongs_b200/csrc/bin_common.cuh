// Shared by the ordering kernels (isect.cu) and the fused projection (project3d.cu): geometry of the binning by
// super-tiles, the per-visible-Gaussian record, the layout of the zero-filled scratch region and the two device
// routines of the ordered compaction (decoupled look-back) and of the super-tile histogram.
#pragma once
#include "hgs_common.cuh"

namespace hgs_bin {

constexpr int BIN_THREADS = 256;
constexpr int CP_TILE = 1024;                    // flat Gaussians per CTA of the compaction
constexpr int ST = 2;                            // tiles per side of a super-tile (the binning / sorting unit)
constexpr int ST2 = ST * ST;
constexpr int SUB = 4;                           // sub-bins per super-tile (by flat index): spreads the atomics
constexpr int PAD = 32;                          // u32 slots per atomic counter: one 128-byte line each (the L2 serialises
                                                 // atomics per line, not per address)
constexpr int BIG_AREA = 16;                     // Gaussians covering more tiles are spread over the whole CTA
#define LB_AGG (1ull << 62)      // look-back flag: block aggregate published
#define LB_PREFIX (2ull << 62)   // look-back flag: inclusive prefix published
#define LB_MASK ((1ull << 62) - 1)
#define KEY_INF (~0ull)
constexpr int ID_BITS = 28;                      // sort key = depth bits << 32 | flat index << 4 | tile mask
static_assert(ST2 <= 4, "the tile mask of a key is 4 bits");

// what the later kernels need to know about a visible Gaussian (written once, in work-list order)
struct __align__(16) VisRec {
    uint32_t g, depth_bits;
    uint32_t xy0, xy1;       // tile box: x0 | y0 << 16,  x1 | y1 << 16
};

struct BinGeom {
    int N, tile_w, tile_h, stw, sth;             // stw x sth super-tiles per camera
    float tile_size, inv_tile_size;              // inv_tile_size > 0: tile_size is a power of two (x / ts == x * inv)
};

// tile box [x0, x1) x [y0, y1) of a visible Gaussian: hgs_tile_bbox's arithmetic (a power-of-two tile size
// divides exactly by multiplication)
__device__ __forceinline__ void tile_box(const BinGeom& G, const float* __restrict__ means2d,
                                         const int32_t* __restrict__ radii, long long g, int& x0, int& y0, int& x1,
                                         int& y1) {
    const float2 m = reinterpret_cast<const float2*>(means2d)[g];
    const float r = (float)radii[g];
    if (G.inv_tile_size > 0.f) {
        const float tr = r * G.inv_tile_size, tx = m.x * G.inv_tile_size, ty = m.y * G.inv_tile_size;
        x0 = (int)fminf(fmaxf(floorf(tx - tr), 0.f), (float)G.tile_w);
        y0 = (int)fminf(fmaxf(floorf(ty - tr), 0.f), (float)G.tile_h);
        x1 = (int)fminf(fmaxf(ceilf(tx + tr), 0.f), (float)G.tile_w);
        y1 = (int)fminf(fmaxf(ceilf(ty + tr), 0.f), (float)G.tile_h);
    } else {
        hgs_tile_bbox(m.x, m.y, r, G.tile_size, G.tile_w, G.tile_h, x0, y0, x1, y1);
    }
}


// exclusive prefix of the block totals of all earlier logical blocks (decoupled look-back).  Call with the 32
// threads of ONE warp after the block published (LB_AGG | total) -- or LB_PREFIX | total for block 0 -- in flags[bid];
// publishes this block's inclusive prefix.  Every lane returns the exclusive prefix.
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long* flags, uint32_t bid, uint32_t tot) {
    const int lane = threadIdx.x & 31;
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
    return excl;
}

// histogram of the super-tiles a visible Gaussian with tile box [x0,x1) x [y0,y1) touches (small boxes: the caller
// spreads boxes of more than BIG_AREA super-tiles over the CTA)
__device__ __forceinline__ void super_hist_add(const BinGeom& G, uint32_t* __restrict__ super_count, uint32_t g, int x0,
                                               int y0, int x1, int y1) {
    const int n_super = G.stw * G.sth;
    uint32_t* srow = super_count + (((long long)(g / (uint32_t)G.N) * n_super) * SUB + (g % SUB)) * PAD;
    for (int y = y0 / ST; y <= (y1 - 1) / ST; ++y)
        for (int x = x0 / ST; x <= (x1 - 1) / ST; ++x) atomicAdd(srow + (long long)(y * G.stw + x) * (SUB * PAD), 1u);
}
__device__ __forceinline__ int super_area(int x0, int y0, int x1, int y1) {
    return ((x1 - 1) / ST - x0 / ST + 1) * ((y1 - 1) / ST - y0 / ST + 1);
}
// the same for one big box, all threads of the CTA cooperating
__device__ __forceinline__ void super_hist_add_cta(const BinGeom& G, uint32_t* __restrict__ super_count, uint32_t g,
                                                   int x0, int y0, int x1, int y1) {
    const int n_super = G.stw * G.sth;
    const int sx0 = x0 / ST, sy0 = y0 / ST, sw = (x1 - 1) / ST - sx0 + 1, sarea = sw * ((y1 - 1) / ST - sy0 + 1);
    uint32_t* srow = super_count + (((long long)(g / (uint32_t)G.N) * n_super) * SUB + (g % SUB)) * PAD;
    for (int k = threadIdx.x; k < sarea; k += blockDim.x)
        atomicAdd(srow + (long long)((sy0 + k / sw) * G.stw + sx0 + k % sw) * (SUB * PAD), 1u);
}

// ---- scratch layout (host side) ------------------------------------------------------------------------
static inline size_t bin_align_up(size_t x) { return (x + 255) & ~(size_t)255; }
static inline BinGeom make_geom(int N, int tile_size, int tile_w, int tile_h) {
    BinGeom G;
    G.N = N; G.tile_w = tile_w; G.tile_h = tile_h;
    G.stw = (tile_w + ST - 1) / ST;
    G.sth = (tile_h + ST - 1) / ST;
    G.tile_size = (float)tile_size;
    G.inv_tile_size = (tile_size & (tile_size - 1)) == 0 ? 1.0f / (float)tile_size : 0.f;
    return G;
}
// zero-filled by phase 1 up to zero_bytes; shared by both phases:
// look-back flags [ceil(CN / CP_TILE)] u64 | tile histogram [C*T] u32 | super-tile sub-bin histogram [S*SUB*PAD] u32 |
// sub-bin cursors [S*SUB*PAD] u32 | ticket || sub-bin offsets [S*SUB] i32 | visible-Gaussian records [CN] x 16 B
struct BinTemp {
    unsigned long long* flags;
    uint32_t *tile_count, *super_count, *cursor, *ticket;
    int32_t* soff;
    VisRec* vrec;
    size_t zero_bytes, bytes;
};
static inline BinTemp bin_temp(void* temp, long long CN, long long total_super, long long total_tiles) {
    BinTemp T;
    const size_t nblk = (size_t)hgs_ceil_div(CN > 0 ? CN : 1, CP_TILE);
    const size_t S = (size_t)(total_super > 0 ? total_super : 1) * SUB, TT = (size_t)(total_tiles > 0 ? total_tiles : 1);
    char* p = (char*)temp;
    T.flags = (unsigned long long*)p; p += bin_align_up(nblk * sizeof(unsigned long long));
    T.tile_count = (uint32_t*)p; p += bin_align_up(TT * sizeof(uint32_t));
    T.super_count = (uint32_t*)p; p += bin_align_up(S * PAD * sizeof(uint32_t));
    T.cursor = (uint32_t*)p; p += bin_align_up(S * PAD * sizeof(uint32_t));
    T.ticket = (uint32_t*)p; p += 256;
    T.zero_bytes = (size_t)(p - (char*)temp);
    T.soff = (int32_t*)p; p += bin_align_up(S * sizeof(int32_t));
    T.vrec = (VisRec*)p; p += bin_align_up((size_t)(CN > 0 ? CN : 1) * sizeof(VisRec));
    T.bytes = (size_t)(p - (char*)temp);
    return T;
}

}  // namespace hgs_bin
