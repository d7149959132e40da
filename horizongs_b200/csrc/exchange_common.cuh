// Building blocks of the peer-memory exchanges (exchange.cu, exchange_vjp.cu): mailbox layout, release / acquire
// flags at system scope, TMA bulk stores, the end-of-push protocol and the bounded wait kernel.
#pragma once
#include "hgs_common.cuh"
#include "../../include/hgs_raster.h"

namespace {

constexpr int EX_MAX_W = HGS_EXCHANGE_MAX_RANKS;
constexpr int EX_MAX_ROW = 128;        // floats per record (all tensors' widths, padded to a multiple of 4)
constexpr int EX_FLAG_STRIDE = 128;    // bytes between the per-source flags
constexpr int EX_COUNTER_OFF = EX_MAX_W * EX_FLAG_STRIDE;   // push-completion counter (local use)
constexpr int EX_CTRL_BYTES = 4096;
constexpr int EX_HDR_BYTES = 128;      // slot header: row count
constexpr int EX_THREADS = 256;
constexpr int EX_IDS = 256;            // Gaussian ids per merge CTA (and per entry of a slot's block index)
constexpr int EX_CHUNK = 64;           // records staged per TMA bulk store

struct ExPeers {
    unsigned char* base[EX_MAX_W];
    unsigned char* mc;      // NVLink multicast mapping of the mailboxes (a store lands in EVERY rank's mailbox), or NULL
    int world, rank;
};
// slot = [header 128 B][block entries: one per block of EX_IDS consecutive ids = {first record, 256-bit presence
//         bitmap}, 48 B each][records: cap x row floats]; mailbox = [control 4 KB][2 parities x world slots].
// The ids themselves never travel: the bitmaps list them, in record order.
constexpr int EX_ENTRY_WORDS = 12;     // lo, bits[8], pad[3]
struct ExLayout {
    long long cap, n_ids;
    int row, world, n_blocks;
    size_t ent_off, rows_off, slot_bytes;
};
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline ExLayout make_layout(int world, long long n_ids, long long cap, int row) {
    ExLayout L;
    L.cap = cap; L.n_ids = n_ids; L.row = row; L.world = world;
    L.n_blocks = (int)((n_ids + EX_IDS - 1) / EX_IDS);
    L.ent_off = EX_HDR_BYTES;
    L.rows_off = L.ent_off + align_up((size_t)L.n_blocks * EX_ENTRY_WORDS * 4, 128);
    L.slot_bytes = L.rows_off + (size_t)cap * row * 4;
    return L;
}
__device__ __forceinline__ size_t slot_offset(const ExLayout& L, int parity, int src) {
    return EX_CTRL_BYTES + (size_t)(parity * L.world + src) * L.slot_bytes;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* addr, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* addr) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
    return v;
}
// stores through the NVSwitch multicast mapping: one store instruction, one NVLink transfer out, delivered to all ranks
__device__ __forceinline__ void mc_st_v4(void* addr, float4 v) {
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void mc_st_f32(void* addr, float v) {
    asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mc_st_u64(void* addr, unsigned long long v) {
    asm volatile("multimem.st.weak.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void mc_st_release_sys_u64(void* addr, unsigned long long v) {
    asm volatile("multimem.st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// TMA bulk copy shared -> global (the destination may be peer memory), tracked by the thread's bulk group
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"((unsigned)__cvta_generic_to_shared(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ int lower_bound_ids(const int32_t* __restrict__ ids, int n, long long key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((long long)ids[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}


// slot entries of this rank's push: for every block of EX_IDS consecutive ids the first record and the presence
// bitmap, written into every peer's slot
__device__ __forceinline__ void write_block_entries(const ExPeers& P, const ExLayout& L, size_t soff,
                                                    const int32_t* __restrict__ ids, int n_rows) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < L.n_blocks; b += gridDim.x * blockDim.x) {
        const long long id0 = (long long)b * EX_IDS;
        const int lo = lower_bound_ids(ids, n_rows, id0);
        unsigned bits[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        for (int j = lo; j < n_rows; ++j) {
            const long long l = (long long)ids[j] - id0;
            if (l >= EX_IDS) break;
#pragma unroll
            for (int w = 0; w < 8; ++w) bits[w] |= ((int)(l >> 5) == w) ? (1u << (l & 31)) : 0u;
        }
        const uint4 e0 = make_uint4((unsigned)lo, bits[0], bits[1], bits[2]);
        const uint4 e1 = make_uint4(bits[3], bits[4], bits[5], bits[6]);
        const uint4 e2 = make_uint4(bits[7], 0u, 0u, 0u);
        if (P.mc != nullptr) {
            uint4* e = reinterpret_cast<uint4*>(P.mc + soff + L.ent_off) + (size_t)b * 3;
            mc_st_v4(e, make_float4(__uint_as_float(e0.x), __uint_as_float(e0.y), __uint_as_float(e0.z), __uint_as_float(e0.w)));
            mc_st_v4(e + 1, make_float4(__uint_as_float(e1.x), __uint_as_float(e1.y), __uint_as_float(e1.z), __uint_as_float(e1.w)));
            mc_st_v4(e + 2, make_float4(__uint_as_float(e2.x), 0.f, 0.f, 0.f));
            continue;
        }
        for (int q = 0; q < P.world; ++q) {
            uint4* e = reinterpret_cast<uint4*>(P.base[q] + soff + L.ent_off) + (size_t)b * 3;
            e[0] = e0; e[1] = e1; e[2] = e2;
        }
    }
}

// end of a push kernel: the threads that issued bulk stores complete them, every thread fences its stores
// system-wide, and the last block to get here raises this rank's flag in every mailbox (release)
__device__ __forceinline__ void publish_push(const ExPeers& P, unsigned long long flag_value) {
    if (threadIdx.x < P.world && P.mc == nullptr) {
        bulk_wait_all();
        fence_proxy_async();
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int* counter = reinterpret_cast<unsigned int*>(P.base[P.rank] + EX_COUNTER_OFF);
        const unsigned int prev = atomicAdd(counter, 1u);
        if (prev == gridDim.x - 1) {
            *counter = 0u;
            __threadfence_system();
            if (P.mc != nullptr) {
                mc_st_release_sys_u64(P.mc + (size_t)P.rank * EX_FLAG_STRIDE, flag_value);
            } else {
                for (int q = 0; q < P.world; ++q)
                    st_release_sys(reinterpret_cast<unsigned long long*>(P.base[q] + (size_t)P.rank * EX_FLAG_STRIDE),
                                   flag_value);
            }
        }
    }
}

// record of local id l of source s in an id block, from the block entry staged in shared memory
struct BlockEntries {
    unsigned bits[EX_MAX_W][8];
    int pre[EX_MAX_W][8];       // records of the block before bitmap word w
    int lo[EX_MAX_W];
};
__device__ __forceinline__ int record_index(const BlockEntries& E, int s, int l) {
    return E.lo[s] + E.pre[s][l >> 5] + __popc(E.bits[s][l >> 5] & ((1u << (l & 31)) - 1u));
}
// one round trip: entries of all sources for id block `blk` -> shared memory (call with all threads, then sync)
__device__ __forceinline__ void load_block_entries(BlockEntries& E, const ExLayout& L, const unsigned char* mailbox,
                                                   int parity, int blk) {
    if (threadIdx.x < L.world * 9) {
        const int src = threadIdx.x / 9, w = threadIdx.x - src * 9;
        const unsigned* e = reinterpret_cast<const unsigned*>(mailbox + slot_offset(L, parity, src) + L.ent_off) +
                            (size_t)blk * EX_ENTRY_WORDS;
        const unsigned v = __ldcg(e + w);
        if (w == 0) E.lo[src] = (int)v;
        else E.bits[src][w - 1] = v;
    }
}
__device__ __forceinline__ void prefix_block_entries(BlockEntries& E, int world) {
    if (threadIdx.x < world) {
        int acc = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { E.pre[threadIdx.x][w] = acc; acc += __popc(E.bits[threadIdx.x][w]); }
    }
}

// Wait (acquire, system scope) until every source has raised its flag for this step; bounded: a dead peer
// fails the step (*status = 1) instead of hanging the GPU.  A source whose contribution did not fit its slot
// pushed nothing and a negative row count: *status = 2 on EVERY rank (overflow is a collective outcome, the step
// counters stay in lock-step).  The merge / reduce kernel follows in stream order.
#define HGS_EX_TIMEOUT 1
#define HGS_EX_OVERFLOW 2
__global__ void exchange_wait_kernel(const unsigned char* mailbox, ExLayout L, int parity, unsigned long long want,
                                     int* __restrict__ status) {
    const int src = threadIdx.x;
    if (src >= L.world) return;
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(mailbox + (size_t)src * EX_FLAG_STRIDE);
    if (ld_acquire_sys(flag) < want) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(flag) < want) {
            if (global_timer_ns() - t0 > 20000000000ull) {
                atomicMax(status, HGS_EX_TIMEOUT);
                return;
            }
            __nanosleep(100);
        }
    }
    const long long rows = __ldcg(reinterpret_cast<const long long*>(mailbox + slot_offset(L, parity, src)));
    if (rows < 0) atomicMax(status, HGS_EX_OVERFLOW);
}


inline bool ex_bad_geometry(int world, int rank, long long n_ids, long long cap_rows) {
    return world < 1 || world > EX_MAX_W || rank < 0 || rank >= world || n_ids < 1 || n_ids >= (1ll << 31) ||
           cap_rows < 1 || cap_rows % 32 != 0 || cap_rows * 32 >= (1ll << 31);
}
inline int ex_fill_peers(ExPeers& P, void* const* mailboxes_host, int world, int rank, void* multicast = nullptr) {
    P.mc = (unsigned char*)multicast;
    for (int q = 0; q < EX_MAX_W; ++q) P.base[q] = q < world ? (unsigned char*)mailboxes_host[q] : nullptr;
    for (int q = 0; q < world; ++q)
        if (P.base[q] == nullptr) return HGS_ERR_INVALID_ARG;
    P.world = world;
    P.rank = rank;
    return 0;
}
inline int ex_push_grid(long long n_rows) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long chunks = (n_rows + EX_CHUNK - 1) / EX_CHUNK;
    return (int)(chunks < 1 ? 1 : (chunks > (long long)sms * 8 ? (long long)sms * 8 : chunks));
}

}  // namespace
