// Stage e: exchange of the view-sharded Gaussian-parameter gradients between the GPUs of
// one NVSwitch domain, written as a SPARSE all-reduce over NVLink peer memory.
//
// Why not the dense NCCL all-reduce: with one view per GPU only the Gaussians a view SEES (13 % of the 6 M scene)
// have a non-zero gradient row; a dense all-reduce moves all 38 floats of all N Gaussians through every link
// (912 MB in and out per GPU and step at 6 M) although 87 % of it is zeros.  Here every rank
//   push   : gathers its visible rows (all parameter gradients of a Gaussian = one contiguous 160-byte record)
//            from the dense gradient tensors into shared memory and stores each staged chunk with one TMA bulk
//            copy (cp.async.bulk shared -> global) per peer straight into a mailbox slot in EVERY peer's memory
//            over NVLink (the local slot is written the same way), then publishes a monotonically increasing
//            flag in each peer with release semantics;
//   reduce : after all sources' flags are acquired, one CTA per block of 256 consecutive Gaussian ids adds the
//            records of sources 0, 1, .., W-1 IN THAT ORDER (ids are ascending per source, so a block's records
//            are one contiguous run per source, located through a per-block entry {first record, presence
//            bitmap} the pusher writes) and stores the whole id block once with coalesced 16-byte writes -- no
//            read-modify-write of the dense tensors.
// Every rank performs the same additions in the same order, so the replicas' gradients are bit-identical
// (what a replicated-parameter optimizer needs), and the volume per GPU is (W-1) x (visible rows) x 160 B
// instead of 2 x (W-1)/W x N x 152 B.
// Mailbox slots are double-buffered on the step parity: a rank can only be one exchange ahead of a peer
// (its reduce of step s waits for that peer's push of step s), so slot (s & 1) is free again at step s + 2.
// Roofline: NVLink inbound bandwidth for push, HBM for reduce.
#include <string.h>

#include "exchange_common.cuh"

namespace {

constexpr int EX_MAX_T = HGS_EXCHANGE_MAX_TENSORS;

struct ExTensors {
    float* p[EX_MAX_T];
    int w[EX_MAX_T];
    int n;      // tensors
    int row;    // floats per record
};
// column -> (pointer to that column of row 0, row stride, position in the merge kernel's staging tile)
struct ColTable {
    float* ptr[EX_MAX_ROW];
    int ld[EX_MAX_ROW];
    int reg[EX_MAX_ROW];
};
__device__ __forceinline__ void build_table(const ExTensors& T, ColTable& tab) {
    for (int col = threadIdx.x; col < T.row; col += blockDim.x) {
        float* p = nullptr;
        int ld = 0, off = 0, reg = 0;
#pragma unroll
        for (int k = 0; k < EX_MAX_T; ++k) {
            if (k < T.n) {
                if (col >= off && col < off + T.w[k]) {
                    p = T.p[k] + (col - off);
                    ld = T.w[k];
                    reg = off * EX_IDS + (col - off);     // tensor k's region starts at EX_IDS * (widths before it)
                }
                off += T.w[k];
            }
        }
        tab.ptr[col] = p;   // nullptr = padding column
        tab.ld[col] = ld;
        tab.reg[col] = reg;
    }
    __syncthreads();
}

// Gather the records of the rows listed in ids[] (ascending) from the dense tensors, stage them in shared memory
// and store each staged chunk with one TMA bulk copy per peer into slot (parity, rank) of that peer's mailbox
// (double buffered: the gather of chunk k+1 overlaps the stores of chunk k); also the slot's block entries.
__global__ void __launch_bounds__(EX_THREADS) exchange_push_kernel(ExTensors T, ExPeers P, ExLayout L,
                                                                  const int32_t* __restrict__ ids, int n_rows, int parity,
                                                                  unsigned long long flag_value) {
    extern __shared__ __align__(128) unsigned char ex_smem[];
    ColTable& tab = *reinterpret_cast<ColTable*>(ex_smem);
    build_table(T, tab);
    const int r4 = T.row >> 2;
    float4* const stage_rows[2] = {reinterpret_cast<float4*>(ex_smem + sizeof(ColTable)),
                                   reinterpret_cast<float4*>(ex_smem + sizeof(ColTable)) + EX_CHUNK * r4};
    const size_t soff = slot_offset(L, parity, P.rank);
    write_block_entries(P, L, soff, ids, n_rows);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int q = 0; q < P.world; ++q) *reinterpret_cast<long long*>(P.base[q] + soff) = n_rows;
    }
    const int n_chunks = (n_rows + EX_CHUNK - 1) / EX_CHUNK;
    int it = 0;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
        const int st = it & 1;
        const int r0 = chunk * EX_CHUNK;
        const int nr = min(EX_CHUNK, n_rows - r0);
        // the bulk stores issued from this stage two iterations ago must have finished READING shared memory
        if (threadIdx.x < P.world) bulk_wait_read<1>();
        __syncthreads();
        for (int f = threadIdx.x; f < nr * r4; f += EX_THREADS) {
            const int r = f / r4;
            const int c = (f - r * r4) << 2;
            const long long id = ids[r0 + r];
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float* p = tab.ptr[c + k];
                v[k] = p != nullptr ? p[id * tab.ld[c + k]] : 0.f;
            }
            stage_rows[st][f] = make_float4(v[0], v[1], v[2], v[3]);
        }
        fence_proxy_async();     // generic-proxy writes of shared memory -> visible to the async proxy (TMA)
        __syncthreads();
        if (threadIdx.x < P.world) {
            bulk_s2g(P.base[threadIdx.x] + soff + L.rows_off + (size_t)r0 * T.row * 4, stage_rows[st],
                     (unsigned)(nr * T.row * 4));
            bulk_commit();
        }
    }
    publish_push(P, flag_value);
}

// Merge of all sources, one CTA per block of EX_IDS consecutive Gaussian ids.  Every source lists its ids in
// ascending order, so the records of an id block are one contiguous run per source: the slot's block entry gives
// its first record and a presence bitmap; the record of local id l is found by popcount.  Thread (touched id,
// column group) loads that Gaussian's record from every source that has one and adds them IN RANK ORDER (the
// same sums on every rank); the sums are staged in shared memory in the dense tensors' own layout and the whole
// id block is written with coalesced 16-byte stores -- rows nobody listed are written as the zeros they are, so
// there are no partial-sector writes and no read-modify-write of the dense tensors.
struct MergeSmem {      // followed by float stage[EX_IDS * row]
    ColTable tab;
    BlockEntries ent;
    unsigned short present[EX_IDS];
    unsigned short touched[EX_IDS];
    int n_touched;
    int pad[3];
};
static_assert(sizeof(MergeSmem) % 16 == 0, "staging tile must stay 16-byte aligned");

__global__ void __launch_bounds__(EX_THREADS) exchange_merge_kernel(ExTensors T, ExLayout L, const unsigned char* mailbox,
                                                                   int parity, const int* __restrict__ status) {
    extern __shared__ __align__(128) unsigned char ex_smem[];
    MergeSmem& S = *reinterpret_cast<MergeSmem*>(ex_smem);
    float* stage = reinterpret_cast<float*>(ex_smem + sizeof(MergeSmem));
    const int r4 = T.row >> 2;
    const int world = L.world;
    const int tid = threadIdx.x;
    if (__ldcg(status) != 0) return;     // a peer never arrived: leave the tensors alone, the host raises
    load_block_entries(S.ent, L, mailbox, parity, blockIdx.x);    // one round trip
    if (tid == 0) S.n_touched = 0;
    {
        float4* z = reinterpret_cast<float4*>(stage);
        for (int i = tid; i < EX_IDS * r4; i += EX_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    build_table(T, S.tab);    // ends with __syncthreads()
    prefix_block_entries(S.ent, world);
    {
        // EX_IDS == EX_THREADS: thread l owns local id l
        unsigned m = 0;
        for (int s = 0; s < world; ++s) m |= ((S.ent.bits[s][tid >> 5] >> (tid & 31)) & 1u) << s;
        S.present[tid] = (unsigned short)m;
        const unsigned ball = __ballot_sync(0xFFFFFFFFu, m != 0);
        int base = 0;
        if ((tid & 31) == 0 && ball) base = atomicAdd(&S.n_touched, __popc(ball));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (m) S.touched[base + __popc(ball & ((1u << (tid & 31)) - 1u))] = (unsigned short)tid;
    }
    __syncthreads();
    const int n_touched = S.n_touched;
    if (n_touched == 0) return;           // nobody listed an id of this block: the rows are (and stay) zero
    const long long id0 = (long long)blockIdx.x * EX_IDS;
    for (int item = tid; item < n_touched * r4; item += EX_THREADS) {
        const int t = item / r4;
        const int c = item - t * r4;
        const int l = S.touched[t];
        unsigned m = S.present[l];
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        while (m) {                        // ascending source rank = the addition order on every replica
            const int s = __ffs(m) - 1;
            m &= m - 1;
            const int p = record_index(S.ent, s, l);
            const float4* rows = reinterpret_cast<const float4*>(mailbox + slot_offset(L, parity, s) + L.rows_off);
            const float4 v = __ldcg(rows + (size_t)p * r4 + c);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        const float out[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int col = c * 4 + k;
            if (S.tab.ptr[col] != nullptr) stage[S.tab.reg[col] + l * S.tab.ld[col]] = out[k];
        }
    }
    __syncthreads();
    // copy the staged id block out, tensor by tensor (same layout as the dense tensors)
    const int n_local = (int)min((long long)EX_IDS, L.n_ids - id0);
    int off = 0;
#pragma unroll 1
    for (int k = 0; k < T.n; ++k) {
        const int w = T.w[k];
        float* dst = T.p[k] + id0 * w;
        const float* src = stage + off * EX_IDS;
        const int n = n_local * w;
        if ((reinterpret_cast<size_t>(dst) & 15) == 0) {
            const int n4 = n >> 2;
            for (int i = tid; i < n4; i += EX_THREADS) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
            for (int i = (n4 << 2) + tid; i < n; i += EX_THREADS) dst[i] = src[i];
        } else {
            for (int i = tid; i < n; i += EX_THREADS) dst[i] = src[i];
        }
        off += w;
    }
}
static_assert(EX_IDS == EX_THREADS, "the merge kernel maps one thread to one local id");

int fill_tensors(ExTensors& T, float* const* tensors, const int* widths, int n_tensors) {
    if (n_tensors < 1 || n_tensors > EX_MAX_T) return HGS_ERR_INVALID_ARG;
    int row = 0;
    for (int k = 0; k < n_tensors; ++k) {
        if (widths[k] < 1 || tensors[k] == nullptr) return HGS_ERR_INVALID_ARG;
        T.p[k] = tensors[k];
        T.w[k] = widths[k];
        row += widths[k];
    }
    for (int k = n_tensors; k < EX_MAX_T; ++k) { T.p[k] = nullptr; T.w[k] = 0; }
    T.n = n_tensors;
    T.row = (row + 3) / 4 * 4;
    return T.row <= EX_MAX_ROW ? 0 : HGS_ERR_INVALID_ARG;
}

}  // namespace

HGS_API int hgs_exchange_row_floats(const int* widths_host, int n_tensors) {
    if (n_tensors < 1 || n_tensors > EX_MAX_T || widths_host == nullptr) return HGS_ERR_INVALID_ARG;
    int row = 0;
    for (int k = 0; k < n_tensors; ++k) {
        if (widths_host[k] < 1) return HGS_ERR_INVALID_ARG;
        row += widths_host[k];
    }
    row = (row + 3) / 4 * 4;
    return row <= EX_MAX_ROW ? row : HGS_ERR_INVALID_ARG;
}

HGS_API size_t hgs_exchange_mailbox_bytes(int world, long long n_ids, long long cap_rows, int row_floats) {
    if (ex_bad_geometry(world, 0, n_ids, cap_rows) || row_floats < 4 || row_floats > EX_MAX_ROW || row_floats % 4) return 0;
    return EX_CTRL_BYTES + (size_t)2 * world * make_layout(world, n_ids, cap_rows, row_floats).slot_bytes;
}

HGS_API int hgs_exchange_push(float* const* tensors_host, const int* widths_host, int n_tensors, long long n_ids,
                              const int32_t* ids, long long n_rows, long long cap_rows, void* const* mailboxes_host,
                              int world, int rank, unsigned long long step, void* stream) {
    ExTensors T;
    if (int e = fill_tensors(T, tensors_host, widths_host, n_tensors)) return e;
    if (ex_bad_geometry(world, rank, n_ids, cap_rows) || n_rows < 0) return HGS_ERR_INVALID_ARG;
    if (n_rows > 0 && ids == nullptr) return HGS_ERR_INVALID_ARG;
    if (n_rows > cap_rows) n_rows = -1;   // overflow: push no records and a negative row count (see exchange_wait_kernel)
    ExPeers P;
    if (int e = ex_fill_peers(P, mailboxes_host, world, rank)) return e;
    const ExLayout L = make_layout(world, n_ids, cap_rows, T.row);
    const int grid = ex_push_grid(n_rows);
    const int smem = (int)(sizeof(ColTable) + (size_t)2 * EX_CHUNK * T.row * 4);
    cudaError_t e = cudaFuncSetAttribute(exchange_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    exchange_push_kernel<<<grid, EX_THREADS, smem, (cudaStream_t)stream>>>(T, P, L, ids, (int)n_rows, (int)(step & 1ull),
                                                                          step + 1ull);
    HGS_LAUNCH_CHECK();
    return 0;
}

HGS_API int hgs_exchange_reduce(float* const* tensors_host, const int* widths_host, int n_tensors, long long n_ids,
                                long long cap_rows, const void* mailbox, int world, int rank, unsigned long long step,
                                int* status_dev, void* stream) {
    ExTensors T;
    if (int e = fill_tensors(T, tensors_host, widths_host, n_tensors)) return e;
    if (ex_bad_geometry(world, rank, n_ids, cap_rows) || mailbox == nullptr || status_dev == nullptr) return HGS_ERR_INVALID_ARG;
    const ExLayout L = make_layout(world, n_ids, cap_rows, T.row);
    cudaStream_t st = (cudaStream_t)stream;
    exchange_wait_kernel<<<1, 32, 0, st>>>((const unsigned char*)mailbox, L, (int)(step & 1ull), step + 1ull, status_dev);
    HGS_LAUNCH_CHECK();
    const int smem = (int)(sizeof(MergeSmem) + (size_t)EX_IDS * T.row * 4);
    cudaError_t e = cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    exchange_merge_kernel<<<L.n_blocks, EX_THREADS, smem, st>>>(T, L, (const unsigned char*)mailbox, (int)(step & 1ull),
                                                                 status_dev);
    HGS_LAUNCH_CHECK();
    return 0;
}

// ---- peer memory: mailbox allocation and CUDA IPC handles (resource management; these DO allocate / synchronise) ----
HGS_API int hgs_peer_alloc(size_t bytes, void** out) {
    if (out == nullptr || bytes == 0) return HGS_ERR_INVALID_ARG;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(p); return (int)e; }
    *out = p;
    return 0;
}
HGS_API int hgs_peer_free(void* p) { return (int)cudaFree(p); }
HGS_API int hgs_peer_export(void* p, unsigned char* handle64_host) {
    static_assert(sizeof(cudaIpcMemHandle_t) == HGS_PEER_HANDLE_BYTES, "IPC handle size");
    if (p == nullptr || handle64_host == nullptr) return HGS_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) return (int)e;
    memcpy(handle64_host, &h, sizeof(h));
    return 0;
}
HGS_API int hgs_peer_import(const unsigned char* handle64_host, void** out) {
    if (handle64_host == nullptr || out == nullptr) return HGS_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return (int)e;
    *out = p;
    return 0;
}
HGS_API int hgs_peer_close(void* p) { return (int)cudaIpcCloseMemHandle(p); }
