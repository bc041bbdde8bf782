// Multi-GPU exchanges over NCCL (NVLink 5 / NVSwitch), one handle per rank.
//
// The reference has no distributed path; SURVEY.md §8(e) defines this one. Reads shard by contiguous read-id ranges
// (each rank scans its own shard against a replicated k-mer table). The only data that has to cross ranks for the
// sparse A * A^T is the inverted index:
//   1. every rank translates its hits from table slots (which differ between ranks: every rank builds its own table
//      with atomics) to the caller's kmer_id, and sorts its local (kmer_id, global row) incidences by kmer_id;
//      kmer_id ranges are owned by ranks (owner = kmer_id / ceil(K / G));
//   2. ALL-TO-ALL (grouped ncclSend / ncclRecv): each owner receives its kmer_id range from every rank, in rank order,
//      which is also global row order, so one stable sort by kmer_id gives the owner's lists with rows ascending;
//   3. ALL-GATHER (grouped ncclBroadcast, one root per rank): the per-owner CSR pieces are concatenated in owner
//      order into a REPLICATED global inverted index (kmer_id ranges are contiguous, so concatenation is the index);
//   4. ALL-GATHER of the by-read incidence (kmer_id per hit + row offsets): with both sides of A * A^T replicated, rank r
//      counts pairs for the pivot rows r, r + G, r + 2G, ... with the single-GPU rule (partner > pivot, list tails
//      only): every unordered pair is produced exactly once, on exactly one rank, with its FINAL score; interleaving
//      balances the ranks (a contiguous shard of early rows would carry most of the y > x work), and no partial
//      scores ever cross the links. (A first version kept pivots on their shard and chose the endpoint by the parity
//      of x + y: balanced too, but it walks whole lists and was no faster on 2 GPUs than one GPU alone.)
// Edge selection needs two small all-reduces (score histograms) and an all-gather of the tie keys; components
// iterate union-find with all-reduce(min) over the replicated label array.
//
// libnccl is bound at run time (dlopen) so that the library loads on machines without NCCL and shares the copy a
// host process (e.g. torch) has already loaded.
#include "hga_internal.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <dlfcn.h>
#include <nccl.h>

struct hga_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1;
    DevBuf d_small;       // counts / scratch
};

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.lib) return HGA_OK;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { hga_set_error("libnccl.so.2 could not be loaded: %s", dlerror()); return HGA_E_NCCL; }
#define HGA_SYM(field, name)                                                                  \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name));                \
    if (!g_nccl.field) { hga_set_error("libnccl: symbol %s missing", name); return HGA_E_NCCL; }
    HGA_SYM(GetUniqueId, "ncclGetUniqueId");
    HGA_SYM(CommInitRank, "ncclCommInitRank");
    HGA_SYM(CommDestroy, "ncclCommDestroy");
    HGA_SYM(AllReduce, "ncclAllReduce");
    HGA_SYM(AllGather, "ncclAllGather");
    HGA_SYM(Broadcast, "ncclBroadcast");
    HGA_SYM(Send, "ncclSend");
    HGA_SYM(Recv, "ncclRecv");
    HGA_SYM(GroupStart, "ncclGroupStart");
    HGA_SYM(GroupEnd, "ncclGroupEnd");
    HGA_SYM(GetErrorString, "ncclGetErrorString");
#undef HGA_SYM
    g_nccl.lib = lib;
    return HGA_OK;
}

#define HGA_NCCL(call)                                                                                                  \
    do {                                                                                                                \
        ncclResult_t _r = (call);                                                                                       \
        if (_r != ncclSuccess) {                                                                                        \
            hga_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(_r));                \
            return HGA_E_NCCL;                                                                                          \
        }                                                                                                               \
    } while (0)

__global__ void slots_to_kids_kernel2(const uint32_t *__restrict__ slot, const uint32_t *__restrict__ slot_kid, uint64_t n, uint32_t *out) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) out[i] = slot_kid[slot[i]];
}

__global__ void shift_u64_kernel(const uint64_t *__restrict__ in, uint64_t n, uint64_t add, uint64_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) out[i] = in[i] + add;
}

__global__ void global_rows_kernel(const uint64_t *__restrict__ row_off, uint64_t n_rows, uint32_t row_base, uint32_t *__restrict__ out_row) {
    const uint64_t warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    const uint64_t w = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t r = w; r < n_rows; r += warps) {
        const uint64_t a = row_off[r], b = row_off[r + 1];
        for (uint64_t i = a + lane; i < b; i += 32) out_row[i] = (uint32_t) r + row_base;
    }
}

// bound[g] = first position of the sorted slots with slot >= g * per_rank, g = 0 .. G
__global__ void owner_bounds_kernel(const uint32_t *__restrict__ sorted_slot, uint64_t n, uint64_t per_rank, int G, unsigned long long *bound) {
    const int g = threadIdx.x;
    if (g > G) return;
    const uint64_t want = (uint64_t) g * per_rank;
    uint64_t lo = 0, hi = n;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (sorted_slot[mid] < want) lo = mid + 1; else hi = mid; }
    bound[g] = (g == G) ? n : lo;
}

// off[s - first] = base + first position of the sorted keys with key >= s, for the owned slots s in [first, last]
__global__ void owned_offsets_kernel(const uint32_t *__restrict__ keys, uint64_t n, uint32_t first, uint32_t count, uint32_t base, uint32_t *__restrict__ off) {
    uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (; i <= n; i += stride) {
        const int64_t cur = (i < n) ? (int64_t) keys[i] - first : (int64_t) count;    // sentinel closes the tail
        const int64_t prev = (i == 0) ? -1 : (int64_t) keys[i - 1] - first;
        for (int64_t s = prev + 1; s <= cur; s++) if (s >= 0 && s <= (int64_t) count) off[s] = base + (uint32_t) i;
    }
}

}  // namespace

int hga_comm_rank(const hga_handle *h) { return h->comm ? h->comm->rank : 0; }
int hga_comm_size(const hga_handle *h) { return h->comm ? h->comm->size : 1; }

void hga_comm_destroy(hga_handle *h) {
    if (!h->comm) return;
    if (h->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm->comm);
    h->comm->d_small.release();
    delete h->comm;
    h->comm = nullptr;
}

int hga_comm_allreduce_u64_sum(hga_handle *h, uint64_t *d_buf, size_t n) {
    HGA_NCCL(g_nccl.AllReduce(d_buf, d_buf, n, ncclUint64, ncclSum, h->comm->comm, h->stream));
    return HGA_OK;
}
int hga_comm_allreduce_u32_min(hga_handle *h, uint32_t *d_buf, size_t n) {
    HGA_NCCL(g_nccl.AllReduce(d_buf, d_buf, n, ncclUint32, ncclMin, h->comm->comm, h->stream));
    return HGA_OK;
}
int hga_comm_allreduce_u32_max(hga_handle *h, uint32_t *d_buf, size_t n) {
    HGA_NCCL(g_nccl.AllReduce(d_buf, d_buf, n, ncclUint32, ncclMax, h->comm->comm, h->stream));
    return HGA_OK;
}

// counts[g] for every rank g: cnt_all[g] = value contributed by rank g (host array of size G)
int hga_comm_allgather_u64(hga_handle *h, uint64_t mine, std::vector<uint64_t> &all) {
    const int G = h->comm->size;
    HGA_TRY(h->comm->d_small.ensure((size_t) (G + 1) * 8 * 2));
    uint64_t *d_in = h->comm->d_small.as<uint64_t>(), *d_out = d_in + 1;
    HGA_CUDA(cudaMemcpyAsync(d_in, &mine, 8, cudaMemcpyHostToDevice, h->stream));
    HGA_NCCL(g_nccl.AllGather(d_in, d_out, 1, ncclUint64, h->comm->comm, h->stream));
    all.assign(G, 0);
    HGA_CUDA(cudaMemcpyAsync(all.data(), d_out, (size_t) G * 8, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));
    return HGA_OK;
}

// variable-size all-gather: rank g contributes counts[g] elements of elem_bytes (4 or 8) from d_mine; d_all receives the
// concatenation in rank order (one ncclBroadcast per root inside a group)
int hga_comm_allgatherv(hga_handle *h, const void *d_mine, void *d_all, const std::vector<uint64_t> &counts, int elem_bytes) {
    const int G = h->comm->size, me = h->comm->rank;
    const ncclDataType_t dt = elem_bytes == 8 ? ncclUint64 : ncclUint32;
    HGA_NCCL(g_nccl.GroupStart());
    uint64_t base = 0;
    for (int g = 0; g < G; g++) {
        char *dst = reinterpret_cast<char *>(d_all) + base * elem_bytes;
        if (counts[g]) HGA_NCCL(g_nccl.Broadcast(g == me ? d_mine : dst, dst, counts[g], dt, g, h->comm->comm, h->stream));
        base += counts[g];
    }
    HGA_NCCL(g_nccl.GroupEnd());
    return HGA_OK;
}

// Steps 1-3 of the header comment. On return h->d_inv_off / h->d_inv_row hold the replicated global inverted index
// (rows are global row numbers = read id - 1) and h->inc_* describe it.
int hga_comm_build_global_index(hga_handle *h) {
    const int G = h->comm->size, me = h->comm->rank;
    const uint32_t n_slots = (uint32_t) h->n_kmers;    // lists of the exchanged index: one per kmer_id
    const uint64_t E_loc = h->n_hits;
    const uint64_t per_rank = ((uint64_t) n_slots + G - 1) / G;
    const uint32_t row_base = h->read_id_base - 1;
    if (E_loc >= (1ull << 32)) { hga_set_error("local incidence of %llu entries exceeds the 32-bit per-GPU limit", (unsigned long long) E_loc); return HGA_E_OVERFLOW; }

    // 1. local sort by slot (stable: rows stay ascending inside a slot)
    HGA_TRY(h->d_sort_a.ensure((E_loc + 1) * 4));      // sorted slots
    HGA_TRY(h->d_sort_b.ensure((E_loc + 1) * 4));      // global rows, unsorted
    HGA_TRY(h->d_x_row.ensure((E_loc + 1) * 4));       // global rows, sorted
    const int end_bit = (int) std::max<uint32_t>(hga_ceil_log2(h->n_kmers + 1), 1);
    HGA_TRY(h->d_hit_kid.ensure((E_loc + 1) * 4));
    if (E_loc) {
        slots_to_kids_kernel2<<<(int) std::min<uint64_t>((E_loc + 255) / 256, (uint64_t) h->sm_count * 16), 256, 0, h->stream>>>(
            h->d_hit_slot.as<uint32_t>(), h->table.slot_kid, E_loc, h->d_hit_kid.as<uint32_t>());
        h->metrics.kernel_launches++;
        const int blocks = (int) std::min<uint64_t>((h->n_reads * 32 + 255) / 256, (uint64_t) h->sm_count * 32);
        global_rows_kernel<<<blocks, 256, 0, h->stream>>>(h->d_row_off.as<uint64_t>(), h->n_reads, row_base, h->d_sort_b.as<uint32_t>());
        size_t tmp = 0;
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, h->d_hit_kid.as<uint32_t>(), h->d_sort_a.as<uint32_t>(), h->d_sort_b.as<uint32_t>(),
                                                 h->d_x_row.as<uint32_t>(), E_loc, 0, end_bit, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp + 16));
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tmp, h->d_hit_kid.as<uint32_t>(), h->d_sort_a.as<uint32_t>(), h->d_sort_b.as<uint32_t>(),
                                                 h->d_x_row.as<uint32_t>(), E_loc, 0, end_bit, h->stream));
        h->metrics.kernel_launches += (uint64_t) (end_bit + 7) / 8 + 3;
    }
    // owner segments of the sorted incidence
    HGA_TRY(h->comm->d_small.ensure((size_t) (G + 2) * 8 * (G + 2)));
    unsigned long long *d_bound = h->comm->d_small.as<unsigned long long>();
    owner_bounds_kernel<<<1, 64, 0, h->stream>>>(h->d_sort_a.as<uint32_t>(), E_loc, per_rank, G, d_bound);
    h->metrics.kernel_launches++;
    HGA_CUDA(cudaGetLastError());
    std::vector<unsigned long long> bound(G + 1);
    HGA_CUDA(cudaMemcpyAsync(bound.data(), d_bound, (size_t) (G + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));

    // 2. all-to-all: counts first (all-gather of every rank's G send counts), then the payload
    unsigned long long *d_cnt_in = d_bound + (G + 2), *d_cnt_all = d_cnt_in + (G + 2);
    std::vector<unsigned long long> send_cnt(G), cnt_all((size_t) G * G);
    for (int g = 0; g < G; g++) send_cnt[g] = bound[g + 1] - bound[g];
    HGA_CUDA(cudaMemcpyAsync(d_cnt_in, send_cnt.data(), (size_t) G * 8, cudaMemcpyHostToDevice, h->stream));
    HGA_NCCL(g_nccl.AllGather(d_cnt_in, d_cnt_all, G, ncclUint64, h->comm->comm, h->stream));
    HGA_CUDA(cudaMemcpyAsync(cnt_all.data(), d_cnt_all, (size_t) G * G * 8, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));
    uint64_t E_own = 0, E_total = 0;
    std::vector<uint64_t> recv_off(G + 1, 0), own_cnt(G, 0);
    for (int src = 0; src < G; src++) { recv_off[src] = E_own; E_own += cnt_all[(size_t) src * G + me]; }
    for (int src = 0; src < G; src++) for (int g = 0; g < G; g++) { own_cnt[g] += cnt_all[(size_t) src * G + g]; E_total += cnt_all[(size_t) src * G + g]; }
    if (E_total >= (1ull << 32)) { hga_set_error("global incidence of %llu entries exceeds the 32-bit limit of the replicated index", (unsigned long long) E_total); return HGA_E_OVERFLOW; }

    StageTimer xt(h, &h->metrics.exchange_ms, true);
    HGA_TRY(h->d_x_slot.ensure((E_own + 1) * 4 * 2));   // received slots | received rows
    uint32_t *rx_slot = h->d_x_slot.as<uint32_t>(), *rx_row = rx_slot + (E_own + 1);
    HGA_NCCL(g_nccl.GroupStart());
    for (int g = 0; g < G; g++) {
        if (send_cnt[g]) {
            HGA_NCCL(g_nccl.Send(h->d_sort_a.as<uint32_t>() + bound[g], send_cnt[g], ncclUint32, g, h->comm->comm, h->stream));
            HGA_NCCL(g_nccl.Send(h->d_x_row.as<uint32_t>() + bound[g], send_cnt[g], ncclUint32, g, h->comm->comm, h->stream));
        }
        const uint64_t rc = cnt_all[(size_t) g * G + me];
        if (rc) {
            HGA_NCCL(g_nccl.Recv(rx_slot + recv_off[g], rc, ncclUint32, g, h->comm->comm, h->stream));
            HGA_NCCL(g_nccl.Recv(rx_row + recv_off[g], rc, ncclUint32, g, h->comm->comm, h->stream));
        }
    }
    HGA_NCCL(g_nccl.GroupEnd());

    // owner: stable sort of the received runs by slot (sources arrive in global row order)
    HGA_TRY(h->d_sort_a.ensure((E_own + 1) * 4));
    HGA_TRY(h->d_sort_b.ensure((E_own + 1) * 4));
    uint32_t *own_slot = h->d_sort_a.as<uint32_t>(), *own_row = h->d_sort_b.as<uint32_t>();
    if (E_own) {
        size_t tmp = 0;
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, rx_slot, own_slot, rx_row, own_row, E_own, 0, end_bit, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp + 16));
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tmp, rx_slot, own_slot, rx_row, own_row, E_own, 0, end_bit, h->stream));
        h->metrics.kernel_launches += (uint64_t) (end_bit + 7) / 8 + 2;
    }

    // 3. replicated global index: rows by grouped broadcast, offsets computed by the owner and broadcast as well
    HGA_TRY(h->d_inv_off.ensure(((size_t) per_rank * G + 2) * 4));
    HGA_TRY(h->d_inv_row.ensure((E_total + 1) * 4));
    uint32_t *inv_off = h->d_inv_off.as<uint32_t>(), *inv_row = h->d_inv_row.as<uint32_t>();
    HGA_CUDA(cudaMemsetAsync(inv_off, 0, ((size_t) per_rank * G + 2) * 4, h->stream));
    uint64_t my_base = 0;
    for (int g = 0; g < me; g++) my_base += own_cnt[g];
    const uint64_t first = (uint64_t) me * per_rank;
    const uint64_t owned = first >= n_slots ? 0 : std::min<uint64_t>(per_rank, n_slots - first);
    {
        // offsets of my slot range go straight to their final place in the global array; entry [first + owned] of the last
        // non-empty range closes the index
        const int blocks = (int) std::min<uint64_t>((E_own + 256) / 256, (uint64_t) h->sm_count * 16);
        if (owned) owned_offsets_kernel<<<blocks, 256, 0, h->stream>>>(own_slot, E_own, (uint32_t) first, (uint32_t) owned, (uint32_t) my_base, inv_off + first);
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
    }
    HGA_NCCL(g_nccl.GroupStart());
    uint64_t base = 0;
    for (int g = 0; g < G; g++) {
        if (own_cnt[g]) HGA_NCCL(g_nccl.Broadcast(g == me ? own_row : inv_row + base, inv_row + base, own_cnt[g], ncclUint32, g, h->comm->comm, h->stream));
        const uint64_t f = (uint64_t) g * per_rank;
        const uint64_t cnt = f >= n_slots ? 0 : std::min<uint64_t>(per_rank, n_slots - f);
        // ranges overlap by one entry (the closing offset of range g is the opening offset of range g + 1, same value)
        if (cnt) HGA_NCCL(g_nccl.Broadcast(inv_off + f, inv_off + f, cnt + (f + cnt == n_slots ? 1 : 0), ncclUint32, g, h->comm->comm, h->stream));
        base += own_cnt[g];
    }
    HGA_NCCL(g_nccl.GroupEnd());
    xt.stop();

    h->inc_rows = h->n_reads_total;
    h->inc_row_first_id = 1;
    h->inc_entries = E_total;
    // 4. replicated by-read incidence
    {
        StageTimer gt(h, &h->metrics.exchange_ms, true);
        std::vector<uint64_t> hit_cnt, row_cnt;
        HGA_TRY(hga_comm_allgather_u64(h, E_loc, hit_cnt));
        HGA_TRY(hga_comm_allgather_u64(h, h->n_reads, row_cnt));
        uint64_t hit_base = 0, n_rows_all = 0;
        for (int g = 0; g < me; g++) hit_base += hit_cnt[g];
        for (int g = 0; g < G; g++) n_rows_all += row_cnt[g];
        uint64_t rows_before = 0;
        for (int g = 0; g < me; g++) rows_before += row_cnt[g];
        if (rows_before != row_base) { hga_set_error("hga_build_index: shards must be contiguous read-id ranges in rank order (rank %d starts at row %u, expected %llu)", me, row_base, (unsigned long long) rows_before); return HGA_E_ARG; }
        if (n_rows_all != h->n_reads_total) { hga_set_error("hga_build_index: the ranks scanned %llu reads, hga_comm_init said %llu", (unsigned long long) n_rows_all, (unsigned long long) h->n_reads_total); return HGA_E_ARG; }
        HGA_TRY(h->d_g_kid.ensure((E_total + 1) * 4));
        HGA_TRY(h->d_g_row_off.ensure((n_rows_all + 2) * 8));
        HGA_TRY(h->d_x_slot.ensure((h->n_reads + 2) * 8));          // my row offsets, shifted to global positions
        shift_u64_kernel<<<(int) std::min<uint64_t>((h->n_reads + 256) / 256, 2048), 256, 0, h->stream>>>(h->d_row_off.as<uint64_t>(), h->n_reads, hit_base,
                                                                                                     h->d_x_slot.as<uint64_t>());
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
        HGA_TRY(hga_comm_allgatherv(h, h->d_hit_kid.p, h->d_g_kid.p, hit_cnt, 4));
        HGA_TRY(hga_comm_allgatherv(h, h->d_x_slot.p, h->d_g_row_off.p, row_cnt, 8));
        const uint64_t e_total = E_total;
        HGA_CUDA(cudaMemcpyAsync(h->d_g_row_off.as<uint64_t>() + n_rows_all, &e_total, 8, cudaMemcpyHostToDevice, h->stream));
        const double first_ms = h->metrics.exchange_ms;
        gt.stop();
        h->metrics.exchange_ms += first_ms;
    }
    h->pair_rows = h->n_reads_total;
    h->pair_pivot_mul = (uint32_t) G; h->pair_pivot_add = (uint32_t) me;
    h->index_by_kid = true;
    h->index_keys = n_slots;
    return HGA_OK;
}

extern "C" int hga_comm_unique_id(void *id128) {
    if (!id128) { hga_set_error("hga_comm_unique_id: NULL buffer"); return HGA_E_ARG; }
    HGA_TRY(load_nccl());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    HGA_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return HGA_OK;
}

extern "C" int hga_comm_init(hga_handle *h, const void *id128, int rank, int nranks, uint64_t n_reads_total) {
    if (!h || !id128) { hga_set_error("hga_comm_init: NULL argument"); return HGA_E_ARG; }
    if (nranks < 1 || rank < 0 || rank >= nranks || nranks > 62) { hga_set_error("hga_comm_init: bad rank %d / %d", rank, nranks); return HGA_E_ARG; }
    if (n_reads_total >= (1ull << 32) - 1) { hga_set_error("hga_comm_init: more than 2^32-2 reads"); return HGA_E_ARG; }
    HGA_TRY(load_nccl());
    HGA_CUDA(cudaSetDevice(h->device));
    hga_comm_destroy(h);
    hga_comm *c = new hga_comm();
    c->rank = rank; c->size = nranks;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) { hga_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); delete c; return HGA_E_NCCL; }
    h->comm = c;
    h->n_reads_total = n_reads_total;
    return HGA_OK;
}
