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
//   4. ALL-TO-ALL of the by-read incidence: row x (its kmer_ids) goes to rank x mod G. Rank r then counts pairs for the
//      pivot rows r, r + G, r + 2G, ... with the single-GPU rule (partner > pivot, list tails only): every unordered
//      pair is produced exactly once, on exactly one rank, with its FINAL score; interleaving balances the ranks (a
//      contiguous shard of early rows would carry most of the y > x work), and no partial scores ever cross the links. (A first version kept pivots on their shard and chose the endpoint by the parity
//      of x + y: balanced too, but it walks whole lists and was no faster on 2 GPUs than one GPU alone.)
// Edge selection needs two small all-reduces (score histograms) and an all-gather of the tie keys; components
// iterate union-find with all-reduce(min) over the replicated label array.
//
// libnccl is bound at run time (dlopen) so that the library loads on machines without NCCL and shares the copy a
// host process (e.g. torch) has already loaded.
#include "hga_internal.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <dlfcn.h>
#include <nccl.h>

struct hga_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1;
    DevBuf d_small;       // counts / scratch
    DevBuf d_stage;       // padded all-gather staging
    DevBuf d_rows;        // per-row lengths / destinations of the by-read exchange
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

// table slot -> index key of the k-mer (kmer_id + kmer_id / per_rank)
__global__ void slots_to_keys_kernel(const uint32_t *__restrict__ slot, const uint32_t *__restrict__ slot_kid, uint64_t n, uint32_t per_rank, uint32_t *out) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t kid = slot_kid[slot[i]];
        out[i] = kid + kid / per_rank;
    }
}

// per row: hit count and destination rank (global row mod G); per hit: the destination of its row. One warp per row.
__global__ void row_dest_kernel(const uint64_t *__restrict__ row_off, uint64_t n_rows, uint32_t row_base, uint32_t G, uint32_t *__restrict__ len,
                                uint8_t *__restrict__ row_dest, uint8_t *__restrict__ hit_dest) {
    const uint64_t warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    const uint64_t w = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t r = w; r < n_rows; r += warps) {
        const uint64_t a = row_off[r], b = row_off[r + 1];
        const uint8_t d = (uint8_t) (((uint32_t) r + row_base) % G);
        if (lane == 0) { len[r] = (uint32_t) (b - a); row_dest[r] = d; }
        for (uint64_t i = a + lane; i < b; i += 32) hit_dest[i] = d;
    }
}

__global__ void shift_u64_kernel(const uint64_t *__restrict__ in, uint64_t n, uint64_t add, uint64_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) out[i] = in[i] + add;
}

// record = index key << 32 | global row, owner = key / keys_per_rank; one warp per row
__global__ void pack_records_kernel(const uint64_t *__restrict__ row_off, uint64_t n_rows, uint32_t row_base, const uint32_t *__restrict__ kid, uint32_t per_rank,
                                    uint64_t *__restrict__ rec, uint8_t *__restrict__ owner) {
    const uint64_t warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    const uint64_t w = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t r = w; r < n_rows; r += warps) {
        const uint64_t a = row_off[r], b = row_off[r + 1];
        for (uint64_t i = a + lane; i < b; i += 32) {
            const uint32_t kk = kid[i];
            rec[i] = ((uint64_t) kk << 32) | ((uint32_t) r + row_base);
            owner[i] = (uint8_t) (kk / per_rank);
        }
    }
}

__global__ void unpack_records_kernel(const uint64_t *__restrict__ rec, uint64_t n, uint32_t *__restrict__ kid, uint32_t *__restrict__ row) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t v = rec[i];
        kid[i] = (uint32_t) (v >> 32); row[i] = (uint32_t) v;
    }
}

// bound[g] = first position of the partitioned records whose owner is >= g, g = 0 .. G
__global__ void owner_bounds_kernel(const uint8_t *__restrict__ sorted_owner, uint64_t n, int G, unsigned long long *bound) {
    const int g = threadIdx.x;
    if (g > G) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (sorted_owner[mid] < g) lo = mid + 1; else hi = mid; }
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
    h->comm->d_stage.release();
    h->comm->d_rows.release();
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
// concatenation in rank order. Large payloads go through ONE ncclAllGather on segments padded to the largest count
// (the ring / NVLS path that reaches NVLink bandwidth) followed by device-to-device compaction copies; small ones use one
// ncclBroadcast per root inside a group.
int hga_comm_allgatherv(hga_handle *h, const void *d_mine, void *d_all, const std::vector<uint64_t> &counts, int elem_bytes) {
    const int G = h->comm->size, me = h->comm->rank;
    const ncclDataType_t dt = elem_bytes == 8 ? ncclUint64 : ncclUint32;
    uint64_t maxc = 0, total = 0;
    for (uint64_t c : counts) { maxc = std::max(maxc, c); total += c; }
    if (total * elem_bytes >= (8ull << 20)) {
        HGA_TRY(h->comm->d_stage.ensure((size_t) G * maxc * elem_bytes + 256));
        char *stage = h->comm->d_stage.as<char>();
        if (counts[me]) HGA_CUDA(cudaMemcpyAsync(stage + (size_t) me * maxc * elem_bytes, d_mine, counts[me] * elem_bytes, cudaMemcpyDeviceToDevice, h->stream));
        HGA_NCCL(g_nccl.AllGather(stage + (size_t) me * maxc * elem_bytes, stage, maxc, dt, h->comm->comm, h->stream));
        uint64_t base = 0;
        for (int g = 0; g < G; g++) {
            if (counts[g]) HGA_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(d_all) + base * elem_bytes, stage + (size_t) g * maxc * elem_bytes, counts[g] * elem_bytes,
                                                    cudaMemcpyDeviceToDevice, h->stream));
            base += counts[g];
        }
        return HGA_OK;
    }
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

// Steps 1-4 of the header comment. On return h->d_inv_off / h->d_inv_row hold the replicated global inverted index (rows
// are global row numbers = read id - 1) keyed by hga_index_key(kmer_id), h->d_g_row_off / h->d_g_kid hold the by-read
// incidence of THIS rank's pivot rows (me, me + G, me + 2G, ...), and h->inc_* / pair_* describe both.
int hga_comm_build_global_index(hga_handle *h) {
    const int G = h->comm->size, me = h->comm->rank;
    const uint64_t K = h->n_kmers;
    const uint64_t E_loc = h->n_hits;
    // index key of a k-mer: kmer_id + kmer_id / per_rank, i.e. every owner's range of per_rank k-mers is followed by one
    // unused key. Its offset entry closes the owner's last list, so the owners' offset arrays AND their row arrays (padded
    // to the longest) can be all-gathered in place with equal counts - no compaction, no per-root broadcasts.
    const uint64_t per_rank = std::max<uint64_t>((K + G - 1) / G, 1), kpr = per_rank + 1, n_keys = kpr * G;
    const uint32_t row_base = h->read_id_base - 1;
    if (E_loc >= (1ull << 32)) { hga_set_error("local incidence of %llu entries exceeds the 32-bit per-GPU limit", (unsigned long long) E_loc); return HGA_E_OVERFLOW; }
    double comm_ms = 0, part_ms = 0;

    // 1. hits keyed by index key; ONE radix pass partitions the packed (key << 32 | global row) records by owner (stable,
    //    so every owner segment keeps global row order)
    const int key_bits = (int) std::max<uint32_t>(hga_ceil_log2(n_keys + 1), 1);
    const int owner_bits = (int) std::max<uint32_t>(hga_ceil_log2((uint64_t) G), 1);
    HGA_TRY(h->d_hit_kid.ensure((E_loc + 1) * 4));
    HGA_TRY(h->d_sort_a.ensure((E_loc + 1) * 8));      // packed records, partitioned
    HGA_TRY(h->d_sort_b.ensure((E_loc + 1) * 8));      // packed records, stream order
    HGA_TRY(h->d_x_row.ensure((E_loc + 1) * 2));       // owner per record: in | out
    uint8_t *own_in = h->d_x_row.as<uint8_t>(), *own_out = own_in + (E_loc + 1);
    uint64_t *rec_in = h->d_sort_b.as<uint64_t>(), *rec_out = h->d_sort_a.as<uint64_t>();
    if (E_loc) {
        slots_to_keys_kernel<<<(int) std::min<uint64_t>((E_loc + 255) / 256, (uint64_t) h->sm_count * 16), 256, 0, h->stream>>>(
            h->d_hit_slot.as<uint32_t>(), h->table.slot_kid, E_loc, (uint32_t) per_rank, h->d_hit_kid.as<uint32_t>());
        const int blocks = (int) std::min<uint64_t>((h->n_reads * 32 + 255) / 256, (uint64_t) h->sm_count * 32);
        pack_records_kernel<<<blocks, 256, 0, h->stream>>>(h->d_row_off.as<uint64_t>(), h->n_reads, row_base, h->d_hit_kid.as<uint32_t>(), (uint32_t) kpr, rec_in, own_in);
        size_t tmp = 0;
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, own_in, own_out, rec_in, rec_out, E_loc, 0, owner_bits, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp + 16));
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tmp, own_in, own_out, rec_in, rec_out, E_loc, 0, owner_bits, h->stream));
        h->metrics.kernel_launches += 5;
        HGA_CUDA(cudaGetLastError());
    }
    HGA_TRY(h->comm->d_small.ensure((size_t) (4 * G + 8) * 8 * (G + 2)));
    unsigned long long *d_bound = h->comm->d_small.as<unsigned long long>();
    owner_bounds_kernel<<<1, 64, 0, h->stream>>>(own_out, E_loc, G, d_bound);
    h->metrics.kernel_launches++;
    HGA_CUDA(cudaGetLastError());
    std::vector<unsigned long long> bound(G + 1);
    HGA_CUDA(cudaMemcpyAsync(bound.data(), d_bound, (size_t) (G + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));

    // 2. all-to-all: counts first (all-gather of every rank's G send counts), then the payload
    unsigned long long *d_cnt_in = d_bound + (G + 2), *d_cnt_all = d_cnt_in + (2 * G + 2);
    std::vector<unsigned long long> send_cnt(G), cnt_all((size_t) G * G);
    for (int g = 0; g < G; g++) send_cnt[g] = bound[g + 1] - bound[g];
    HGA_CUDA(cudaMemcpyAsync(d_cnt_in, send_cnt.data(), (size_t) G * 8, cudaMemcpyHostToDevice, h->stream));
    HGA_NCCL(g_nccl.AllGather(d_cnt_in, d_cnt_all, G, ncclUint64, h->comm->comm, h->stream));
    HGA_CUDA(cudaMemcpyAsync(cnt_all.data(), d_cnt_all, (size_t) G * G * 8, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));
    uint64_t E_own = 0, E_total = 0, maxc = 0;
    std::vector<uint64_t> recv_off(G + 1, 0), own_cnt(G, 0);
    for (int src = 0; src < G; src++) { recv_off[src] = E_own; E_own += cnt_all[(size_t) src * G + me]; }
    for (int src = 0; src < G; src++) for (int g = 0; g < G; g++) { own_cnt[g] += cnt_all[(size_t) src * G + g]; E_total += cnt_all[(size_t) src * G + g]; }
    for (int g = 0; g < G; g++) maxc = std::max(maxc, own_cnt[g]);
    if (maxc * G >= (1ull << 32)) { hga_set_error("global incidence of %llu entries exceeds the 32-bit limit of the replicated index", (unsigned long long) E_total); return HGA_E_OVERFLOW; }

    HGA_TRY(h->d_x_slot.ensure((E_own + 1) * 8 * 2));   // received records | sorted records
    uint64_t *rx = h->d_x_slot.as<uint64_t>(), *rx_sorted = rx + (E_own + 1);
    {
        StageTimer xt(h, &part_ms, true);
        HGA_NCCL(g_nccl.GroupStart());
        for (int g = 0; g < G; g++) {
            if (send_cnt[g]) HGA_NCCL(g_nccl.Send(rec_out + bound[g], send_cnt[g], ncclUint64, g, h->comm->comm, h->stream));
            const uint64_t rc = cnt_all[(size_t) g * G + me];
            if (rc) HGA_NCCL(g_nccl.Recv(rx + recv_off[g], rc, ncclUint64, g, h->comm->comm, h->stream));
        }
        HGA_NCCL(g_nccl.GroupEnd());
        xt.stop();
        comm_ms += part_ms;
    }

    // owner: stable sort of the received runs by key (sources arrive in global row order); the rows go straight to this
    // owner's segment of the replicated row array, the offsets to its segment of the replicated offset array
    HGA_TRY(h->d_inv_off.ensure((n_keys + 2) * 4));
    HGA_TRY(h->d_inv_row.ensure(((size_t) G * maxc + 4) * 4));
    HGA_TRY(h->d_sort_a.ensure((E_own + 1) * 4));
    uint32_t *inv_off = h->d_inv_off.as<uint32_t>(), *inv_row = h->d_inv_row.as<uint32_t>();
    uint32_t *own_key = h->d_sort_a.as<uint32_t>(), *own_row = inv_row + (size_t) me * maxc;
    if (E_own) {
        size_t tmp = 0;
        HGA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp, rx, rx_sorted, E_own, 32, 32 + key_bits, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp + 16));
        HGA_CUDA(cub::DeviceRadixSort::SortKeys(h->d_sort_tmp.p, tmp, rx, rx_sorted, E_own, 32, 32 + key_bits, h->stream));
        unpack_records_kernel<<<(int) std::min<uint64_t>((E_own + 255) / 256, (uint64_t) h->sm_count * 16), 256, 0, h->stream>>>(rx_sorted, E_own, own_key, own_row);
        h->metrics.kernel_launches += (uint64_t) (key_bits + 7) / 8 + 3;
        HGA_CUDA(cudaGetLastError());
    }
    {
        const int blocks = (int) std::min<uint64_t>((E_own + 256) / 256, (uint64_t) h->sm_count * 16);
        owned_offsets_kernel<<<blocks, 256, 0, h->stream>>>(own_key, E_own, (uint32_t) (me * kpr), (uint32_t) per_rank, (uint32_t) (me * maxc), inv_off + me * kpr);
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
    }
    // 3. replicated global index: two in-place all-gathers with equal counts
    {
        StageTimer xt(h, &part_ms, true);
        if (maxc) HGA_NCCL(g_nccl.AllGather(own_row, inv_row, maxc, ncclUint32, h->comm->comm, h->stream));
        HGA_NCCL(g_nccl.AllGather(inv_off + me * kpr, inv_off, kpr, ncclUint32, h->comm->comm, h->stream));
        xt.stop();
        comm_ms += part_ms;
    }
    h->inc_rows = h->n_reads_total;
    h->inc_row_first_id = 1;
    h->inc_entries = E_total;

    // 4. by-read incidence of the pivot rows: row x goes to rank x mod G (all-to-all, E / G entries per rank)
    {
        const uint64_t R = h->n_reads;
        HGA_TRY(h->comm->d_rows.ensure((R + 1) * (2 + 8) + 64));
        uint32_t *len_in = h->comm->d_rows.as<uint32_t>(), *len_out = len_in + (R + 1);
        uint8_t *rd_in = reinterpret_cast<uint8_t *>(len_out + (R + 1)), *rd_out = rd_in + (R + 1);
        uint8_t *hd_in = own_in, *hd_out = own_out;                                  // per-hit destination (reuses the owner bytes)
        uint32_t *kid_part = h->d_sort_b.as<uint32_t>();                             // hits partitioned by destination
        if (R) {
            const int blocks = (int) std::min<uint64_t>((R * 32 + 255) / 256, (uint64_t) h->sm_count * 32);
            row_dest_kernel<<<blocks, 256, 0, h->stream>>>(h->d_row_off.as<uint64_t>(), R, row_base, (uint32_t) G, len_in, rd_in, hd_in);
            size_t tmp = 0, tmp2 = 0;
            HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, hd_in, hd_out, h->d_hit_kid.as<uint32_t>(), kid_part, E_loc, 0, owner_bits, h->stream));
            HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, rd_in, rd_out, len_in, len_out, R, 0, owner_bits, h->stream));
            HGA_TRY(h->d_sort_tmp.ensure(std::max(tmp, tmp2) + 16));
            if (E_loc) HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tmp, hd_in, hd_out, h->d_hit_kid.as<uint32_t>(), kid_part, E_loc, 0, owner_bits, h->stream));
            HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tmp2, rd_in, rd_out, len_in, len_out, R, 0, owner_bits, h->stream));
            h->metrics.kernel_launches += 7;
            HGA_CUDA(cudaGetLastError());
        }
        unsigned long long *d_hb = d_bound, *d_rb = d_bound + (G + 2);               // (d_cnt_in region is free again)
        owner_bounds_kernel<<<1, 64, 0, h->stream>>>(hd_out, E_loc, G, d_hb);
        owner_bounds_kernel<<<1, 64, 0, h->stream>>>(rd_out, R, G, d_rb);
        std::vector<unsigned long long> hb(G + 1), rb(G + 1), mine(2 * G), all((size_t) 2 * G * G);
        HGA_CUDA(cudaMemcpyAsync(hb.data(), d_hb, (size_t) (G + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaMemcpyAsync(rb.data(), d_rb, (size_t) (G + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        for (int g = 0; g < G; g++) { mine[g] = hb[g + 1] - hb[g]; mine[G + g] = rb[g + 1] - rb[g]; }
        unsigned long long *d_mine = d_bound + 2 * (G + 2), *d_all = d_mine + (2 * G + 2);
        HGA_CUDA(cudaMemcpyAsync(d_mine, mine.data(), (size_t) 2 * G * 8, cudaMemcpyHostToDevice, h->stream));
        HGA_NCCL(g_nccl.AllGather(d_mine, d_all, 2 * G, ncclUint64, h->comm->comm, h->stream));
        HGA_CUDA(cudaMemcpyAsync(all.data(), d_all, (size_t) 2 * G * G * 8, cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        uint64_t my_hits = 0, my_rows = 0, rows_before = 0, rows_all = 0;
        std::vector<uint64_t> h_off(G), r_off(G);
        for (int src = 0; src < G; src++) {
            h_off[src] = my_hits; r_off[src] = my_rows;
            my_hits += all[(size_t) src * 2 * G + me]; my_rows += all[(size_t) src * 2 * G + G + me];
            uint64_t rows_src = 0;
            for (int g = 0; g < G; g++) rows_src += all[(size_t) src * 2 * G + G + g];
            if (src < me) rows_before += rows_src;
            rows_all += rows_src;
        }
        if (rows_before != row_base) { hga_set_error("hga_build_index: shards must be contiguous read-id ranges in rank order (rank %d starts at row %u, expected %llu)", me, row_base, (unsigned long long) rows_before); return HGA_E_ARG; }
        if (rows_all != h->n_reads_total) { hga_set_error("hga_build_index: the ranks scanned %llu reads, hga_comm_init said %llu", (unsigned long long) rows_all, (unsigned long long) h->n_reads_total); return HGA_E_ARG; }
        const uint64_t expect_rows = h->n_reads_total > (uint64_t) me ? (h->n_reads_total - me + G - 1) / G : 0;
        if (my_rows != expect_rows) { hga_set_error("hga_build_index: received %llu pivot rows, expected %llu (internal error)", (unsigned long long) my_rows, (unsigned long long) expect_rows); return HGA_E_STATE; }
        HGA_TRY(h->d_g_kid.ensure((my_hits + 1) * 4));
        HGA_TRY(h->d_g_row_off.ensure((my_rows + 2) * 8 + (my_rows + 2) * 4));
        uint64_t *g_row_off = h->d_g_row_off.as<uint64_t>();
        uint32_t *len_rx = reinterpret_cast<uint32_t *>(g_row_off + (my_rows + 2));
        {
            StageTimer xt(h, &part_ms, true);
            HGA_NCCL(g_nccl.GroupStart());
            for (int g = 0; g < G; g++) {
                if (mine[g]) HGA_NCCL(g_nccl.Send(kid_part + hb[g], mine[g], ncclUint32, g, h->comm->comm, h->stream));
                if (mine[G + g]) HGA_NCCL(g_nccl.Send(len_out + rb[g], mine[G + g], ncclUint32, g, h->comm->comm, h->stream));
                const uint64_t hc = all[(size_t) g * 2 * G + me], rc = all[(size_t) g * 2 * G + G + me];
                if (hc) HGA_NCCL(g_nccl.Recv(h->d_g_kid.as<uint32_t>() + h_off[g], hc, ncclUint32, g, h->comm->comm, h->stream));
                if (rc) HGA_NCCL(g_nccl.Recv(len_rx + r_off[g], rc, ncclUint32, g, h->comm->comm, h->stream));
            }
            HGA_NCCL(g_nccl.GroupEnd());
            xt.stop();
            comm_ms += part_ms;
        }
        // CSR offsets of the received rows (sources arrive in rank order = ascending row number)
        HGA_CUDA(cudaMemsetAsync(len_rx + my_rows, 0, 4, h->stream));
        size_t tmp = 0;
        HGA_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, tmp, len_rx, g_row_off, cub::Sum(), 0ull, my_rows + 1, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp + 16));
        HGA_CUDA(cub::DeviceScan::ExclusiveScan(h->d_sort_tmp.p, tmp, len_rx, g_row_off, cub::Sum(), 0ull, my_rows + 1, h->stream));
        h->metrics.kernel_launches += 4;
        h->pair_rows = my_rows;
    }
    h->pair_pivot_mul = (uint32_t) G; h->pair_pivot_add = (uint32_t) me;
    h->index_by_kid = true;
    h->index_keys = (uint32_t) n_keys;
    h->index_key_div = (uint32_t) per_rank;
    h->metrics.exchange_ms = comm_ms;       // NCCL calls only (the sorts between them belong to index_ms)
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
