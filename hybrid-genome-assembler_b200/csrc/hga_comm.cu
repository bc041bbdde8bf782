// Multi-GPU exchanges over NCCL (NVLink 5 / NVSwitch), one handle per rank.
//
// The reference has no distributed path; SURVEY.md §8(e) defines this one. Reads shard by contiguous read-id ranges (each rank
// scans its own shard against a replicated k-mer table); the inverted index is PARTITIONED by k-mer owner and the pair scores
// are reduced at the owner of x:
//   1. the table layout is a function of the k-mer array (hga_table.cu), so a slot number means the same k-mer on every rank: whole
//      32-slot buckets are dealt round robin, owner(slot) = (slot / 32) mod G, list number at the owner = (slot / 32) / G * 32 + slot mod 32
//      (the hits of a minimizer run keep neighbouring list numbers: the pair counter's locality survives the partition);
//   2. ALL-TO-ALL 1 (grouped ncclSend / ncclRecv): (list number, global row) records go to their owner. Every source sends in
//      row order and the sources arrive in rank order = global row order, so the received stream IS the owner's by-row
//      incidence (row offsets from run boundaries, no sort), and one stable sort by list number gives its inverted lists with
//      rows ascending. Nothing is replicated: a rank holds E / G incidence entries whatever G is;
//   3. every rank runs the single-GPU pair kernels over ALL rows as pivots, each row restricted to the hits of the rank's own
//      k-mers (y > x, list tails only): PARTIAL scores, work = the increments of the owned lists = 1 / G of the total;
//   4. ALL-TO-ALL 2: partial (x, y, score) records go to owner(x) = x mod G, where one radix sort by (x, y) and a segmented sum
//      give the final scores: every unordered pair ends up on exactly one rank, in canonical order.
// Edge selection needs two small all-reduces (score histograms) and an all-gather of the tie keys; components iterate
// union-find with all-reduce(min) over the replicated label array.
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

// record = list number at the owner << 32 | global row; one warp per row
__global__ void pack_records_kernel(const uint64_t *__restrict__ row_off, uint64_t n_rows, uint32_t row_base, const uint32_t *__restrict__ slot,
                                    uint32_t G, uint64_t *__restrict__ rec, uint8_t *__restrict__ owner) {
    const uint64_t warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    const uint64_t w = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t r = w; r < n_rows; r += warps) {
        const uint64_t a = row_off[r], b = row_off[r + 1];
        for (uint64_t i = a + lane; i < b; i += 32) {
            const uint32_t s = slot[i];
            rec[i] = ((uint64_t) hga_list_of_slot(s, G) << 32) | ((uint32_t) r + row_base);
            owner[i] = (uint8_t) hga_owner_of_slot(s, G);
        }
    }
}

__global__ void unpack_records_kernel(const uint64_t *__restrict__ rec, uint64_t n, uint32_t *__restrict__ hi, uint32_t *__restrict__ lo) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t v = rec[i];
        if (hi) hi[i] = (uint32_t) (v >> 32);
        if (lo) lo[i] = (uint32_t) v;
    }
}

// send counts of a partitioned array: cnt[g] = number of elements whose (sorted) destination byte is g, g = 0 .. G - 1
__global__ void dest_counts_kernel(const uint8_t *__restrict__ sorted_dest, uint64_t n, int G, unsigned long long *cnt) {
    const int g = threadIdx.x;
    if (g >= G) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (sorted_dest[mid] < g) lo = mid + 1; else hi = mid; }
    uint64_t lo2 = lo, hi2 = n;
    while (lo2 < hi2) { const uint64_t mid = (lo2 + hi2) >> 1; if (sorted_dest[mid] <= g) lo2 = mid + 1; else hi2 = mid; }
    cnt[g] = lo2 - lo;
}

// off[s] = first position of the ascending values (field of a 64-bit record) with value >= s, s = 0 .. count
template<typename OFF, int SHIFT>
__global__ void field_offsets_kernel(const uint64_t *__restrict__ rec, uint64_t n, uint64_t count, OFF *__restrict__ off) {
    uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (; i <= n; i += stride) {
        const int64_t cur = (i < n) ? (int64_t) (uint32_t) (rec[i] >> SHIFT) : (int64_t) count;    // sentinel closes the tail
        const int64_t prev = (i == 0) ? -1 : (int64_t) (uint32_t) (rec[i - 1] >> SHIFT);
        for (int64_t s = prev + 1; s <= cur && s <= (int64_t) count; s++) off[s] = (OFF) i;
    }
}

__global__ void pair_dest_kernel(const uint64_t *__restrict__ key, uint64_t n, uint32_t G, uint8_t *__restrict__ dest) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) dest[i] = (uint8_t) ((uint32_t) (key[i] >> 32) % G);
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

// counts all-to-all: every rank's G send counts + one extra value (device array of G + 1) -> the G x (G + 1) matrix on the host
// (cnt_all[src * (G + 1) + dst], extra at [src * (G + 1) + G])
static int exchange_counts(hga_handle *h, const unsigned long long *d_send_cnt, unsigned long long *d_all, std::vector<unsigned long long> &cnt_all) {
    const int G = h->comm->size;
    cnt_all.assign((size_t) G * (G + 1), 0);
    HGA_NCCL(g_nccl.AllGather(d_send_cnt, d_all, G + 1, ncclUint64, h->comm->comm, h->stream));
    HGA_CUDA(cudaMemcpyAsync(cnt_all.data(), d_all, (size_t) G * (G + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));
    return HGA_OK;
}

// Steps 1-2 of the header comment. On return h->d_inv_off / h->d_inv_row hold the inverted lists of THIS rank's k-mers (list number
// = kmer_id / G, rows = global row numbers = read id - 1) and h->d_g_row_off / h->d_g_kid the by-row incidence of all rows
// restricted to those k-mers.
int hga_comm_build_owner_index(hga_handle *h) {
    const int G = h->comm->size, me = h->comm->rank;
    const uint64_t E_loc = h->n_hits, R_all = h->n_reads_total;
    const uint32_t n_buckets_all = h->table.n_slots / HGA_BUCKET_SLOTS;                         // n_slots is a multiple of 32
    const uint32_t n_lists = (n_buckets_all + G - 1) / G * HGA_BUCKET_SLOTS;                    // of the fullest owner; a multiple of 32
    const uint32_t row_base = h->read_id_base - 1;
    if (E_loc >= (1ull << 32)) { hga_set_error("local incidence of %llu entries exceeds the 32-bit per-GPU limit", (unsigned long long) E_loc); return HGA_E_OVERFLOW; }
    double comm_ms = 0, part_ms = 0;
    Trace tr(h);

    // 1. packed records partitioned by owner (ONE stable radix pass: every owner segment keeps global row order)
    const int owner_bits = (int) std::max<uint32_t>(hga_ceil_log2((uint64_t) G), 1);
    HGA_TRY(h->d_sort_a.ensure((E_loc + 1) * 8));      // packed records, partitioned
    HGA_TRY(h->d_sort_b.ensure((E_loc + 1) * 8));      // packed records, stream order
    HGA_TRY(h->d_x_row.ensure((E_loc + 1) * 2));       // owner per record: in | out
    uint8_t *own_in = h->d_x_row.as<uint8_t>(), *own_out = own_in + (E_loc + 1);
    uint64_t *rec_in = h->d_sort_b.as<uint64_t>(), *rec_out = h->d_sort_a.as<uint64_t>();
    if (E_loc) {
        const int blocks = (int) std::min<uint64_t>((h->n_reads * 32 + 255) / 256, (uint64_t) h->sm_count * 32);
        pack_records_kernel<<<blocks, 256, 0, h->stream>>>(h->d_row_off.as<uint64_t>(), h->n_reads, row_base, h->d_hit_slot.as<uint32_t>(), (uint32_t) G, rec_in, own_in);
        size_t tmp = 0;
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, own_in, own_out, rec_in, rec_out, E_loc, 0, owner_bits, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp + 16));
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tmp, own_in, own_out, rec_in, rec_out, E_loc, 0, owner_bits, h->stream));
        h->metrics.kernel_launches += 4;
        HGA_CUDA(cudaGetLastError());
    }
    HGA_TRY(h->comm->d_small.ensure((size_t) (G + 2) * 8 * (G + 2)));
    unsigned long long *d_cnt = h->comm->d_small.as<unsigned long long>(), *d_cnt_all = d_cnt + (G + 2);
    dest_counts_kernel<<<1, 64, 0, h->stream>>>(own_out, E_loc, G, d_cnt);
    const unsigned long long my_rows = h->n_reads;                                  // rides along: the shard layout check below needs every rank's row count
    HGA_CUDA(cudaMemcpyAsync(d_cnt + G, &my_rows, 8, cudaMemcpyHostToDevice, h->stream));
    h->metrics.kernel_launches++;
    HGA_CUDA(cudaGetLastError());
    tr.mark("pack+partition");
    std::vector<unsigned long long> cnt_all;
    HGA_TRY(exchange_counts(h, d_cnt, d_cnt_all, cnt_all));                         // the only host synchronisation of the stage
    tr.mark("counts");
    const size_t CS = (size_t) G + 1;
    {   // shards must be contiguous read-id ranges in rank order: the row stream of source s must lie in s's range
        uint64_t rows_before = 0, tot = 0;
        for (int g = 0; g < G; g++) { if (g < me) rows_before += cnt_all[g * CS + G]; tot += cnt_all[g * CS + G]; }
        if (rows_before != row_base) { hga_set_error("hga_build_index: shards must be contiguous read-id ranges in rank order (rank %d starts at row %u, expected %llu)", me, row_base, (unsigned long long) rows_before); return HGA_E_ARG; }
        if (tot != R_all) { hga_set_error("hga_build_index: the ranks scanned %llu reads, hga_comm_init said %llu", (unsigned long long) tot, (unsigned long long) R_all); return HGA_E_ARG; }
    }

    // 2. all-to-all: the records of my lists from every rank, in rank order = global row order
    uint64_t E_own = 0;
    std::vector<uint64_t> recv_off(G + 1, 0), send_off(G + 1, 0);
    for (int src = 0; src < G; src++) { recv_off[src] = E_own; E_own += cnt_all[src * CS + me]; }
    for (int g = 0; g < G; g++) send_off[g + 1] = send_off[g] + cnt_all[me * CS + g];
    if (E_own >= (1ull << 32)) { hga_set_error("this rank's share of the incidence (%llu entries) exceeds the 32-bit per-GPU limit", (unsigned long long) E_own); return HGA_E_OVERFLOW; }
    HGA_TRY(h->d_x_slot.ensure((E_own + 1) * 8));       // received records
    uint64_t *rx = h->d_x_slot.as<uint64_t>();
    {
        StageTimer xt(h, &part_ms, true);
        HGA_NCCL(g_nccl.GroupStart());
        for (int g = 0; g < G; g++) {
            const uint64_t sc = cnt_all[me * CS + g], rc = cnt_all[g * CS + me];
            if (sc) HGA_NCCL(g_nccl.Send(rec_out + send_off[g], sc, ncclUint64, g, h->comm->comm, h->stream));
            if (rc) HGA_NCCL(g_nccl.Recv(rx + recv_off[g], rc, ncclUint64, g, h->comm->comm, h->stream));
        }
        HGA_NCCL(g_nccl.GroupEnd());
        xt.stop();
        comm_ms += part_ms;
    }
    tr.mark("alltoall");

    // by-row incidence = the received stream (list numbers in row order + row offsets from the run boundaries); inverted lists = the
    // single-GPU list builder over (list number, row)
    HGA_TRY(h->d_g_row_off.ensure((R_all + 2) * 8));
    HGA_TRY(h->d_g_kid.ensure((E_own + 1) * 4));
    HGA_TRY(h->d_sort_b.ensure((E_own + 1) * 4));       // rows in stream order (scratch of the list builder)
    {
        const int blocks = (int) std::min<uint64_t>((E_own + 256) / 256, (uint64_t) h->sm_count * 16);
        field_offsets_kernel<uint64_t, 0><<<blocks, 256, 0, h->stream>>>(rx, E_own, R_all, h->d_g_row_off.as<uint64_t>());
        if (E_own) unpack_records_kernel<<<blocks, 256, 0, h->stream>>>(rx, E_own, h->d_g_kid.as<uint32_t>(), h->d_sort_b.as<uint32_t>());
        h->metrics.kernel_launches += 2;
        HGA_CUDA(cudaGetLastError());
    }
    tr.mark("unpack");
    HGA_TRY(hga_build_lists(h, h->d_g_kid.as<uint32_t>(), h->d_sort_b.as<uint32_t>(), E_own, n_lists));
    tr.mark("lists");
    h->inc_rows = R_all;
    h->inc_row_first_id = 1;
    h->inc_entries = E_own;
    h->pair_rows = R_all;
    h->pair_pivot_mul = 1; h->pair_pivot_add = 0;
    h->index_by_kid = true;
    h->index_keys = n_lists;
    h->index_key_div = 0;
    h->metrics.exchange_ms = comm_ms;       // NCCL payload calls only (the sorts between them belong to index_ms)
    tr.dump("index", me);
    return HGA_OK;
}

// Step 4 of the header comment: n partial (key = x << 32 | y, score) records of this rank -> owner(x) = x mod G. On return
// *out_n records (unsorted, pairs may repeat: one partial per contributing rank) are in h->d_pair_key2 / h->d_pair_score2.
int hga_comm_exchange_partials(hga_handle *h, uint64_t n, uint64_t *out_n) {
    const int G = h->comm->size, me = h->comm->rank;
    const int owner_bits = (int) std::max<uint32_t>(hga_ceil_log2((uint64_t) G), 1);
    double part_ms = 0;
    Trace tr(h);
    HGA_TRY(h->d_x_row.ensure((n + 1) * 2));
    HGA_TRY(h->d_pair_key2.ensure((n + 1) * 8));
    HGA_TRY(h->d_pair_score2.ensure((n + 1) * 4));
    uint8_t *d_in = h->d_x_row.as<uint8_t>(), *d_out = d_in + (n + 1);
    uint64_t *key_part = h->d_pair_key2.as<uint64_t>();
    uint32_t *score_part = h->d_pair_score2.as<uint32_t>();
    if (n) {
        pair_dest_kernel<<<(int) std::min<uint64_t>((n + 255) / 256, (uint64_t) h->sm_count * 16), 256, 0, h->stream>>>(h->d_pair_key.as<uint64_t>(), n, (uint32_t) G, d_in);
        size_t t1 = 0, t2 = 0;
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, d_in, d_out, h->d_pair_key.as<uint64_t>(), key_part, n, 0, owner_bits, h->stream));
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t2, d_in, d_out, h->d_pair_score.as<uint32_t>(), score_part, n, 0, owner_bits, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(std::max(t1, t2) + 16));
        // two stable passes with the same keys: the same permutation for both value arrays
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, t1, d_in, d_out, h->d_pair_key.as<uint64_t>(), key_part, n, 0, owner_bits, h->stream));
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, t2, d_in, d_out, h->d_pair_score.as<uint32_t>(), score_part, n, 0, owner_bits, h->stream));
        h->metrics.kernel_launches += 7;
        HGA_CUDA(cudaGetLastError());
    }
    HGA_TRY(h->comm->d_small.ensure((size_t) (G + 2) * 8 * (G + 2)));
    unsigned long long *d_cnt = h->comm->d_small.as<unsigned long long>(), *d_cnt_all = d_cnt + (G + 2);
    dest_counts_kernel<<<1, 64, 0, h->stream>>>(d_out, n, G, d_cnt);
    h->metrics.kernel_launches++;
    HGA_CUDA(cudaGetLastError());
    tr.mark("partition");
    std::vector<unsigned long long> cnt_all;
    HGA_TRY(exchange_counts(h, d_cnt, d_cnt_all, cnt_all));
    tr.mark("counts");
    uint64_t n_recv = 0;
    std::vector<uint64_t> recv_off(G + 1, 0), send_off(G + 1, 0);
    const size_t CS = (size_t) G + 1;
    for (int src = 0; src < G; src++) { recv_off[src] = n_recv; n_recv += cnt_all[src * CS + me]; }
    for (int g = 0; g < G; g++) send_off[g + 1] = send_off[g] + cnt_all[me * CS + g];
    // the receive buffers: the (now free) primary pair arrays
    HGA_TRY(h->d_pair_key.ensure((n_recv + 1) * 8));
    HGA_TRY(h->d_pair_score.ensure((n_recv + 1) * 4));
    {
        StageTimer xt(h, &part_ms, true);
        HGA_NCCL(g_nccl.GroupStart());
        for (int g = 0; g < G; g++) {
            const uint64_t sc = cnt_all[me * CS + g], rc = cnt_all[g * CS + me];
            if (sc) {
                HGA_NCCL(g_nccl.Send(key_part + send_off[g], sc, ncclUint64, g, h->comm->comm, h->stream));
                HGA_NCCL(g_nccl.Send(score_part + send_off[g], sc, ncclUint32, g, h->comm->comm, h->stream));
            }
            if (rc) {
                HGA_NCCL(g_nccl.Recv(h->d_pair_key.as<uint64_t>() + recv_off[g], rc, ncclUint64, g, h->comm->comm, h->stream));
                HGA_NCCL(g_nccl.Recv(h->d_pair_score.as<uint32_t>() + recv_off[g], rc, ncclUint32, g, h->comm->comm, h->stream));
            }
        }
        HGA_NCCL(g_nccl.GroupEnd());
        xt.stop();
        h->metrics.exchange_ms += part_ms;
    }
    *out_n = n_recv;
    tr.mark("alltoall");
    if (tr.on) fprintf(stderr, "[hga trace r%d partials] sent=%llu received=%llu\n", me, (unsigned long long) n, (unsigned long long) n_recv);
    tr.dump("partials", me);
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
