// Inverted index (k-mer -> rows) from the by-read incidence.
// Replaces the kmer_component_index half of construct_indices
// (clustering/ReadClusteringEngine.cpp:262-269 push_back per occurrence, :282-284 sort of every list).
//
// A stable LSD radix sort of (slot, row) pairs by slot keeps rows ascending inside every list because the
// incidence is already ordered by row; offsets come from run boundaries of the sorted keys.
#include "hga_internal.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace {

// one warp per row: out_row[row_off[r] .. row_off[r+1]) = r
__global__ void expand_rows_kernel(const uint64_t *__restrict__ row_off, uint64_t n_rows, uint32_t *__restrict__ out_row) {
    const uint64_t warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    const uint64_t w = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t r = w; r < n_rows; r += warps) {
        const uint64_t a = row_off[r], b = row_off[r + 1];
        for (uint64_t i = a + lane; i < b; i += 32) out_row[i] = (uint32_t) r;
    }
}

// inv_off[s] = first position whose key is >= s (keys sorted ascending)
__global__ void run_offsets_kernel(const uint32_t *__restrict__ keys, uint64_t n, uint32_t n_slots, uint32_t *__restrict__ inv_off) {
    uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (; i <= n; i += stride) {
        const uint32_t cur = (i < n) ? keys[i] : n_slots;          // sentinel closes the tail
        const int64_t prev = (i == 0) ? -1 : (int64_t) keys[i - 1];
        for (int64_t s = prev + 1; s <= (int64_t) cur; s++) inv_off[s] = (uint32_t) i;
    }
}

}  // namespace

int hga_index_run(hga_handle *h) {
    if (!h->have_scan) { hga_set_error("hga_build_index: no scan result"); return HGA_E_STATE; }
    h->have_index = h->have_pairs = h->have_selection = h->have_components = false;
    const uint32_t n_slots = h->table.n_slots;

    if (h->comm && hga_comm_size(h) > 1) {
        // sharded reads: all-to-all by k-mer owner into a PARTITIONED index (hga_comm.cu)
        StageTimer timer(h, &h->metrics.index_ms);
        HGA_TRY(hga_comm_build_owner_index(h));
        timer.stop();
        h->have_index = true;
        return HGA_OK;
    }
    h->inc_rows = h->n_reads;
    h->pair_rows = h->n_reads; h->pair_pivot_mul = 1; h->pair_pivot_add = 0;
    h->index_by_kid = false; h->index_keys = n_slots;
    h->inc_row_first_id = h->read_id_base;
    h->inc_entries = h->n_hits;
    const uint64_t E = h->inc_entries;
    if (E >= (1ull << 32)) { hga_set_error("incidence of %llu entries exceeds the 32-bit per-GPU limit", (unsigned long long) E); return HGA_E_OVERFLOW; }

    StageTimer timer(h, &h->metrics.index_ms);
    HGA_TRY(h->d_inv_off.ensure(((size_t) n_slots + 2) * 4));
    HGA_TRY(h->d_inv_row.ensure((E + 1) * 4));
    HGA_TRY(h->d_sort_a.ensure((E + 1) * 4));     // sorted keys
    uint32_t *inv_off = h->d_inv_off.as<uint32_t>();

    if (E > 0) {
        HGA_TRY(h->d_sort_b.ensure((E + 1) * 4));
        uint32_t *d_rows = h->d_sort_b.as<uint32_t>();
        const int blocks = (int) std::min<uint64_t>((h->inc_rows * 32 + 255) / 256, (uint64_t) h->sm_count * 32);
        expand_rows_kernel<<<blocks, 256, 0, h->stream>>>(h->d_row_off.as<uint64_t>(), h->inc_rows, d_rows);
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
        size_t tmp_bytes = 0;
        const int end_bit = (int) std::max<uint32_t>(h->table.slot_bits, 1);
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->d_hit_slot.as<uint32_t>(), h->d_sort_a.as<uint32_t>(), d_rows, h->d_inv_row.as<uint32_t>(), E, 0,
                                                 end_bit, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp_bytes + 16));
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tmp_bytes, h->d_hit_slot.as<uint32_t>(), h->d_sort_a.as<uint32_t>(), d_rows,
                                                 h->d_inv_row.as<uint32_t>(), E, 0, end_bit, h->stream));
        h->metrics.kernel_launches += (uint64_t) (end_bit + 7) / 8 + 2;   // CUB: histogram + onesweep passes
        const int oblocks = (int) std::min<uint64_t>((E + 256) / 256, (uint64_t) h->sm_count * 16);
        run_offsets_kernel<<<oblocks, 256, 0, h->stream>>>(h->d_sort_a.as<uint32_t>(), E, n_slots, inv_off);
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
    } else {
        HGA_CUDA(cudaMemsetAsync(inv_off, 0, ((size_t) n_slots + 1) * 4, h->stream));
    }
    timer.stop();
    h->have_index = true;
    return HGA_OK;
}
