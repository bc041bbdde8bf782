// Device construction of the discriminative k-mer membership structures (layout: hga_internal.cuh).
// Replaces std::unordered_set<Kmer>::contains + KmerIndex of the reference
// (clustering/ReadClusteringEngine.cpp:237-241, :251, :263).
#include "hga_internal.cuh"

#include <cstdlib>

namespace {

// pass 1: filter bits + main (locality-bucketed) key table; keys whose bucket is full are listed for pass 2
__global__ void table_insert_main_kernel(const uint64_t *__restrict__ kmers, uint64_t n, KmerTable t, uint32_t *over_list, unsigned int *over_count, int *flags) {
    uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(t.keys);
    for (; i < n; i += stride) {
        const unsigned long long key = kmers[i];
        if (key == HGA_EMPTY_KEY) { atomicOr(flags, 2); continue; }   // never a canonical k-mer
        const uint32_t hb = hga_bits_hash(key, t.geom);
        const uint32_t B = hga_locality_from_min(hga_minimizer(key, hb, t.geom));
        atomicOr(&t.filter[(size_t) hga_scale(B, t.n_blocks) * 8 + hga_bits_word(hb)], hga_bits_mask(hb, t.filter_k));
        const uint32_t home = hga_scale(B, t.n_buckets) * HGA_BUCKET_SLOTS, start = hga_start_sector(B, hb, t.sector_by_min) * HGA_SECTOR_SLOTS;
        bool done = false;
        for (uint32_t j = 0; j < HGA_BUCKET_SLOTS && !done; j++) {
            const uint32_t slot = home + ((start + j) & (HGA_BUCKET_SLOTS - 1));
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&keys[slot]);
            if (cur == HGA_EMPTY_KEY) cur = atomicCAS(&keys[slot], (unsigned long long) HGA_EMPTY_KEY, key);
            if (cur == HGA_EMPTY_KEY) { t.kid_slot[i] = slot; done = true; }
            else if (cur == key) { atomicOr(flags, 1); t.kid_slot[i] = slot; done = true; }   // duplicate
        }
        if (!done) over_list[atomicAdd(over_count, 1u)] = (uint32_t) i;
    }
}

// pass 2: overflow region, plain open addressing by k-mer hash (n_over is a power of two, load <= 0.5)
__global__ void table_insert_over_kernel(const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ over_list, uint32_t n_list, KmerTable t, int *flags) {
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(t.keys) + t.n_main;
    const uint32_t mask = t.n_over - 1;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_list; q += gridDim.x * blockDim.x) {
        const uint32_t i = over_list[q];
        const unsigned long long key = kmers[i];
        uint32_t pos = hga_plain_hash(key) & mask;
        for (;;) {
            const unsigned long long old = atomicCAS(&keys[pos], (unsigned long long) HGA_EMPTY_KEY, key);
            if (old == HGA_EMPTY_KEY) { t.kid_slot[i] = t.n_main + pos; break; }
            if (old == key) { atomicOr(flags, 1); t.kid_slot[i] = t.n_main + pos; break; }
            pos = (pos + 1) & mask;
        }
    }
}

__global__ void table_slot_kid_kernel(const uint32_t *__restrict__ kid_slot, uint64_t n, uint32_t *slot_kid) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) slot_kid[kid_slot[i]] = (uint32_t) i;
}

}  // namespace

int hga_table_build(hga_handle *h, const uint64_t *host_kmers) {
    const uint64_t n = h->n_kmers;
    if (n >= (1ull << 29)) { hga_set_error("too many k-mers (%llu): the slot id space is 32 bit", (unsigned long long) n); return HGA_E_ARG; }
    // filter: sized to stay L2 resident (<= 48 MB, see hga_internal.cuh); HGA_FILTER_* / HGA_TABLE_LOAD are experiment switches
    double bits_per_key = 16.0, max_mb = 48.0, load = 0.25;
    if (const char *e = getenv("HGA_FILTER_BITS_PER_KEY")) bits_per_key = atof(e);
    if (const char *e = getenv("HGA_FILTER_MAX_MB")) max_mb = atof(e);
    if (const char *e = getenv("HGA_TABLE_LOAD")) load = std::min(0.9, std::max(0.05, atof(e)));
    uint64_t n_blocks = (uint64_t) (n * bits_per_key / 256.0) + 64;
    const uint64_t max_blocks = (uint64_t) (max_mb * 1024 * 1024 / 32);
    if (n_blocks > max_blocks) n_blocks = max_blocks;

    KmerTable &t = h->table;
    t.geom = hga_make_geom(h->k, n);
    t.n_buckets = (uint32_t) ((uint64_t) (n / load) / HGA_BUCKET_SLOTS + 1);
    t.n_main = t.n_buckets * HGA_BUCKET_SLOTS;
    t.n_over = 0;
    t.n_blocks = (uint32_t) n_blocks;
    if (const char *e = getenv("HGA_SECTOR_BY_MIN")) t.sector_by_min = atoi(e) != 0;
    if (const char *e = getenv("HGA_FILTER_K")) t.filter_k = atoi(e) == 2 ? 2 : 3;
    HGA_TRY(h->d_keys.ensure((size_t) t.n_main * 8));
    HGA_TRY(h->d_kid_slot.ensure((size_t) (n + 1) * 4));
    HGA_TRY(h->d_filter.ensure((size_t) n_blocks * 32));
    t.keys = h->d_keys.as<uint64_t>();
    t.kid_slot = h->d_kid_slot.as<uint32_t>();
    t.filter = h->d_filter.as<uint32_t>();

    DevBuf d_in, d_flags, d_over;
    HGA_TRY(d_in.ensure((size_t) (n + 1) * 8));
    HGA_TRY(d_over.ensure((size_t) (n + 1) * 4));
    HGA_TRY(d_flags.ensure(16));
    int *d_fl = d_flags.as<int>();
    unsigned int *d_cnt = reinterpret_cast<unsigned int *>(d_fl + 1);
    StageTimer timer(h, &h->metrics.table_build_ms);
    HGA_CUDA(cudaMemcpyAsync(d_in.p, host_kmers, n * 8, cudaMemcpyHostToDevice, h->stream));
    HGA_CUDA(cudaMemsetAsync(t.keys, 0xff, (size_t) t.n_main * 8, h->stream));
    HGA_CUDA(cudaMemsetAsync(t.filter, 0, (size_t) n_blocks * 32, h->stream));
    HGA_CUDA(cudaMemsetAsync(d_flags.p, 0, 16, h->stream));
    int blocks = (int) std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t) h->sm_count * 16));
    if (n) {
        table_insert_main_kernel<<<blocks, 256, 0, h->stream>>>(d_in.as<uint64_t>(), n, t, d_over.as<uint32_t>(), d_cnt, d_fl);
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
    }
    int host_flags[2] = {0, 0};
    HGA_CUDA(cudaMemcpyAsync(host_flags, d_flags.p, 8, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));
    const uint32_t n_list = (uint32_t) host_flags[1];
    if (n_list > 0) {
        // grow the key array by the overflow region (main region preserved)
        uint32_t n_over = 64;
        while (n_over < 2 * (uint64_t) n_list) n_over <<= 1;
        DevBuf grown;
        HGA_TRY(grown.ensure(((size_t) t.n_main + n_over) * 8));
        HGA_CUDA(cudaMemcpyAsync(grown.p, t.keys, (size_t) t.n_main * 8, cudaMemcpyDeviceToDevice, h->stream));
        HGA_CUDA(cudaMemsetAsync(grown.as<uint64_t>() + t.n_main, 0xff, (size_t) n_over * 8, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        h->d_keys.release();
        h->d_keys = grown;
        t.keys = h->d_keys.as<uint64_t>();
        t.n_over = n_over;
        table_insert_over_kernel<<<(int) std::min<uint32_t>((n_list + 255) / 256, (uint32_t) h->sm_count * 16), 256, 0, h->stream>>>(
            d_in.as<uint64_t>(), d_over.as<uint32_t>(), n_list, t, d_fl);
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
        HGA_CUDA(cudaMemcpyAsync(host_flags, d_flags.p, 4, cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
    }
    t.n_slots = t.n_main + t.n_over;
    t.slot_bits = hga_ceil_log2(t.n_slots);
    HGA_TRY(h->d_slot_kid.ensure((size_t) t.n_slots * 4));
    t.slot_kid = h->d_slot_kid.as<uint32_t>();
    HGA_CUDA(cudaMemsetAsync(t.slot_kid, 0xff, (size_t) t.n_slots * 4, h->stream));
    if (n) {
        table_slot_kid_kernel<<<blocks, 256, 0, h->stream>>>(t.kid_slot, n, t.slot_kid);
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
    }
    timer.stop();
    d_in.release(); d_flags.release(); d_over.release();
    if (host_flags[0] & 2) { hga_set_error("k-mer value 0xFFFFFFFFFFFFFFFF is not a canonical k-mer"); return HGA_E_ARG; }
    if (host_flags[0] & 1) { hga_set_error("duplicate k-mer in the set handed to hga_create"); return HGA_E_DUPLICATE; }
    h->metrics.table_bytes = (uint64_t) t.n_slots * 8;
    h->metrics.filter_bytes = n_blocks * 32;
    h->metrics.table_overflow_keys = n_list;
    return HGA_OK;
}
