// Device construction of the discriminative k-mer membership structures.
// Replaces std::unordered_set<Kmer>::contains + KmerIndex of the reference
// (clustering/ReadClusteringEngine.cpp:237-241, :251, :263).
#include "hga_internal.cuh"

#include <cstdlib>

namespace {

__global__ void table_insert_kernel(const uint64_t *__restrict__ kmers, uint64_t n, unsigned long long *keys, uint32_t *slot_kid, uint32_t *kid_slot,
                                    uint32_t n_groups, unsigned long long *filter, uint32_t n_words, int *flags) {
    uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t key = kmers[i];
        if (key == HGA_EMPTY_KEY) { atomicOr(flags, 2); continue; }   // never a canonical k-mer
        KmerHash hs = hga_hash(key);
        uint32_t g = hga_scale(hs.hi, n_groups);
        bool done = false;
        for (uint32_t probes = 0; probes < n_groups && !done; probes++) {
            for (int j = 0; j < 4 && !done; j++) {
                uint32_t slot = g * 4 + j;
                unsigned long long old = atomicCAS(&keys[slot], (unsigned long long) HGA_EMPTY_KEY, (unsigned long long) key);
                if (old == HGA_EMPTY_KEY) {
                    slot_kid[slot] = (uint32_t) i;
                    kid_slot[i] = slot;
                    done = true;
                } else if (old == key) {
                    atomicOr(flags, 1);   // duplicate
                    kid_slot[i] = slot;
                    done = true;
                }
            }
            g = (g + 1 == n_groups) ? 0 : g + 1;
        }
        uint32_t m0, m1;
        hga_filter_mask(hs.lo, m0, m1);
        atomicOr(&filter[hga_scale(hs.hi, n_words)], ((unsigned long long) m1 << 32) | m0);
    }
}

}  // namespace

int hga_table_build(hga_handle *h, const uint64_t *host_kmers) {
    uint64_t n = h->n_kmers;
    if (n >= (1ull << 30)) { hga_set_error("too many k-mers (%llu): the slot id space is 32 bit", (unsigned long long) n); return HGA_E_ARG; }
    double bits_per_key = 16.0, max_mb = 64.0;
    if (const char *e = getenv("HGA_FILTER_BITS_PER_KEY")) bits_per_key = atof(e);
    if (const char *e = getenv("HGA_FILTER_MAX_MB")) max_mb = atof(e);
    uint64_t n_groups = (2 * n + 3) / 4 + 1;           // load factor <= 0.5
    uint64_t n_words = (uint64_t) (n * bits_per_key / 64.0) + 1024;
    uint64_t max_words = (uint64_t) (max_mb * 1024 * 1024 / 8);
    if (n_words > max_words) n_words = max_words;

    KmerTable &t = h->table;
    t.n_groups = (uint32_t) n_groups;
    t.n_slots = (uint32_t) (n_groups * 4);
    t.n_words = (uint32_t) n_words;
    t.slot_bits = hga_ceil_log2(t.n_slots);
    HGA_TRY(h->d_keys.ensure((size_t) t.n_slots * 8));
    HGA_TRY(h->d_slot_kid.ensure((size_t) t.n_slots * 4));
    HGA_TRY(h->d_kid_slot.ensure((size_t) (n + 1) * 4));
    HGA_TRY(h->d_filter.ensure((size_t) n_words * 8));
    t.keys = h->d_keys.as<uint64_t>();
    t.slot_kid = h->d_slot_kid.as<uint32_t>();
    t.kid_slot = h->d_kid_slot.as<uint32_t>();
    t.filter = h->d_filter.as<uint64_t>();

    DevBuf d_in, d_flags;
    HGA_TRY(d_in.ensure((size_t) (n + 1) * 8));
    HGA_TRY(d_flags.ensure(16));
    StageTimer timer(h, &h->metrics.table_build_ms);
    HGA_CUDA(cudaMemcpyAsync(d_in.p, host_kmers, n * 8, cudaMemcpyHostToDevice, h->stream));
    HGA_CUDA(cudaMemsetAsync(t.keys, 0xff, (size_t) t.n_slots * 8, h->stream));
    HGA_CUDA(cudaMemsetAsync(t.slot_kid, 0xff, (size_t) t.n_slots * 4, h->stream));
    HGA_CUDA(cudaMemsetAsync(t.filter, 0, (size_t) n_words * 8, h->stream));
    HGA_CUDA(cudaMemsetAsync(d_flags.p, 0, 16, h->stream));
    if (n) {
        int blocks = (int) ((n + 255) / 256);
        if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
        table_insert_kernel<<<blocks, 256, 0, h->stream>>>(d_in.as<uint64_t>(), n, (unsigned long long *) t.keys, t.slot_kid, t.kid_slot, t.n_groups,
                                                          (unsigned long long *) t.filter, t.n_words, d_flags.as<int>());
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
    }
    int flags = 0;
    HGA_CUDA(cudaMemcpyAsync(&flags, d_flags.p, 4, cudaMemcpyDeviceToHost, h->stream));
    timer.stop();
    d_in.release(); d_flags.release();
    if (flags & 2) { hga_set_error("k-mer value 0xFFFFFFFFFFFFFFFF is not a canonical k-mer"); return HGA_E_ARG; }
    if (flags & 1) { hga_set_error("duplicate k-mer in the set handed to hga_create"); return HGA_E_DUPLICATE; }
    h->metrics.table_bytes = (uint64_t) t.n_slots * 8;
    h->metrics.filter_bytes = n_words * 8;
    return HGA_OK;
}
