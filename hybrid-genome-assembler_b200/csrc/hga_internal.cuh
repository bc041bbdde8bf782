// Internal declarations shared by the CUDA translation units of libhga_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hga_b200.h"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
void hga_set_error(const char *fmt, ...);

#define HGA_CUDA(call)                                                                                          \
    do {                                                                                                        \
        cudaError_t _e = (call);                                                                                \
        if (_e != cudaSuccess) {                                                                                \
            hga_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));           \
            return HGA_E_CUDA;                                                                                  \
        }                                                                                                       \
    } while (0)

#define HGA_TRY(call)                   \
    do {                                \
        int _rc = (call);               \
        if (_rc != HGA_OK) return _rc;  \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device / pinned buffers that only grow
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return HGA_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + (bytes >> 4) + 256;   // a little slack so repeated runs do not realloc
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            hga_set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            cudaGetLastError();
            return HGA_E_NOMEM;
        }
        cap = want;
        return HGA_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template<typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return HGA_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + 64;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) {
            hga_set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
            cudaGetLastError();
            return HGA_E_NOMEM;
        }
        cap = want;
        return HGA_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template<typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ------------------------------------------------------------------------------------------------
// k-mer membership structures (device)
//   key table : groups of 4 x u64 canonical k-mers (one 32 B sector), group-wise linear probing, load <= 0.5.
//               The internal k-mer id ("slot") is the index of the key in this table; slot_kid maps it back
//               to the caller's kmer_id.
//   filter    : blocked Bloom filter, one 64-bit word per k-mer, 4 bits (2 per 32-bit half). Sized to stay
//               L2-resident; consulted before the key table so that non-members cost one 8 B probe.
// ------------------------------------------------------------------------------------------------
struct KmerTable {
    uint64_t *keys = nullptr;       // n_groups * 4
    uint32_t *slot_kid = nullptr;   // n_groups * 4
    uint32_t *kid_slot = nullptr;   // n_kmers
    uint64_t *filter = nullptr;     // n_words
    uint32_t n_groups = 0;
    uint32_t n_words = 0;
    uint32_t n_slots = 0;
    uint32_t slot_bits = 0;         // ceil(log2(n_slots))
};

#define HGA_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define HGA_HASH_MULT 0x9E3779B97F4A7C15ull

struct KmerHash {
    uint32_t hi;     // selects filter word and home group
    uint32_t lo;     // selects the bits inside the filter word
};

__host__ __device__ __forceinline__ KmerHash hga_hash(uint64_t kmer) {
    uint64_t t = kmer * HGA_HASH_MULT;
    KmerHash h;
    h.hi = (uint32_t) (t >> 32);
    h.lo = (uint32_t) t;
    return h;
}

// 2 bits in each 32-bit half of the filter word, taken from the top 20 bits of h.lo
__host__ __device__ __forceinline__ void hga_filter_mask(uint32_t lo, uint32_t &m0, uint32_t &m1) {
    m0 = (1u << ((lo >> 27) & 31)) | (1u << ((lo >> 22) & 31));
    m1 = (1u << ((lo >> 17) & 31)) | (1u << ((lo >> 12) & 31));
}

__host__ __device__ __forceinline__ uint32_t hga_scale(uint32_t h, uint32_t n) {
#ifdef __CUDA_ARCH__
    return __umulhi(h, n);
#else
    return (uint32_t) (((uint64_t) h * n) >> 32);
#endif
}

// ------------------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------------------
struct hga_comm;   // NCCL state, hga_comm.cu

struct hga_handle {
    int device = 0;
    int k = 0;
    uint64_t n_kmers = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 148;

    // table
    KmerTable table;
    DevBuf d_keys, d_slot_kid, d_kid_slot, d_filter;

    // scan results (rows = reads of this rank's shard)
    uint64_t n_reads = 0, n_bases = 0, n_hits = 0;
    uint32_t read_id_base = 1;
    DevBuf d_bases, d_read_off;           // staging for hga_scan (host entry)
    DevBuf d_row_off;                     // u64[n_reads+1]
    DevBuf d_hit_slot, d_hit_pos;         // u32[E]
    DevBuf d_tile_state, d_scan_scalars;
    bool have_scan = false;

    // incidence the pair counter works on (local rows for 1 GPU; global rows x owned k-mers with a comm)
    uint64_t inc_rows = 0;                // number of rows in the incidence
    uint32_t inc_row_first_id = 1;        // read id of row 0
    uint64_t inc_entries = 0;
    uint64_t *inc_row_off = nullptr;      // u64[inc_rows+1]  (aliases d_row_off or d_x_row_off)
    uint32_t *inc_slot = nullptr;         // u32[inc_entries] (aliases d_hit_slot or d_x_slot)
    DevBuf d_x_row_off, d_x_slot, d_x_row;   // exchanged incidence (multi-GPU)

    // inverted index
    DevBuf d_inv_off;                     // u64[n_slots+1]
    DevBuf d_inv_row;                     // u32[inc_entries]: ROW numbers (0-based), ascending inside a list
    DevBuf d_sort_a, d_sort_b, d_sort_tmp;
    bool have_index = false;

    // pairs
    uint64_t n_pairs = 0, n_increments = 0;
    DevBuf d_pair_key, d_pair_score;      // u64 key = (x_row << 32 | y_row), u32 score; sorted by key
    DevBuf d_pair_key2, d_pair_score2, d_pair_scalars, d_heavy_list, d_heavy_tab, d_pivot_flag;
    uint64_t pair_capacity = 0;
    bool have_pairs = false;
    uint32_t pair_min_score = 1;

    // selection
    uint64_t sel_n_directed = 0, sel_cut = 0, n_selected = 0;
    DevBuf d_hist, d_sel_key, d_sel_score, d_sel_scalars;
    bool have_selection = false;

    // components
    DevBuf d_parent, d_comp_size, d_comp_label, d_comp_scalars;
    uint64_t n_components = 0;
    bool have_components = false;

    // host mirrors for hga_get_*
    PinBuf h_row_off, h_kid, h_pos, h_inv_off, h_inv_read, h_px, h_py, h_ps, h_sx, h_sy, h_ss, h_label, h_clabel, h_csize, h_scalars;
    DevBuf d_export_a, d_export_b, d_export_c;

    hga_metrics_t metrics;
    hga_comm *comm = nullptr;
    uint64_t n_reads_total = 0;           // over all ranks (== n_reads without a comm)
};

// stage launchers (each in its own .cu)
int hga_table_build(hga_handle *h, const uint64_t *host_kmers);
int hga_scan_run(hga_handle *h, const char *d_bases, const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases);
int hga_index_run(hga_handle *h);
int hga_pairs_run(hga_handle *h, uint32_t min_score, const uint32_t *pivots, uint64_t n_pivots);
int hga_select_run(hga_handle *h, double fraction, uint32_t score_threshold);
int hga_cc_run(hga_handle *h, int min_size);

// multi-GPU hooks (hga_comm.cu); all are no-ops / never called without a communicator
int hga_comm_exchange_incidence(hga_handle *h);
int hga_comm_reduce_pairs(hga_handle *h);
int hga_comm_allreduce_u64_sum(hga_handle *h, uint64_t *d_buf, size_t n);
int hga_comm_allreduce_u32_min(hga_handle *h, uint32_t *d_buf, size_t n);
int hga_comm_allreduce_u32_max(hga_handle *h, uint32_t *d_buf, size_t n);
int hga_comm_rank(const hga_handle *h);
int hga_comm_size(const hga_handle *h);
void hga_comm_destroy(hga_handle *h);

static inline uint32_t hga_ceil_log2(uint64_t n) {
    uint32_t b = 0;
    while (b < 63 && (1ull << b) < n) b++;
    return b;
}

struct StageTimer {
    hga_handle *h;
    double *slot;
    StageTimer(hga_handle *hh, double *s) : h(hh), slot(s) { cudaEventRecord(h->ev0, h->stream); }
    void stop() {
        cudaEventRecord(h->ev1, h->stream);
        cudaEventSynchronize(h->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        *slot = ms;
    }
};

#define HGA_STR2(x) #x
#define HGA_STR(x) HGA_STR2(x)
