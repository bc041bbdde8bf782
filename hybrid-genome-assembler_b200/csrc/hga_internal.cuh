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
//
// Every canonical k-mer gets a 32-bit LOCALITY hash B that is shared by most consecutive windows of a read:
// B = rehash of the minimum, over the W central m-mers of the k-mer, of the hashed canonical m-mer (a strand-
// symmetric minimizer). Runs of ~(W + 1) / 2 consecutive windows share B, so their probes fall into the same
// 32 B sector / 128 B bucket and coalesce inside a warp. m is chosen from the size of the set: the number of
// distinct m-mers (4^m / 2) must be well above the number of keys, otherwise unrelated loci share a minimizer
// VALUE and pile into the same block (4^m / 2 >= 3 n), and W = k - m + 1 <= HGA_MIN_W. m <= 16 so that m-mers
// fit 32 bits. For k < HGA_MIN_K_FOR_MIN there is no room for minimizers and B is a plain hash of the k-mer.
//
//   filter    : blocked Bloom filter; block = 32 B (8 words) selected by B, word + 2 bits inside the block
//               selected by an independent hash of the k-mer. Sized to stay L2-resident; consulted first so
//               that a non-member costs one 4 B probe that its neighbours share.
//   key table : buckets of 16 x u64 keys (one 128 B line) selected by B, load factor 1/3; a lookup reads the whole home
//               bucket in one round trip (eight independent 16 B loads). Keys that find no room in their chain of
//               chain_buckets buckets (default 1: the home bucket only; 3 % of the keys of config 4) go to a plain
//               open-addressing overflow table hashed by k-mer, which only lookups that meet a FULL bucket consult. The internal k-mer id ("slot") is the index of the key in the (main | overflow) key
//               array; slot_kid maps it back to the caller's kmer_id.
// ------------------------------------------------------------------------------------------------
#define HGA_MIN_W 8
#define HGA_MIN_K_FOR_MIN 12
#define HGA_BUCKET_SLOTS 16      // one 128 B line of u64 keys
#define HGA_CHAIN_BUCKETS 4      // a key lives within this many buckets of its home bucket, else in the overflow region
#define HGA_CHAIN_SLOTS (HGA_BUCKET_SLOTS * HGA_CHAIN_BUCKETS)

struct KmerGeom {
    int k = 0;
    int use_min = 0;      // 1: B from the minimizer, 0: B from the k-mer hash
    int m = 0;            // m-mer length (<= 16)
    int W = 0;            // m-mers the minimum is taken over (2 .. HGA_MIN_W)
    int skip = 0;         // m-mers skipped at each end of the window (> 0 only when k - m + 1 > HGA_MIN_W)
    uint32_t mmask = 0;   // 2m low bits
    int rc_shift = 0;     // 2 (k - m): the reverse-complement k-mer's first m-mer
};

struct KmerTable {
    uint64_t *keys = nullptr;       // n_slots (main region, then overflow region)
    uint32_t *slot_kid = nullptr;   // n_slots
    uint32_t *kid_slot = nullptr;   // n_kmers
    uint32_t *filter = nullptr;     // n_blocks * 8 words
    uint32_t n_buckets = 0;         // main region: (n_buckets + HGA_CHAIN_BUCKETS) * HGA_BUCKET_SLOTS slots
    uint32_t n_main = 0;            // slots in the main region
    uint32_t n_over = 0;            // slots in the overflow region (0: none; else a power of two)
    uint32_t n_blocks = 0;
    uint32_t n_slots = 0;           // n_main + n_over
    uint32_t slot_bits = 0;         // ceil(log2(n_slots))
    uint32_t chain_buckets = 1;     // buckets a key may live in (home bucket first, at most HGA_CHAIN_BUCKETS); experiment switch HGA_CHAIN_BUCKETS
    KmerGeom geom;
};

#define HGA_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define HGA_C1 0x9E3779B1u
#define HGA_C2 0x85EBCA77u
#define HGA_C3 0xC2B2AE3Du
#define HGA_C4 0x27D4EB2Fu

static inline KmerGeom hga_make_geom(int k, uint64_t n_kmers) {
    KmerGeom g;
    g.k = k;
    if (k < HGA_MIN_K_FOR_MIN) return g;
    g.use_min = 1;
    int m = 4;
    while (m < 16 && (1ull << (2 * m - 1)) < 3 * n_kmers) m++;      // 4^m / 2 >= 3 n
    if (m < k - HGA_MIN_W + 1) m = k - HGA_MIN_W + 1;                // W <= HGA_MIN_W where k leaves room
    if (m > 16) m = 16;
    if (m > k - 1) m = k - 1;                                         // W >= 2
    int W = k - m + 1;
    if (W > HGA_MIN_W) W = ((k - m + 1 - HGA_MIN_W) % 2 == 0) ? HGA_MIN_W : HGA_MIN_W - 1;   // centred: strand symmetric
    g.m = m; g.W = W;
    g.skip = (k - m + 1 - W) / 2;
    g.mmask = (m == 16) ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1);
    g.rc_shift = 2 * (k - m);
    return g;
}

// hash of the k-mer that picks the word and the two bits inside a filter block (independent of B)
__host__ __device__ __forceinline__ uint32_t hga_bits_hash(uint64_t kmer) {
    return ((uint32_t) kmer ^ ((uint32_t) (kmer >> 32) * HGA_C3)) * HGA_C1;
}
__host__ __device__ __forceinline__ uint32_t hga_bits_word(uint32_t h) { return h >> 29; }
__host__ __device__ __forceinline__ uint32_t hga_bits_mask(uint32_t h) { return (1u << ((h >> 24) & 31)) | (1u << ((h >> 19) & 31)); }
// Probe order of a key: the 4-slot SECTOR (32 B) of its home bucket picked by the k-mer hash, then the bucket's other three
// sectors in cyclic order, then the next bucket the same way. A lookup therefore usually ends after ONE 32 B load.
#define HGA_SECTOR_SLOTS 4
__host__ __device__ __forceinline__ uint32_t hga_bits_sector(uint32_t h) { return (h >> 15) & (HGA_BUCKET_SLOTS / HGA_SECTOR_SLOTS - 1); }
// slot (relative to the home bucket) of the j-th probe, j = 0 .. HGA_CHAIN_SLOTS - 1
__host__ __device__ __forceinline__ uint32_t hga_chain_slot(uint32_t sector, uint32_t j) {
    const uint32_t in_bucket = j & (HGA_BUCKET_SLOTS - 1);
    const uint32_t sec = (sector + in_bucket / HGA_SECTOR_SLOTS) & (HGA_BUCKET_SLOTS / HGA_SECTOR_SLOTS - 1);
    return (j & ~(uint32_t) (HGA_BUCKET_SLOTS - 1)) | (sec * HGA_SECTOR_SLOTS) | (in_bucket & (HGA_SECTOR_SLOTS - 1));
}

// plain k-mer hash: B for small k, and the overflow table's home position
__host__ __device__ __forceinline__ uint32_t hga_plain_hash(uint64_t kmer) {
    uint32_t h = (uint32_t) kmer * HGA_C2 + (uint32_t) (kmer >> 32) * HGA_C1;
    h ^= h >> 16;
    return h * HGA_C3;
}

__host__ __device__ __forceinline__ uint32_t hga_scale(uint32_t h, uint32_t n) {
#ifdef __CUDA_ARCH__
    return __umulhi(h, n);
#else
    return (uint32_t) (((uint64_t) h * n) >> 32);
#endif
}

// reverse complement of an m-mer held in the low 2m bits (codes A0 C1 G2 T3)
__host__ __device__ __forceinline__ uint32_t hga_revcomp32(uint32_t x, int m) {
#ifdef __CUDA_ARCH__
    x = __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    x = (x >> 16) | (x << 16);
#endif
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);   // bits reversed -> 2-bit groups reversed
    return (~x) >> (32 - 2 * m);
}

__host__ __device__ __forceinline__ uint32_t hga_mmer_hash(uint32_t fwd_m, uint32_t rc_m) {
    return (fwd_m < rc_m ? fwd_m : rc_m) * HGA_C1;
}
// the minimum of 8 hashes is biased towards small values; the multiply carries its low bits up into the bits hga_scale uses
__host__ __device__ __forceinline__ uint32_t hga_locality_from_min(uint32_t gmin) { return gmin * HGA_C2; }

// B computed from the k-mer VALUE alone (table build; scan windows that contain a non-ACGT byte, whose two
// strands are not reverse complements of each other)
__host__ __device__ __forceinline__ uint32_t hga_locality_hash(uint64_t kmer, const KmerGeom &g) {
    if (!g.use_min) return hga_plain_hash(kmer);
    uint32_t gmin = 0xFFFFFFFFu;
    for (int o = 0; o < g.W; o++) {
        const int sh = 2 * (g.k - g.m - g.skip - o);
        const uint32_t fm = (uint32_t) (kmer >> sh) & g.mmask;
        const uint32_t h = hga_mmer_hash(fm, hga_revcomp32(fm, g.m));
        gmin = h < gmin ? h : gmin;
    }
    return hga_locality_from_min(gmin);
}

// ------------------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------------------
struct hga_comm;   // NCCL state, hga_comm.cu

// host-side result of hga_enrich (small: cores, enrichment connections, final membership)
struct EnrichResult {
    std::vector<uint32_t> core_id, core_read;           // survivor ids ascending; members ascending inside a core
    std::vector<uint64_t> core_off;
    std::vector<uint32_t> conn_x, conn_y, conn_score;   // directed (core survivor -> partner), canonical order
    std::vector<uint32_t> final_id, final_read, assignment;
    std::vector<uint64_t> final_off;
    // tail / spectral block (hga_enrich_full): what it saw and what it merged
    bool tail_block_ran = false;                        // more than two scaffold components (:768)
    uint64_t n_scaffold_cores = 0;                      // cores before the merge of the spectral clusters
    std::vector<uint32_t> tconn_x, tconn_y;             // tail connections, x < y, canonical order (all, not only score > 5)
    std::vector<uint64_t> tconn_score;
    std::vector<uint32_t> cluster_member;               // spectral clusters, element [0] of a cluster = the surviving id; empty clusters dropped
    std::vector<uint64_t> cluster_off;
};

// hga_enrich_full: the parameters of the tail / spectral block (:768-777)
struct TailParams {
    const uint64_t *read_off;       // HOST, n_reads + 1, as given to hga_scan
    uint32_t amplification_min_score;
    int spectral_dims;
};

struct hga_handle {
    int device = 0;
    int k = 0;
    uint64_t n_kmers = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // H2D of the bases, overlapped with the scan (hga_scan)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;   // stage timer, nested (exchange) timer
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
    DevBuf d_tile_state, d_tile_dir, d_scan_scalars;
    bool have_scan = false;

    // row space of the inverted index, pairs and components (this GPU's reads; ALL reads with a communicator)
    uint64_t inc_rows = 0;                // number of rows
    uint32_t inc_row_first_id = 1;        // read id of row 0
    uint64_t inc_entries = 0;             // entries of the inverted index
    DevBuf d_x_slot, d_x_row;             // exchange staging (multi-GPU)
    DevBuf d_hit_kid;                     // multi-GPU: hits keyed by the caller's kmer_id (slots differ between ranks: the table is built with atomics)
    DevBuf d_g_kid, d_g_row_off;          // multi-GPU: by-read incidence of this rank's pivot rows (index key per hit, u64 row offsets)
    bool index_by_kid = false;            // the inverted index is keyed by kmer_id (multi-GPU) instead of table slot
    uint32_t index_keys = 0;              // number of lists in the inverted index: n_slots, or the index-key space when keyed by kmer_id
    uint32_t index_key_div = 0;           // multi-GPU: index key of kmer_id = kmer_id + kmer_id / index_key_div (one unused key closes every owner's range)

    // inverted index
    DevBuf d_inv_off;                     // u32[n_slots+1] (the incidence of one GPU has < 2^32 entries)
    DevBuf d_inv_row;                     // u32[inc_entries]: ROW numbers (0-based), ascending inside a list
    DevBuf d_sort_a, d_sort_b, d_sort_tmp;
    bool have_index = false;

    // pairs
    uint64_t n_pairs = 0, n_increments = 0;
    DevBuf d_pair_key, d_pair_score;      // u64 key = (x_row << 32 | y_row), u32 score; sorted by key
    DevBuf d_pair_key2, d_pair_score2, d_pair_scalars, d_heavy_list, d_mid_list, d_redo_list, d_heavy_tab, d_pivot_flag, d_pivot_order;
    uint64_t pair_capacity = 0;
    uint64_t pair_rows = 0;               // rows of the by-read incidence the pair counter walks (this rank's pivots with a communicator)
    uint32_t pair_pivot_mul = 1, pair_pivot_add = 0;   // local row t is global row t * mul + add (rank, rank + G, ...)
    bool have_pairs = false;
    uint32_t pair_min_score = 1;
    bool pair_subset = false;             // the pairs come from a pivot subset (d_pivot_flag marks the pivots)

    // selection
    uint64_t sel_n_directed = 0, sel_cut = 0, n_selected = 0;
    DevBuf d_hist, d_sel_key, d_sel_score, d_sel_scalars;
    bool have_selection = false;

    // components
    DevBuf d_parent, d_comp_size, d_comp_label, d_comp_scalars;
    uint64_t n_components = 0;
    bool have_components = false;

    // merge + enrichment (hga_enrich.cu)
    DevBuf d_enr_parent, d_enr_core_of, d_enr_surv, d_enr_R, d_enr_scalars, d_enr_keys, d_enr_keys2, d_enr_core_koff, d_purged_off, d_purged_row;
    DevBuf d_purged2_off, d_purged2_row;  // second purge (merge of the spectral clusters); swapped into d_purged_* when it ran
    uint64_t n_purged = 0, n_core_kmers = 0;
    EnrichResult enrich;
    bool have_enrichment = false;

    // host mirrors for hga_get_*
    PinBuf h_row_off, h_kid, h_pos, h_inv_off, h_inv_read, h_px, h_py, h_ps, h_sx, h_sy, h_ss, h_label, h_clabel, h_csize, h_scalars;
    DevBuf d_export_a, d_export_b, d_export_c;

    hga_metrics_t metrics;
    hga_comm *comm = nullptr;
    uint64_t n_reads_total = 0;           // over all ranks (== n_reads without a comm)
};

// stage launchers (each in its own .cu)
int hga_table_build(hga_handle *h, const uint64_t *host_kmers);
// h_bases != nullptr: the bases are still on the host; the run copies them (in chunks, overlapped with the scan when large)
int hga_scan_run(hga_handle *h, const char *d_bases, const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases, const char *h_bases);
int hga_index_run(hga_handle *h);
int hga_pairs_run(hga_handle *h, uint32_t min_score, const uint32_t *pivots, uint64_t n_pivots);
int hga_select_run(hga_handle *h, double fraction, uint32_t score_threshold);
int hga_cc_run(hga_handle *h, int min_size);
// tail == nullptr: without the tail / spectral block (hga_enrich, hga_enrich_ex)
int hga_enrich_run(hga_handle *h, int min_size, int max_size, uint32_t min_score, const TailParams *tail);
// CSR by index key (off u32[keys + 1], rows) -> CSR by the caller's kmer_id with read ids, in the pinned export buffers (hga_capi.cu)
int hga_export_index(hga_handle *h, const uint32_t *d_off, const uint32_t *d_row, uint64_t E, hga_index *out);

// multi-GPU hooks (hga_comm.cu); all are no-ops / never called without a communicator
int hga_comm_build_global_index(hga_handle *h);
int hga_comm_allgather_u64(hga_handle *h, uint64_t mine, std::vector<uint64_t> &all);
int hga_comm_allgatherv(hga_handle *h, const void *d_mine, void *d_all, const std::vector<uint64_t> &counts, int elem_bytes);
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
    cudaEvent_t a, b;
    StageTimer(hga_handle *hh, double *s, bool nested = false) : h(hh), slot(s), a(nested ? hh->ev2 : hh->ev0), b(nested ? hh->ev3 : hh->ev1) {
        cudaEventRecord(a, h->stream);
    }
    void stop() {
        cudaEventRecord(b, h->stream);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        *slot = ms;
    }
};

#define HGA_STR2(x) #x
#define HGA_STR(x) HGA_STR2(x)
