// enrich: merge of the scaffold components, enrichment connections, restricted union-find, final merge.
//
// Replaces, in ReadClusteringEngine::run_clustering (clustering/ReadClusteringEngine.cpp):
//   :763-764  union_find(...) roots + merge_components(scaffold_components)      (merge_components: :349-422)
//   :785-794  get_connections(core_component_ids, enrichment_connections_min_score), union_find(conns, restricted = cores,
//             2, -1), merge_components, get_component_ids(scaffold_component_min_size)
// The tail / spectral block in between (:768-777) runs when the caller asks for it (hga_enrich_full -> `tail`): the
// state after the scaffold merge goes to the host stages hga_host_tail_connections / hga_spectral_clustering (the tail amplification
// comes back to the GPU: gpu_tail_amplify) and the clusters they return are merged by a SECOND merge_components on the GPU (enr_merge2_keys_kernel, enr_purge2_kernel; rule at the kernels).
// Without it (hga_enrich, hga_enrich_ex) what runs here is the reference's own path when it has at most two scaffold components
// or finds no strong tail connection.
//
// What the reference's merge does to the engine state, restated as data-parallel rules (oracle: orc_engine_merge):
//   * the survivor of a component (element [0] = the root the sequential union_find ended with) receives the sorted UNIQUE
//     union U_c of its members' k-mer id lists (:379, merge_n_vectors(unique = true));
//   * every k-mer gets a removal list: each non-surviving member once per occurrence, the survivor ONCE if the k-mer is in
//     U_c (:385-389). The purge (:395-419) is a two-pointer walk that STOPS when the removal list is exhausted and keeps only
//     what it has copied so far. With R(k) = the largest removed id of k-mer k, the new list is therefore
//         { e in list(k) : e < R(k), e in no core, or e a survivor and not the first of its copies }
//     (entries >= R(k) are dropped whoever they belong to; a survivor holding the k-mer m times keeps m - 1 stale entries);
//     k-mers no merged component holds are untouched.
//   * enrichment score(c, y) = sum over k in U_c of the copies of y in the purged list of k, y != survivor(c) (:311-317).
// The survivor ids matter (they enter R(k)), and they depend on the order the sequential union_find saw the edges in: the
// roots are replayed on the host over the selected edges in the canonical order (score desc, x asc, y asc), union by size with
// ties to y's root (:459-466), after the GPU has dropped the edges whose endpoints were already connected. Everything
// proportional to the incidence (E) runs on the GPU:
//   enr_edge_filter/hook     the ~N of M selected edges that can still join two components (the host replays only those)
//   enr_list_cores_kernel    (core, slot) candidates, one per run of same-core entries of a list, + member part of R(k)
//   CUB radix sort + unique  U_c for every core at once
//   enr_survivor_max_kernel  survivor part of R(k)
//   enr_purge_kernel         count / fill passes over the inverted index -> purged CSR           2 x 4 E B read, <= 4 E written
//   enr_emit_kernel          (core, partner row) per (k in U_c, entry of purged list(k)), then radix sort + run-length encode
//   enr_filter_kernel        score >= min, partner != survivor
// The connection list that comes out is small (cores x partners); its canonical order, the restricted union-find over it and the
// final membership are host code (sequential in the reference as well).
#include "hga_internal.cuh"

#include <algorithm>
#include <chrono>
#include <numeric>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

namespace {

inline int grid_for(const hga_handle *h, uint64_t n, int per_block = 256) {
    return (int) std::max<uint64_t>(1, std::min<uint64_t>((n + per_block - 1) / per_block, (uint64_t) h->sm_count * 16));
}

// (core, slot) candidates from the inverted index: one thread per list; a run of consecutive entries of the same core yields ONE
// candidate (a discriminative k-mer's reads mostly come from one haplotype, hence one core: ~1 candidate per list instead of
// one per hit), the radix sort + unique that follows removes the rest. FILL = false also leaves the member part of R(k): the
// list is ascending, so the last entry that belongs to a core is the largest removed member.
template<bool FILL>
__global__ void enr_list_cores_kernel(const uint32_t *__restrict__ inv_off, const uint32_t *__restrict__ inv_row, uint32_t n_slots,
                                      const int32_t *__restrict__ core_of, uint32_t *__restrict__ cnt, const unsigned long long *__restrict__ out_off,
                                      uint64_t *__restrict__ out, uint32_t *__restrict__ R) {
    for (uint64_t s = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; s < n_slots; s += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t lo = inv_off[s], hi = inv_off[s + 1];
        uint32_t n = 0, last1 = 0;
        unsigned long long w = FILL ? out_off[s] : 0;
        int32_t prev = -1;
        for (uint32_t i = lo; i < hi; i++) {
            const uint32_t e = inv_row[i];
            const int32_t c = core_of[e];
            if (c < 0) continue;
            last1 = e + 1;
            if (c != prev) { n++; if (FILL) out[w++] = ((uint64_t) (uint32_t) c << 32) | s; }
            prev = c;
        }
        if (!FILL) { cnt[s] = n; R[s] = last1; }
    }
}

// ---- union_find roots: the edges that can still join two components ---------------------------------------------------------
__device__ __forceinline__ uint32_t enr_find(uint32_t *parent, uint32_t v) {
    uint32_t p = parent[v];
    while (p != v) {
        const uint32_t gp = parent[p];
        if (gp != p) parent[v] = gp;     // path halving (benign race: only ever points further up)
        v = p; p = gp;
    }
    return v;
}

__global__ void enr_iota_kernel(uint32_t *parent, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) parent[i] = (uint32_t) i;
}

// flag[i] = the endpoints of edge i are not connected by the edges of the EARLIER batches (parent is not modified apart from
// path halving, so an edge never sees the unions of its own batch)
__global__ void enr_edge_filter_kernel(const uint64_t *__restrict__ key, uint64_t lo, uint64_t hi, uint32_t *parent, uint8_t *__restrict__ flag) {
    for (uint64_t i = lo + blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < hi; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t kk = key[i];
        flag[i] = enr_find(parent, (uint32_t) (kk >> 32)) != enr_find(parent, (uint32_t) kk);
    }
}

__global__ void enr_edge_hook_kernel(const uint64_t *__restrict__ key, uint64_t lo, uint64_t hi, uint32_t *parent, const uint8_t *__restrict__ flag) {
    for (uint64_t i = lo + blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < hi; i += (uint64_t) gridDim.x * blockDim.x) {
        if (!flag[i]) continue;
        const uint64_t kk = key[i];
        uint32_t a = (uint32_t) (kk >> 32), b = (uint32_t) kk;
        for (;;) {
            a = enr_find(parent, a); b = enr_find(parent, b);
            if (a == b) break;
            const uint32_t top = max(a, b), bot = min(a, b);
            if (atomicCAS(&parent[top], top, bot) == top) break;
        }
    }
}

// the survivor is scheduled for removal from every k-mer of its union, whether it holds the k-mer itself or not
__global__ void enr_survivor_max_kernel(const uint64_t *__restrict__ ukeys, uint64_t n, const uint32_t *__restrict__ surv_row, uint32_t *R) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t k = ukeys[i];
        atomicMax(&R[(uint32_t) k], surv_row[k >> 32] + 1);
    }
}

// first position of every core's run in the sorted unique keys
__global__ void enr_core_bounds_kernel(const uint64_t *__restrict__ ukeys, uint64_t n, uint32_t n_cores, unsigned long long *core_koff) {
    for (uint64_t c = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; c <= n_cores; c += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t want = c << 32;
        uint64_t lo = 0, hi = n;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (ukeys[mid] < want) lo = mid + 1; else hi = mid; }
        core_koff[c] = lo;
    }
}

// The purge of one inverted list. FILL = false: count the surviving entries; FILL = true: write them at out_off[slot].
template<bool FILL>
__global__ void enr_purge_kernel(const uint32_t *__restrict__ inv_off, const uint32_t *__restrict__ inv_row, uint32_t n_slots,
                                 const uint32_t *__restrict__ R, const int32_t *__restrict__ core_of, const uint32_t *__restrict__ surv_row,
                                 uint32_t *__restrict__ cnt, const uint32_t *__restrict__ out_off, uint32_t *__restrict__ out_row) {
    for (uint64_t s = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; s < n_slots; s += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t lo = inv_off[s], hi = inv_off[s + 1];
        uint32_t n = 0;
        uint32_t w = FILL ? out_off[s] : 0;
        if (lo != hi) {
            const uint32_t r1 = R[s];                 // largest removed row + 1; 0: no merged component holds this k-mer
            if (r1 == 0) {
                n = hi - lo;
                if (FILL) for (uint32_t i = lo; i < hi; i++) out_row[w++] = inv_row[i];
            } else {
                uint32_t prev = 0xFFFFFFFFu;
                for (uint32_t i = lo; i < hi; i++) {
                    const uint32_t e = inv_row[i];
                    if (e + 1 >= r1) break;           // the walk ends with the removal list: e >= R(k) is dropped (ascending list)
                    const int32_t c = core_of[e];
                    const bool keep = c < 0 || (surv_row[c] == e && prev == e);   // stale copies of a survivor stay
                    prev = e;
                    if (keep) { n++; if (FILL) out_row[w++] = e; }
                }
            }
        }
        if (!FILL) cnt[s] = n;
    }
}

__global__ void enr_emit_len_kernel(const uint64_t *__restrict__ ukeys, uint64_t n, const uint32_t *__restrict__ p_off, unsigned long long *len) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t slot = (uint32_t) ukeys[i];
        len[i] = p_off[slot + 1] - p_off[slot];
    }
}

// (core, partner row) for every entry of the purged list of every k-mer of every core
__global__ void enr_emit_kernel(const uint64_t *__restrict__ ukeys, uint64_t n, const uint32_t *__restrict__ p_off, const uint32_t *__restrict__ p_row,
                                const unsigned long long *__restrict__ at, uint64_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t k = ukeys[i];
        const uint32_t slot = (uint32_t) k;
        const uint64_t hi_bits = k & 0xFFFFFFFF00000000ull;
        unsigned long long w = at[i];
        for (uint32_t j = p_off[slot]; j < p_off[slot + 1]; j++) out[w++] = hi_bits | p_row[j];
    }
}

// :317 erase(pivot), :320 score >= min_score
__global__ void enr_filter_kernel(const uint64_t *__restrict__ run_key, const uint32_t *__restrict__ run_len, uint64_t n_runs, const uint32_t *__restrict__ surv_row,
                                  uint32_t min_score, uint32_t first_id, uint32_t *cx, uint32_t *cy, uint32_t *cs, unsigned long long *count) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_runs; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t k = run_key[i];
        const uint32_t s = surv_row[k >> 32], y = (uint32_t) k, v = run_len[i];
        if (y != s && v >= min_score) {
            const unsigned long long w = atomicAdd(count, 1ull);
            cx[w] = s + first_id; cy[w] = y + first_id; cs[w] = v;
        }
    }
}

// ---- tail amplification (amplify_component -> get_connections(tail, min), ReadClusteringEngine.cpp:583-592, :301-333) on the GPU ---------------------------
// A pivot = one tail vertex. Its k-mer list is, as in the engine state after the first merge: the unique union U_c for a survivor, the read's own hits
// (duplicates included) for every other id. amp_list_kernel flattens the lists of all pivots into keys (pivot index << 32 | slot); the enrichment's
// enr_emit_len / enr_emit kernels, a sort and a run-length encode turn them into (pivot, partner, count); amp_filter_kernel keeps count >= min, partner != pivot.
__global__ void amp_list_kernel(const unsigned long long *__restrict__ key_off, const unsigned long long *__restrict__ src_begin, const uint8_t *__restrict__ from_union,
                                uint64_t n_pivots, uint64_t n_keys, const uint32_t *__restrict__ hit_slot, const uint64_t *__restrict__ ukeys, uint64_t *__restrict__ out) {
    for (uint64_t j = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; j < n_keys; j += (uint64_t) gridDim.x * blockDim.x) {
        uint64_t lo = 0, hi = n_pivots - 1;                       // the pivot whose key range holds j: last pi with key_off[pi] <= j
        while (lo < hi) { const uint64_t mid = (lo + hi + 1) >> 1; if (key_off[mid] <= j) lo = mid; else hi = mid - 1; }
        const uint64_t at = src_begin[lo] + (j - key_off[lo]);
        const uint32_t slot = from_union[lo] ? (uint32_t) ukeys[at] : hit_slot[at];
        out[j] = (lo << 32) | slot;
    }
}

__global__ void amp_filter_kernel(const uint64_t *__restrict__ run_key, const uint32_t *__restrict__ run_len, uint64_t n_runs, const uint32_t *__restrict__ pivot_row,
                                  uint32_t min_score, uint32_t first_id, uint64_t *out, unsigned long long *count) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_runs; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t k = run_key[i];
        const uint32_t y = (uint32_t) k;
        if (y != pivot_row[k >> 32] && run_len[i] >= min_score) out[atomicAdd(count, 1ull)] = (k & 0xFFFFFFFF00000000ull) | (y + first_id);
    }
}

// sort keys of the connection list, read through the current permutation. STAGE 0: x (and perm = identity); 1: min << 32 | max;
// 2: ~score
template<int STAGE>
__global__ void enr_conn_key_kernel(const uint32_t *__restrict__ cx, const uint32_t *__restrict__ cy, const uint32_t *__restrict__ cs, const uint32_t *__restrict__ perm,
                                    uint64_t n, uint32_t *__restrict__ k32, uint64_t *__restrict__ k64, uint32_t *__restrict__ perm_out) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t j = STAGE == 0 ? (uint32_t) i : perm[i];
        if (STAGE == 0) { k32[i] = cx[j]; perm_out[i] = (uint32_t) i; }
        else if (STAGE == 1) { const uint32_t x = cx[j], y = cy[j]; k64[i] = ((uint64_t) min(x, y) << 32) | max(x, y); }
        else k32[i] = ~cs[j];
    }
}

__global__ void enr_conn_gather_kernel(const uint32_t *__restrict__ cx, const uint32_t *__restrict__ cy, const uint32_t *__restrict__ cs, const uint32_t *__restrict__ perm,
                                       uint64_t n, uint32_t *__restrict__ ox, uint32_t *__restrict__ oy, uint32_t *__restrict__ os) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t j = perm[i];
        ox[i] = cx[j]; oy[i] = cy[j]; os[i] = cs[j];
    }
}

__global__ void enr_key_rows_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint32_t *x, uint32_t *y) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        x[i] = (uint32_t) (keys[i] >> 32); y[i] = (uint32_t) keys[i];
    }
}

// ---- second merge: merge_components(spectral clusters) on the state the first merge left (:774) ------------------------------------
// A cluster is a set of cores (merged scaffold components); element [0] survives. What merge_components (:349-422) does then, as rules:
//   * the survivor's list becomes the unique union of the members' lists, which after the first merge are the unions U_c: every key
//     (core, slot) of the sorted unique array moves to the core its cluster merges into (sort + unique afterwards);
//   * removal list of k-mer k (:385-389): every member c of a multi-member cluster with k in U_c, once (U_c is unique), and the
//     cluster's survivor for every k of the new union. R2(k) = the largest of them; the two-pointer purge (:405-416) again stops
//     with the removal list. A purged list holds reads outside every core and stale copies of first-merge survivors (a survivor that
//     held k m times kept m - 1 entries, and then k is in its U_c), so: new list = { e < R2(k) } minus the FIRST copy of every
//     entry that is the survivor of a core in a multi-member cluster.
// map[c] = index of the core that c merges into, in the new (compacted) core numbering, | 1 << 31 when c's cluster has several members.
__global__ void enr_merge2_keys_kernel(const uint64_t *__restrict__ ukeys, uint64_t n, const uint32_t *__restrict__ map, const uint32_t *__restrict__ surv_old,
                                       const uint32_t *__restrict__ surv_new, uint32_t *R2, uint64_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t k = ukeys[i];
        const uint32_t slot = (uint32_t) k, c = (uint32_t) (k >> 32), m = map[c], nc = m & 0x7FFFFFFFu;
        if (m >> 31) {
            atomicMax(&R2[slot], surv_old[c] + 1);
            atomicMax(&R2[slot], surv_new[nc] + 1);
        }
        out[i] = ((uint64_t) nc << 32) | slot;
    }
}

template<bool FILL>
__global__ void enr_purge2_kernel(const uint32_t *__restrict__ p_off, const uint32_t *__restrict__ p_row, uint32_t n_slots, const uint32_t *__restrict__ R2,
                                  const uint8_t *__restrict__ removed_once, uint32_t *__restrict__ cnt, const uint32_t *__restrict__ out_off,
                                  uint32_t *__restrict__ out_row) {
    for (uint64_t s = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; s < n_slots; s += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t lo = p_off[s], hi = p_off[s + 1];
        uint32_t n = 0;
        uint32_t w = FILL ? out_off[s] : 0;
        if (lo != hi) {
            const uint32_t r1 = R2[s];                // largest removed row + 1; 0: no merged cluster lists this k-mer
            if (r1 == 0) {
                n = hi - lo;
                if (FILL) for (uint32_t i = lo; i < hi; i++) out_row[w++] = p_row[i];
            } else {
                uint32_t prev = 0xFFFFFFFFu;
                for (uint32_t i = lo; i < hi; i++) {
                    const uint32_t e = p_row[i];
                    if (e + 1 >= r1) break;
                    const bool keep = !removed_once[e] || prev == e;
                    prev = e;
                    if (keep) { n++; if (FILL) out_row[w++] = e; }
                }
            }
        }
        if (!FILL) cnt[s] = n;
    }
}

uint32_t dsu_find(std::vector<uint32_t> &parent, uint32_t v) {
    uint32_t r = v;
    while (parent[r] != r) r = parent[r];
    while (parent[v] != r) { const uint32_t nx = parent[v]; parent[v] = r; v = nx; }
    return r;
}

struct Conn { uint32_t x, y, s; };

// Wall time of the phases (the stream is synchronised at each mark), summed into the stage the reference's timers attribute them to
// (hga_metrics_t::enrich_phase_ms); HGA_ENRICH_TIMING=1 also prints every phase on stderr.
struct PhaseClock {
    bool on;
    cudaStream_t st;
    double *stage_ms;
    std::chrono::steady_clock::time_point t0;
    PhaseClock(cudaStream_t s, double *stages) : on(getenv("HGA_ENRICH_TIMING") != nullptr), st(s), stage_ms(stages), t0(std::chrono::steady_clock::now()) {
        for (int i = 0; i < 6; i++) stage_ms[i] = 0;
    }
    void mark(const char *what, int stage) {
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        stage_ms[stage] += ms;
        if (on) fprintf(stderr, "hga_enrich: %-28s %8.2f ms\n", what, ms);
        t0 = t1;
    }
};

// The GPU side of hga_host_tail_connections_impl's amplification callback (hga_tails.cpp): the state hga_enrich_run holds after the first merge
struct AmplifyCtx {
    hga_handle *h;
    const std::vector<int32_t> *core_of;          // row -> core (or -1)
    const std::vector<uint32_t> *surv_row;        // core -> surviving row
    const uint64_t *row_off;                      // HOST row offsets of the hits (hga_get_hits)
    uint32_t first_id;
    uint64_t C;
    const uint64_t *d_u;                          // sorted unique (core << 32 | slot)
    const unsigned long long *d_core_koff;        // core -> first key of its union in d_u
    const uint32_t *d_poff, *d_prow;              // purged index by slot
    unsigned long long *d_count;
};

int gpu_tail_amplify(void *vctx, const uint32_t *pivot_id, uint64_t n_pivots, uint32_t min_score, std::vector<std::pair<uint32_t, uint32_t>> &out) {
    AmplifyCtx &c = *static_cast<AmplifyCtx *>(vctx);
    hga_handle *h = c.h;
    if (n_pivots == 0) return HGA_OK;
    if (n_pivots >= (1ull << 32)) { hga_set_error("tail amplification: too many tail vertices"); return HGA_E_OVERFLOW; }
    std::vector<unsigned long long> koff(c.C + 1, 0);
    HGA_CUDA(cudaMemcpyAsync(koff.data(), c.d_core_koff, (c.C + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));
    // per pivot: where its k-mer list lives (the union of its core for a survivor, its own hits otherwise) and where its keys go
    std::vector<unsigned long long> key_off(n_pivots + 1, 0), src_begin(n_pivots, 0);
    std::vector<uint8_t> from_union(n_pivots, 0);
    std::vector<uint32_t> pivot_row(n_pivots, 0);
    for (uint64_t i = 0; i < n_pivots; i++) {
        const uint32_t r = pivot_id[i] - c.first_id;
        pivot_row[i] = r;
        const int32_t core = (*c.core_of)[r];
        uint64_t len;
        if (core >= 0 && (*c.surv_row)[core] == r) { from_union[i] = 1; src_begin[i] = koff[core]; len = koff[core + 1] - koff[core]; }
        else { src_begin[i] = c.row_off[r]; len = c.row_off[r + 1] - c.row_off[r]; }
        key_off[i + 1] = key_off[i] + len;
    }
    const uint64_t n_keys = key_off[n_pivots];
    if (n_keys == 0) return HGA_OK;
    DevBuf b_desc, b_keys, b_len, b_emit, b_sorted, b_runlen, b_out;
    struct Release { DevBuf *b[7]; ~Release() { for (DevBuf *x : b) x->release(); } } rel{{&b_desc, &b_keys, &b_len, &b_emit, &b_sorted, &b_runlen, &b_out}};
    HGA_TRY(b_desc.ensure((n_pivots + 1) * (8 + 8 + 4 + 1) + 64));
    unsigned long long *d_key_off = b_desc.as<unsigned long long>(), *d_src = d_key_off + (n_pivots + 1);
    uint32_t *d_pivot_row = reinterpret_cast<uint32_t *>(d_src + n_pivots);
    uint8_t *d_from = reinterpret_cast<uint8_t *>(d_pivot_row + n_pivots);
    HGA_CUDA(cudaMemcpyAsync(d_key_off, key_off.data(), (n_pivots + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    HGA_CUDA(cudaMemcpyAsync(d_src, src_begin.data(), n_pivots * 8, cudaMemcpyHostToDevice, h->stream));
    HGA_CUDA(cudaMemcpyAsync(d_pivot_row, pivot_row.data(), n_pivots * 4, cudaMemcpyHostToDevice, h->stream));
    HGA_CUDA(cudaMemcpyAsync(d_from, from_union.data(), n_pivots, cudaMemcpyHostToDevice, h->stream));
    HGA_TRY(b_keys.ensure((n_keys + 1) * 8));
    HGA_TRY(b_len.ensure((n_keys + 2) * 8 * 2));
    uint64_t *d_keys = b_keys.as<uint64_t>();
    unsigned long long *d_len = b_len.as<unsigned long long>(), *d_at = d_len + (n_keys + 2);
    amp_list_kernel<<<grid_for(h, n_keys), 256, 0, h->stream>>>(d_key_off, d_src, d_from, n_pivots, n_keys, h->d_hit_slot.as<uint32_t>(), c.d_u, d_keys);
    enr_emit_len_kernel<<<grid_for(h, n_keys), 256, 0, h->stream>>>(d_keys, n_keys, c.d_poff, d_len);
    HGA_CUDA(cudaMemsetAsync(d_len + n_keys, 0, 8, h->stream));
    size_t tmp = 0;
    HGA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_len, d_at, n_keys + 1, h->stream));
    HGA_TRY(h->d_sort_tmp.ensure(tmp + 16));
    HGA_CUDA(cub::DeviceScan::ExclusiveSum(h->d_sort_tmp.p, tmp, d_len, d_at, n_keys + 1, h->stream));
    unsigned long long n_emit = 0;
    HGA_CUDA(cudaMemcpyAsync(&n_emit, d_at + n_keys, 8, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));
    h->metrics.kernel_launches += 4;
    if (n_emit == 0) return HGA_OK;
    HGA_TRY(b_emit.ensure((n_emit + 1) * 8));
    HGA_TRY(b_sorted.ensure((n_emit + 1) * 8));
    HGA_TRY(b_runlen.ensure((n_emit + 1) * 4));
    uint64_t *d_emit = b_emit.as<uint64_t>(), *d_sorted = b_sorted.as<uint64_t>();
    uint32_t *d_run_len = b_runlen.as<uint32_t>();
    enr_emit_kernel<<<grid_for(h, n_keys), 256, 0, h->stream>>>(d_keys, n_keys, c.d_poff, c.d_prow, d_at, d_emit);
    const int bits = 32 + (int) std::max<uint32_t>(hga_ceil_log2(n_pivots + 1), 1);
    size_t t1 = 0, t2 = 0;
    HGA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, t1, d_emit, d_sorted, n_emit, 0, bits, h->stream));
    HGA_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, t2, d_sorted, d_emit, d_run_len, c.d_count, n_emit, h->stream));
    HGA_TRY(h->d_sort_tmp.ensure(std::max(t1, t2) + 16));
    HGA_CUDA(cub::DeviceRadixSort::SortKeys(h->d_sort_tmp.p, t1, d_emit, d_sorted, n_emit, 0, bits, h->stream));
    HGA_CUDA(cub::DeviceRunLengthEncode::Encode(h->d_sort_tmp.p, t2, d_sorted, d_emit, d_run_len, c.d_count, n_emit, h->stream));   // run keys in d_emit
    unsigned long long n_runs = 0;
    HGA_CUDA(cudaMemcpyAsync(&n_runs, c.d_count, 8, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));
    HGA_TRY(b_out.ensure((n_runs + 1) * 8));
    HGA_CUDA(cudaMemsetAsync(c.d_count, 0, 8, h->stream));
    amp_filter_kernel<<<grid_for(h, n_runs), 256, 0, h->stream>>>(d_emit, d_run_len, n_runs, d_pivot_row, min_score, c.first_id, b_out.as<uint64_t>(), c.d_count);
    h->metrics.kernel_launches += (uint64_t) (bits + 7) / 8 + 5;
    HGA_CUDA(cudaGetLastError());
    unsigned long long n_out = 0;
    HGA_CUDA(cudaMemcpyAsync(&n_out, c.d_count, 8, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<uint64_t> got(n_out);
    if (n_out) HGA_CUDA(cudaMemcpy(got.data(), b_out.p, n_out * 8, cudaMemcpyDeviceToHost));
    out.reserve(out.size() + n_out);
    for (uint64_t v : got) out.push_back({(uint32_t) (v >> 32), (uint32_t) v});
    return HGA_OK;
}

}  // namespace

// hga_tails.cpp
typedef int (*hga_tail_amplify_fn)(void *ctx, const uint32_t *pivot_id, uint64_t n_pivots, uint32_t min_score, std::vector<std::pair<uint32_t, uint32_t>> &out);
int hga_host_tail_connections_impl(uint64_t n_reads, const uint64_t *row_off, const uint32_t *kmer_id, const uint32_t *pos, const uint32_t *read_len,
                                   uint64_t avg_read_length, uint32_t read_id_first, uint64_t n_comp, const uint64_t *comp_off,
                                   const uint32_t *comp_member, const uint64_t *tree_off, const uint32_t *tree_x, const uint32_t *tree_y,
                                   const uint64_t *purged_off, const uint32_t *purged_read, uint32_t amplification_min_score, uint32_t *out_x,
                                   uint32_t *out_y, uint64_t *out_score, uint64_t *out_n, hga_tail_amplify_fn amplify_fn, void *amplify_ctx);

int hga_enrich_run(hga_handle *h, int min_size, int max_size, uint32_t min_score, const TailParams *tail) {
    if (!h->have_selection || !h->have_index || !h->have_scan) { hga_set_error("hga_enrich: needs hga_scan, hga_build_index and hga_select_edges"); return HGA_E_STATE; }
    if (h->comm && hga_comm_size(h) > 1) { hga_set_error("hga_enrich: not available with a communicator yet (single GPU only)"); return HGA_E_STATE; }
    if (min_size < 2) { hga_set_error("hga_enrich: min_size must be >= 2 (a core is a merged component)"); return HGA_E_ARG; }
    if (max_size != -1 && max_size < 1) { hga_set_error("hga_enrich: max_size must be -1 (no limit) or positive"); return HGA_E_ARG; }
    // --sc_max_size: a union is skipped when the merged component would exceed the limit (:457), so "already connected by earlier
    // edges" no longer means "no-op" and the GPU pre-filter does not apply: the host replays every selected edge
    const bool limited = max_size != -1;
    const uint64_t max_comp = limited ? (uint64_t) max_size : ~0ull;
    h->have_enrichment = false;
    EnrichResult &res = h->enrich;
    res = EnrichResult();
    const uint64_t n = h->inc_rows, M = h->n_selected, E = h->n_hits;
    const uint32_t first_id = h->inc_row_first_id;
    const uint32_t n_slots = h->index_keys;
    StageTimer timer(h, &h->metrics.enrich_ms);
    PhaseClock pc(h->stream, h->metrics.enrich_phase_ms);
    HGA_TRY(h->d_enr_scalars.ensure(64));
    unsigned long long *d_count = h->d_enr_scalars.as<unsigned long long>();

    // ---- 1. roots of the sequential union_find (:453-478) over the selected edges in canonical order ----------------------
    // Only an edge whose endpoints are not yet connected by the edges before it changes the union_find state; every other edge
    // is a no-op in the sequential loop (:455). The GPU walks the canonical list in batches of doubling size and keeps the edges
    // whose endpoints are not connected by the EARLIER batches (a superset of the state-changing edges, ~N of the M edges); the
    // host replays those, in order.
    uint64_t M2 = 0;
    const uint32_t *ex = nullptr, *ey = nullptr;
    if (M) {
        // the selection is stored in (x, y) order; a stable descending sort by score gives (score desc, x asc, y asc)
        HGA_TRY(h->d_export_a.ensure((M + 1) * 8 * 2));
        HGA_TRY(h->d_export_b.ensure((M + 1) * 4 * 3));
        HGA_TRY(h->d_export_c.ensure(M + 64));
        HGA_TRY(h->d_enr_parent.ensure((n + 1) * 4));
        uint64_t *d_key = h->d_export_a.as<uint64_t>(), *d_kept = d_key + (M + 1);
        uint32_t *d_score = h->d_export_b.as<uint32_t>(), *d_x = d_score + (M + 1), *d_y = d_x + (M + 1);
        uint8_t *d_flag = h->d_export_c.as<uint8_t>();
        uint32_t *d_par = h->d_enr_parent.as<uint32_t>();
        size_t tmp = 0, tmp2 = 0;
        HGA_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, h->d_sel_score.as<uint32_t>(), d_score, h->d_sel_key.as<uint64_t>(), d_key, M, 0, 32, h->stream));
        HGA_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp2, d_key, d_flag, d_kept, d_count, M, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(std::max(tmp, tmp2) + 16));
        HGA_CUDA(cub::DeviceRadixSort::SortPairsDescending(h->d_sort_tmp.p, tmp, h->d_sel_score.as<uint32_t>(), d_score, h->d_sel_key.as<uint64_t>(), d_key, M, 0, 32, h->stream));
        enr_iota_kernel<<<grid_for(h, n), 256, 0, h->stream>>>(d_par, n);
        h->metrics.kernel_launches += 6;
        unsigned long long kept = M;
        if (limited) {
            std::swap(d_key, d_kept);          // every edge, in canonical order
        } else {
        for (uint64_t lo = 0, batch = 1 << 16; lo < M; lo += batch, batch *= 2) {
            const uint64_t hi = std::min(M, lo + batch);
            enr_edge_filter_kernel<<<grid_for(h, hi - lo), 256, 0, h->stream>>>(d_key, lo, hi, d_par, d_flag);
            enr_edge_hook_kernel<<<grid_for(h, hi - lo), 256, 0, h->stream>>>(d_key, lo, hi, d_par, d_flag);
            h->metrics.kernel_launches += 2;
        }
        HGA_CUDA(cub::DeviceSelect::Flagged(h->d_sort_tmp.p, tmp2, d_key, d_flag, d_kept, d_count, M, h->stream));
        HGA_CUDA(cudaMemcpyAsync(&kept, d_count, 8, cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        }
        // second level: the survivors of the doubling batches (a large batch admits many edges between the same two components:
        // 2.56 M of 14.0 M at config 4) once more, in small fixed batches; what the host replays is then close to the N - 1 edges
        // that actually join components
        if (!limited && kept > (1u << 16)) {
            enr_iota_kernel<<<grid_for(h, n), 256, 0, h->stream>>>(d_par, n);
            const uint64_t small = 1 << 14;
            for (uint64_t lo = 0; lo < kept; lo += small) {
                const uint64_t hi = std::min<uint64_t>(kept, lo + small);
                enr_edge_filter_kernel<<<grid_for(h, hi - lo), 256, 0, h->stream>>>(d_kept, lo, hi, d_par, d_flag);
                enr_edge_hook_kernel<<<grid_for(h, hi - lo), 256, 0, h->stream>>>(d_kept, lo, hi, d_par, d_flag);
                h->metrics.kernel_launches += 2;
            }
            HGA_CUDA(cub::DeviceSelect::Flagged(h->d_sort_tmp.p, tmp2, d_kept, d_flag, d_key, d_count, kept, h->stream));   // d_key is free by now
            HGA_CUDA(cudaMemcpyAsync(&kept, d_count, 8, cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            std::swap(d_key, d_kept);
        }
        M2 = kept;
        enr_key_rows_kernel<<<grid_for(h, M2), 256, 0, h->stream>>>(d_kept, M2, d_x, d_y);
        h->metrics.kernel_launches += 3;
        HGA_CUDA(cudaGetLastError());
        HGA_TRY(h->h_sx.ensure((M2 + 1) * 4)); HGA_TRY(h->h_sy.ensure((M2 + 1) * 4));
        HGA_CUDA(cudaMemcpyAsync(h->h_sx.p, d_x, M2 * 4, cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaMemcpyAsync(h->h_sy.p, d_y, M2 * 4, cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        ex = h->h_sx.as<uint32_t>(); ey = h->h_sy.as<uint32_t>();
    }
    // pivot subsets (--sc_score): a pair with ONE pivot endpoint is in the reference's list in one direction only, pivot first
    std::vector<uint8_t> is_pivot;
    if (h->pair_subset && M2) {
        is_pivot.resize(n + 1);
        HGA_CUDA(cudaMemcpyAsync(is_pivot.data(), h->d_pivot_flag.p, n, cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
    }
    pc.mark("selection sort + D2H", 0);
    std::vector<uint32_t> parent(n + 1), size(n + 1, 1);
    std::vector<uint8_t> touched(n + 1, 0);
    std::iota(parent.begin(), parent.end(), 0u);
    std::vector<std::pair<uint32_t, uint32_t>> tree_edges;    // the edges that performed a union = the spanning forest (:474-476); tail block only
    for (uint64_t i = 0; i < M2; i++) {
        // the reference's list holds (x, y) and (y, x) back to back; the second is a no-op after the first
        uint32_t x = ex[i], y = ey[i];
        if (!is_pivot.empty() && !is_pivot[x]) std::swap(x, y);                                // (pivot, non-pivot) as get_connections emits it
        touched[x] = touched[y] = 1;                                                           // :427-431
        const uint32_t px = dsu_find(parent, x), py = dsu_find(parent, y);                     // :453-454
        if (px == py) continue;                                                                // :455
        if ((uint64_t) size[px] + size[py] > max_comp) continue;                               // :457
        const uint32_t bigger = size[px] > size[py] ? px : py, smaller = bigger == px ? py : px;   // :459-466 (ties: y's root)
        parent[smaller] = bigger;                                                              // :468-470
        size[bigger] += size[smaller];                                                         // :471
        if (tail) tree_edges.push_back({x, y});                                                // :474
    }
    if (pc.on) fprintf(stderr, "hga_enrich: %llu of %llu selected edges replayed\n", (unsigned long long) M2, (unsigned long long) M);
    pc.mark("  replay loop", 0);
    // cores = components with >= min_size vertices (:482-486), identified by their root = element [0] = the survivor (:366)
    std::vector<int32_t> core_of(n + 1, -1);
    std::vector<uint32_t> surv_row;
    for (uint64_t r = 0; r < n; r++)
        if (touched[r] && parent[r] == r && size[r] >= (uint32_t) min_size) { core_of[r] = (int32_t) surv_row.size(); surv_row.push_back((uint32_t) r); }
    uint32_t C = (uint32_t) surv_row.size();
    for (uint64_t r = 0; r < n; r++)
        if (touched[r] && parent[r] != r) core_of[r] = core_of[dsu_find(parent, (uint32_t) r)];
    // cores of the result: survivor ids ascending, members ascending (again after the tail block, which merges cores)
    auto fill_cores = [&]() {
        res.core_id.resize(C);
        res.core_off.assign(C + 1, 0);
        for (uint64_t r = 0; r < n; r++) if (core_of[r] >= 0) res.core_off[core_of[r] + 1]++;
        for (uint32_t c = 0; c < C; c++) { res.core_id[c] = surv_row[c] + first_id; res.core_off[c + 1] += res.core_off[c]; }
        res.core_read.resize(res.core_off[C]);
        std::vector<uint64_t> cur(res.core_off.begin(), res.core_off.end() - 1);
        for (uint64_t r = 0; r < n; r++) if (core_of[r] >= 0) res.core_read[cur[core_of[r]]++] = (uint32_t) r + first_id;
    };
    fill_cores();
    res.n_scaffold_cores = C;

    pc.mark("host root replay + cores", 0);
    // ---- 2. GPU: unions, removal bounds, purged index ----------------------------------------------------------------------
    HGA_TRY(h->d_enr_core_of.ensure((n + 1) * 4));
    HGA_TRY(h->d_enr_surv.ensure(((size_t) C + 1) * 4 * 2));        // second half: the survivors after the tail block's merge
    HGA_TRY(h->d_enr_R.ensure(((size_t) n_slots + 1) * 4));
    int32_t *d_core_of = h->d_enr_core_of.as<int32_t>();
    uint32_t *d_surv = h->d_enr_surv.as<uint32_t>(), *d_R = h->d_enr_R.as<uint32_t>();
    HGA_CUDA(cudaMemcpyAsync(d_core_of, core_of.data(), (n + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    if (C) HGA_CUDA(cudaMemcpyAsync(d_surv, surv_row.data(), (size_t) C * 4, cudaMemcpyHostToDevice, h->stream));
    HGA_CUDA(cudaMemsetAsync(d_R, 0, ((size_t) n_slots + 1) * 4, h->stream));

    const uint32_t *inv_off = h->d_inv_off.as<uint32_t>(), *inv_row = h->d_inv_row.as<uint32_t>();
    uint64_t n_u = 0;
    HGA_TRY(h->d_purged_off.ensure(((size_t) n_slots + 2) * 4 * 2));
    uint32_t *d_cnt = h->d_purged_off.as<uint32_t>() + (n_slots + 2), *d_poff = h->d_purged_off.as<uint32_t>();
    HGA_TRY(h->d_enr_keys.ensure(64));
    if (E && C) {
        HGA_TRY(h->d_export_a.ensure(((size_t) n_slots + 2) * 8));
        unsigned long long *d_cand_off = h->d_export_a.as<unsigned long long>();
        enr_list_cores_kernel<false><<<grid_for(h, n_slots), 256, 0, h->stream>>>(inv_off, inv_row, n_slots, d_core_of, d_cnt, nullptr, nullptr, d_R);
        HGA_CUDA(cudaMemsetAsync(d_cnt + n_slots, 0, 4, h->stream));
        size_t tmp = 0;
        HGA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_cnt, d_cand_off, (uint64_t) n_slots + 1, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp + 16));
        HGA_CUDA(cub::DeviceScan::ExclusiveSum(h->d_sort_tmp.p, tmp, d_cnt, d_cand_off, (uint64_t) n_slots + 1, h->stream));
        unsigned long long n_cand = 0;
        HGA_CUDA(cudaMemcpyAsync(&n_cand, d_cand_off + n_slots, 8, cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        HGA_TRY(h->d_enr_keys.ensure((n_cand + 1) * 8));
        HGA_TRY(h->d_enr_keys2.ensure((n_cand + 1) * 8));
        uint64_t *d_keys = h->d_enr_keys.as<uint64_t>(), *d_sorted = h->d_enr_keys2.as<uint64_t>();
        h->metrics.kernel_launches += 3;
        if (n_cand) {
            enr_list_cores_kernel<true><<<grid_for(h, n_slots), 256, 0, h->stream>>>(inv_off, inv_row, n_slots, d_core_of, nullptr, d_cand_off, d_keys, nullptr);
            const int bits = 32 + (int) std::max<uint32_t>(hga_ceil_log2((uint64_t) C), 1);
            size_t t1 = 0, t2 = 0;
            HGA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, t1, d_keys, d_sorted, n_cand, 0, bits, h->stream));
            HGA_CUDA(cub::DeviceSelect::Unique(nullptr, t2, d_sorted, d_keys, d_count, n_cand, h->stream));
            HGA_TRY(h->d_sort_tmp.ensure(std::max(t1, t2) + 16));
            HGA_CUDA(cub::DeviceRadixSort::SortKeys(h->d_sort_tmp.p, t1, d_keys, d_sorted, n_cand, 0, bits, h->stream));
            HGA_CUDA(cub::DeviceSelect::Unique(h->d_sort_tmp.p, t2, d_sorted, d_keys, d_count, n_cand, h->stream));   // unique keys back in d_keys
            h->metrics.kernel_launches += (uint64_t) (bits + 7) / 8 + 4;
            HGA_CUDA(cudaGetLastError());
            unsigned long long nu = 0;
            HGA_CUDA(cudaMemcpyAsync(&nu, d_count, 8, cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            n_u = nu;
        }
    }
    const uint64_t *d_keys = h->d_enr_keys.as<uint64_t>();
    pc.mark("core k-mer unions", 0);
    const uint64_t *d_u = d_keys;      // sorted unique (core, slot)
    HGA_TRY(h->d_enr_core_koff.ensure(((size_t) C + 2) * 8));
    unsigned long long *d_core_koff = h->d_enr_core_koff.as<unsigned long long>();
    enr_core_bounds_kernel<<<grid_for(h, (uint64_t) C + 1), 256, 0, h->stream>>>(d_u, n_u, C, d_core_koff);
    if (n_u) enr_survivor_max_kernel<<<grid_for(h, n_u), 256, 0, h->stream>>>(d_u, n_u, d_surv, d_R);
    h->metrics.kernel_launches += 2;

    enr_purge_kernel<false><<<grid_for(h, n_slots), 256, 0, h->stream>>>(inv_off, inv_row, n_slots, d_R, d_core_of, d_surv, d_cnt, nullptr, nullptr);
    HGA_CUDA(cudaMemsetAsync(d_cnt + n_slots, 0, 4, h->stream));
    {
        size_t tmp = 0;
        HGA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_cnt, d_poff, (uint64_t) n_slots + 1, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp + 16));
        HGA_CUDA(cub::DeviceScan::ExclusiveSum(h->d_sort_tmp.p, tmp, d_cnt, d_poff, (uint64_t) n_slots + 1, h->stream));
    }
    uint32_t n_purged = 0;
    HGA_CUDA(cudaMemcpyAsync(&n_purged, d_poff + n_slots, 4, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));
    HGA_TRY(h->d_purged_row.ensure(((size_t) n_purged + 1) * 4));
    uint32_t *d_prow = h->d_purged_row.as<uint32_t>();
    enr_purge_kernel<true><<<grid_for(h, n_slots), 256, 0, h->stream>>>(inv_off, inv_row, n_slots, d_R, d_core_of, d_surv, nullptr, d_poff, d_prow);
    h->metrics.kernel_launches += 4;
    HGA_CUDA(cudaGetLastError());
    h->n_purged = n_purged;
    h->n_core_kmers = n_u;

    pc.mark("purge", 0);
    // ---- 2b. tail / spectral block (:768-777), on request: host stages on the merged state, second merge on the GPU -------------
    if (tail && C > 2) {
        res.tail_block_ran = true;
        // the engine state the host stages read: hits by read sorted by (kmer_id, pos), purged index by kmer_id, read lengths, the
        // scaffold components (survivor first) with the spanning trees of the replayed union_find
        hga_hits hits;
        HGA_TRY(hga_get_hits(h, 1, &hits));
        if (hits.n_reads != n) { hga_set_error("hga_enrich_full: the scan covers %llu reads, the index %llu rows", (unsigned long long) hits.n_reads, (unsigned long long) n); return HGA_E_STATE; }
        // (the purged index stays on the GPU: the one step that reads it, the tail amplification, runs there - gpu_tail_amplify)
        std::vector<uint32_t> read_len(n + 1);
        for (uint64_t r = 0; r < n; r++) read_len[r] = (uint32_t) (tail->read_off[r + 1] - tail->read_off[r]);
        const uint64_t avg_read_length = n ? (tail->read_off[n] - tail->read_off[0]) / n : 0;      // SequenceRecordIterator.cpp:64
        std::vector<uint64_t> comp_off(C + 1, 0), tree_off(C + 1, 0);
        std::vector<uint32_t> comp_member(res.core_read.size());
        for (uint32_t c = 0; c < C; c++) {
            comp_off[c] = res.core_off[c];
            uint64_t w = res.core_off[c];
            comp_member[w++] = surv_row[c] + first_id;                                              // element [0] = the root (:366)
            for (uint64_t i = res.core_off[c]; i < res.core_off[c + 1]; i++) if (res.core_read[i] != surv_row[c] + first_id) comp_member[w++] = res.core_read[i];
        }
        comp_off[C] = res.core_off[C];
        for (const auto &e : tree_edges) if (core_of[e.first] >= 0) tree_off[core_of[e.first] + 1]++;
        for (uint32_t c = 0; c < C; c++) tree_off[c + 1] += tree_off[c];
        std::vector<uint32_t> tree_x(tree_off[C] + 1), tree_y(tree_off[C] + 1);
        {
            std::vector<uint64_t> cur(tree_off.begin(), tree_off.end() - 1);
            for (const auto &e : tree_edges) if (core_of[e.first] >= 0) { const uint64_t w = cur[core_of[e.first]]++; tree_x[w] = e.first + first_id; tree_y[w] = e.second + first_id; }
        }
        const uint64_t cap = (uint64_t) C * (C - 1) / 2;
        res.tconn_x.resize(cap); res.tconn_y.resize(cap); res.tconn_score.resize(cap);
        uint64_t n_t = 0;
        AmplifyCtx actx{h, &core_of, &surv_row, hits.row_off, first_id, C, d_u, d_core_koff, d_poff, d_prow, d_count};
        HGA_TRY(hga_host_tail_connections_impl(n, hits.row_off, hits.kmer_id, hits.pos, read_len.data(), avg_read_length, first_id, C, comp_off.data(), comp_member.data(),
                                               tree_off.data(), tree_x.data(), tree_y.data(), nullptr, nullptr, tail->amplification_min_score,
                                               res.tconn_x.data(), res.tconn_y.data(), res.tconn_score.data(), &n_t, gpu_tail_amplify, &actx));
        res.tconn_x.resize(n_t); res.tconn_y.resize(n_t); res.tconn_score.resize(n_t);
        pc.mark("  tail connections (host)", 1);
        uint64_t n_strong = 0;                                                                      // :770 score > 5; the list is score-descending
        while (n_strong < n_t && res.tconn_score[n_strong] > 5) n_strong++;
        res.cluster_off.assign(1, 0);
        if (n_strong) {
            std::vector<uint32_t> member(2 * n_strong + 1);
            std::vector<uint64_t> coff((size_t) tail->spectral_dims + 2, 0);
            uint64_t n_nodes = 0, n_cl = 0;
            HGA_TRY(hga_spectral_clustering(res.tconn_x.data(), res.tconn_y.data(), res.tconn_score.data(), n_strong, tail->spectral_dims, member.data(), coff.data(),
                                            &n_nodes, &n_cl));
            for (uint64_t i = 0; i < n_cl; i++) {
                if (coff[i + 1] == coff[i]) continue;      // an empty cluster (a rotated dimension no point prefers) would make merge_components read element [0] of an empty vector
                res.cluster_member.insert(res.cluster_member.end(), member.begin() + (ptrdiff_t) coff[i], member.begin() + (ptrdiff_t) coff[i + 1]);
                res.cluster_off.push_back(res.cluster_member.size());
            }
        }
        pc.mark("  spectral clustering (host)", 2);
        // which core every core merges into (itself unless its cluster has several members), new compact numbering of the survivors
        std::vector<uint32_t> into(C), multi(C, 0);
        std::iota(into.begin(), into.end(), 0u);
        bool any_multi = false;
        for (size_t i = 0; i + 1 < res.cluster_off.size(); i++) {
            const uint64_t a = res.cluster_off[i], b = res.cluster_off[i + 1];
            if (b - a < 2) continue;                                                                // :360-365
            const int32_t s = core_of[res.cluster_member[a] - first_id];
            for (uint64_t j = a; j < b; j++) {
                const int32_t c = core_of[res.cluster_member[j] - first_id];
                if (c < 0 || s < 0 || surv_row[c] + first_id != res.cluster_member[j]) { hga_set_error("hga_enrich_full: a spectral cluster names %u, which is no scaffold survivor", res.cluster_member[j]); return HGA_E_STATE; }
                into[c] = (uint32_t) s; multi[c] = 1; any_multi = true;
            }
        }
        if (any_multi) {
            std::vector<uint32_t> new_idx(C, 0), surv_new;
            for (uint32_t c = 0; c < C; c++) if (into[c] == c) { new_idx[c] = (uint32_t) surv_new.size(); surv_new.push_back(surv_row[c]); }
            const uint32_t C2 = (uint32_t) surv_new.size();
            std::vector<uint32_t> map(C);
            for (uint32_t c = 0; c < C; c++) map[c] = new_idx[into[c]] | (multi[c] << 31);
            std::vector<uint8_t> removed_once(n + 1, 0);
            for (uint32_t c = 0; c < C; c++) if (multi[c]) removed_once[surv_row[c]] = 1;
            // device: map (in the core_koff buffer's place is too small: use d_enr_parent, free since the root replay), flags, survivors
            HGA_TRY(h->d_enr_parent.ensure((size_t) C * 4 + n + 64));
            uint32_t *d_map = h->d_enr_parent.as<uint32_t>();
            uint8_t *d_once = reinterpret_cast<uint8_t *>(d_map + C);
            uint32_t *d_surv_new = d_surv + (C + 1);
            HGA_CUDA(cudaMemcpyAsync(d_map, map.data(), (size_t) C * 4, cudaMemcpyHostToDevice, h->stream));
            HGA_CUDA(cudaMemcpyAsync(d_once, removed_once.data(), n, cudaMemcpyHostToDevice, h->stream));
            HGA_CUDA(cudaMemcpyAsync(d_surv_new, surv_new.data(), (size_t) C2 * 4, cudaMemcpyHostToDevice, h->stream));
            HGA_CUDA(cudaMemsetAsync(d_R, 0, ((size_t) n_slots + 1) * 4, h->stream));
            // unions of the clusters: relabel, sort, unique (back into d_enr_keys)
            uint64_t n_u2 = 0;
            if (n_u) {
                HGA_TRY(h->d_enr_keys2.ensure((n_u + 1) * 8));
                HGA_TRY(h->d_export_a.ensure((n_u + 1) * 8));
                uint64_t *d_rel = h->d_enr_keys2.as<uint64_t>(), *d_srt = h->d_export_a.as<uint64_t>(), *d_uniq = h->d_enr_keys.as<uint64_t>();
                enr_merge2_keys_kernel<<<grid_for(h, n_u), 256, 0, h->stream>>>(d_u, n_u, d_map, d_surv, d_surv_new, d_R, d_rel);
                const int bits = 32 + (int) std::max<uint32_t>(hga_ceil_log2((uint64_t) C2), 1);
                size_t t1 = 0, t2 = 0;
                HGA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, t1, d_rel, d_srt, n_u, 0, bits, h->stream));
                HGA_CUDA(cub::DeviceSelect::Unique(nullptr, t2, d_srt, d_uniq, d_count, n_u, h->stream));
                HGA_TRY(h->d_sort_tmp.ensure(std::max(t1, t2) + 16));
                HGA_CUDA(cub::DeviceRadixSort::SortKeys(h->d_sort_tmp.p, t1, d_rel, d_srt, n_u, 0, bits, h->stream));
                HGA_CUDA(cub::DeviceSelect::Unique(h->d_sort_tmp.p, t2, d_srt, d_uniq, d_count, n_u, h->stream));
                h->metrics.kernel_launches += (uint64_t) (bits + 7) / 8 + 4;
                HGA_CUDA(cudaGetLastError());
                unsigned long long nu = 0;
                HGA_CUDA(cudaMemcpyAsync(&nu, d_count, 8, cudaMemcpyDeviceToHost, h->stream));
                HGA_CUDA(cudaStreamSynchronize(h->stream));
                n_u2 = nu;
            }
            // second purge: purged index -> d_purged2_*, then the buffers change places
            HGA_TRY(h->d_purged2_off.ensure(((size_t) n_slots + 2) * 4 * 2));
            uint32_t *d_poff2 = h->d_purged2_off.as<uint32_t>(), *d_cnt2 = d_poff2 + (n_slots + 2);
            enr_purge2_kernel<false><<<grid_for(h, n_slots), 256, 0, h->stream>>>(d_poff, d_prow, n_slots, d_R, d_once, d_cnt2, nullptr, nullptr);
            HGA_CUDA(cudaMemsetAsync(d_cnt2 + n_slots, 0, 4, h->stream));
            {
                size_t tmp = 0;
                HGA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_cnt2, d_poff2, (uint64_t) n_slots + 1, h->stream));
                HGA_TRY(h->d_sort_tmp.ensure(tmp + 16));
                HGA_CUDA(cub::DeviceScan::ExclusiveSum(h->d_sort_tmp.p, tmp, d_cnt2, d_poff2, (uint64_t) n_slots + 1, h->stream));
            }
            uint32_t n_purged2 = 0;
            HGA_CUDA(cudaMemcpyAsync(&n_purged2, d_poff2 + n_slots, 4, cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            HGA_TRY(h->d_purged2_row.ensure(((size_t) n_purged2 + 1) * 4));
            uint32_t *d_prow2 = h->d_purged2_row.as<uint32_t>();
            enr_purge2_kernel<true><<<grid_for(h, n_slots), 256, 0, h->stream>>>(d_poff, d_prow, n_slots, d_R, d_once, nullptr, d_poff2, d_prow2);
            h->metrics.kernel_launches += 4;
            HGA_CUDA(cudaGetLastError());
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            std::swap(h->d_purged_off, h->d_purged2_off);
            std::swap(h->d_purged_row, h->d_purged2_row);
            d_poff = d_poff2; d_prow = d_prow2; n_purged = n_purged2;
            h->n_purged = n_purged2;
            // the merged state in the new core numbering
            for (uint64_t r = 0; r < n; r++) if (core_of[r] >= 0) core_of[r] = (int32_t) new_idx[into[core_of[r]]];
            surv_row = surv_new;
            C = C2;
            d_surv = d_surv_new;
            n_u = n_u2;
            h->n_core_kmers = n_u;
            enr_core_bounds_kernel<<<grid_for(h, (uint64_t) C + 1), 256, 0, h->stream>>>(d_u, n_u, C, d_core_koff);
            h->metrics.kernel_launches++;
            fill_cores();
        }
        pc.mark("  merge of the spectral clusters", 3);
    }
    // ---- 3. GPU: enrichment connections = run lengths of the sorted (core, partner) emissions ------------------------------
    std::vector<Conn> conns;
    if (n_u) {
        HGA_TRY(h->d_export_a.ensure((n_u + 2) * 8 * 2));
        unsigned long long *d_len = h->d_export_a.as<unsigned long long>(), *d_at = d_len + (n_u + 2);
        enr_emit_len_kernel<<<grid_for(h, n_u), 256, 0, h->stream>>>(d_u, n_u, d_poff, d_len);
        HGA_CUDA(cudaMemsetAsync(d_len + n_u, 0, 8, h->stream));
        size_t tmp = 0;
        HGA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_len, d_at, n_u + 1, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp + 16));
        HGA_CUDA(cub::DeviceScan::ExclusiveSum(h->d_sort_tmp.p, tmp, d_len, d_at, n_u + 1, h->stream));
        unsigned long long n_emit = 0;
        HGA_CUDA(cudaMemcpyAsync(&n_emit, d_at + n_u, 8, cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        h->metrics.kernel_launches += 3;
        if (n_emit) {
            HGA_TRY(h->d_enr_keys2.ensure((n_emit + 1) * 8));          // free again (the unique keys live in d_enr_keys)
            HGA_TRY(h->d_export_b.ensure((n_emit + 1) * 8));
            HGA_TRY(h->d_export_c.ensure((n_emit + 1) * 4 * 4));
            uint64_t *d_emit = h->d_enr_keys2.as<uint64_t>(), *d_sorted = h->d_export_b.as<uint64_t>();
            enr_emit_kernel<<<grid_for(h, n_u), 256, 0, h->stream>>>(d_u, n_u, d_poff, d_prow, d_at, d_emit);
            const int bits = 32 + (int) std::max<uint32_t>(hga_ceil_log2((uint64_t) C + 1), 1);
            size_t t1 = 0, t2 = 0;
            uint32_t *d_run_len = h->d_export_c.as<uint32_t>();
            HGA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, t1, d_emit, d_sorted, n_emit, 0, bits, h->stream));
            HGA_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, t2, d_sorted, d_emit, d_run_len, d_count, n_emit, h->stream));
            HGA_TRY(h->d_sort_tmp.ensure(std::max(t1, t2) + 16));
            HGA_CUDA(cub::DeviceRadixSort::SortKeys(h->d_sort_tmp.p, t1, d_emit, d_sorted, n_emit, 0, bits, h->stream));
            HGA_CUDA(cub::DeviceRunLengthEncode::Encode(h->d_sort_tmp.p, t2, d_sorted, d_emit, d_run_len, d_count, n_emit, h->stream));   // run keys in d_emit
            unsigned long long n_runs = 0;
            HGA_CUDA(cudaMemcpyAsync(&n_runs, d_count, 8, cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            uint32_t *d_cx = d_run_len + (n_emit + 1), *d_cy = d_cx + (n_emit + 1), *d_cs = d_cy + (n_emit + 1);
            HGA_CUDA(cudaMemsetAsync(d_count, 0, 8, h->stream));
            enr_filter_kernel<<<grid_for(h, n_runs), 256, 0, h->stream>>>(d_emit, d_run_len, n_runs, d_surv, min_score, first_id, d_cx, d_cy, d_cs, d_count);
            h->metrics.kernel_launches += (uint64_t) (bits + 7) / 8 + 6;
            HGA_CUDA(cudaGetLastError());
            unsigned long long n_conn = 0;
            HGA_CUDA(cudaMemcpyAsync(&n_conn, d_count, 8, cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            std::vector<uint32_t> cx(n_conn), cy(n_conn), cs(n_conn);
            if (n_conn) {
                // canonical order (score desc, min asc, max asc, x asc) = three stable radix sorts of a permutation, least significant
                // key first: x, then (min << 32 | max), then ~score
                HGA_TRY(h->d_enr_keys2.ensure((n_conn + 1) * (8 + 8 + 4 + 4 + 4 + 4 + 4 * 3)));
                uint64_t *k64_in = h->d_enr_keys2.as<uint64_t>(), *k64_out = k64_in + (n_conn + 1);
                uint32_t *k32_in = reinterpret_cast<uint32_t *>(k64_out + (n_conn + 1)), *k32_out = k32_in + (n_conn + 1);
                uint32_t *perm_a = k32_out + (n_conn + 1), *perm_b = perm_a + (n_conn + 1);
                uint32_t *ox = perm_b + (n_conn + 1), *oy = ox + (n_conn + 1), *os = oy + (n_conn + 1);
                size_t t32 = 0, t64 = 0;
                HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t32, k32_in, k32_out, perm_a, perm_b, n_conn, 0, 32, h->stream));
                HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t64, k64_in, k64_out, perm_a, perm_b, n_conn, 0, 64, h->stream));
                HGA_TRY(h->d_sort_tmp.ensure(std::max(t32, t64) + 16));
                const int g = grid_for(h, n_conn);
                enr_conn_key_kernel<0><<<g, 256, 0, h->stream>>>(d_cx, d_cy, d_cs, nullptr, n_conn, k32_in, k64_in, perm_a);            // x, perm = iota
                HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, t32, k32_in, k32_out, perm_a, perm_b, n_conn, 0, 32, h->stream));
                enr_conn_key_kernel<1><<<g, 256, 0, h->stream>>>(d_cx, d_cy, d_cs, perm_b, n_conn, k32_in, k64_in, nullptr);            // (min, max) in perm order
                HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, t64, k64_in, k64_out, perm_b, perm_a, n_conn, 0, 64, h->stream));
                enr_conn_key_kernel<2><<<g, 256, 0, h->stream>>>(d_cx, d_cy, d_cs, perm_a, n_conn, k32_in, k64_in, nullptr);            // ~score in perm order
                HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, t32, k32_in, k32_out, perm_a, perm_b, n_conn, 0, 32, h->stream));
                enr_conn_gather_kernel<<<g, 256, 0, h->stream>>>(d_cx, d_cy, d_cs, perm_b, n_conn, ox, oy, os);
                h->metrics.kernel_launches += 20;
                HGA_CUDA(cudaGetLastError());
                HGA_CUDA(cudaMemcpyAsync(cx.data(), ox, n_conn * 4, cudaMemcpyDeviceToHost, h->stream));
                HGA_CUDA(cudaMemcpyAsync(cy.data(), oy, n_conn * 4, cudaMemcpyDeviceToHost, h->stream));
                HGA_CUDA(cudaMemcpyAsync(cs.data(), os, n_conn * 4, cudaMemcpyDeviceToHost, h->stream));
                HGA_CUDA(cudaStreamSynchronize(h->stream));
            }
            conns.resize(n_conn);
            for (uint64_t i = 0; i < n_conn; i++) conns[i] = {cx[i], cy[i], cs[i]};
        }
    }
    pc.mark("enrichment connections", 4);

    // ---- 4. host: canonical order, restricted union_find (:424-489 with restricted = cores, min 2, max -1), final merge -----
    // the connections arrive in canonical order (sorted on the GPU above)
    res.conn_x.resize(conns.size()); res.conn_y.resize(conns.size()); res.conn_score.resize(conns.size());
    for (size_t i = 0; i < conns.size(); i++) { res.conn_x[i] = conns[i].x; res.conn_y[i] = conns[i].y; res.conn_score[i] = conns[i].s; }
    pc.mark("  canonical sort of the connections", 4);
    std::iota(parent.begin(), parent.end(), 0u);
    std::fill(size.begin(), size.end(), 1u);
    std::vector<uint8_t> restricted(n + 1, 0), affected(n + 1, 0);
    for (uint32_t c = 0; c < C; c++) restricted[surv_row[c]] = 1;
    for (const Conn &cn : conns) {
        const uint32_t x = cn.x - first_id, y = cn.y - first_id;
        affected[x] = affected[y] = 1;
        const uint32_t px = dsu_find(parent, x), py = dsu_find(parent, y);
        if (px == py) continue;
        if (restricted[px] && restricted[py]) continue;                                        // :456
        const uint32_t bigger = size[px] > size[py] ? px : py, smaller = bigger == px ? py : px;
        parent[smaller] = bigger;
        size[bigger] += size[smaller];
        restricted[bigger] |= restricted[smaller];                                             // :478
    }
    pc.mark("  restricted union-find loop", 5);
    // final id of every read: the root of its enrichment component when that has >= 2 vertices (union_find's min_size = 2),
    // otherwise the id it had; a core's members follow their survivor. get_component_ids keeps ids with >= min_size reads.
    std::vector<uint32_t> final_of(n + 1, 0xFFFFFFFFu);    // row -> final survivor row
    for (uint64_t r = 0; r < n; r++) {
        uint32_t v;
        if (core_of[r] >= 0) v = surv_row[core_of[r]];
        else if (affected[r]) v = (uint32_t) r;
        else continue;
        if (affected[v]) v = dsu_find(parent, v);
        final_of[r] = v;
    }
    std::vector<uint32_t> fcount(n + 1, 0);
    for (uint64_t r = 0; r < n; r++) if (final_of[r] != 0xFFFFFFFFu) fcount[final_of[r]]++;
    // components ordered by their smallest member
    std::vector<uint32_t> order_of(n + 1, 0xFFFFFFFFu);
    uint32_t n_final = 0;
    for (uint64_t r = 0; r < n; r++) {
        const uint32_t f = final_of[r];
        if (f == 0xFFFFFFFFu || fcount[f] < (uint32_t) min_size) continue;
        if (order_of[f] == 0xFFFFFFFFu) { order_of[f] = n_final++; res.final_id.push_back(f + first_id); res.final_off.push_back(fcount[f]); }
    }
    {
        uint64_t acc = 0;
        for (auto &v : res.final_off) { const uint64_t c = v; v = acc; acc += c; }
        res.final_off.push_back(acc);
        res.final_read.resize(acc);
        std::vector<uint64_t> cur(res.final_off.begin(), res.final_off.end() - 1);
        res.assignment.assign(n, 0);
        for (uint64_t r = 0; r < n; r++) {
            const uint32_t f = final_of[r];
            if (f == 0xFFFFFFFFu || order_of[f] == 0xFFFFFFFFu) continue;
            res.final_read[cur[order_of[f]]++] = (uint32_t) r + first_id;
            res.assignment[r] = f + first_id;
        }
    }
    pc.mark("host restricted union-find", 5);
    timer.stop();
    h->metrics.n_cores = C; h->metrics.n_enrich_connections = conns.size(); h->metrics.n_final_components = n_final;
    h->have_enrichment = true;
    return HGA_OK;
}
