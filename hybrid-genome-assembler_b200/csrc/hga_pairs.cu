// pair_count: score(x,y) = sum_k mult_x(k) * mult_y(k) for all read pairs sharing a discriminative k-mer.
//
// Replaces ReadClusteringEngine::get_connections / get_all_connections
// (clustering/ReadClusteringEngine.cpp:301-339): per pivot, the reference walks the pivot's k-mer id list
// (duplicates included) and every entry of each k-mer's inverted list (duplicates included) and does
// ++count[candidate] in a tsl::robin_map. This is a sparse integer A * A^T; here it is row-wise Gustavson:
//   * one CTA per pivot row, a shared-memory open-addressing accumulator sized from the row's work,
//   * one warp per incidence entry streams that k-mer's inverted list with coalesced loads,
//   * every unordered pair is produced once (candidate > pivot), not twice as in the reference,
//   * rows whose partner set does not fit shared memory go to a second kernel with a per-CTA table in HBM.
// Output: (key = min_row << 32 | max_row, score) appended through one atomic cursor per row, then one
// radix sort by key puts the pairs in the canonical (x asc, y asc) order that the tie rule of
// hga_select_edges needs.
#include "hga_internal.cuh"

#include <cub/device/device_radix_sort.cuh>

#define PC_THREADS 128
#define PC_WARPS (PC_THREADS / 32)
#define PC_CMAX 4096
#define PC_EMPTY 0xFFFFFFFFu
#define HV_THREADS 256

struct PairScalars {
    unsigned long long ticket;
    unsigned long long cursor;        // entries wanted (may exceed capacity)
    unsigned long long heavy_count;
    unsigned long long heavy_ticket;
    unsigned long long increments;
    unsigned int overflow;
    unsigned int pad;
};

namespace {

struct PairParams {
    const uint64_t *row_off;
    const uint32_t *row_slot;
    uint64_t n_rows;
    const uint64_t *inv_off;
    const uint32_t *inv_row;
    const uint32_t *pivot_rows;     // nullptr: every row is a pivot
    const uint8_t *pivot_flag;      // nullptr when every row is a pivot
    uint64_t n_pivots;
    uint32_t min_score;
    uint64_t *out_key;
    uint32_t *out_score;
    uint64_t capacity;
    uint32_t *heavy_list;
    uint32_t *heavy_tab;            // per-CTA tables of heavy_cap (key,val) pairs
    uint32_t heavy_cap;             // power of two
    PairScalars *sc;
};

__device__ __forceinline__ bool keep_candidate(uint32_t x, uint32_t y, const uint8_t *pivot_flag) {
    if (y == x) return false;                         // :317 erase(pivot)
    if (pivot_flag == nullptr) return y > x;          // all rows are pivots: count each unordered pair once
    return !pivot_flag[y] || y > x;                   // pivot subset: pair (pivot, non-pivot) or ordered pivot pair
}

__device__ __forceinline__ uint32_t hash_row(uint32_t y) { return y * 2654435761u; }

template<int THREADS>
__device__ __forceinline__ unsigned long long block_sum(unsigned long long v, unsigned long long *s_red) {
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
    #pragma unroll
    for (int i = 0; i < THREADS / 32; i++) t += s_red[i];
    __syncthreads();
    return t;
}

// exclusive prefix of v over the CTA, total in *total
template<int THREADS>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *s_red, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_red[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    #pragma unroll
    for (int i = 0; i < THREADS / 32; i++) { if (i < warp) base += s_red[i]; tot += s_red[i]; }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

__global__ void __launch_bounds__(PC_THREADS) pair_count_kernel(PairParams p) {
    __shared__ uint32_t s_key[PC_CMAX];
    __shared__ uint32_t s_val[PC_CMAX];
    __shared__ uint32_t s_red[PC_WARPS];
    __shared__ unsigned long long s_red64[PC_WARPS];
    __shared__ unsigned long long s_ticket, s_base;
    __shared__ uint32_t s_distinct;
    __shared__ volatile uint32_t s_overflow;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (;;) {
        if (tid == 0) { s_ticket = atomicAdd(&p.sc->ticket, 1ull); s_distinct = 0; s_overflow = 0; }
        __syncthreads();
        const uint64_t t = s_ticket;
        if (t >= p.n_pivots) break;
        const uint32_t x = p.pivot_rows ? p.pivot_rows[t] : (uint32_t) t;
        const uint64_t a = p.row_off[x], b = p.row_off[x + 1];
        if (a == b) { __syncthreads(); continue; }

        // pass A: total list length W bounds the number of distinct partners
        unsigned long long w = 0;
        for (uint64_t j = a + tid; j < b; j += PC_THREADS) {
            const uint32_t slot = __ldg(&p.row_slot[j]);
            w += __ldg(&p.inv_off[slot + 1]) - __ldg(&p.inv_off[slot]);
        }
        const unsigned long long W = block_sum<PC_THREADS>(w, s_red64);
        uint32_t C = 64;
        while (C < PC_CMAX && C < 2 * W) C <<= 1;
        const uint32_t limit = (C / 4) * 3;   // never reached unless C was clamped to PC_CMAX (distinct <= W <= C/2)
        const uint32_t cmask = C - 1;
        int cbits = 31 - __clz(C);
        for (uint32_t i = tid; i < C; i += PC_THREADS) { s_key[i] = PC_EMPTY; s_val[i] = 0; }
        __syncthreads();

        // pass B: warp per incidence entry, lanes over the inverted list
        for (uint64_t j = a + warp; j < b && !s_overflow; j += PC_WARPS) {
            const uint32_t slot = __ldg(&p.row_slot[j]);
            const uint64_t lo = __ldg(&p.inv_off[slot]), hi = __ldg(&p.inv_off[slot + 1]);
            for (uint64_t i = lo + lane; i < hi; i += 32) {
                const uint32_t y = __ldg(&p.inv_row[i]);
                if (!keep_candidate(x, y, p.pivot_flag)) continue;
                uint32_t hsh = hash_row(y) >> (32 - cbits);
                for (;;) {
                    const uint32_t old = atomicCAS(&s_key[hsh], PC_EMPTY, y);
                    if (old == PC_EMPTY) {
                        if (atomicAdd(&s_distinct, 1u) + 1 > limit) s_overflow = 1;
                    }
                    if (old == PC_EMPTY || old == y) { atomicAdd(&s_val[hsh], 1u); break; }
                    if (s_overflow) break;
                    hsh = (hsh + 1) & cmask;
                }
            }
        }
        __syncthreads();
        if (s_overflow) {
            if (tid == 0) {
                const unsigned long long hi = atomicAdd(&p.sc->heavy_count, 1ull);
                p.heavy_list[hi] = x;
            }
            __syncthreads();
            continue;
        }

        // compaction: entries with score >= min_score
        uint32_t mine = 0;
        for (uint32_t i = tid; i < C; i += PC_THREADS) mine += (s_key[i] != PC_EMPTY && s_val[i] >= p.min_score);
        uint32_t total;
        uint32_t off = block_excl_scan<PC_THREADS>(mine, s_red, &total);
        if (tid == 0) {
            s_base = total ? atomicAdd(&p.sc->cursor, (unsigned long long) total) : 0ull;
            if (total && s_base + total > p.capacity) p.sc->overflow = 1;
        }
        __syncthreads();
        const uint64_t base = s_base;
        if (total && base + total <= p.capacity) {
            for (uint32_t i = tid; i < C; i += PC_THREADS) {
                const uint32_t y = s_key[i], v = s_val[i];
                if (y != PC_EMPTY && v >= p.min_score) {
                    const uint32_t lo = min(x, y), hi = max(x, y);
                    p.out_key[base + off] = ((uint64_t) lo << 32) | hi;
                    p.out_score[base + off] = v;
                    off++;
                }
            }
        }
        __syncthreads();
    }
}

// Rows whose partner set overflowed the shared-memory accumulator: same algorithm, table in HBM.
__global__ void __launch_bounds__(HV_THREADS) pair_count_heavy_kernel(PairParams p) {
    __shared__ uint32_t s_red[HV_THREADS / 32];
    __shared__ unsigned long long s_ticket, s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *tab_key = p.heavy_tab + (size_t) blockIdx.x * 2 * p.heavy_cap;
    uint32_t *tab_val = tab_key + p.heavy_cap;
    const uint32_t cmask = p.heavy_cap - 1;
    const int cbits = 31 - __clz(p.heavy_cap);
    const uint64_t n_heavy = p.sc->heavy_count;

    for (;;) {
        if (tid == 0) s_ticket = atomicAdd(&p.sc->heavy_ticket, 1ull);
        __syncthreads();
        const uint64_t t = s_ticket;
        if (t >= n_heavy) break;
        const uint32_t x = p.heavy_list[t];
        const uint64_t a = p.row_off[x], b = p.row_off[x + 1];
        for (uint32_t i = tid; i < p.heavy_cap; i += HV_THREADS) { tab_key[i] = PC_EMPTY; tab_val[i] = 0; }
        __syncthreads();
        for (uint64_t j = a + warp; j < b; j += HV_THREADS / 32) {
            const uint32_t slot = __ldg(&p.row_slot[j]);
            const uint64_t lo = __ldg(&p.inv_off[slot]), hi = __ldg(&p.inv_off[slot + 1]);
            for (uint64_t i = lo + lane; i < hi; i += 32) {
                const uint32_t y = __ldg(&p.inv_row[i]);
                if (!keep_candidate(x, y, p.pivot_flag)) continue;
                uint32_t hsh = hash_row(y) >> (32 - cbits);
                for (;;) {
                    const uint32_t old = atomicCAS(&tab_key[hsh], PC_EMPTY, y);
                    if (old == PC_EMPTY || old == y) { atomicAdd(&tab_val[hsh], 1u); break; }
                    hsh = (hsh + 1) & cmask;
                }
            }
        }
        __syncthreads();
        uint32_t mine = 0;
        for (uint32_t i = tid; i < p.heavy_cap; i += HV_THREADS) mine += (tab_key[i] != PC_EMPTY && tab_val[i] >= p.min_score);
        uint32_t total;
        uint32_t off = block_excl_scan<HV_THREADS>(mine, s_red, &total);
        if (tid == 0) {
            s_base = total ? atomicAdd(&p.sc->cursor, (unsigned long long) total) : 0ull;
            if (total && s_base + total > p.capacity) p.sc->overflow = 1;
        }
        __syncthreads();
        const uint64_t base = s_base;
        if (total && base + total <= p.capacity) {
            for (uint32_t i = tid; i < p.heavy_cap; i += HV_THREADS) {
                const uint32_t y = tab_key[i], v = tab_val[i];
                if (y != PC_EMPTY && v >= p.min_score) {
                    const uint32_t lo = min(x, y), hi = max(x, y);
                    p.out_key[base + off] = ((uint64_t) lo << 32) | hi;
                    p.out_score[base + off] = v;
                    off++;
                }
            }
        }
        __syncthreads();
    }
}

// work measure: sum over k-mers of occ*(occ-1)/2
__global__ void increments_kernel(const uint64_t *__restrict__ inv_off, uint32_t n_slots, unsigned long long *out) {
    unsigned long long acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t) gridDim.x * blockDim.x) {
        const unsigned long long len = inv_off[i + 1] - inv_off[i];
        acc += len * (len - (len ? 1 : 0)) / 2;
    }
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

__global__ void mark_pivots_kernel(const uint32_t *__restrict__ pivot_rows, uint64_t n, uint8_t *flag) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) flag[pivot_rows[i]] = 1;
}

}  // namespace

int hga_pairs_run(hga_handle *h, uint32_t min_score, const uint32_t *pivots, uint64_t n_pivots) {
    if (!h->have_index) { hga_set_error("hga_pair_count: no index (call hga_build_index)"); return HGA_E_STATE; }
    h->have_pairs = h->have_selection = h->have_components = false;
    const bool multi = h->comm && hga_comm_size(h) > 1;
    const uint64_t n_rows = h->inc_rows;
    h->pair_min_score = min_score;

    HGA_TRY(h->d_pair_scalars.ensure(sizeof(PairScalars)));
    PairScalars *d_sc = h->d_pair_scalars.as<PairScalars>();
    HGA_TRY(h->d_heavy_list.ensure((n_rows + 1) * 4));

    PairParams p;
    p.row_off = h->inc_row_off; p.row_slot = h->inc_slot; p.n_rows = n_rows;
    p.inv_off = h->d_inv_off.as<uint64_t>(); p.inv_row = h->d_inv_row.as<uint32_t>();
    p.pivot_rows = nullptr; p.pivot_flag = nullptr; p.n_pivots = n_rows;
    p.min_score = multi ? 1u : min_score;     // partial scores are thresholded after the cross-rank reduction
    p.heavy_list = h->d_heavy_list.as<uint32_t>();
    p.heavy_tab = nullptr; p.heavy_cap = 0;
    p.sc = d_sc;

    DevBuf d_pivots;
    if (pivots) {
        // caller passes read ids; rows are id - first id
        std::vector<uint32_t> rows(n_pivots);
        for (uint64_t i = 0; i < n_pivots; i++) {
            const uint64_t r = (uint64_t) pivots[i] - h->inc_row_first_id;
            if (pivots[i] < h->inc_row_first_id || r >= n_rows) { hga_set_error("pivot read id %u out of range", pivots[i]); return HGA_E_ARG; }
            rows[i] = (uint32_t) r;
        }
        HGA_TRY(d_pivots.ensure((n_pivots + 1) * 4));
        HGA_TRY(h->d_pivot_flag.ensure(n_rows + 1));
        HGA_CUDA(cudaMemcpyAsync(d_pivots.p, rows.data(), n_pivots * 4, cudaMemcpyHostToDevice, h->stream));
        HGA_CUDA(cudaMemsetAsync(h->d_pivot_flag.p, 0, n_rows + 1, h->stream));
        if (n_pivots) {
            mark_pivots_kernel<<<(int) std::min<uint64_t>((n_pivots + 255) / 256, 1024), 256, 0, h->stream>>>(d_pivots.as<uint32_t>(), n_pivots,
                                                                                                            h->d_pivot_flag.as<uint8_t>());
            h->metrics.kernel_launches++;
        }
        HGA_CUDA(cudaStreamSynchronize(h->stream));   // rows vector goes out of scope below
        p.pivot_rows = d_pivots.as<uint32_t>();
        p.pivot_flag = h->d_pivot_flag.as<uint8_t>();
        p.n_pivots = n_pivots;
    }

    StageTimer timer(h, &h->metrics.pair_ms);
    uint64_t capacity = std::max<uint64_t>(h->pair_capacity, std::max<uint64_t>(64 * n_rows, 1ull << 20));
    int occ = 0;
    HGA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pair_count_kernel, PC_THREADS, 0));
    if (occ < 1) occ = 1;
    const int grid = (int) std::max<uint64_t>(1, std::min<uint64_t>((uint64_t) h->sm_count * occ, p.n_pivots));

    PairScalars sc;
    h->metrics.pair_retries = 0;
    for (int attempt = 0;; attempt++) {
        HGA_TRY(h->d_pair_key.ensure((capacity + 1) * 8));
        HGA_TRY(h->d_pair_score.ensure((capacity + 1) * 4));
        p.out_key = h->d_pair_key.as<uint64_t>(); p.out_score = h->d_pair_score.as<uint32_t>(); p.capacity = capacity;
        HGA_CUDA(cudaMemsetAsync(d_sc, 0, sizeof(PairScalars), h->stream));
        if (p.n_pivots) {
            pair_count_kernel<<<grid, PC_THREADS, 0, h->stream>>>(p);
            h->metrics.kernel_launches++;
            HGA_CUDA(cudaGetLastError());
        }
        HGA_CUDA(cudaMemcpyAsync(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        if (sc.heavy_count) {
            uint32_t cap = 1024;
            while (cap < 2 * n_rows && cap < (1u << 30)) cap <<= 1;
            uint64_t budget = 4ull << 30;
            int hgrid = (int) std::min<uint64_t>(sc.heavy_count, (uint64_t) h->sm_count);
            while (hgrid > 1 && (uint64_t) hgrid * cap * 8 > budget) hgrid--;
            HGA_TRY(h->d_heavy_tab.ensure((size_t) hgrid * cap * 8));
            p.heavy_tab = h->d_heavy_tab.as<uint32_t>(); p.heavy_cap = cap;
            pair_count_heavy_kernel<<<hgrid, HV_THREADS, 0, h->stream>>>(p);
            h->metrics.kernel_launches++;
            HGA_CUDA(cudaGetLastError());
            HGA_CUDA(cudaMemcpyAsync(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
        }
        h->metrics.heavy_pivots = sc.heavy_count;
        if (!sc.overflow) break;
        if (attempt >= 1) { hga_set_error("pair_count: output overflow after exact resize (internal error)"); return HGA_E_OVERFLOW; }
        capacity = sc.cursor;
        h->metrics.pair_retries++;
    }
    h->pair_capacity = capacity;
    uint64_t P = sc.cursor;

    // canonical physical order: sort by (x_row, y_row)
    HGA_TRY(h->d_pair_key2.ensure((P + 1) * 8));
    HGA_TRY(h->d_pair_score2.ensure((P + 1) * 4));
    if (P > 0) {
        const int row_bits = (int) std::max<uint32_t>(hga_ceil_log2(n_rows + 1), 1);
        size_t tmp_bytes = 0;
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->d_pair_key.as<uint64_t>(), h->d_pair_key2.as<uint64_t>(),
                                                 h->d_pair_score.as<uint32_t>(), h->d_pair_score2.as<uint32_t>(), P, 0, 32 + row_bits, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp_bytes + 16));
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tmp_bytes, h->d_pair_key.as<uint64_t>(), h->d_pair_key2.as<uint64_t>(),
                                                 h->d_pair_score.as<uint32_t>(), h->d_pair_score2.as<uint32_t>(), P, 0, 32 + row_bits, h->stream));
        h->metrics.kernel_launches += (uint64_t) (32 + row_bits + 7) / 8 + 2;
    }
    std::swap(h->d_pair_key, h->d_pair_key2);
    std::swap(h->d_pair_score, h->d_pair_score2);
    h->n_pairs = P;

    {   // work measure
        HGA_CUDA(cudaMemsetAsync(&d_sc->increments, 0, 8, h->stream));
        increments_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(h->d_inv_off.as<uint64_t>(), h->table.n_slots, &d_sc->increments);
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaMemcpyAsync(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        h->n_increments = sc.increments;
    }
    if (multi) HGA_TRY(hga_comm_reduce_pairs(h));   // all-to-all of partial scores + reduce by key + min_score filter
    timer.stop();
    h->metrics.n_pairs = h->n_pairs; h->metrics.n_increments = h->n_increments;
    h->have_pairs = true;
    return HGA_OK;
}
