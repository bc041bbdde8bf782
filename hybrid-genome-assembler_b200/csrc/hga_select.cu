// select_edges: the scaffold-forming slice of the sorted connection list.
//
// Replaces clustering/ReadClusteringEngine.cpp:748-756: the reference sorts the DIRECTED list (every pair
// twice) by score descending and keeps the first n = (size_t)(size * fraction) entries (or score > S with
// --sc_score). The order inside a score is unspecified there; here it is the canonical total order
// (score desc, x asc, y asc), so the kept set is: every pair with score > s*, plus the first ceil(r/2) pairs
// with score == s* in (x, y) order, where s* is the score of the n-th directed entry and r = n - 2*|score > s*|.
//
// s* comes from a two-level (16 + 16 bit) radix select over the locally sorted scores; the two histograms are
// the only data that crosses ranks (all-reduce sum) in the multi-GPU case.
#include "hga_internal.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>

#define SEL_CHUNK 2048
#define SEL_THREADS 256

namespace {

// hist[b] = number of sorted scores s with (s >> shift) == b (and, for the low level, s >> 16 == hi_bin)
__global__ void hist_from_sorted_kernel(const uint32_t *__restrict__ sorted, uint64_t n, int level, uint32_t hi_bin, unsigned long long *hist) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= 65536) return;
    uint32_t lo_v, hi_v;   // scores in [lo_v, hi_v]
    if (level == 0) { lo_v = b << 16; hi_v = lo_v | 0xFFFFu; }
    else { lo_v = (hi_bin << 16) | b; hi_v = lo_v; }
    uint64_t lo = 0, hi = n;          // lower_bound(lo_v)
    while (lo < hi) { uint64_t m = (lo + hi) >> 1; if (sorted[m] < lo_v) lo = m + 1; else hi = m; }
    const uint64_t first = lo;
    hi = n;                            // upper_bound(hi_v)
    while (lo < hi) { uint64_t m = (lo + hi) >> 1; if (sorted[m] <= hi_v) lo = m + 1; else hi = m; }
    hist[b] = lo - first;
}

__global__ void count_chunks_kernel(const uint32_t *__restrict__ score, uint64_t n, uint32_t cut, unsigned long long *blk_ties,
                                    unsigned long long *blk_above) {
    __shared__ uint32_t s_t[SEL_THREADS / 32], s_a[SEL_THREADS / 32];
    const uint64_t base = (uint64_t) blockIdx.x * SEL_CHUNK;
    uint32_t t = 0, a = 0;
    for (uint32_t i = threadIdx.x; i < SEL_CHUNK; i += SEL_THREADS) {
        const uint64_t g = base + i;
        if (g < n) { const uint32_t s = score[g]; t += (s == cut); a += (s > cut); }
    }
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) { t += __shfl_xor_sync(0xFFFFFFFFu, t, d); a += __shfl_xor_sync(0xFFFFFFFFu, a, d); }
    if ((threadIdx.x & 31) == 0) { s_t[threadIdx.x >> 5] = t; s_a[threadIdx.x >> 5] = a; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tt = 0, aa = 0;
        for (int i = 0; i < SEL_THREADS / 32; i++) { tt += s_t[i]; aa += s_a[i]; }
        blk_ties[blockIdx.x] = tt; blk_above[blockIdx.x] = aa;
    }
}

// blk_sel[b] = above_b + number of this chunk's ties that fall inside the quota
__global__ void chunk_selected_kernel(const unsigned long long *__restrict__ blk_ties, const unsigned long long *__restrict__ tie_base,
                                      const unsigned long long *__restrict__ blk_above, uint64_t nb, unsigned long long tie_offset,
                                      unsigned long long quota, unsigned long long *blk_sel) {
    const uint64_t b = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const unsigned long long before = tie_offset + tie_base[b];
    const unsigned long long room = quota > before ? quota - before : 0;
    blk_sel[b] = blk_above[b] + min(blk_ties[b], room);
}

__global__ void write_selected_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ score, uint64_t n, uint32_t cut,
                                      const unsigned long long *__restrict__ tie_base, const unsigned long long *__restrict__ sel_base,
                                      unsigned long long tie_offset, unsigned long long quota, uint64_t *out_key, uint32_t *out_score) {
    __shared__ uint32_t s_t[SEL_THREADS / 32], s_s[SEL_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t base = (uint64_t) blockIdx.x * SEL_CHUNK;
    unsigned long long ties_before = tie_offset + tie_base[blockIdx.x];
    unsigned long long out = sel_base[blockIdx.x];
    // contiguous sub-chunks of SEL_THREADS keep the (x, y) order
    for (uint32_t c = 0; c < SEL_CHUNK; c += SEL_THREADS) {
        const uint64_t g = base + c + threadIdx.x;
        uint32_t s = 0; uint64_t kk = 0;
        bool tie = false, above = false;
        if (g < n) { s = score[g]; kk = key[g]; tie = (s == cut); above = (s > cut); }
        // exclusive rank among ties in this sub-chunk
        const uint32_t tb = __ballot_sync(0xFFFFFFFFu, tie);
        if (lane == 0) s_t[warp] = __popc(tb);
        __syncthreads();
        uint32_t tprefix = __popc(tb & ((1u << lane) - 1)), ttotal = 0;
        for (int i = 0; i < SEL_THREADS / 32; i++) { if (i < warp) tprefix += s_t[i]; ttotal += s_t[i]; }
        const bool sel = above || (tie && ties_before + tprefix < quota);
        const uint32_t sb = __ballot_sync(0xFFFFFFFFu, sel);
        if (lane == 0) s_s[warp] = __popc(sb);
        __syncthreads();
        uint32_t sprefix = __popc(sb & ((1u << lane) - 1)), stotal = 0;
        for (int i = 0; i < SEL_THREADS / 32; i++) { if (i < warp) sprefix += s_s[i]; stotal += s_s[i]; }
        if (sel) { out_key[out + sprefix] = kk; out_score[out + sprefix] = s; }
        ties_before += ttotal; out += stotal;
        __syncthreads();
    }
}

// keys of the pairs whose score equals the cut, in (x, y) order (tie_base = exclusive scan of the chunk tie counts)
__global__ void write_ties_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ score, uint64_t n, uint32_t cut,
                                  const unsigned long long *__restrict__ tie_base, uint64_t *out_key) {
    __shared__ uint32_t s_t[SEL_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t base = (uint64_t) blockIdx.x * SEL_CHUNK;
    unsigned long long out = tie_base[blockIdx.x];
    for (uint32_t c = 0; c < SEL_CHUNK; c += SEL_THREADS) {
        const uint64_t g = base + c + threadIdx.x;
        const bool tie = g < n && score[g] == cut;
        const uint32_t tb = __ballot_sync(0xFFFFFFFFu, tie);
        if (lane == 0) s_t[warp] = __popc(tb);
        __syncthreads();
        uint32_t prefix = __popc(tb & ((1u << lane) - 1)), total = 0;
        for (int i = 0; i < SEL_THREADS / 32; i++) { if (i < warp) prefix += s_t[i]; total += s_t[i]; }
        if (tie) out_key[out + prefix] = key[g];
        out += total;
        __syncthreads();
    }
}

// number of entries of the ascending array that are <= limit
__global__ void count_le_kernel(const uint64_t *__restrict__ sorted, uint64_t n, const uint64_t *limit, unsigned long long *out) {
    const uint64_t lim = *limit;
    uint64_t lo = 0, hi = n;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (sorted[mid] <= lim) lo = mid + 1; else hi = mid; }
    *out = lo;
}

}  // namespace

int hga_select_run(hga_handle *h, double fraction, uint32_t score_threshold) {
    if (!h->have_pairs) { hga_set_error("hga_select_edges: no pairs (call hga_pair_count)"); return HGA_E_STATE; }
    h->have_selection = h->have_components = false;
    const bool multi = h->comm && hga_comm_size(h) > 1;
    const uint64_t P = h->n_pairs;
    StageTimer timer(h, &h->metrics.select_ms);

    HGA_TRY(h->d_hist.ensure(65536 * 8 * 2 + 64));
    unsigned long long *d_hist = h->d_hist.as<unsigned long long>();
    std::vector<unsigned long long> hist(65536);

    uint64_t P_total = P;
    if (multi) {
        HGA_CUDA(cudaMemcpyAsync(d_hist, &P, 8, cudaMemcpyHostToDevice, h->stream));
        HGA_TRY(hga_comm_allreduce_u64_sum(h, (uint64_t *) d_hist, 1));
        HGA_CUDA(cudaMemcpyAsync(&P_total, d_hist, 8, cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
    }

    uint64_t n_directed = 0, cut = 0, quota = 0;
    if (score_threshold > 0) {
        cut = score_threshold;      // keep score > S  (.cpp:752)
        quota = 0;
    } else {
        n_directed = (uint64_t) ((double) (2 * P_total) * fraction);    // .cpp:755
        if (n_directed > 2 * P_total) n_directed = 2 * P_total;
        if (n_directed > 0) {
            // locally sorted copy of the scores
            HGA_TRY(h->d_sort_a.ensure((P + 1) * 4));
            uint32_t *d_sorted = h->d_sort_a.as<uint32_t>();
            if (P > 0) {
                // scores are small (a few thousand at most at config 4): sort only the bits the largest one needs
                size_t tmp_bytes = 0, red_bytes = 0;
                uint32_t *d_max = reinterpret_cast<uint32_t *>(d_hist);
                HGA_CUDA(cub::DeviceReduce::Max(nullptr, red_bytes, h->d_pair_score.as<uint32_t>(), d_max, P, h->stream));
                HGA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, h->d_pair_score.as<uint32_t>(), d_sorted, P, 0, 32, h->stream));
                HGA_TRY(h->d_sort_tmp.ensure(std::max(tmp_bytes, red_bytes) + 16));
                HGA_CUDA(cub::DeviceReduce::Max(h->d_sort_tmp.p, red_bytes, h->d_pair_score.as<uint32_t>(), d_max, P, h->stream));
                uint32_t max_score = 0;
                HGA_CUDA(cudaMemcpyAsync(&max_score, d_max, 4, cudaMemcpyDeviceToHost, h->stream));
                HGA_CUDA(cudaStreamSynchronize(h->stream));
                const int end_bit = (int) std::max<uint32_t>(hga_ceil_log2((uint64_t) max_score + 1), 1);
                HGA_CUDA(cub::DeviceRadixSort::SortKeys(h->d_sort_tmp.p, tmp_bytes, h->d_pair_score.as<uint32_t>(), d_sorted, P, 0, end_bit, h->stream));
                h->metrics.kernel_launches += (uint64_t) (end_bit + 7) / 8 + 4;
            }
            const uint64_t rank = (n_directed + 1) / 2;   // the n-th directed entry belongs to the rank-th pair (descending)
            uint32_t hi_bin = 0;
            uint64_t above = 0;                            // pairs with score strictly above the current bin
            for (int level = 0; level < 2; level++) {
                hist_from_sorted_kernel<<<65536 / 256, 256, 0, h->stream>>>(d_sorted, P, level, hi_bin, d_hist);
                h->metrics.kernel_launches++;
                HGA_CUDA(cudaGetLastError());
                if (multi) HGA_TRY(hga_comm_allreduce_u64_sum(h, (uint64_t *) d_hist, 65536));
                HGA_CUDA(cudaMemcpyAsync(hist.data(), d_hist, 65536 * 8, cudaMemcpyDeviceToHost, h->stream));
                HGA_CUDA(cudaStreamSynchronize(h->stream));
                int b = 65535;
                for (; b >= 0; b--) {
                    if (above + hist[b] >= rank) break;
                    above += hist[b];
                }
                if (b < 0) { hga_set_error("select: rank beyond the pair count (internal error)"); return HGA_E_STATE; }
                if (level == 0) hi_bin = (uint32_t) b; else cut = ((uint64_t) hi_bin << 16) | (uint32_t) b;
            }
            const uint64_t r = n_directed - 2 * above;     // directed slots left for the tie group
            quota = (r + 1) / 2;
        }
    }

    uint64_t n_sel = 0;
    // with a communicator every rank goes through this block, whatever its own pair count: the tie gather below is collective
    // (a rank that holds no pair contributes no tie; skipping it would leave the others waiting in NCCL)
    if ((P > 0 || multi) && (score_threshold > 0 || n_directed > 0)) {
        const uint64_t nb = (P + SEL_CHUNK - 1) / SEL_CHUNK;
        HGA_TRY(h->d_sel_scalars.ensure((nb + 1) * 8 * 5 + 64));
        unsigned long long *blk_ties = h->d_sel_scalars.as<unsigned long long>();
        unsigned long long *blk_above = blk_ties + (nb + 1), *tie_base = blk_above + (nb + 1), *blk_sel = tie_base + (nb + 1), *sel_base = blk_sel + (nb + 1);
        if (nb) count_chunks_kernel<<<(unsigned) nb, SEL_THREADS, 0, h->stream>>>(h->d_pair_score.as<uint32_t>(), P, (uint32_t) cut, blk_ties, blk_above);
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
        size_t tmp_bytes = 0;
        HGA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, blk_ties, tie_base, nb + 1, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp_bytes + 16));
        HGA_CUDA(cudaMemsetAsync(blk_ties + nb, 0, 8, h->stream));
        HGA_CUDA(cub::DeviceScan::ExclusiveSum(h->d_sort_tmp.p, tmp_bytes, blk_ties, tie_base, nb + 1, h->stream));
        unsigned long long tie_offset = 0;
        if (multi && quota > 0) {
            // The canonical order interleaves the ranks' pairs, so the global quota of tied pairs is turned into a local
            // one: gather every rank's tie keys, find the quota-th smallest, count the local ties up to it.
            unsigned long long my_ties = 0;
            HGA_CUDA(cudaMemcpyAsync(&my_ties, tie_base + nb, 8, cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            std::vector<uint64_t> counts;
            HGA_TRY(hga_comm_allgather_u64(h, my_ties, counts));
            uint64_t T = 0;
            for (uint64_t c : counts) T += c;
            if (quota > T) { hga_set_error("select: tie quota beyond the number of ties (internal error)"); return HGA_E_STATE; }
            HGA_TRY(h->d_sel_key.ensure((my_ties + 1) * 8));
            HGA_TRY(h->d_export_a.ensure((T + 1) * 8 * 2 + 64));
            uint64_t *d_mine = h->d_sel_key.as<uint64_t>(), *d_all = h->d_export_a.as<uint64_t>(), *d_all_sorted = d_all + (T + 1);
            if (nb) write_ties_kernel<<<(unsigned) nb, SEL_THREADS, 0, h->stream>>>(h->d_pair_key.as<uint64_t>(), h->d_pair_score.as<uint32_t>(), P, (uint32_t) cut, tie_base, d_mine);
            HGA_CUDA(cudaGetLastError());
            HGA_TRY(hga_comm_allgatherv(h, d_mine, d_all, counts, 8));
            size_t tb2 = 0;
            HGA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb2, d_all, d_all_sorted, T, 0, 64, h->stream));
            HGA_TRY(h->d_sort_tmp.ensure(std::max(tb2, tmp_bytes) + 16));
            HGA_CUDA(cub::DeviceRadixSort::SortKeys(h->d_sort_tmp.p, tb2, d_all, d_all_sorted, T, 0, 64, h->stream));
            count_le_kernel<<<1, 1, 0, h->stream>>>(d_mine, my_ties, d_all_sorted + (quota - 1), d_hist);
            h->metrics.kernel_launches += 12;
            unsigned long long local_quota = 0;
            HGA_CUDA(cudaMemcpyAsync(&local_quota, d_hist, 8, cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            quota = local_quota;
        }
        if (nb) chunk_selected_kernel<<<(unsigned) ((nb + 255) / 256), 256, 0, h->stream>>>(blk_ties, tie_base, blk_above, nb, tie_offset, quota, blk_sel);
        HGA_CUDA(cudaMemsetAsync(blk_sel + nb, 0, 8, h->stream));
        HGA_CUDA(cub::DeviceScan::ExclusiveSum(h->d_sort_tmp.p, tmp_bytes, blk_sel, sel_base, nb + 1, h->stream));
        h->metrics.kernel_launches += 3;
        unsigned long long total_sel = 0;
        HGA_CUDA(cudaMemcpyAsync(&total_sel, sel_base + nb, 8, cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        n_sel = total_sel;
        HGA_TRY(h->d_sel_key.ensure((n_sel + 1) * 8));
        HGA_TRY(h->d_sel_score.ensure((n_sel + 1) * 4));
        if (nb) write_selected_kernel<<<(unsigned) nb, SEL_THREADS, 0, h->stream>>>(h->d_pair_key.as<uint64_t>(), h->d_pair_score.as<uint32_t>(), P, (uint32_t) cut,
                                                                          tie_base, sel_base, tie_offset, quota, h->d_sel_key.as<uint64_t>(),
                                                                          h->d_sel_score.as<uint32_t>());
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
    } else {
        HGA_TRY(h->d_sel_key.ensure(8));
        HGA_TRY(h->d_sel_score.ensure(4));
    }
    timer.stop();
    h->n_selected = n_sel;
    h->sel_cut = (score_threshold > 0 || n_directed > 0) ? cut : 0;
    h->sel_n_directed = n_directed;
    if (score_threshold > 0) {
        uint64_t tot = n_sel;
        if (multi) {
            HGA_CUDA(cudaMemcpyAsync(d_hist, &tot, 8, cudaMemcpyHostToDevice, h->stream));
            HGA_TRY(hga_comm_allreduce_u64_sum(h, (uint64_t *) d_hist, 1));
            HGA_CUDA(cudaMemcpyAsync(&tot, d_hist, 8, cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
        }
        h->sel_n_directed = 2 * tot;
    }
    h->metrics.n_selected = n_sel;
    h->have_selection = true;
    return HGA_OK;
}
