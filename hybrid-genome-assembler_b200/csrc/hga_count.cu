// count: exact canonical k-mer counts of a set of reads (the GPU half of the SDK selection).
//
// Replaces what occurrences/run_jellyfish.sh asks jellyfish for, per read file (occurrences/JellyfishOccurrenceReader.cpp:16-38):
//   jellyfish bc -C -m k ...; jellyfish count -C -m k --bc ...; jellyfish dump -c; LC_ALL=C sort
// i.e. the canonical k-mers (-C: the smaller of a k-mer and its reverse complement; with A < C < G < T that is the smaller 2-bit
// value, the same canonical form as KmerIterator.cpp:69) that occur at least twice (the two-pass Bloom-counter filter, here without
// its false positives), with their exact counts, in ascending order (string order = value order for equal-length ACGT strings).
// jellyfish itself is not in this image and is not vendored by the reference: the counting is pinned with the reference's own KmerIterator +
// std::map on ACGT-only reads (oracle/_ref/occ_driver count) and checked against exact counting in numpy elsewhere.
// Windows containing a byte other than A C G T a c g t are skipped, as jellyfish does (NOT the code-0 rule of KmerIterator).
//
// Small inputs (at most one budget of windows, 2^29): count_emit_kernel (one thread per window start: binary search of the read, k byte
// loads, canonical value or a sentinel), CUB radix sort + run-length encode, count_flag_kernel + CUB select for count >= min_count.
// Large inputs run in KEY-RANGE passes, so that the scratch is one budget of keys whatever the input size: a strided sample of the
// windows gives the quantiles of the key distribution; pass r emits only the k-mers of range [q_r, q_r+1) (count_emit_range_kernel: the
// same per-window code, matches compacted with one atomic per warp), sorts, run-length encodes, filters and appends its (k-mer, count)
// runs to the host result; ranges ascend, so the result is sorted without a merge and every k-mer is counted in exactly one pass. A
// range that overflows the budget is split at its midpoint and redone. HBM-bound by the sort passes of the emitted keys plus one read
// of the bases per pass.
#include "hga_internal.cuh"

#include <algorithm>
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_select.cuh>

namespace {

constexpr unsigned long long COUNT_SENTINEL = ~0ull;      // no canonical k-mer has this value (the reverse complement of T...T is 0)

__device__ __forceinline__ int count_base_code(unsigned char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

// canonical k-mer of the window starting at position p, or COUNT_SENTINEL (window crosses a read end / holds a byte other than ACGTacgt)
__device__ __forceinline__ unsigned long long count_window_key(const char *__restrict__ bases, const uint64_t *__restrict__ read_off, uint64_t n_reads, int k, uint64_t p) {
    // the read that holds position p: the last r with read_off[r] <= p
    uint64_t lo = 0, hi = n_reads;
    while (lo < hi) { const uint64_t mid = (lo + hi + 1) >> 1; if (read_off[mid] <= p) lo = mid; else hi = mid - 1; }
    if (p + (uint64_t) k > read_off[lo + 1]) return COUNT_SENTINEL;
    unsigned long long fwd = 0, rc = 0;
    for (int j = 0; j < k; j++) {
        const int c = count_base_code((unsigned char) bases[p + j]);
        if (c < 0) return COUNT_SENTINEL;
        fwd = (fwd << 2) | (unsigned long long) c;
        rc = (rc >> 2) | ((unsigned long long) (3 - c) << (2 * (k - 1)));
    }
    if (k < 32) fwd &= (1ull << (2 * k)) - 1;
    return fwd < rc ? fwd : rc;
}

__global__ void count_emit_kernel(const char *__restrict__ bases, const uint64_t *__restrict__ read_off, uint64_t n_reads, int k, uint64_t p0, uint64_t p1,
                                  unsigned long long *__restrict__ out) {
    for (uint64_t p = p0 + blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; p < p1; p += (uint64_t) gridDim.x * blockDim.x)
        out[p - p0] = count_window_key(bases, read_off, n_reads, k, p);
}

// Key-range pass: the canonical k-mers in [key_lo, key_hi) of the windows of every span_stride-th span of COUNT_SPAN positions, compacted into out
// (unordered). cursor counts every match, stored or not (capacity overflow is detected from it); one atomic per warp on the device.
#define COUNT_SPAN 8
__global__ void count_emit_range_kernel(const char *__restrict__ bases, const uint64_t *__restrict__ read_off, uint64_t n_reads, int k, uint64_t n_bases,
                                        unsigned long long key_lo, unsigned long long key_hi, uint64_t span_stride, unsigned long long *__restrict__ out,
                                        uint64_t capacity, unsigned long long *cursor) {
    const uint64_t n_spans = (n_bases + COUNT_SPAN - 1) / COUNT_SPAN, n_sel = (n_spans + span_stride - 1) / span_stride;
    const uint64_t threads = (uint64_t) gridDim.x * blockDim.x, rounds = (n_sel + threads - 1) / threads;
    for (uint64_t round = 0; round < rounds; round++) {                      // the same trip count for every thread: the warp stays together
        const uint64_t t = round * threads + blockIdx.x * (uint64_t) blockDim.x + threadIdx.x;
        unsigned long long keys[COUNT_SPAN];
        uint32_t cnt = 0;
        if (t < n_sel) {
            const uint64_t p0 = t * span_stride * COUNT_SPAN;
            #pragma unroll
            for (int j = 0; j < COUNT_SPAN; j++) {
                unsigned long long key = COUNT_SENTINEL;
                if (p0 + j < n_bases) key = count_window_key(bases, read_off, n_reads, k, p0 + j);
                if (key != COUNT_SENTINEL && key >= key_lo && key < key_hi) keys[cnt++] = key;
            }
        }
        unsigned long long at;
#ifdef __CUDA_ARCH__
        const int lane = threadIdx.x & 31;
        uint32_t incl = cnt;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
        unsigned long long base = 0;
        if (lane == 31 && incl) base = atomicAdd(cursor, (unsigned long long) incl);
        base = __shfl_sync(0xFFFFFFFFu, base, 31);
        at = base + incl - cnt;
#else
        at = atomicAdd(cursor, (unsigned long long) cnt);
#endif
        for (uint32_t j = 0; j < cnt; j++) if (at + j < capacity) out[at + j] = keys[j];
    }
}

__global__ void count_flag_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ cnt, uint64_t n, uint32_t min_count, uint8_t *__restrict__ flag) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
        flag[i] = keys[i] != COUNT_SENTINEL && cnt[i] >= min_count;
}

struct Bufs {
    DevBuf bases, off, keys, sorted, run_key, run_len, acc_key2, acc_cnt2, tmp, scalars, flag;
    ~Bufs() { for (DevBuf *b : {&bases, &off, &keys, &sorted, &run_key, &run_len, &acc_key2, &acc_cnt2, &tmp, &scalars, &flag}) b->release(); }
};

}  // namespace

extern "C" int hga_count_kmers(int device, int k, const char *bases, const uint64_t *read_off, uint64_t n_reads, uint32_t min_count, hga_kmer_counts_t *out) {
    if (!out) { hga_set_error("hga_count_kmers: NULL argument"); return HGA_E_ARG; }
    out->n = 0; out->kmer = nullptr; out->count = nullptr;
    if (k < 1 || k > 32) { hga_set_error("Kmer size is too big or too small (k = %d)", k); return HGA_E_ARG; }
    if (n_reads && (!bases || !read_off)) { hga_set_error("hga_count_kmers: NULL argument"); return HGA_E_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); hga_set_error("hga_count_kmers: no usable CUDA device (this library has no CPU path)"); return HGA_E_CUDA; }
    HGA_CUDA(cudaSetDevice(device));
    if (n_reads == 0) return HGA_OK;
    const uint64_t base0 = read_off[0], n_bases = read_off[n_reads] - base0;
    if (n_bases == 0) return HGA_OK;
    cudaStream_t st = nullptr;      // legacy default stream: every call below is ordered on it
    Bufs b;
    HGA_TRY(b.bases.ensure(n_bases + 64));
    HGA_TRY(b.off.ensure((n_reads + 1) * 8));
    HGA_TRY(b.scalars.ensure(64));
    std::vector<uint64_t> off(n_reads + 1);
    for (uint64_t i = 0; i <= n_reads; i++) off[i] = read_off[i] - base0;
    HGA_CUDA(cudaMemcpyAsync(b.bases.p, bases + base0, n_bases, cudaMemcpyHostToDevice, st));
    HGA_CUDA(cudaMemcpyAsync(b.off.p, off.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    unsigned long long *d_scal = b.scalars.as<unsigned long long>();
    // keys per pass (the scratch is ~6 arrays of this many entries); HGA_COUNT_CHUNK overrides it (tests force the multi-pass path on small inputs with it)
    const char *ch_env = getenv("HGA_COUNT_CHUNK");
    const uint64_t B = ch_env && std::strtoull(ch_env, nullptr, 10) >= 256 ? std::strtoull(ch_env, nullptr, 10) : (1ull << 29);
    // 2k key bits are enough: no canonical k-mer has all of them set (T...T is not canonical), so the sentinels still sort behind
    // every k-mer and stay together
    const int end_bit = 2 * k;
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
    auto grid = [&](uint64_t n) { return (int) std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t) sm * 16)); };
    std::vector<uint64_t> res_k;
    std::vector<uint32_t> res_c;

    // m unsorted keys in b.keys -> sorted runs -> (k-mer, count >= min_count, sentinel dropped) appended to the host result
    auto finish_pass = [&](uint64_t m) -> int {
        if (m == 0) return HGA_OK;
        HGA_TRY(b.sorted.ensure(m * 8)); HGA_TRY(b.run_key.ensure(m * 8)); HGA_TRY(b.run_len.ensure(m * 4));
        unsigned long long *d_keys = b.keys.as<unsigned long long>(), *d_sorted = b.sorted.as<unsigned long long>(), *d_rk = b.run_key.as<unsigned long long>();
        uint32_t *d_rl = b.run_len.as<uint32_t>();
        size_t t1 = 0, t2 = 0;
        HGA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, t1, d_keys, d_sorted, m, 0, end_bit, st));
        HGA_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, t2, d_sorted, d_rk, d_rl, d_scal, m, st));
        HGA_TRY(b.tmp.ensure(std::max(t1, t2) + 16));
        HGA_CUDA(cub::DeviceRadixSort::SortKeys(b.tmp.p, t1, d_keys, d_sorted, m, 0, end_bit, st));
        HGA_CUDA(cub::DeviceRunLengthEncode::Encode(b.tmp.p, t2, d_sorted, d_rk, d_rl, d_scal, m, st));
        HGA_CUDA(cudaGetLastError());
        unsigned long long runs = 0;
        HGA_CUDA(cudaMemcpyAsync(&runs, d_scal, 8, cudaMemcpyDeviceToHost, st));
        HGA_CUDA(cudaStreamSynchronize(st));
        if (runs == 0) return HGA_OK;
        HGA_TRY(b.flag.ensure(runs + 16));
        HGA_TRY(b.acc_key2.ensure(runs * 8)); HGA_TRY(b.acc_cnt2.ensure(runs * 4));
        uint8_t *d_flag = b.flag.as<uint8_t>();
        count_flag_kernel<<<grid(runs), 256, 0, st>>>(d_rk, d_rl, runs, min_count, d_flag);
        HGA_CUDA(cub::DeviceSelect::Flagged(nullptr, t1, d_rk, d_flag, b.acc_key2.as<unsigned long long>(), d_scal, runs, st));
        HGA_CUDA(cub::DeviceSelect::Flagged(nullptr, t2, d_rl, d_flag, b.acc_cnt2.as<uint32_t>(), d_scal, runs, st));
        HGA_TRY(b.tmp.ensure(std::max(t1, t2) + 16));
        HGA_CUDA(cub::DeviceSelect::Flagged(b.tmp.p, t1, d_rk, d_flag, b.acc_key2.as<unsigned long long>(), d_scal, runs, st));
        HGA_CUDA(cub::DeviceSelect::Flagged(b.tmp.p, t2, d_rl, d_flag, b.acc_cnt2.as<uint32_t>(), d_scal, runs, st));
        HGA_CUDA(cudaGetLastError());
        unsigned long long ns = 0;
        HGA_CUDA(cudaMemcpyAsync(&ns, d_scal, 8, cudaMemcpyDeviceToHost, st));
        HGA_CUDA(cudaStreamSynchronize(st));
        if (ns) {
            const size_t at = res_k.size();
            res_k.resize(at + ns); res_c.resize(at + ns);
            HGA_CUDA(cudaMemcpy(res_k.data() + at, b.acc_key2.p, ns * 8, cudaMemcpyDeviceToHost));
            HGA_CUDA(cudaMemcpy(res_c.data() + at, b.acc_cnt2.p, ns * 4, cudaMemcpyDeviceToHost));
        }
        return HGA_OK;
    };

    if (n_bases <= B) {
        // one pass: a key (or the sentinel) per window start
        HGA_TRY(b.keys.ensure(n_bases * 8));
        count_emit_kernel<<<grid(n_bases), 256, 0, st>>>(b.bases.as<char>(), b.off.as<uint64_t>(), n_reads, k, 0, n_bases, b.keys.as<unsigned long long>());
        HGA_TRY(finish_pass(n_bases));
    } else {
        // key-range passes. Quantiles of the key distribution from a strided sample of the windows (about 2^24 of them, all of them when the input is small)
        HGA_TRY(b.keys.ensure(B * 8));
        const uint64_t n_spans = (n_bases + COUNT_SPAN - 1) / COUNT_SPAN;
        const uint64_t sample_stride = std::max<uint64_t>(1, n_bases / std::min<uint64_t>(B, 1ull << 24));
        const uint64_t sample_cap = std::min<uint64_t>(B, (n_spans + sample_stride - 1) / sample_stride * COUNT_SPAN);
        HGA_CUDA(cudaMemsetAsync(d_scal + 1, 0, 8, st));
        count_emit_range_kernel<<<grid((n_spans + sample_stride - 1) / sample_stride), 256, 0, st>>>(b.bases.as<char>(), b.off.as<uint64_t>(), n_reads, k, n_bases, 0ull, COUNT_SENTINEL,
                                                                                                    sample_stride, b.keys.as<unsigned long long>(), sample_cap, d_scal + 1);
        unsigned long long n_sample = 0;
        HGA_CUDA(cudaMemcpyAsync(&n_sample, d_scal + 1, 8, cudaMemcpyDeviceToHost, st));
        HGA_CUDA(cudaStreamSynchronize(st));
        n_sample = std::min<unsigned long long>(n_sample, sample_cap);
        // ranges that are expected to hold ~0.6 B keys each
        const double est_keys = (double) n_sample * (double) sample_stride;
        const uint64_t n_ranges = std::max<uint64_t>(1, (uint64_t) (est_keys / (0.6 * (double) B)) + 1);
        std::vector<unsigned long long> cut(n_ranges + 1, 0ull);
        cut[n_ranges] = COUNT_SENTINEL;                                     // exclusive: no canonical k-mer has this value
        if (n_sample && n_ranges > 1) {
            HGA_TRY(b.sorted.ensure(n_sample * 8));
            size_t t1 = 0;
            HGA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, t1, b.keys.as<unsigned long long>(), b.sorted.as<unsigned long long>(), n_sample, 0, end_bit, st));
            HGA_TRY(b.tmp.ensure(t1 + 16));
            HGA_CUDA(cub::DeviceRadixSort::SortKeys(b.tmp.p, t1, b.keys.as<unsigned long long>(), b.sorted.as<unsigned long long>(), n_sample, 0, end_bit, st));
            for (uint64_t r = 1; r < n_ranges; r++)
                HGA_CUDA(cudaMemcpy(&cut[r], b.sorted.as<unsigned long long>() + (size_t) ((double) n_sample * (double) r / (double) n_ranges), 8, cudaMemcpyDeviceToHost));
        }
        // the ranges, ascending; one that overflows the budget is split at its midpoint and redone (a single value that overflows IS its own count)
        std::vector<std::pair<unsigned long long, unsigned long long>> todo;
        for (uint64_t r = n_ranges; r-- > 0;) if (cut[r] < cut[r + 1]) todo.push_back({cut[r], cut[r + 1]});
        while (!todo.empty()) {
            const unsigned long long lo = todo.back().first, hi = todo.back().second;
            todo.pop_back();
            HGA_CUDA(cudaMemsetAsync(d_scal + 1, 0, 8, st));
            count_emit_range_kernel<<<grid(n_spans), 256, 0, st>>>(b.bases.as<char>(), b.off.as<uint64_t>(), n_reads, k, n_bases, lo, hi, 1, b.keys.as<unsigned long long>(), B, d_scal + 1);
            HGA_CUDA(cudaGetLastError());
            unsigned long long m = 0;
            HGA_CUDA(cudaMemcpyAsync(&m, d_scal + 1, 8, cudaMemcpyDeviceToHost, st));
            HGA_CUDA(cudaStreamSynchronize(st));
            if (m > B) {
                if (hi - lo == 1) {                                          // one k-mer, seen m times
                    if (m >= min_count) { res_k.push_back(lo); res_c.push_back((uint32_t) std::min<unsigned long long>(m, 0xFFFFFFFFull)); }
                    continue;
                }
                const unsigned long long mid = lo + (hi - lo) / 2;
                todo.push_back({mid, hi});
                todo.push_back({lo, mid});
                continue;
            }
            HGA_TRY(finish_pass(m));
        }
    }
    const uint64_t n_out = res_k.size();
    uint64_t *hk = (uint64_t *) std::malloc((n_out + 1) * 8);
    uint32_t *hc = (uint32_t *) std::malloc((n_out + 1) * 4);
    if (!hk || !hc) { std::free(hk); std::free(hc); hga_set_error("hga_count_kmers: out of host memory"); return HGA_E_NOMEM; }
    if (n_out) { std::memcpy(hk, res_k.data(), n_out * 8); std::memcpy(hc, res_c.data(), n_out * 4); }
    out->n = n_out; out->kmer = hk; out->count = hc;
    return HGA_OK;
}

extern "C" void hga_free_kmer_counts(hga_kmer_counts_t *c) {
    if (!c) return;
    std::free((void *) c->kmer); std::free((void *) c->count);
    c->kmer = nullptr; c->count = nullptr; c->n = 0;
}
