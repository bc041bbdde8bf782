// count: exact canonical k-mer counts of a set of reads (the GPU half of the SDK selection).
//
// Replaces what occurrences/run_jellyfish.sh asks jellyfish for, per read file (occurrences/JellyfishOccurrenceReader.cpp:16-38):
//   jellyfish bc -C -m k ...; jellyfish count -C -m k --bc ...; jellyfish dump -c; LC_ALL=C sort
// i.e. the canonical k-mers (-C: the smaller of a k-mer and its reverse complement; with A < C < G < T that is the smaller 2-bit
// value, the same canonical form as KmerIterator.cpp:69) that occur at least twice (the two-pass Bloom-counter filter, here without
// its false positives), with their exact counts, in ascending order (string order = value order for equal-length ACGT strings).
// jellyfish itself is not in this image and is not vendored by the reference, so the counting has no golden output to pin against:
// "parity unpinned" for this stage; the checker is exact counting in numpy (tests/test_sdk_selection_cpu.py).
// Windows containing a byte other than A C G T a c g t are skipped, as jellyfish does (NOT the code-0 rule of KmerIterator).
//
// Kernels: count_emit_kernel (one thread per window start: binary search of the read, k byte loads, canonical value or a sentinel),
// CUB radix sort + run-length encode per chunk of 2^28 positions, partial (k-mer, count) lists of the chunks merged by one more
// sort + reduce-by-key, count_flag_kernel + CUB select for count >= min_count. Simple on purpose (first version, 8 B per position of
// scratch); HBM-bound by the sort passes: 2k bits -> ceil(2k / 8) passes x 16 B per position.
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

__global__ void count_emit_kernel(const char *__restrict__ bases, const uint64_t *__restrict__ read_off, uint64_t n_reads, int k, uint64_t p0, uint64_t p1,
                                  unsigned long long *__restrict__ out) {
    for (uint64_t p = p0 + blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; p < p1; p += (uint64_t) gridDim.x * blockDim.x) {
        // the read that holds position p: the last r with read_off[r] <= p
        uint64_t lo = 0, hi = n_reads;
        while (lo < hi) { const uint64_t mid = (lo + hi + 1) >> 1; if (read_off[mid] <= p) lo = mid; else hi = mid - 1; }
        unsigned long long key = COUNT_SENTINEL;
        if (p + (uint64_t) k <= read_off[lo + 1]) {
            unsigned long long fwd = 0, rc = 0;
            bool ok = true;
            for (int j = 0; j < k; j++) {
                const int c = count_base_code((unsigned char) bases[p + j]);
                if (c < 0) { ok = false; break; }
                fwd = (fwd << 2) | (unsigned long long) c;
                rc = (rc >> 2) | ((unsigned long long) (3 - c) << (2 * (k - 1)));
            }
            if (ok) {
                if (k < 32) fwd &= (1ull << (2 * k)) - 1;
                key = fwd < rc ? fwd : rc;
            }
        }
        out[p - p0] = key;
    }
}

__global__ void count_flag_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ cnt, uint64_t n, uint32_t min_count, uint8_t *__restrict__ flag) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
        flag[i] = keys[i] != COUNT_SENTINEL && cnt[i] >= min_count;
}

struct SumU32 {
    __host__ __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a + b; }
};

struct Bufs {
    DevBuf bases, off, keys, sorted, run_key, run_len, acc_key, acc_cnt, acc_key2, acc_cnt2, tmp, scalars, flag;
    ~Bufs() { for (DevBuf *b : {&bases, &off, &keys, &sorted, &run_key, &run_len, &acc_key, &acc_cnt, &acc_key2, &acc_cnt2, &tmp, &scalars, &flag}) b->release(); }
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
    // positions per chunk; HGA_COUNT_CHUNK overrides it (tests force the multi-chunk merge on small inputs with it)
    const char *ch_env = getenv("HGA_COUNT_CHUNK");
    const uint64_t CH = ch_env && std::strtoull(ch_env, nullptr, 10) >= 256 ? std::strtoull(ch_env, nullptr, 10) : (1ull << 28);
    // 2k key bits are enough: no canonical k-mer has all of them set (T...T is not canonical), so the sentinels still sort behind
    // every k-mer and stay together
    const int end_bit = 2 * k;
    uint64_t n_acc = 0;
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
    auto grid = [&](uint64_t n) { return (int) std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t) sm * 16)); };
    for (uint64_t p0 = 0; p0 < n_bases; p0 += CH) {
        const uint64_t p1 = std::min(n_bases, p0 + CH), m = p1 - p0;
        HGA_TRY(b.keys.ensure(m * 8)); HGA_TRY(b.sorted.ensure(m * 8)); HGA_TRY(b.run_key.ensure(m * 8)); HGA_TRY(b.run_len.ensure(m * 4));
        unsigned long long *d_keys = b.keys.as<unsigned long long>(), *d_sorted = b.sorted.as<unsigned long long>(), *d_rk = b.run_key.as<unsigned long long>();
        uint32_t *d_rl = b.run_len.as<uint32_t>();
        count_emit_kernel<<<grid(m), 256, 0, st>>>(b.bases.as<char>(), b.off.as<uint64_t>(), n_reads, k, p0, p1, d_keys);
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
        // append the chunk's (k-mer, count) runs to the accumulated list (the sentinel run, if any, goes along and is dropped at the end)
        if (b.acc_key.cap < (n_acc + runs) * 8) {
            DevBuf nk, nc;
            HGA_TRY(nk.ensure((n_acc + runs) * 8 * 2)); 
            if (nc.ensure((n_acc + runs) * 4 * 2) != HGA_OK) { nk.release(); return HGA_E_NOMEM; }
            if (n_acc) {
                HGA_CUDA(cudaMemcpyAsync(nk.p, b.acc_key.p, n_acc * 8, cudaMemcpyDeviceToDevice, st));
                HGA_CUDA(cudaMemcpyAsync(nc.p, b.acc_cnt.p, n_acc * 4, cudaMemcpyDeviceToDevice, st));
                HGA_CUDA(cudaStreamSynchronize(st));
            }
            b.acc_key.release(); b.acc_cnt.release();
            b.acc_key = nk; b.acc_cnt = nc;
        }
        HGA_CUDA(cudaMemcpyAsync(b.acc_key.as<unsigned long long>() + n_acc, d_rk, runs * 8, cudaMemcpyDeviceToDevice, st));
        HGA_CUDA(cudaMemcpyAsync(b.acc_cnt.as<uint32_t>() + n_acc, d_rl, runs * 4, cudaMemcpyDeviceToDevice, st));
        n_acc += runs;
    }
    unsigned long long *d_k = b.acc_key.as<unsigned long long>();
    uint32_t *d_c = b.acc_cnt.as<uint32_t>();
    if (n_bases > CH && n_acc) {
        // several chunks: the same k-mer can head a run in more than one of them
        if (n_acc > 0x7FFFFFF0ull) { hga_set_error("hga_count_kmers: %llu partial runs exceed this version's merge limit", (unsigned long long) n_acc); return HGA_E_OVERFLOW; }
        HGA_TRY(b.acc_key2.ensure(n_acc * 8)); HGA_TRY(b.acc_cnt2.ensure(n_acc * 4));
        unsigned long long *d_k2 = b.acc_key2.as<unsigned long long>();
        uint32_t *d_c2 = b.acc_cnt2.as<uint32_t>();
        size_t t1 = 0, t2 = 0;
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, d_k, d_k2, d_c, d_c2, n_acc, 0, 64, st));
        HGA_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, t2, d_k2, d_k, d_c2, d_c, d_scal, SumU32(), (int) n_acc, st));
        HGA_TRY(b.tmp.ensure(std::max(t1, t2) + 16));
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(b.tmp.p, t1, d_k, d_k2, d_c, d_c2, n_acc, 0, 64, st));
        HGA_CUDA(cub::DeviceReduce::ReduceByKey(b.tmp.p, t2, d_k2, d_k, d_c2, d_c, d_scal, SumU32(), (int) n_acc, st));
        unsigned long long nu = 0;
        HGA_CUDA(cudaMemcpyAsync(&nu, d_scal, 8, cudaMemcpyDeviceToHost, st));
        HGA_CUDA(cudaStreamSynchronize(st));
        n_acc = nu;
    }
    // count >= min_count, sentinel dropped
    uint64_t n_out = 0;
    if (n_acc) {
        HGA_TRY(b.flag.ensure(n_acc + 16));
        HGA_TRY(b.acc_key2.ensure(n_acc * 8)); HGA_TRY(b.acc_cnt2.ensure(n_acc * 4));
        uint8_t *d_flag = b.flag.as<uint8_t>();
        count_flag_kernel<<<grid(n_acc), 256, 0, st>>>(d_k, d_c, n_acc, min_count, d_flag);
        size_t t1 = 0, t2 = 0;
        HGA_CUDA(cub::DeviceSelect::Flagged(nullptr, t1, d_k, d_flag, b.acc_key2.as<unsigned long long>(), d_scal, n_acc, st));
        HGA_CUDA(cub::DeviceSelect::Flagged(nullptr, t2, d_c, d_flag, b.acc_cnt2.as<uint32_t>(), d_scal, n_acc, st));
        HGA_TRY(b.tmp.ensure(std::max(t1, t2) + 16));
        HGA_CUDA(cub::DeviceSelect::Flagged(b.tmp.p, t1, d_k, d_flag, b.acc_key2.as<unsigned long long>(), d_scal, n_acc, st));
        HGA_CUDA(cub::DeviceSelect::Flagged(b.tmp.p, t2, d_c, d_flag, b.acc_cnt2.as<uint32_t>(), d_scal, n_acc, st));
        HGA_CUDA(cudaGetLastError());
        unsigned long long ns = 0;
        HGA_CUDA(cudaMemcpyAsync(&ns, d_scal, 8, cudaMemcpyDeviceToHost, st));
        HGA_CUDA(cudaStreamSynchronize(st));
        n_out = ns;
    }
    uint64_t *hk = (uint64_t *) std::malloc((n_out + 1) * 8);
    uint32_t *hc = (uint32_t *) std::malloc((n_out + 1) * 4);
    if (!hk || !hc) { std::free(hk); std::free(hc); hga_set_error("hga_count_kmers: out of host memory"); return HGA_E_NOMEM; }
    if (n_out) {
        cudaError_t e1 = cudaMemcpy(hk, b.acc_key2.p, n_out * 8, cudaMemcpyDeviceToHost);
        cudaError_t e2 = cudaMemcpy(hc, b.acc_cnt2.p, n_out * 4, cudaMemcpyDeviceToHost);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { std::free(hk); std::free(hc); hga_set_error("hga_count_kmers: D2H failed"); return HGA_E_CUDA; }
    }
    out->n = n_out; out->kmer = hk; out->count = hc;
    return HGA_OK;
}

extern "C" void hga_free_kmer_counts(hga_kmer_counts_t *c) {
    if (!c) return;
    std::free((void *) c->kmer); std::free((void *) c->count);
    c->kmer = nullptr; c->count = nullptr; c->n = 0;
}
