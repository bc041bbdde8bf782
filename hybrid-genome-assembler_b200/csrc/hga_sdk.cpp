// SDK selection, host half: what occurrences/JellyfishOccurrenceReader.cpp does with the per-file sorted k-mer
// dumps once jellyfish has counted them. HOST code, as in the reference; the counting itself is hga_count_kmers (csrc/hga_count.cu).
//   hga_host_sdk_merge        <- get_next_kmer (:63-86): k-way merge of the files' ascending (k-mer, count) lists; per distinct k-mer
//                                the total count, the largest per-file count and the number of files that hold it
//   hga_host_sdk_specificity  <- get_specificity (:88-108): table [smallest threshold > 100 * max / total][total] = number of k-mers
//   hga_host_sdk_select       <- export_kmers (:110-134): lower <= total <= upper, each kept with probability `percent`; the
//                                reference draws from std::random_device (not reproducible), here from a seeded generator and
//                                percent >= 1 keeps everything; also the number of selected k-mers present in exactly ONE file
// Pinned against the reference's own reader compiled unmodified (oracle/_ref/occ_driver, tests/test_sdk_selection_cpu.py).
#include <algorithm>
#include <cstdint>
#include <map>
#include <queue>
#include <random>
#include <vector>

#include "../../include/hga_b200.h"

void hga_set_error(const char *fmt, ...);

extern "C" int hga_host_sdk_merge(int n_files, const uint64_t *file_off, const uint64_t *kmer, const uint32_t *count, uint64_t *out_kmer, uint32_t *out_total,
                                  uint32_t *out_max, uint32_t *out_files, uint64_t *out_n) {
    if (n_files < 0 || !file_off || !out_n || (file_off[n_files] && (!kmer || !count || !out_kmer || !out_total || !out_max || !out_files))) {
        hga_set_error("hga_host_sdk_merge: bad argument");
        return HGA_E_ARG;
    }
    using Head = std::pair<uint64_t, int>;                                  // (k-mer, file)
    std::priority_queue<Head, std::vector<Head>, std::greater<Head>> heads;
    std::vector<uint64_t> at(n_files);
    for (int f = 0; f < n_files; f++) {
        at[f] = file_off[f];
        for (uint64_t i = file_off[f] + 1; i < file_off[f + 1]; i++)
            if (kmer[i] <= kmer[i - 1]) { hga_set_error("hga_host_sdk_merge: the k-mers of file %d are not strictly ascending", f); return HGA_E_ARG; }
        if (at[f] < file_off[f + 1]) heads.push({kmer[at[f]], f});
    }
    uint64_t n = 0;
    while (!heads.empty()) {
        const uint64_t cur = heads.top().first;
        uint64_t total = 0;
        uint32_t mx = 0, files = 0;
        while (!heads.empty() && heads.top().first == cur) {
            const int f = heads.top().second;
            heads.pop();
            const uint32_t c = count[at[f]];
            total += c; mx = std::max(mx, c); files += c > 0;
            if (++at[f] < file_off[f + 1]) heads.push({kmer[at[f]], f});
        }
        if (total > 0xFFFFFFFFull) { hga_set_error("hga_host_sdk_merge: total count overflows 32 bits"); return HGA_E_OVERFLOW; }
        out_kmer[n] = cur; out_total[n] = (uint32_t) total; out_max[n] = mx; out_files[n] = files;
        n++;
    }
    *out_n = n;
    return HGA_OK;
}

extern "C" int hga_host_sdk_specificity(uint64_t n, const uint32_t *total, const uint32_t *max, const double *thresholds, int n_thresholds, double *out_threshold,
                                        uint32_t *out_occurrences, uint64_t *out_unique, uint64_t capacity, uint64_t *out_n) {
    if (!out_n || n_thresholds < 1 || !thresholds || (n && (!total || !max))) { hga_set_error("hga_host_sdk_specificity: bad argument"); return HGA_E_ARG; }
    std::vector<double> thr(thresholds, thresholds + n_thresholds);
    std::sort(thr.begin(), thr.end());
    std::map<double, std::map<uint32_t, uint64_t>> table;
    for (double t : thr) table[t];
    for (uint64_t i = 0; i < n; i++) {
        const double v = ((double) max[i] / (double) total[i]) * 100;       // :101
        auto it = std::upper_bound(thr.begin(), thr.end(), v);
        if (it == thr.end()) { hga_set_error("hga_host_sdk_specificity: no threshold above %.4f (the reference dereferences end() here)", v); return HGA_E_ARG; }
        table[*it][total[i]] += 1;
    }
    uint64_t m = 0;
    for (const auto &t : table)
        for (const auto &oc : t.second) {
            if (m < capacity && out_threshold && out_occurrences && out_unique) { out_threshold[m] = t.first; out_occurrences[m] = oc.first; out_unique[m] = oc.second; }
            m++;
        }
    *out_n = m;                                                              // > capacity: call again with more room
    return HGA_OK;
}

extern "C" int hga_host_sdk_select(uint64_t n, const uint32_t *total, const uint32_t *files, uint32_t lower, uint32_t upper, double percent, uint64_t seed,
                                   uint8_t *out_selected, uint64_t *out_n_selected, uint64_t *out_n_discriminative) {
    if (!out_n_selected || !out_n_discriminative || (n && (!total || !files || !out_selected))) { hga_set_error("hga_host_sdk_select: bad argument"); return HGA_E_ARG; }
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<> dis(0.0, 1.0);
    uint64_t sel = 0, disc = 0;
    for (uint64_t i = 0; i < n; i++) {
        const bool keep = lower <= total[i] && total[i] <= upper && (percent >= 1.0 || dis(rng) < percent);      // :124
        out_selected[i] = keep;
        sel += keep;
        disc += keep && files[i] == 1;                                                                            // :128
    }
    *out_n_selected = sel; *out_n_discriminative = disc;
    return HGA_OK;
}
