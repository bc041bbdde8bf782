// jf_occurrences - the reference's SDK-selection front end (jellyfish_occurrences.cpp) on top of libhga_b200.so: it produces the
// --kmers file `categorization` takes.
//
//   jf_occurrences <read files...> [-k K] [-o OUT]           then, on stdin:  <lower> <upper> <percent>
//
// Same options as jellyfish_occurrences.cpp:22-30 (read_paths, --k-size / -k, --output / -o, --help). Where the reference shells out
// to jellyfish once per read file (occurrences/run_jellyfish.sh) this program counts the canonical k-mers of every file on the GPU
// (hga_count_kmers: exact counts, k-mers seen at least twice); the merge of the per-file lists, the specificity table and the
// export of a count range are the reader's host arithmetic (hga_host_sdk_*; JellyfishOccurrenceReader.cpp:63-134). Differences,
// on purpose: the specificity table is printed as text (the reference pipes it into a Python plot, Plotting.cpp); without -k the k
// sweep of get_unique_k_length (occurrences/KmerAnalysis.cpp:41-56: k = 11, 13, ... until the number of distinct k-mers changes by
// less than 10 %) uses EXACT distinct counts instead of a HyperLogLog estimate; the `percent` sampling draws from a generator
// seeded with --seed (default 0) instead of std::random_device. No GPU, no result: there is no CPU path.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "hga_b200.h"
#include "hga_host.h"

namespace {

void check(int rc, const char *what) {
    if (rc != HGA_OK) {
        std::cerr << "jf_occurrences: " << what << " failed (" << rc << "): " << hga_last_error() << "\n";
        std::exit(2);
    }
}

std::string number_to_sequence(uint64_t v, int k) {            // KmerIterator::number_to_sequence
    std::string s((size_t) k, 'A');
    for (int i = k - 1; i >= 0; i--) { s[(size_t) i] = "ACGT"[v & 3]; v >>= 2; }
    return s;
}

}  // namespace

int main(int argc, char **argv) {
    std::vector<std::string> read_paths;
    std::string output_path;
    int k = 0, device = 0, threads = 0;
    uint64_t seed = 0;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto need = [&]() -> const char * { if (i + 1 >= argc) throw std::invalid_argument("the required argument for option '" + a + "' is missing"); return argv[++i]; };
        if (a == "-h" || a == "--help") {
            std::cout << "Options:\n  -h [ --help ]         Help screen\n  --read_paths arg      Path to file with reads (FASTA or FASTQ)\n"
                         "  -k [ --k-size ] arg   Size of kmer to analyze & select\n  -o [ --output ] arg   Output path for the exported kmers\n";
            return 0;
        } else if (a == "-k" || a == "--k-size") k = std::atoi(need());
        else if (a == "-o" || a == "--output") output_path = need();
        else if (a == "--read_paths") read_paths.push_back(need());
        else if (a == "--device") device = std::atoi(need());
        else if (a == "--threads" || a == "-t") threads = std::atoi(need());
        else if (a == "--seed") seed = std::strtoull(need(), nullptr, 10);
        else if (a.size() > 1 && a[0] == '-') throw std::invalid_argument("unrecognised option '" + a + "'");
        else read_paths.push_back(a);
    }
    if (read_paths.empty()) throw std::invalid_argument("You need to specify paths to read files");      // jellyfish_occurrences.cpp:38

    hga_host::SequenceRecords reads(read_paths, threads);
    std::vector<uint64_t> file_first{0};                       // first read of every file
    for (const auto &m : reads.file_meta) file_first.push_back(file_first.back() + m.records);

    if (k == 0) {
        // get_unique_k_length (KmerAnalysis.cpp:41-56) with exact distinct counts
        auto distinct = [&](int kk) {
            hga_kmer_counts_t c;
            check(hga_count_kmers(device, kk, reads.bases_data, reads.seq_off.data(), reads.n_reads(), 1, &c), "hga_count_kmers");
            const uint64_t n = c.n;
            hga_free_kmer_counts(&c);
            return (long long) n;
        };
        k = 11;
        long long previous = distinct(k);
        std::cout << "k=11 : " << previous << " kmers" << std::endl;
        while (k < 33 && k + 2 <= 32) {
            const long long count = distinct(k + 2);
            std::cout << "k=" << (k + 2) << " : " << count << " kmers" << std::endl;
            if (((double) std::llabs(count - previous) / ((double) (count + previous) / 2.0)) < 0.1) break;
            k += 2;
            previous = count;
        }
    }

    // per file: exact canonical counts, k-mers seen at least twice (run_jellyfish.sh: bc + count --bc + dump + sort)
    std::vector<uint64_t> file_off{0}, kmers;
    std::vector<uint32_t> counts;
    for (size_t f = 0; f + 1 < file_first.size(); f++) {
        hga_kmer_counts_t c;
        check(hga_count_kmers(device, k, reads.bases_data, reads.seq_off.data() + file_first[f], file_first[f + 1] - file_first[f], 2, &c), "hga_count_kmers");
        kmers.insert(kmers.end(), c.kmer, c.kmer + c.n);
        counts.insert(counts.end(), c.count, c.count + c.n);
        file_off.push_back(kmers.size());
        hga_free_kmer_counts(&c);
    }
    const uint64_t cap = kmers.size() + 1;
    std::vector<uint64_t> m_kmer(cap);
    std::vector<uint32_t> m_total(cap), m_max(cap), m_files(cap);
    uint64_t n = 0;
    check(hga_host_sdk_merge((int) (file_off.size() - 1), file_off.data(), kmers.data(), counts.data(), m_kmer.data(), m_total.data(), m_max.data(), m_files.data(), &n),
          "hga_host_sdk_merge");

    const double thresholds[] = {70, 85, 90, 95, 99, 100, 100.01};      // jellyfish_occurrences.cpp:47
    uint64_t rows = 0;
    check(hga_host_sdk_specificity(n, m_total.data(), m_max.data(), thresholds, 7, nullptr, nullptr, nullptr, 0, &rows), "hga_host_sdk_specificity");
    std::vector<double> s_thr(rows + 1);
    std::vector<uint32_t> s_occ(rows + 1);
    std::vector<uint64_t> s_cnt(rows + 1);
    check(hga_host_sdk_specificity(n, m_total.data(), m_max.data(), thresholds, 7, s_thr.data(), s_occ.data(), s_cnt.data(), rows, &rows), "hga_host_sdk_specificity");
    std::cout << "# k = " << k << ": upper specificity, occurrences, unique k-mers\n";
    for (uint64_t i = 0; i < rows; i++) std::printf("%.2f %u %llu\n", s_thr[i], s_occ[i], (unsigned long long) s_cnt[i]);
    std::fflush(stdout);

    int lower = 0, upper = 0;
    double percent = 0;
    std::cout << "Enter lower and upper bounds for exported kmers as well as percentage\n";      // :52
    if (!(std::cin >> lower >> upper >> percent)) { std::cerr << "jf_occurrences: expected <lower> <upper> <percent> on stdin\n"; return 2; }
    if (output_path.empty()) {
        char buf[128];
        std::snprintf(buf, sizeof buf, "%d-mers_%d_%d_%g%%.txt", k, lower, upper, percent * 100);      // :57
        output_path = buf;
    }
    std::vector<uint8_t> sel(n + 1);
    uint64_t n_sel = 0, n_disc = 0;
    check(hga_host_sdk_select(n, m_total.data(), m_files.data(), (uint32_t) std::max(lower, 0), (uint32_t) std::max(upper, 0), percent, seed, sel.data(), &n_sel, &n_disc),
          "hga_host_sdk_select");
    std::ofstream out(output_path);
    for (uint64_t i = 0; i < n; i++) if (sel[i]) out << number_to_sequence(m_kmer[i], k) << "\n";
    out.close();
    std::cout << n_disc << " out of " << n_sel << " exported kmers are discriminative\n";      // :132
    return 0;
}
