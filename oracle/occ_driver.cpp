// ORACLE / TEST INFRASTRUCTURE ONLY: the reference's SDK-selection back end, compiled unmodified.
//
// occurrences/JellyfishOccurrenceReader.cpp (k-way merge of per-file sorted k-mer dumps :63-86, specificity table :88-108,
// export of the k-mers whose total count lies in a range :110-134) is what jellyfish_occurrences.cpp runs after jellyfish has
// counted every read file (occurrences/run_jellyfish.sh: `jellyfish bc -C` + `count -C --bc` + `dump -c` + `LC_ALL=C sort`).
// jellyfish is not in this image; the reader skips it when `<read file>_<k>-mers_sorted` already exists (:19-24), so the tests
// write those dumps themselves (exact canonical counts >= 2, the two-pass Bloom-counter result without false positives) and this
// driver pins everything downstream of the counting.
//
//   occ_driver specificity <k> <read paths...>                        -> "<upper specificity> <occurrences> <unique k-mers>" lines
//   occ_driver export <k> <lower> <upper> <percent> <out> <read paths...>
//   occ_driver count <k> <min_count> <read paths...>                  -> "<k-mer value> <count>" lines, ascending
//
// `count` pins the COUNTING (what jellyfish does upstream of the reader) with reference code: the reference's own record stream
// (common/SequenceRecordIterator.cpp) and rolling canonical k-mer (common/KmerIterator.cpp) feeding a std::map. On inputs made of
// A C G T only the jellyfish rule (windows with another byte are skipped) and KmerIterator's rule (another byte reads as code 0)
// cannot differ, so on such inputs this IS what `jellyfish count -C` + `dump` give, computed by the reference's code.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <set>
#include <string>
#include <vector>

#include <map>

#include "common/KmerIterator.h"
#include "common/SequenceRecordIterator.h"
#include "occurrences/JellyfishOccurrenceReader.h"

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: occ_driver specificity <k> <paths...> | export <k> <lower> <upper> <percent> <out> <paths...>\n"); return 2; }
    const std::string mode = argv[1];
    const int k = std::atoi(argv[2]);
    if (mode == "count" && argc >= 5) {
        const unsigned long min_count = std::strtoul(argv[3], nullptr, 10);
        std::vector<std::string> paths(argv + 4, argv + argc);
        SequenceRecordIterator reader(paths, false);
        std::map<Kmer, unsigned long> counts;
        std::optional<GenomeReadData> rec;
        while ((rec = reader.get_next_record()) != std::nullopt) {
            KmerIterator it(rec->sequence, k);
            while (it.next_kmer()) counts[it.current_kmer]++;
        }
        for (auto &kv : counts) if (kv.second >= min_count) printf("%llu %lu\n", (unsigned long long) kv.first, kv.second);
        return 0;
    }
    if (mode == "specificity") {
        std::vector<std::string> paths(argv + 3, argv + argc);
        JellyfishOccurrenceReader reader(paths, k);
        std::set<double> thresholds = {70, 85, 90, 95, 99, 100, 100.01};          // jellyfish_occurrences.cpp:47
        KmerSpecificity spec = reader.get_specificity(thresholds);
        for (auto &t : spec) for (auto &oc : t.second) printf("%.2f %d %d\n", t.first, oc.first, oc.second);
        return 0;
    }
    if (mode == "export" && argc >= 8) {
        const int lower = std::atoi(argv[3]), upper = std::atoi(argv[4]);
        const double percent = std::atof(argv[5]);
        std::string out = argv[6];
        std::vector<std::string> paths(argv + 7, argv + argc);
        JellyfishOccurrenceReader reader(paths, k);
        reader.export_kmers(lower, upper, percent, out);
        std::cout << "\n";
        return 0;
    }
    return 2;
}
