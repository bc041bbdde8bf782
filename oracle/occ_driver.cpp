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
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <set>
#include <string>
#include <vector>

#include "occurrences/JellyfishOccurrenceReader.h"

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: occ_driver specificity <k> <paths...> | export <k> <lower> <upper> <percent> <out> <paths...>\n"); return 2; }
    const std::string mode = argv[1];
    const int k = std::atoi(argv[2]);
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
