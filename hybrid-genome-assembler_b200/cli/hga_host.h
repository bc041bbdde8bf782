// Host side of the drop-in `categorization` executable: the k-mer file loader and the FASTA/FASTQ record loader,
// with the reference's rules (paths relative to /root/reference/src):
//   load_text_file_kmers   read_clustering.cpp:18-33      one KmerIterator per line, first window only, k = last line
//   KmerIterator tables    common/KmerIterator.cpp:6-21   A0 C1 G2 T3 / complement A3 C2 G1 T0, any other byte 0 in BOTH
//   SequenceRecords        common/SequenceRecordIterator.cpp:73-173   format sniffed per file from its first record,
//                          FASTQ = 4 lines / record, FASTA = 2 lines / record, the record reader runs on across file
//                          boundaries, header = line minus its first character, ReadID = 1, 2, ... over all files,
//                          per-file and aggregate MetaData (the aggregate max_read_length is never updated: prints 0)
// Nothing here touches the GPU; the arithmetic on reads happens behind include/hga_b200.h.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace hga_host {

inline uint64_t canonical_first_window(const std::string &s, int k) {
    if (k > 32) throw std::invalid_argument("Kmer size is too big");     // KmerIterator.cpp:24-26
    if (k == 0 || (int) s.size() < k) return 0;                          // no window: current_kmer stays 0
    uint64_t fwd = 0, rev = 0;
    for (int i = 0; i < k; i++) {
        uint64_t c = 0, cc = 0;
        switch (s[i]) {
            case 'A': c = 0; cc = 3; break;
            case 'C': c = 1; cc = 2; break;
            case 'G': c = 2; cc = 1; break;
            case 'T': c = 3; cc = 0; break;
            default: c = 0; cc = 0; break;                                // operator[] default-inserts 0 in both tables
        }
        fwd |= c << (2 * (k - 1 - i));
        rev |= cc << (2 * i);
    }
    return std::min(fwd, rev);
}

struct KmerSet {
    std::vector<uint64_t> kmers;   // sorted, unique canonical values
    int k = 0;
};

inline KmerSet load_text_file_kmers(const std::string &path) {
    KmerSet out;
    std::ifstream in(path, std::ios::binary);
    std::string line;
    while (std::getline(in, line)) {                                      // a missing file yields an empty set, as in the reference
        out.k = (int) line.size();
        out.kmers.push_back(canonical_first_window(line, out.k));
    }
    std::sort(out.kmers.begin(), out.kmers.end());
    out.kmers.erase(std::unique(out.kmers.begin(), out.kmers.end()), out.kmers.end());
    return out;
}

struct MetaData {
    std::string filename;
    uint64_t records = 0, min_read_length = UINT64_MAX, max_read_length = 0, avg_read_length = 0, total_bases = 0;
    std::string repr() const {                                            // SequenceRecordIterator.h:60-63
        std::ostringstream o;
        o << filename << ":\n- " << records << " reads\n- " << total_bases << " total bases\n- " << avg_read_length << " average read length\n- "
          << max_read_length << " max read length\n- " << min_read_length << " min read length\n\n";
        return o.str();
    }
};

struct SequenceRecords {
    std::vector<std::string> paths;
    std::vector<MetaData> file_meta;
    MetaData meta;
    std::string bases;                   // sequences back to back
    std::vector<uint64_t> seq_off;       // n_reads + 1
    std::vector<std::string> headers, qualities;
    std::vector<int> file_index;
    bool fastq_last = true;

    size_t n_reads() const { return headers.size(); }

    std::string fastx_string(size_t i) const {                            // SequenceRecordIterator.h:36-48
        const std::string seq = bases.substr(seq_off[i], seq_off[i + 1] - seq_off[i]);
        if (!qualities[i].empty()) return "@" + headers[i] + "\n" + seq + "\n+\n" + qualities[i];
        return ">" + headers[i] + "\n" + seq;
    }

    explicit SequenceRecords(const std::vector<std::string> &read_paths) : paths(read_paths), file_meta(read_paths.size()) {
        std::vector<std::string> lines;
        size_t at = 0;
        int cur = -1, method = 4;
        auto open_file = [&](int pos) {
            std::ifstream f(paths[pos], std::ios::binary);
            if (!f) throw std::invalid_argument("File with path \"" + paths[pos] + "\" does not exist");
            lines.clear();
            std::string l;
            while (std::getline(f, l)) lines.push_back(l);
            at = 0;
            // sniff (load_file_at_position :85-99)
            if (lines.size() < 2) throw std::logic_error("File is empty");
            const char h0 = lines[0].empty() ? '\0' : lines[0][0];
            if (h0 == '@') {
                if (lines.size() < 3) throw std::logic_error("File is empty");
                if (!lines[2].empty() && lines[2][0] == '+') method = 4;
            } else if (h0 == '>') {
                method = 2;
            } else {
                throw std::logic_error("Unrecognized file format");
            }
        };
        auto next_line = [&](std::string &out) -> bool {
            while (at >= lines.size()) {
                if (cur + 1 >= (int) paths.size()) return false;
                cur++;
                open_file(cur);
            }
            out = lines[at++];
            return true;
        };
        cur = 0;
        open_file(0);
        int prev_file = -1;
        std::vector<std::string> names;
        seq_off.push_back(0);
        for (;;) {
            const int n = method;
            std::string rec[4];
            int got = 0;
            for (; got < n; got++) if (!next_line(rec[got])) break;
            if (got < n) break;
            if (rec[0].empty()) throw std::out_of_range("basic_string::substr");    // header.substr(1) on an empty header
            headers.push_back(rec[0].substr(1));
            bases += rec[1];
            seq_off.push_back(bases.size());
            qualities.push_back(n == 4 ? rec[3] : std::string());
            file_index.push_back(cur);
            if (cur != prev_file) {
                file_meta[cur] = MetaData();
                const size_t slash = paths[cur].find_last_of('/');
                file_meta[cur].filename = slash == std::string::npos ? paths[cur] : paths[cur].substr(slash + 1);
                names.push_back(file_meta[cur].filename);
                prev_file = cur;
            }
            MetaData &fm = file_meta[cur];
            const uint64_t L = rec[1].size();
            fm.total_bases += L; fm.min_read_length = std::min(fm.min_read_length, L); fm.max_read_length = std::max(fm.max_read_length, L);
            fm.records++; fm.avg_read_length += L;
            meta.total_bases += L; meta.min_read_length = std::min(meta.min_read_length, L); meta.records++; meta.avg_read_length += L;
        }
        fastq_last = method == 4;
        if (meta.records == 0) throw std::logic_error("File is empty");
        meta.avg_read_length /= meta.records;          // aggregate max_read_length stays 0 (never updated, :59-62)
        for (size_t i = 0; i < names.size(); i++) meta.filename += (i ? "__" : "") + names[i];
        for (auto &fm : file_meta) if (fm.records) fm.avg_read_length /= fm.records;
    }
};

}  // namespace hga_host
