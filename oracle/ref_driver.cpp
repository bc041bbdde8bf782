// ORACLE / TEST INFRASTRUCTURE ONLY — never linked into or called by the product path.
//
// Driver around the UNMODIFIED reference sources (compiled in place from /root/reference/src by
// oracle/Makefile, outputs only under oracle/_ref/). It reaches the protected hot-path stages of
// ReadClusteringEngine (clustering/ReadClusteringEngine.h:151-190) through a subclass and dumps every
// stage result as raw little-endian arrays so that tests can pin the C restatement (oracle/hga_oracle.c)
// and, through it, the CUDA path.
//
// Stages exercised (all real reference code):
//   KmerIterator                      common/KmerIterator.cpp:23-76
//   SequenceRecordIterator            common/SequenceRecordIterator.cpp
//   construct_indices                 clustering/ReadClusteringEngine.cpp:234-299
//   get_all_connections               clustering/ReadClusteringEngine.cpp:301-339
//   union_find                        clustering/ReadClusteringEngine.cpp:424-489
//   merge_components (--enrich)       clustering/ReadClusteringEngine.cpp:349-422
//   get_connections  (--enrich)       clustering/ReadClusteringEngine.cpp:301-333  (pivots = the merged cores)
//   get_core_component_connections, spectral_clustering (--enrich N --full)   clustering/ReadClusteringEngine.cpp:491-697, lib/clustering/*
//                                     (compiled against the Eigen2 stand-in of oracle/shim/eigen2: Jacobi eigen-solver)
// Restated here because read_clustering.cpp cannot be compiled (boost::program_options is absent):
//   load_text_file_kmers              read_clustering.cpp:18-33   (7 lines, uses the real KmerIterator)
//   the 15 % cut                      clustering/ReadClusteringEngine.cpp:754-755
// The reference leaves the order inside a score tie group unspecified (hash iteration order + unstable
// std::sort, .cpp:331). The driver imposes the canonical total order of SURVEY.md §8a-6
//   (score desc, min(x,y) asc, max(x,y) asc, x asc)
// before slicing, so that n, s*, the selected edge set, the components and the spanning forest are all
// well defined.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <set>
#include <string>
#include <unordered_set>
#include <vector>

#include "clustering/ReadClusteringEngine.h"
#include "common/KmerIterator.h"

// external-linkage free function of the reference (clustering/ReadClusteringEngine.cpp:424)
std::vector<std::pair<Component, SpanningTree>>
union_find(std::vector<ComponentConnection> &connections, std::set<ComponentID> &restricted, int min_component_size, int max_component_size);

// free functions of the reference with external linkage (clustering/ReadClusteringEngine.cpp:126, :653)
std::vector<ComponentConnection> filter_connections(std::vector<ComponentConnection> &original, const std::function<bool(ComponentConnection &)> &func);
std::vector<Component> spectral_clustering(std::vector<ComponentConnection> &connections, int dims);

// lib/clustering/{SpectralClustering,ClusterRotate,Evrot}.cpp are compiled unmodified against the Eigen2 stand-in of
// oracle/shim/eigen2 (Jacobi eigen-solver). Kmeans.cpp is not compiled: clusterKmeans is never called by categorization.
#include "lib/clustering/Kmeans.h"
std::vector<std::vector<int>> Kmeans::cluster(Eigen::MatrixXd &, int) { throw std::logic_error("k-means is not part of the categorization path"); }

namespace {

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template<typename T>
void dump(const std::string &dir, const std::string &name, const std::vector<T> &v) {
    std::string path = dir + "/" + name;
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); exit(2); }
    if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
    fclose(f);
}

// read_clustering.cpp:18-33 — k is the length of the LAST line; each line is canonicalised through the
// first window of a KmerIterator; duplicates collapse in the unordered_set.
std::pair<std::unordered_set<Kmer>, int> load_text_file_kmers(const std::string &path) {
    std::ifstream in;
    in.open(path);
    std::string kmer;
    int k = 0;
    std::unordered_set<Kmer> s;
    while (std::getline(in, kmer)) {
        k = kmer.length();
        KmerIterator k_it(kmer, k);
        k_it.next_kmer();
        s.insert(k_it.current_kmer);
    }
    return {s, k};
}

bool canonical_less(const ComponentConnection &a, const ComponentConnection &b) {
    if (a.score != b.score) return a.score > b.score;
    ComponentID amin = std::min(a.component_x_id, a.component_y_id), amax = std::max(a.component_x_id, a.component_y_id);
    ComponentID bmin = std::min(b.component_x_id, b.component_y_id), bmax = std::max(b.component_x_id, b.component_y_id);
    if (amin != bmin) return amin < bmin;
    if (amax != bmax) return amax < bmax;
    return a.component_x_id < b.component_x_id;
}

class Probe : public ReadClusteringEngine {
public:
    Probe(SequenceRecordIterator &it, ReadClusteringConfig cfg) : ReadClusteringEngine(it, cfg) {}

    int run(std::unordered_set<Kmer> &kmers, int k, const std::string &out, bool do_dump, double fraction, int min_size,
            ConnectionScore min_score, int stop_after, ConnectionScore enrich_min = 0, ConnectionScore sc_score = 0, bool full = false, bool force_spectral = false) {
        // KmerID assignment order = iteration order of the very same unordered_set object (.cpp:237-241)
        std::vector<Kmer> id2kmer;
        for (auto kmer : kmers) id2kmer.push_back(kmer);

        double t0 = now_ms();
        construct_indices(kmers, k);
        double t_index = now_ms() - t0;

        uint64_t E = 0;
        for (auto &lst : kmer_component_index) E += lst.size();

        FILE *meta = fopen((out + "/meta.txt").c_str(), "w");
        fprintf(meta, "k=%d\nn_kmers=%zu\nn_reads=%lu\ntotal_bases=%lu\nreads_with_hits=%zu\nincidences=%lu\nindex_ms=%.3f\n",
                k, id2kmer.size(), (unsigned long) reader->meta.records, (unsigned long) reader->meta.total_bases,
                component_index.size(), (unsigned long) E, t_index);

        if (do_dump) {
            // per-read hit multiset (k-mer VALUES, sorted) + first-occurrence positions
            std::vector<ReadID> ids;
            for (auto p : component_index) ids.push_back(p.first);
            std::sort(ids.begin(), ids.end());
            std::vector<uint32_t> hit_read, fp_read, fp_pos, read_len_id, read_len;
            std::vector<uint64_t> hit_kmer, fp_kmer;
            for (auto id : ids) {
                std::vector<uint64_t> vals;
                for (KmerID kid : component_index[id]->discriminative_kmer_ids) vals.push_back(id2kmer[kid]);
                std::sort(vals.begin(), vals.end());
                for (auto v : vals) { hit_read.push_back(id); hit_kmer.push_back(v); }
                std::vector<std::pair<uint64_t, uint32_t>> fp;
                for (auto kp : read_metas[id].kmer_positions) fp.push_back({id2kmer[kp.first], kp.second});
                std::sort(fp.begin(), fp.end());
                for (auto &e : fp) { fp_read.push_back(id); fp_kmer.push_back(e.first); fp_pos.push_back(e.second); }
                read_len_id.push_back(id);
                read_len.push_back(read_metas[id].length);
            }
            dump(out, "hit_read.u32", hit_read);
            dump(out, "hit_kmer.u64", hit_kmer);
            dump(out, "firstpos_read.u32", fp_read);
            dump(out, "firstpos_kmer.u64", fp_kmer);
            dump(out, "firstpos_pos.u32", fp_pos);
            dump(out, "readlen_read.u32", read_len_id);
            dump(out, "readlen_len.u32", read_len);

            // inverted index ordered by k-mer value
            std::vector<std::pair<uint64_t, KmerID>> order;
            for (KmerID i = 0; i < id2kmer.size(); i++) order.push_back({id2kmer[i], i});
            std::sort(order.begin(), order.end());
            std::vector<uint64_t> inv_kmer, inv_off{0};
            std::vector<uint32_t> inv_read;
            for (auto &o : order) {
                inv_kmer.push_back(o.first);
                for (auto r : kmer_component_index[o.second]) inv_read.push_back(r);
                inv_off.push_back(inv_read.size());
            }
            dump(out, "inv_kmer.u64", inv_kmer);
            dump(out, "inv_off.u64", inv_off);
            dump(out, "inv_read.u32", inv_read);
        }
        if (stop_after == 1) { fclose(meta); return 0; }

        if (force_spectral) {
            // --spectral (run_clustering :739-746), all real reference code: get_all_connections(5), spectral_clustering of the whole
            // data set, merge_components, get_component_ids(min size). The driver imposes the canonical order on the connections
            // (their order decides the row order of the affinity matrix, :657-663) and skips empty clusters (see below).
            auto conns = get_all_connections(5);
            std::sort(conns.begin(), conns.end(), canonical_less);
            t0 = now_ms();
            auto spectral = spectral_clustering(conns, config.spectral_dims);
            fprintf(meta, "directed_connections=%zu\nspectral_clusters=%zu\nspectral_ms=%.3f\n", conns.size(), spectral.size(), now_ms() - t0);
            std::vector<Component> nonempty;
            for (auto &c : spectral) if (!c.empty()) nonempty.push_back(c);
            merge_components(nonempty);
            std::vector<ComponentID> final_ids;
            for (auto p : component_index) if (p.second->size() >= (uint64_t) min_size) final_ids.push_back(p.first);
            fprintf(meta, "final_components=%zu\n", final_ids.size());
            std::vector<std::pair<uint32_t, ComponentID>> forder;
            for (auto id : final_ids) {
                auto &r = component_index[id]->contained_read_ids;
                forder.push_back({*std::min_element(r.begin(), r.end()), id});
            }
            std::sort(forder.begin(), forder.end());
            std::vector<uint32_t> final_id, final_read, cx, cy;
            std::vector<uint64_t> final_off{0}, cs;
            for (auto &o : forder) {
                final_id.push_back(o.second);
                auto reads = component_index[o.second]->contained_read_ids;
                std::sort(reads.begin(), reads.end());
                final_read.insert(final_read.end(), reads.begin(), reads.end());
                final_off.push_back(final_read.size());
            }
            for (auto &c : conns) { cx.push_back(c.component_x_id); cy.push_back(c.component_y_id); cs.push_back(c.score); }
            dump(out, "conn_x.u32", cx);
            dump(out, "conn_y.u32", cy);
            dump(out, "conn_score.u64", cs);
            dump(out, "final_id.u32", final_id);
            dump(out, "final_off.u64", final_off);
            dump(out, "final_read.u32", final_read);
            fclose(meta);
            return 0;
        }

        t0 = now_ms();
        std::vector<ComponentConnection> connections;
        if (sc_score > 0) {
            // --sc_score S (run_clustering :749-752): pivots = components with at least S discriminative k-mers, min score S
            std::vector<ComponentID> ids;
            for (auto p : component_index) if (p.second->discriminative_kmer_ids.size() >= sc_score) ids.push_back(p.first);
            connections = get_connections(ids, sc_score);
        } else {
            connections = get_all_connections(min_score);
        }
        double t_conn = now_ms() - t0;
        uint64_t score_sum = 0;
        for (auto &c : connections) score_sum += c.score;
        fprintf(meta, "directed_connections=%zu\nscore_sum=%lu\nconnections_ms=%.3f\n", connections.size(), (unsigned long) score_sum, t_conn);

        t0 = now_ms();
        std::sort(connections.begin(), connections.end(), canonical_less);
        double t_sort = now_ms() - t0;
        // clustering/ReadClusteringEngine.cpp:755
        size_t n = connections.size() * fraction;
        uint64_t cut_score = n > 0 ? connections[n - 1].score : 0;
        if (sc_score > 0) {
            // :752 filter_connections(score > S): a prefix of the sorted list
            n = 0;
            while (n < connections.size() && connections[n].score > sc_score) n++;
            cut_score = sc_score;
        }
        size_t above = 0, tied = 0;
        for (auto &c : connections) { if (c.score > cut_score) above++; else if (c.score == cut_score) tied++; }
        fprintf(meta, "cut_n=%zu\ncut_score=%lu\ndirected_above_cut=%zu\ndirected_tied_at_cut=%zu\ncanonical_sort_ms=%.3f\n", n,
                (unsigned long) cut_score, above, tied, t_sort);
        if (do_dump) {
            std::vector<uint32_t> cx, cy;
            std::vector<uint64_t> cs;
            for (auto &c : connections) { cx.push_back(c.component_x_id); cy.push_back(c.component_y_id); cs.push_back(c.score); }
            dump(out, "conn_x.u32", cx);
            dump(out, "conn_y.u32", cy);
            dump(out, "conn_score.u64", cs);
        }
        if (stop_after == 2) { fclose(meta); return 0; }

        std::vector<ComponentConnection> selected(connections.begin(), connections.begin() + n);
        std::set<ComponentID> restricted;
        t0 = now_ms();
        auto comps = union_find(selected, restricted, min_size, config.scaffold_component_max_size);
        double t_uf = now_ms() - t0;
        fprintf(meta, "scaffold_components=%zu\nunion_find_ms=%.3f\n", comps.size(), t_uf);
        if (do_dump) {
            // components sorted by their smallest member; members sorted; root (= element [0], .cpp:366) kept apart
            std::vector<std::pair<uint32_t, size_t>> order;
            for (size_t i = 0; i < comps.size(); i++) order.push_back({*std::min_element(comps[i].first.begin(), comps[i].first.end()), i});
            std::sort(order.begin(), order.end());
            std::vector<uint64_t> comp_off{0}, tree_off{0};
            std::vector<uint32_t> comp_read, comp_root, tree_x, tree_y;
            for (auto &o : order) {
                auto members = comps[o.second].first;
                comp_root.push_back(members[0]);
                std::sort(members.begin(), members.end());
                for (auto m : members) comp_read.push_back(m);
                comp_off.push_back(comp_read.size());
                std::vector<std::pair<uint32_t, uint32_t>> edges;
                for (auto e : comps[o.second].second) edges.push_back({std::min(e.first, e.second), std::max(e.first, e.second)});
                std::sort(edges.begin(), edges.end());
                for (auto &e : edges) { tree_x.push_back(e.first); tree_y.push_back(e.second); }
                tree_off.push_back(tree_x.size());
            }
            dump(out, "comp_off.u64", comp_off);
            dump(out, "comp_read.u32", comp_read);
            dump(out, "comp_root.u32", comp_root);
            dump(out, "tree_off.u64", tree_off);
            dump(out, "tree_x.u32", tree_x);
            dump(out, "tree_y.u32", tree_y);
        }
        if (enrich_min == 0) { fclose(meta); return 0; }

        // ---- merge of the scaffold components + enrichment, all real reference code --------------------
        // run_clustering :764 (merge_components), :785-794 (get_connections over the cores, restricted union_find,
        // merge_components). The tail / spectral block (:768-777) is skipped, which is the reference's own path when it finds
        // no strong tail connections (:771) or at most two scaffold components (:768).
        std::vector<Component> scaffold_components;
        for (auto &p : comps) scaffold_components.push_back(p.first);
        t0 = now_ms();
        auto scaffold_ids = merge_components(scaffold_components);
        double t_merge = now_ms() - t0;
        if (full && scaffold_ids.size() > 2) {
            // ---- SURVEY §8f-2, run_clustering :768-777, all real reference code (with the Eigen2 stand-in underneath): tails of the
            // spanning trees, tail amplification, tail connections, spectral clustering of the scaffold components, merge.
            // The driver imposes the canonical order on the tail connections before the spectral stage (their order decides the
            // row order of the affinity matrix, :657-663).
            if (do_dump) {
                // per scaffold component (in the order of comp_off: by smallest member): the tail vertices of its spanning tree
                // (:505-581), left then right, and the amplified tails (:583-592); all sorted
                std::vector<std::pair<uint32_t, size_t>> order;
                for (size_t i = 0; i < comps.size(); i++) order.push_back({*std::min_element(comps[i].first.begin(), comps[i].first.end()), i});
                std::sort(order.begin(), order.end());
                std::vector<uint64_t> tail_off{0}, amp_off{0};
                std::vector<uint32_t> tail_v, amp_v;
                for (auto &o : order) {
                    auto tails = get_spanning_tree_tails(comps[o.second].second);
                    for (auto *side : {&tails.first, &tails.second}) {
                        std::vector<uint32_t> v(side->begin(), side->end());
                        std::sort(v.begin(), v.end());
                        tail_v.insert(tail_v.end(), v.begin(), v.end());
                        tail_off.push_back(tail_v.size());
                        auto amp = amplify_component(*side, config.tail_amplification_min_score);
                        amp_v.insert(amp_v.end(), amp.begin(), amp.end());
                        amp_off.push_back(amp_v.size());
                    }
                }
                dump(out, "tail_off.u64", tail_off);
                dump(out, "tail_vertex.u32", tail_v);
                dump(out, "amp_off.u64", amp_off);
                dump(out, "amp_vertex.u32", amp_v);
            }
            t0 = now_ms();
            auto core_forming = get_core_component_connections(comps);
            double t_tail = now_ms() - t0;
            std::sort(core_forming.begin(), core_forming.end(), canonical_less);
            auto strong = filter_connections(core_forming, [](ComponentConnection &conn) { return conn.score > 5; });
            fprintf(meta, "tail_connections=%zu\nstrong_tail_connections=%zu\ntail_connections_ms=%.3f\n", core_forming.size(), strong.size(), t_tail);
            if (do_dump) {
                std::vector<uint32_t> cx, cy;
                std::vector<uint64_t> cs;
                for (auto &c : core_forming) { cx.push_back(c.component_x_id); cy.push_back(c.component_y_id); cs.push_back(c.score); }
                dump(out, "tconn_x.u32", cx);
                dump(out, "tconn_y.u32", cy);
                dump(out, "tconn_score.u64", cs);
            }
            if (!strong.empty()) {
                t0 = now_ms();
                auto spectral = spectral_clustering(strong, config.spectral_dims);
                double t_spec = now_ms() - t0;
                fprintf(meta, "spectral_clusters=%zu\nspectral_ms=%.3f\n", spectral.size(), t_spec);
                if (do_dump) {
                    // clusters ordered by their smallest component id; members sorted
                    std::vector<std::pair<std::vector<uint32_t>, uint32_t>> cl;      // (sorted members, element [0] = survivor of the merge)
                    for (auto &c : spectral) { if (c.empty()) continue; std::vector<uint32_t> m(c.begin(), c.end()); std::sort(m.begin(), m.end()); cl.push_back({m, c[0]}); }
                    std::sort(cl.begin(), cl.end());
                    std::vector<uint64_t> off{0};
                    std::vector<uint32_t> mem, first;
                    for (auto &c : cl) { mem.insert(mem.end(), c.first.begin(), c.first.end()); off.push_back(mem.size()); first.push_back(c.second); }
                    dump(out, "spectral_off.u64", off);
                    dump(out, "spectral_member.u32", mem);
                    dump(out, "spectral_first.u32", first);
                }
                // empty clusters (a rotated dimension that no point prefers) would make merge_components read element [0] of an
                // empty vector (:362-366); the reference has the same hazard, the driver skips them
                std::vector<Component> nonempty;
                for (auto &c : spectral) if (!c.empty()) nonempty.push_back(c);
                merge_components(nonempty);
            }
            for (auto it = component_index.begin(); it != component_index.end();) {       // remove_merged_components (:718-725), without the free()
                if (it->second->size() == 0) it = component_index.erase(it); else ++it;
            }
        }
        std::vector<ComponentID> core_ids;
        for (auto p : component_index) if (p.second->size() >= (uint64_t) min_size) core_ids.push_back(p.first);
        std::sort(core_ids.begin(), core_ids.end());
        fprintf(meta, "merged_scaffolds=%zu\ncores=%zu\nmerge_ms=%.3f\n", scaffold_ids.size(), core_ids.size(), t_merge);
        if (do_dump) {
            // merged cores: survivor id, k-mer VALUES of the merged list (sorted), member reads (sorted)
            std::vector<uint32_t> core_id, core_read;
            std::vector<uint64_t> core_kmer_off{0}, core_kmer, core_read_off{0};
            for (auto id : core_ids) {
                core_id.push_back(id);
                std::vector<uint64_t> vals;
                for (KmerID kid : component_index[id]->discriminative_kmer_ids) vals.push_back(id2kmer[kid]);
                std::sort(vals.begin(), vals.end());
                core_kmer.insert(core_kmer.end(), vals.begin(), vals.end());
                core_kmer_off.push_back(core_kmer.size());
                auto reads = component_index[id]->contained_read_ids;
                std::sort(reads.begin(), reads.end());
                core_read.insert(core_read.end(), reads.begin(), reads.end());
                core_read_off.push_back(core_read.size());
            }
            dump(out, "core_id.u32", core_id);
            dump(out, "core_kmer_off.u64", core_kmer_off);
            dump(out, "core_kmer.u64", core_kmer);
            dump(out, "core_read_off.u64", core_read_off);
            dump(out, "core_read.u32", core_read);
            // the purged inverted index, ordered by k-mer value (the reference's purge loop, :395-419)
            std::vector<std::pair<uint64_t, KmerID>> order;
            for (KmerID i = 0; i < id2kmer.size(); i++) order.push_back({id2kmer[i], i});
            std::sort(order.begin(), order.end());
            std::vector<uint64_t> inv_off{0};
            std::vector<uint32_t> inv_read;
            for (auto &o : order) {
                for (auto r : kmer_component_index[o.second]) inv_read.push_back(r);
                inv_off.push_back(inv_read.size());
            }
            dump(out, "purged_off.u64", inv_off);
            dump(out, "purged_read.u32", inv_read);
        }

        t0 = now_ms();
        auto econn = get_connections(core_ids, enrich_min);
        double t_econn = now_ms() - t0;
        std::sort(econn.begin(), econn.end(), canonical_less);
        fprintf(meta, "enrichment_connections=%zu\nenrichment_connections_ms=%.3f\n", econn.size(), t_econn);
        if (do_dump) {
            std::vector<uint32_t> cx, cy;
            std::vector<uint64_t> cs;
            for (auto &c : econn) { cx.push_back(c.component_x_id); cy.push_back(c.component_y_id); cs.push_back(c.score); }
            dump(out, "econn_x.u32", cx);
            dump(out, "econn_y.u32", cy);
            dump(out, "econn_score.u64", cs);
        }
        std::set<ComponentID> core_set(core_ids.begin(), core_ids.end());
        auto enriched = union_find(econn, core_set, 2, -1);
        std::vector<Component> enriched_components;
        for (auto &p : enriched) enriched_components.push_back(p.first);
        merge_components(enriched_components);
        std::vector<ComponentID> final_ids;
        for (auto p : component_index) if (p.second->size() >= (uint64_t) min_size) final_ids.push_back(p.first);
        std::sort(final_ids.begin(), final_ids.end());
        fprintf(meta, "final_components=%zu\n", final_ids.size());
        if (do_dump) {
            // final components ordered by their smallest member
            std::vector<std::pair<uint32_t, ComponentID>> forder;
            for (auto id : final_ids) {
                auto &r = component_index[id]->contained_read_ids;
                forder.push_back({*std::min_element(r.begin(), r.end()), id});
            }
            std::sort(forder.begin(), forder.end());
            std::vector<uint32_t> final_id, final_read;
            std::vector<uint64_t> final_off{0};
            for (auto &o : forder) {
                final_id.push_back(o.second);
                auto reads = component_index[o.second]->contained_read_ids;
                std::sort(reads.begin(), reads.end());
                final_read.insert(final_read.end(), reads.begin(), reads.end());
                final_off.push_back(final_read.size());
            }
            dump(out, "final_id.u32", final_id);
            dump(out, "final_off.u64", final_off);
            dump(out, "final_read.u32", final_read);
        }
        fclose(meta);
        return 0;
    }
};

int usage() {
    fprintf(stderr,
            "ref_driver kmeriter <k> <sequence>\n"
            "ref_driver canon <kmer file> <out dir>\n"
            "ref_driver records <out dir> <reads...>\n"
            "ref_driver run --kmers F --out DIR [--threads T] [--fraction 0.15] [--min-size 30] [--min-score 1]\n"
            "               [--no-dump] [--stop-after 1|2] [--enrich MIN_SCORE] [--sc-score S] [--full] [--force-spectral] [--spectral-dims D] [--max-size N] <reads...>\n");
    return 2;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 2) return usage();
    std::string mode = argv[1];
    if (mode == "kmeriter") {
        if (argc != 4) return usage();
        int k = atoi(argv[2]);
        std::string seq = argv[3];
        KmerIterator it(seq, k);
        while (it.next_kmer()) printf("%lu %lu\n", it.position_in_sequence, (unsigned long) it.current_kmer);
        return 0;
    }
    if (mode == "canon") {
        if (argc != 4) return usage();
        auto p = load_text_file_kmers(argv[2]);
        std::vector<uint64_t> v(p.first.begin(), p.first.end());
        std::sort(v.begin(), v.end());
        dump(argv[3], "canon_kmers.u64", v);
        printf("k=%d\nn_kmers=%zu\n", p.second, v.size());
        return 0;
    }
    if (mode == "records") {
        // record stream exactly as SequenceRecordIterator yields it: id, header, sequence, qualities
        if (argc < 4) return usage();
        std::vector<std::string> paths;
        for (int i = 3; i < argc; i++) paths.push_back(argv[i]);
        SequenceRecordIterator reader(paths, false);
        reader.show_progress = false;
        std::ofstream out(std::string(argv[2]) + "/records.txt");
        for (auto m : reader.file_meta) out << "#META " << m.filename << " " << m.records << " " << m.total_bases << " " << m.avg_read_length << " "
                                            << m.max_read_length << " " << m.min_read_length << "\n";
        out << "#AGG " << reader.meta.filename << " " << reader.meta.records << " " << reader.meta.total_bases << " " << reader.meta.avg_read_length
            << " " << reader.meta.max_read_length << " " << reader.meta.min_read_length << "\n";
        reader.rewind();
        std::optional<GenomeReadData> r;
        while ((r = reader.get_next_record()) != std::nullopt) {
            out << r->id << "\t" << r->header << "\t" << r->sequence << "\t" << r->qualities << "\n";
        }
        return 0;
    }
    if (mode != "run") return usage();

    std::string kmer_path, out;
    std::vector<std::string> paths;
    ReadClusteringConfig config;
    bool do_dump = true;
    double fraction = config.scaffold_forming_fraction;
    int min_size = config.scaffold_component_min_size;
    ConnectionScore min_score = 1;
    int stop_after = 0;
    ConnectionScore enrich_min = 0, sc_score = 0;
    bool full = false, force_spectral = false;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) exit(usage()); return argv[++i]; };
        if (a == "--kmers") kmer_path = next();
        else if (a == "--out") out = next();
        else if (a == "--threads") config.threads = std::stoi(next());
        else if (a == "--fraction") fraction = std::stod(next());
        else if (a == "--min-size") min_size = std::stoi(next());
        else if (a == "--min-score") min_score = std::stoul(next());
        else if (a == "--stop-after") stop_after = std::stoi(next());
        else if (a == "--enrich") enrich_min = std::stoul(next());
        else if (a == "--sc-score") sc_score = std::stoul(next());
        else if (a == "--full") full = true;
        else if (a == "--force-spectral") force_spectral = true;
        else if (a == "--spectral-dims") config.spectral_dims = std::stoi(next());
        else if (a == "--max-size") config.scaffold_component_max_size = std::stoi(next());
        else if (a == "--no-dump") do_dump = false;
        else paths.push_back(a);
    }
    if (kmer_path.empty() || out.empty() || paths.empty()) return usage();

    double t0 = now_ms();
    auto kk = load_text_file_kmers(kmer_path);
    double t_load = now_ms() - t0;
    t0 = now_ms();
    SequenceRecordIterator reader(paths, false);
    reader.show_progress = false;
    double t_meta = now_ms() - t0;
    Probe engine(reader, config);
    int rc = engine.run(kk.first, kk.second, out, do_dump, fraction, min_size, min_score, stop_after, enrich_min, sc_score, full, force_spectral);
    FILE *meta = fopen((out + "/meta.txt").c_str(), "a");
    fprintf(meta, "kmer_load_ms=%.3f\nmeta_pass_ms=%.3f\nthreads=%d\n", t_load, t_meta, config.threads);
    fclose(meta);
    return rc;
}
