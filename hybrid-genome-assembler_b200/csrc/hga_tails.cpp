// Tail connections between scaffold components (HOST code; SURVEY.md §8f-2, second piece).
//
// Replaces ReadClusteringEngine::get_core_component_connections (clustering/ReadClusteringEngine.cpp:594-651) and what it calls:
// approximate_read_overlap (:491-508), get_spanning_tree_tails (:510-581), amplify_component (:583-592), accumulate_kmer_ids
// (:340-346). It runs on the engine state AFTER merge_components(scaffold_components) (:764), which is what the caller hands over:
//   * hits by read, every row sorted by (kmer_id, pos)   (hga_get_hits(h, 1, ..): discriminative_kmer_ids with duplicates and
//     kmer_positions = first occurrence per k-mer);
//   * the scaffold components, element [0] = the surviving id, and their spanning trees (the edges that performed a union);
//   * the PURGED inverted index (hga_get_purged_index).
// The k-mer list of a component id is, as in the reference after the merge: the sorted UNIQUE union of all members' lists for a
// survivor (:379), the read's own sorted list with duplicates for every other id (merged-away members keep theirs).
//
// Per component: the two ends of the spanning tree by a double sweep (farthest vertex from an arbitrary start, farthest from
// that, farthest from that; distance = read lengths minus approximate overlaps along the tree path), the vertices within two
// average read lengths of each end (the "tails"), every tail amplified by the reads it reaches with at least
// `amplification_min_score` shared k-mers, the unique k-mer union of every amplified tail; per pair of components the largest of
// the four tail-to-tail intersections. The reference starts the first sweep at adjacency_map.begin() (whatever its hash map puts
// first). That start matters: the survivor's list is the merged union while its positions are its own read's, so an "overlap" on
// a tree edge at the survivor can exceed the read lengths, the uint64 distances wrap for some start vertices and not for others,
// and the tails (even whether they are empty: a wrapped maximum plus the tail length wraps again and nothing is "within" it)
// differ. Canonical choice, imposed on the reference too (oracle/shim/tsl/robin_map.h orders that one map): the first sweep starts
// at the SMALLEST vertex id; distance ties go to the smallest id. The arithmetic is the reference's, wrap-around included.
//
// Small host arithmetic (a few hundred tree vertices per component): sequential in the reference, sequential here - except the
// amplification (get_connections(tail, min) through the purged index, :583-592), which is the one data-parallel step: inside
// hga_enrich_full it runs on the GPU (hga_enrich.cu hands hga_host_tail_connections_impl a callback that counts the partners of ALL
// tail vertices of all components in one go, with the enrichment's emit / sort / run-length kernels); the plain C entry point walks
// the purged lists on the host.
#include <algorithm>
#include <cstdint>
#include <map>
#include <queue>
#include <unordered_map>
#include <vector>

#include "../../include/hga_b200.h"

void hga_set_error(const char *fmt, ...);

namespace {

struct TailState {
    uint64_t n_reads;
    uint32_t first_id;
    const uint64_t *row_off;
    const uint32_t *kid, *pos, *read_len;
    const uint64_t *purged_off;
    const uint32_t *purged_read;
    std::unordered_map<uint32_t, std::vector<uint32_t>> survivor_list;   // survivor id -> unique union

    // discriminative_kmer_ids of a component id in the current engine state
    void list_of(uint32_t id, const uint32_t *&p, size_t &n) const {
        auto it = survivor_list.find(id);
        if (it != survivor_list.end()) { p = it->second.data(); n = it->second.size(); return; }
        const uint64_t r = id - first_id;
        p = kid + row_off[r]; n = (size_t) (row_off[r + 1] - row_off[r]);
    }
    // read_metas[id].kmer_positions[k]: first occurrence in READ id; operator[] default-inserts 0 for a k-mer the read does not hold
    uint32_t first_pos(uint32_t id, uint32_t k) const {
        const uint64_t r = id - first_id, a = row_off[r], b = row_off[r + 1];
        const uint32_t *lo = std::lower_bound(kid + a, kid + b, k);
        return (lo != kid + b && *lo == k) ? pos[lo - kid] : 0u;
    }
    // :491-508
    int overlap(uint32_t x, uint32_t y) const {
        const uint32_t *px, *py;
        size_t nx, ny;
        list_of(x, px, nx); list_of(y, py, ny);
        int max_x = 0, max_y = 0, min_x = 0, min_y = 0;
        bool any = false;
        size_t i = 0, j = 0;
        while (i < nx && j < ny) {                           // get_vectors_intersection (Utils.h:125-141): one for one
            if (px[i] < py[j]) i++;
            else if (py[j] < px[i]) j++;
            else {
                const int a = (int) first_pos(x, px[i]), b = (int) first_pos(y, px[i]);
                if (!any) { max_x = min_x = a; max_y = min_y = b; any = true; }
                else { max_x = std::max(max_x, a); min_x = std::min(min_x, a); max_y = std::max(max_y, b); min_y = std::min(min_y, b); }
                i++; j++;
            }
        }
        return std::max(max_x - min_x, max_y - min_y);
    }
};

}  // namespace

// amplification callback: for every pivot id (n_pivots of them, in tail order) the ids it reaches through the purged index with at least
// min_score shared k-mers (duplicates of the pivot's list count, the pivot itself excluded): appends (pivot index, partner id) to out
typedef int (*hga_tail_amplify_fn)(void *ctx, const uint32_t *pivot_id, uint64_t n_pivots, uint32_t min_score, std::vector<std::pair<uint32_t, uint32_t>> &out);

int hga_host_tail_connections_impl(uint64_t n_reads, const uint64_t *row_off, const uint32_t *kmer_id, const uint32_t *pos, const uint32_t *read_len,
                                   uint64_t avg_read_length, uint32_t read_id_first, uint64_t n_comp, const uint64_t *comp_off,
                                   const uint32_t *comp_member, const uint64_t *tree_off, const uint32_t *tree_x, const uint32_t *tree_y,
                                   const uint64_t *purged_off, const uint32_t *purged_read, uint32_t amplification_min_score, uint32_t *out_x,
                                   uint32_t *out_y, uint64_t *out_score, uint64_t *out_n, hga_tail_amplify_fn amplify_fn, void *amplify_ctx) {
    if (!row_off || !comp_off || !tree_off || (!purged_off && !amplify_fn) || !out_n || (n_comp > 1 && (!out_x || !out_y || !out_score))) {
        hga_set_error("hga_host_tail_connections: NULL argument");
        return HGA_E_ARG;
    }
    *out_n = 0;
    TailState S{n_reads, read_id_first, row_off, kmer_id, pos, read_len, purged_off, purged_read, {}};
    // the survivors' lists after merge_components: sorted unique union of the members' lists (:379)
    for (uint64_t c = 0; c < n_comp; c++) {
        std::vector<uint32_t> u;
        for (uint64_t i = comp_off[c]; i < comp_off[c + 1]; i++) {
            const uint64_t r = comp_member[i] - read_id_first;
            u.insert(u.end(), kmer_id + row_off[r], kmer_id + row_off[r + 1]);
        }
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
        S.survivor_list.emplace(comp_member[comp_off[c]], std::move(u));
    }
    const uint64_t tail_length = avg_read_length * 2;                       // :543

    std::vector<std::pair<uint32_t, std::pair<std::vector<uint32_t>, std::vector<uint32_t>>>> comp_tails;   // (survivor, (tail vertices of one end, of the other)), components ascending
    for (uint64_t c = 0; c < n_comp; c++) {
        // adjacency with the approximate overlap as edge attribute (:511-516)
        std::map<uint32_t, std::vector<std::pair<uint32_t, int>>> adj;
        for (uint64_t e = tree_off[c]; e < tree_off[c + 1]; e++) {
            const int d = S.overlap(tree_x[e], tree_y[e]);
            adj[tree_x[e]].push_back({tree_y[e], d});
            adj[tree_y[e]].push_back({tree_x[e], d});
        }
        if (adj.empty()) continue;
        auto bfs = [&](uint32_t start) {                                    // :518-538 (uint64 arithmetic as in the reference)
            std::map<uint32_t, uint64_t> dist;
            std::map<uint32_t, bool> visited;
            std::queue<uint32_t> q;
            dist[start] = read_len[start - read_id_first];
            q.push(start);
            while (!q.empty()) {
                const uint32_t v = q.front();
                visited[v] = true;
                q.pop();
                for (const auto &nb : adj[v])
                    if (!visited.count(nb.first)) {
                        dist[nb.first] = dist[v] + read_len[nb.first - read_id_first] - (uint64_t) (int64_t) nb.second;
                        q.push(nb.first);
                    }
            }
            return dist;
        };
        auto farthest = [](const std::map<uint32_t, uint64_t> &dist) {      // :540-544; ties: smallest id
            std::pair<uint32_t, uint64_t> best = *dist.begin();
            for (const auto &d : dist) if (d.second > best.second) best = d;
            return best;
        };
        const auto d0 = bfs(adj.begin()->first);
        const auto far0 = farthest(d0);
        const auto d_right = bfs(far0.first);
        const auto far_right = farthest(d_right);
        std::vector<uint32_t> end_a, end_b;
        for (const auto &d : d_right) if (d.second + tail_length > far_right.second) end_a.push_back(d.first);
        const auto d_left = bfs(far_right.first);
        const auto far_left = farthest(d_left);
        for (const auto &d : d_left) if (d.second + tail_length > far_left.second) end_b.push_back(d.first);

        comp_tails.push_back({comp_member[comp_off[c]], {std::move(end_b), std::move(end_a)}});
    }

    // amplify_component (:583-592) for ALL tails at once: the tail plus every id a tail vertex reaches with score >= min
    std::vector<uint32_t> pivot_id;                                         // the tail vertices, tail by tail
    std::vector<uint64_t> tail_first;                                       // tail t = pivots [tail_first[t], tail_first[t + 1])
    for (const auto &ct : comp_tails)
        for (const std::vector<uint32_t> *tl : {&ct.second.first, &ct.second.second}) {
            tail_first.push_back(pivot_id.size());
            pivot_id.insert(pivot_id.end(), tl->begin(), tl->end());
        }
    tail_first.push_back(pivot_id.size());
    std::vector<std::pair<uint32_t, uint32_t>> reached;                    // (pivot index, partner id)
    if (amplify_fn) {
        const int rc = amplify_fn(amplify_ctx, pivot_id.data(), pivot_id.size(), amplification_min_score, reached);
        if (rc != HGA_OK) return rc;
    } else {
        std::unordered_map<uint32_t, uint64_t> count;
        for (size_t pi = 0; pi < pivot_id.size(); pi++) {
            const uint32_t pivot = pivot_id[pi];
            count.clear();
            const uint32_t *pl;
            size_t nl;
            S.list_of(pivot, pl, nl);
            for (size_t i = 0; i < nl; i++)                                 // duplicates of the pivot's list count (:311-316)
                for (uint64_t j = purged_off[pl[i]]; j < purged_off[pl[i] + 1]; j++) count[purged_read[j]]++;
            count.erase(pivot);                                              // :317
            for (const auto &cn : count) if (cn.second >= amplification_min_score) reached.push_back({(uint32_t) pi, cn.first});
        }
    }
    std::vector<std::vector<uint32_t>> amplified(tail_first.size() - 1);
    for (size_t t = 0; t + 1 < tail_first.size(); t++) amplified[t].assign(pivot_id.begin() + (ptrdiff_t) tail_first[t], pivot_id.begin() + (ptrdiff_t) tail_first[t + 1]);
    for (const auto &pr : reached) {
        const size_t t = (size_t) (std::upper_bound(tail_first.begin(), tail_first.end(), (uint64_t) pr.first) - tail_first.begin()) - 1;
        amplified[t].push_back(pr.second);
    }
    // accumulate_kmer_ids (:340-346): sorted unique union of the components' current lists
    auto kmers_of = [&](std::vector<uint32_t> &ids) {
        std::sort(ids.begin(), ids.end());
        ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        std::vector<uint32_t> u;
        for (uint32_t id : ids) { const uint32_t *pl; size_t nl; S.list_of(id, pl, nl); u.insert(u.end(), pl, pl + nl); }
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
        return u;
    };
    std::map<uint32_t, std::pair<std::vector<uint32_t>, std::vector<uint32_t>>> tails;   // survivor -> (k-mers of one end, of the other)
    for (size_t i = 0; i < comp_tails.size(); i++) tails[comp_tails[i].first] = {kmers_of(amplified[2 * i]), kmers_of(amplified[2 * i + 1])};

    // :621-640: for every pair of components the largest of the four tail-to-tail intersections; :650 keep score > 0. The reference
    // intersects the sorted unions pair by pair (S^2 / 2 pairs x 4 merges); the unions are sets, so |A n B| is the number of
    // k-mers that list both tails: one pass over a k-mer -> tails posting list counts all pairs at once.
    struct Conn { uint32_t x, y; uint64_t s; };
    std::vector<Conn> conns;
    {
        std::vector<uint32_t> comp_id;                                     // tail t belongs to component comp_id[t / 2]
        std::vector<std::pair<uint32_t, uint32_t>> posting;                // (k-mer, tail index)
        uint32_t t = 0;
        for (const auto &tc : tails) {
            comp_id.push_back(tc.first);
            for (uint32_t k : tc.second.first) posting.push_back({k, t});
            for (uint32_t k : tc.second.second) posting.push_back({k, t + 1});
            t += 2;
        }
        std::sort(posting.begin(), posting.end());
        std::unordered_map<uint64_t, uint64_t> shared;                     // (tail a << 32 | tail b), a < b, different components
        for (size_t i = 0; i < posting.size();) {
            size_t j = i;
            while (j < posting.size() && posting[j].first == posting[i].first) j++;
            for (size_t p = i; p < j; p++)
                for (size_t q = p + 1; q < j; q++)
                    if (posting[p].second / 2 != posting[q].second / 2) shared[((uint64_t) posting[p].second << 32) | posting[q].second]++;
            i = j;
        }
        std::unordered_map<uint64_t, uint64_t> best;                       // (component index a << 32 | b) -> max over the four combinations
        for (const auto &e : shared) {
            const uint64_t key = ((e.first >> 33) << 32) | ((e.first & 0xFFFFFFFFu) >> 1);
            uint64_t &m = best[key];
            m = std::max(m, e.second);
        }
        for (const auto &e : best) conns.push_back({comp_id[e.first >> 32], comp_id[e.first & 0xFFFFFFFFu], e.second});
    }
    // canonical order (score desc, min asc, max asc); x < y by construction
    std::sort(conns.begin(), conns.end(), [](const Conn &p, const Conn &q) { return p.s != q.s ? p.s > q.s : (p.x != q.x ? p.x < q.x : p.y < q.y); });
    for (size_t i = 0; i < conns.size(); i++) { out_x[i] = conns[i].x; out_y[i] = conns[i].y; out_score[i] = conns[i].s; }
    *out_n = conns.size();
    return HGA_OK;
}

extern "C" int hga_host_tail_connections(uint64_t n_reads, const uint64_t *row_off, const uint32_t *kmer_id, const uint32_t *pos, const uint32_t *read_len,
                                         uint64_t avg_read_length, uint32_t read_id_first, uint64_t n_comp, const uint64_t *comp_off,
                                         const uint32_t *comp_member, const uint64_t *tree_off, const uint32_t *tree_x, const uint32_t *tree_y,
                                         const uint64_t *purged_off, const uint32_t *purged_read, uint32_t amplification_min_score, uint32_t *out_x,
                                         uint32_t *out_y, uint64_t *out_score, uint64_t *out_n) {
    if (!purged_off) { hga_set_error("hga_host_tail_connections: NULL argument"); return HGA_E_ARG; }
    return hga_host_tail_connections_impl(n_reads, row_off, kmer_id, pos, read_len, avg_read_length, read_id_first, n_comp, comp_off, comp_member, tree_off, tree_x, tree_y,
                                          purged_off, purged_read, amplification_min_score, out_x, out_y, out_score, out_n, nullptr, nullptr);
}
