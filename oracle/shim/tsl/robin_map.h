// Oracle build shim (test infrastructure only): the subset of tsl::robin_map the reference touches
// (clustering/ReadClusteringEngine.{h,cpp}): insert -> pair<iterator,bool>, iterator.value(), operator[],
// contains, erase(key|iterator), range-for, initializer-list construction, and a value_type with a
// NON-const key (ConcurrentQueue<IndexRemovalMap::value_type> needs it assignable, .cpp:395).
// Backed by std::unordered_map, so iteration order is not robin-hood order; the oracle only exposes
// order-independent results.
// ONE instantiation is ordered instead: robin_map<ReadID, robin_map<ReadID, int>>, the adjacency map of get_spanning_tree_tails
// (clustering/ReadClusteringEngine.cpp:510). The reference starts its first distance sweep at adjacency_map.begin() (:545), i.e. at
// whatever vertex its hash map puts first, and with the survivor's merged k-mer list the "overlaps" along the tree can exceed the
// read lengths: the uint64 distances then wrap for some start vertices and not for others, and the tails (even whether they are
// empty) depend on that start. Like the tie order of the edge list, this is pinned to a canonical choice on both sides: the sweep
// starts at the SMALLEST vertex id (std::map here, hga_tails.cpp in the product).
#pragma once
#include <cstdint>
#include <map>
#include <unordered_map>
#include <initializer_list>
#include <utility>
namespace tsl {
template<typename K, typename V> class robin_map;
template<typename K, typename V> struct robin_map_storage { using type = std::unordered_map<K, V>; };
template<> struct robin_map_storage<uint32_t, robin_map<uint32_t, int>> { using type = std::map<uint32_t, robin_map<uint32_t, int>>; };
#ifdef HGA_SHIM_ORDERED_DISTANCES
// experiment switch (not used by the committed build): also order robin_map<uint32_t, uint64_t>, the type of the distance maps of the
// tail search (ties of std::max_element then go to the smallest id) - and of the hot counting map of get_connections, which is why
// it is not the default: it would slow the CPU baseline down
template<> struct robin_map_storage<uint32_t, uint64_t> { using type = std::map<uint32_t, uint64_t>; };
#endif

template<typename K, typename V>
class robin_map {
    using base_t = typename robin_map_storage<K, V>::type;
    base_t m;
public:
    using key_type = K;
    using mapped_type = V;
    using value_type = std::pair<K, V>;
    using size_type = std::size_t;

    template<typename It>
    class iter_t {
        It it;
        friend class robin_map;
    public:
        using iterator_category = std::forward_iterator_tag;
        using value_type = typename It::value_type;
        using difference_type = std::ptrdiff_t;
        using pointer = typename It::pointer;
        using reference = typename It::reference;
        iter_t() = default;
        iter_t(It i) : it(i) {}
        reference operator*() const { return *it; }
        pointer operator->() const { return it.operator->(); }
        iter_t &operator++() { ++it; return *this; }
        iter_t operator++(int) { iter_t t = *this; ++it; return t; }
        bool operator==(const iter_t &o) const { return it == o.it; }
        bool operator!=(const iter_t &o) const { return it != o.it; }
        const K &key() const { return it->first; }
        auto &value() const { return it->second; }
    };
    using iterator = iter_t<typename base_t::iterator>;
    using const_iterator = iter_t<typename base_t::const_iterator>;

    robin_map() = default;
    robin_map(std::initializer_list<value_type> il) { for (auto &p : il) m.insert({p.first, p.second}); }

    iterator begin() { return m.begin(); }
    iterator end() { return m.end(); }
    const_iterator begin() const { return m.begin(); }
    const_iterator end() const { return m.end(); }
    size_type size() const { return m.size(); }
    bool empty() const { return m.empty(); }

    std::pair<iterator, bool> insert(const value_type &p) {
        auto r = m.insert({p.first, p.second});
        return {iterator(r.first), r.second};
    }
    V &operator[](const K &k) { return m[k]; }
    bool contains(const K &k) const { return m.find(k) != m.end(); }
    size_type erase(const K &k) { return m.erase(k); }
    iterator erase(iterator pos) { return m.erase(pos.it); }
    iterator find(const K &k) { return m.find(k); }
    void clear() { m.clear(); }
};
}
