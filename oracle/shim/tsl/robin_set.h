// Oracle build shim (test infrastructure only): tsl::robin_set -> std::unordered_set.
// Iteration order differs from the real robin-hood set; every comparison made with this oracle is
// order-independent (see oracle/README.md).
#pragma once
#include <unordered_set>
namespace tsl {
template<typename K> using robin_set = std::unordered_set<K>;
}
