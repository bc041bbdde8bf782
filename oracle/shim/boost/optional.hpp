// Oracle build shim (test infrastructure only; NOT product code).
// Stands in for <boost/optional.hpp>, which the reference includes from common/KmerIterator.h:3 but never uses.
// The real header transitively provides <unordered_map>/<algorithm>, which common/KmerIterator.cpp relies on.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <climits>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <optional>
