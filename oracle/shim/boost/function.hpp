// Oracle build shim (test infrastructure only): occurrences/JellyfishOccurrenceReader.h includes <boost/function.hpp> but uses
// std::function only; the real header also brings <cstdint> and the stream headers in, which that file relies on.
#pragma once
#include <cstdint>
#include <fstream>
#include <functional>
#include <iostream>
