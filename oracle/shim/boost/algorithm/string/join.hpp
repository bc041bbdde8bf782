// Oracle build shim (test infrastructure only): boost::algorithm::join over a container of std::string.
#pragma once
#include <string>
namespace boost { namespace algorithm {
template<typename C>
std::string join(const C &parts, const std::string &sep) {
    std::string out;
    bool first = true;
    for (const auto &p : parts) {
        if (!first) out += sep;
        out += p;
        first = false;
    }
    return out;
}
}}
