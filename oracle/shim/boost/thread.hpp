// Oracle build shim (test infrastructure only): the reference uses <boost/thread.hpp> solely to reach
// boost::posix_time wall-clock helpers in common/Utils.h:17-35 (the "<label> took <ms>ms" timers).
#pragma once
#include <chrono>
#include <optional>
namespace boost { namespace posix_time {
struct time_duration {
    std::chrono::steady_clock::duration d;
    long total_milliseconds() const { return (long) std::chrono::duration_cast<std::chrono::milliseconds>(d).count(); }
};
struct ptime {
    std::chrono::steady_clock::time_point t;
    time_duration operator-(const ptime &o) const { return {t - o.t}; }
};
struct microsec_clock {
    static ptime local_time() { return {std::chrono::steady_clock::now()}; }
};
}}
