// Oracle build shim (test infrastructure only): boost::regex -> std::regex.
// Used by common/SequenceRecordIterator.h:96-103 for simulator header parsing (debug metadata only).
#pragma once
#include <regex>
namespace boost {
using std::regex;
using std::smatch;
using std::regex_search;
}
