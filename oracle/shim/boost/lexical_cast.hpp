// Minimal stand-in for <boost/lexical_cast.hpp> (Boost is not installed in this image).
// TEST INFRASTRUCTURE ONLY: lets oracle/Makefile compile the reference's contig.cpp in place.
// The reference only uses lexical_cast<string>(int | uint64_t | double)
// (/root/reference/DBG_contig/contig.cpp:1006,1021-1029). Boost formats doubles with
// max_digits10 (17) significant digits in general notation; the golden header
// "avgDepth: 22.271739130434781" (test/02.build_contig/*.small.fa:1) confirms that.
#pragma once
#include <sstream>
#include <string>
#include <limits>
#include <type_traits>

namespace boost {
template <typename Target, typename Source>
inline Target lexical_cast(const Source &v)
{
    static_assert(std::is_same<Target, std::string>::value, "shim supports string targets only");
    std::ostringstream os;
    if (std::is_floating_point<Source>::value)
        os.precision(std::numeric_limits<Source>::max_digits10);
    os << v;
    return os.str();
}
}  // namespace boost
