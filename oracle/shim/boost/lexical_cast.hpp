/* oracle/shim -- TEST INFRASTRUCTURE.  boost::lexical_cast<std::string>(size_t)
 * as used by /root/reference/src/options.hpp:119. */
#ifndef BMAGWA_ORACLE_SHIM_BOOST_LEXICAL_CAST_HPP
#define BMAGWA_ORACLE_SHIM_BOOST_LEXICAL_CAST_HPP
#include <sstream>
#include <string>
namespace boost {
template <class Target, class Source>
inline Target lexical_cast(const Source& v) { std::ostringstream o; o << v; return o.str(); }
}
#endif
