/* oracle/shim -- TEST INFRASTRUCTURE.  boost::split / is_any_of / trim as used
 * by /root/reference/src/options.hpp:236-239,298-301 (single-character
 * delimiter sets, whitespace trim, token compression off). */
#ifndef BMAGWA_ORACLE_SHIM_BOOST_ALGORITHM_STRING_HPP
#define BMAGWA_ORACLE_SHIM_BOOST_ALGORITHM_STRING_HPP
#include <cctype>
#include <string>
#include <vector>
namespace boost {
struct shim_any_of { std::string set; };
inline shim_any_of is_any_of(const char* s) { shim_any_of a; a.set = s; return a; }
inline void split(std::vector<std::string>& out, const std::string& in, const shim_any_of& pred)
{
  out.clear();
  std::string cur;
  for (size_t i = 0; i < in.size(); ++i) {
    if (pred.set.find(in[i]) != std::string::npos) { out.push_back(cur); cur.clear(); }
    else cur += in[i];
  }
  out.push_back(cur);
}
inline void trim(std::string& s)
{
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) ++a;
  while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
  s = s.substr(a, b - a);
}
}
#endif
