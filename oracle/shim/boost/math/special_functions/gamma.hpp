/* oracle/shim -- TEST INFRASTRUCTURE.  boost::math::lgamma, the one Boost.Math
 * function the reference calls (/root/reference/src/utils.cpp:25,105), mapped
 * onto the C library's lgamma (both are accurate to a few ulp; the reference's
 * test at this boundary is an identity to 1e-6, src/tests/prior_tests.hpp). */
#ifndef BMAGWA_ORACLE_SHIM_BOOST_MATH_GAMMA_HPP
#define BMAGWA_ORACLE_SHIM_BOOST_MATH_GAMMA_HPP
#include <cmath>
namespace boost { namespace math {
inline double lgamma(double x) { int sign; return ::lgamma_r(x, &sign); }
}}
#endif
