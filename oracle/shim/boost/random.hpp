/* oracle/shim/boost/random.hpp -- TEST INFRASTRUCTURE, not product code.
 *
 * The reference takes its random-number stream from Boost.Random
 * (/root/reference/src/rand.hpp:26,45-51,87-89,174-186).  Boost is neither
 * vendored nor pinned by the reference ("use Boost", README.markdown:39-40) and
 * is absent from this image, so this header restates, from the published
 * Boost.Random 1.47-1.55 sources (the releases contemporary with BMAGWA v2.0,
 * 2012), exactly the five class templates the reference instantiates:
 *
 *   mt19937              32-bit Mersenne Twister, identical to std::mt19937
 *   uniform_01<double>   (eng() - eng.min()) * 2^-32, redrawn if it rounds to 1
 *   normal_distribution  Box-Muller with the second variate cached
 *                        (r1, r2 uniforms; rho = sqrt(-2 log(1 - r2));
 *                        first call rho*cos(2 pi r1), second rho*sin(2 pi r1))
 *   gamma_distribution   alpha == 1: exponential; alpha > 1: the tan-rejection
 *                        method (Numerical Recipes gamdev / Knuth 3.4.1);
 *                        alpha < 1: the rejection method with p = e/(alpha+e)
 *   variate_generator    owns a COPY of the distribution, engine by reference
 *
 * PARITY UNPINNED: later Boost releases switched normal_distribution to a
 * ziggurat; the reference's own tests at this boundary check moments only
 * (src/tests/rand_tests.hpp:30-176), so draw-level equality with an upstream
 * binary cannot be asserted.  What this shim guarantees is that the reference
 * build in oracle/_ref and the product's host sampler (which carries its own,
 * independently written copy of these algorithms) consume one and the same
 * stream.
 */
#ifndef BMAGWA_ORACLE_SHIM_BOOST_RANDOM_HPP
#define BMAGWA_ORACLE_SHIM_BOOST_RANDOM_HPP

#include <cmath>
#include <random>
#include <stdint.h>

namespace boost {

typedef std::mt19937 mt19937;

template <class RealType = double>
class uniform_01
{
  public:
    typedef RealType result_type;
    uniform_01() {}
    template <class Engine>
    result_type operator()(Engine& eng)
    {
      for (;;) {
        result_type r = static_cast<result_type>(eng() - (eng.min)()) * (1.0 / 4294967296.0);
        if (r < result_type(1)) return r;
      }
    }
};

template <class RealType = double>
class exponential_distribution
{
  public:
    explicit exponential_distribution(RealType lambda = RealType(1)) : lambda_(lambda) {}
    template <class Engine>
    RealType operator()(Engine& eng)
    {
      return -RealType(1) / lambda_ * std::log(RealType(1) - uniform_01<RealType>()(eng));
    }
  private:
    RealType lambda_;
};

template <class RealType = double>
class normal_distribution
{
  public:
    typedef RealType result_type;
    explicit normal_distribution(RealType mean = RealType(0), RealType sigma = RealType(1))
    : mean_(mean), sigma_(sigma), r1_(0), r2_(0), cached_rho_(0), valid_(false) {}
    template <class Engine>
    result_type operator()(Engine& eng)
    {
      if (!valid_) {
        r1_ = uniform_01<RealType>()(eng);
        r2_ = uniform_01<RealType>()(eng);
        cached_rho_ = std::sqrt(-result_type(2) * std::log(result_type(1) - r2_));
        valid_ = true;
      } else {
        valid_ = false;
      }
      const result_type pi = result_type(3.14159265358979323846);
      return cached_rho_ * (valid_ ? std::cos(result_type(2) * pi * r1_)
                                   : std::sin(result_type(2) * pi * r1_)) * sigma_ + mean_;
    }
  private:
    RealType mean_, sigma_, r1_, r2_, cached_rho_;
    bool valid_;
};

template <class RealType = double>
class gamma_distribution
{
  public:
    typedef RealType result_type;
    explicit gamma_distribution(RealType alpha = RealType(1), RealType beta = RealType(1))
    : exp_(RealType(1)), alpha_(alpha), beta_(beta)
    {
      p_ = std::exp(result_type(1)) / (alpha_ + std::exp(result_type(1)));
    }
    template <class Engine>
    result_type operator()(Engine& eng)
    {
      using std::tan; using std::sqrt; using std::exp; using std::log; using std::pow;
      if (alpha_ == result_type(1)) {
        return exp_(eng) * beta_;
      } else if (alpha_ > result_type(1)) {
        const result_type pi = result_type(3.14159265358979323846);
        for (;;) {
          result_type y = tan(pi * uniform_01<RealType>()(eng));
          result_type x = sqrt(result_type(2) * alpha_ - result_type(1)) * y
                          + alpha_ - result_type(1);
          if (x <= result_type(0)) continue;
          if (uniform_01<RealType>()(eng) >
              (result_type(1) + y * y) * exp((alpha_ - result_type(1))
                                             * log(x / (alpha_ - result_type(1)))
                                             - sqrt(result_type(2) * alpha_ - result_type(1)) * y))
            continue;
          return x * beta_;
        }
      } else {
        for (;;) {
          result_type u = uniform_01<RealType>()(eng);
          result_type y = exp_(eng);
          result_type x, q;
          if (u < p_) {
            x = exp(-y / alpha_);
            q = p_ * exp(-x);
          } else {
            x = result_type(1) + y;
            q = p_ + (result_type(1) - p_) * pow(x, alpha_ - result_type(1));
          }
          if (u >= q) continue;
          return x * beta_;
        }
      }
    }
  private:
    exponential_distribution<RealType> exp_;
    result_type alpha_, beta_, p_;
};

template <class Engine, class Distribution> class variate_generator;

template <class Engine, class Distribution>
class variate_generator<Engine&, Distribution>
{
  public:
    typedef typename Distribution::result_type result_type;
    variate_generator(Engine& e, Distribution d) : eng_(e), dist_(d) {}
    result_type operator()() { return dist_(eng_); }
  private:
    Engine& eng_;
    Distribution dist_;
};

} // namespace boost

#endif
