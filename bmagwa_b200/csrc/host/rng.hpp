// rng.hpp -- the chain's random-number stream (host side).
//
// Mirrors bmagwa::Rand (src/rand.hpp:36-191): one 32-bit Mersenne Twister per chain feeding
// uniform, standard-normal, Gamma and scaled-inverse-chi^2 variates.  The reference takes the
// variate algorithms from Boost.Random (not vendored, version not pinned); the ones below are
// the Boost 1.47-1.55 era algorithms (the releases contemporary with BMAGWA 2.0):
//   uniform      u32 * 2^-32
//   normal       polar-free Box-Muller on (u1,u2), caching the sine branch
//   gamma(a)     a == 1: -log(1-u);  a > 1: tangent rejection (Numerical Recipes "gamdev");
//                a < 1: two-piece rejection with p = e / (a + e)
// Keeping the stream (not just the distribution) identical is what lets a fixed-seed chain of this
// sampler reproduce the accepted-move sequence of the reference build in oracle/_ref.
#pragma once
#include <cmath>
#include <cstdint>

namespace bmg {

class ChainRng {
 public:
  ChainRng(uint32_t seed, double sinvchi2_nu) : fixed_nu_(sinvchi2_nu)
  {
    state_[0] = seed;
    for (int i = 1; i < kN; ++i) state_[i] = 1812433253u * (state_[i - 1] ^ (state_[i - 1] >> 30)) + (uint32_t)i;
    pos_ = kN;
  }

  // uniform on [0,1) with 32 random bits
  double u01() { return (double)next_u32() * (1.0 / 4294967296.0); }

  double normal()
  {
    have_cached_ = !have_cached_;
    if (have_cached_) {
      angle_u_ = u01();
      const double u2 = u01();
      radius_ = std::sqrt(-2.0 * std::log(1.0 - u2));
      return radius_ * std::cos(kTwoPi * angle_u_);
    }
    return radius_ * std::sin(kTwoPi * angle_u_);
  }

  double gamma(double shape)
  {
    if (shape == 1.0) return exponential();
    if (shape > 1.0) {
      const double root = std::sqrt(2.0 * shape - 1.0), am1 = shape - 1.0;
      while (true) {
        const double tn = std::tan(kPi * u01());
        const double x = root * tn + shape - 1.0;  // evaluation order of the published algorithm
        if (x <= 0.0) continue;
        const double bound = (1.0 + tn * tn) * std::exp(am1 * std::log(x / am1) - root * tn);
        if (u01() > bound) continue;
        return x;
      }
    }
    const double e = std::exp(1.0), p = e / (shape + e);
    while (true) {
      const double u = u01();
      const double y = exponential();
      double x, q;
      if (u < p) {
        x = std::exp(-y / shape);
        q = p * std::exp(-x);
      } else {
        x = 1.0 + y;
        q = p + (1.0 - p) * std::pow(x, shape - 1.0);
      }
      if (u >= q) continue;
      return x;
    }
  }

  // scaled inverse chi^2 with the degrees of freedom fixed at construction (sigma2 draws, rand.hpp:61-78)
  double sinvchi2_fixed(double s2)
  {
    double v = -1.0;
    while (!(v > 0.0) || !std::isfinite(v)) v = fixed_nu_ * s2 / (2.0 * gamma(0.5 * fixed_nu_));
    return v;
  }
  // scaled inverse chi^2 (tau2 draws, rand.hpp:85-97)
  double sinvchi2(double nu, double s2)
  {
    double v = -1.0;
    while (!(v > 0.0) || !std::isfinite(v)) v = nu * s2 / (2.0 * gamma(0.5 * nu));
    return v;
  }

 private:
  static constexpr int kN = 624, kM = 397;
  static constexpr double kPi = 3.14159265358979323846, kTwoPi = 2.0 * 3.14159265358979323846;

  double exponential() { return -std::log(1.0 - u01()); }

  uint32_t next_u32()
  {
    if (pos_ >= kN) regenerate();
    uint32_t y = state_[pos_++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  void regenerate()
  {
    for (int i = 0; i < kN; ++i) {
      const uint32_t y = (state_[i] & 0x80000000u) | (state_[(i + 1) % kN] & 0x7fffffffu);
      state_[i] = state_[(i + kM) % kN] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    pos_ = 0;
  }

  uint32_t state_[kN];
  int pos_;
  double fixed_nu_;
  bool have_cached_ = false;
  double angle_u_ = 0.0, radius_ = 0.0;
};

}  // namespace bmg
