// philox.cuh -- counter-based RNG for the device-side draws (throughput mode of the per-SNP
// tau draw; probit latent uniforms).  Philox4x32-10 (Salmon et al., SC'11).
#pragma once
#include <stdint.h>

namespace bmg {

struct Philox {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t out[4];
  int have;

  __device__ Philox(uint64_t seed, uint64_t stream, uint64_t index)
  {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
    ctr[0] = 0;
    ctr[1] = (uint32_t)stream ^ (uint32_t)(stream >> 32) * 0x9E3779B9u;
    ctr[2] = (uint32_t)index;
    ctr[3] = (uint32_t)(index >> 32);
    have = 0;
  }
  __device__ void round(uint32_t* c, const uint32_t* k)
  {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ void refill()
  {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, k);
      k[0] += 0x9E3779B9u;
      k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    ++ctr[0];
    have = 4;
  }
  __device__ uint32_t next32()
  {
    if (have == 0) refill();
    return out[--have];
  }
  // uniform in (0,1) with 52 random bits, never 0 or 1: (x + 1/2) 2^-52 is exact for x < 2^52 (with 53 bits the half is
  // rounded away above 2^52 and the largest x gives exactly 1)
  __device__ double u01()
  {
    const uint64_t a = next32(), b = next32();
    const uint64_t x = ((a << 32) | b) >> 12;  // 52 bits
    return ((double)x + 0.5) * (1.0 / 4503599627370496.0);
  }
  __device__ double normal()
  {
    const double u1 = u01(), u2 = u01();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
  // Gamma(shape a, scale 1), Marsaglia & Tsang (2000)
  __device__ double gamma(double a)
  {
    double boost = 1.0;
    if (a < 1.0) {
      boost = pow(u01(), 1.0 / a);
      a += 1.0;
    }
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
      double x, v;
      do {
        x = normal();
        v = 1.0 + c * x;
      } while (v <= 0.0);
      v = v * v * v;
      const double u = u01();
      if (u < 1.0 - 0.0331 * x * x * x * x) return boost * d * v;
      if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
    }
  }
};

}  // namespace bmg
