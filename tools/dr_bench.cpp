// dr_bench.cpp -- host-only timing of the delayed-rejection enumeration (exhaustive.hpp) per move size.
//   g++ -O3 -march=x86-64-v3 -ffp-contract=off -std=c++17 -I bmagwa_b200/csrc/host tools/dr_bench.cpp -o /tmp/dr_bench
#include <chrono>
#include <cstdio>
#include <random>
#include "model.hpp"
#include "exhaustive.hpp"
using namespace bmg;
int main()
{
  const int n = 5000, m_e = 3, k0 = 20;
  const double types_prior[5] = {1, 1, 1, 1, 1};
  std::vector<double> ite(m_e, 1.0); ite[0] = 0.0;
  std::mt19937 gen(5);
  std::normal_distribution<double> N01;
  for (int ms = 2; ms <= 10; ++ms) {
    const int k = k0 + ms, cols = m_e + k;
    // random design: Gram of `cols` random vectors of length 200
    const int L = 200;
    std::vector<double> X((size_t)L * cols), y(L);
    for (auto& v : X) v = N01(gen);
    for (auto& v : y) v = N01(gen);
    double yy = 0; for (double v : y) yy += v * v;
    Prior prior(n, 100000, m_e, yy, types_prior, 20, 300, ite, 1.0, 0.5, 5.0, 0.05, 0.0, true);
    auto dot = [&](const double* a, const double* b) { double s = 0; for (int i = 0; i < L; ++i) s += a[i] * b[i]; return s; };
    UpperMat exx; exx.resize(m_e);
    std::vector<double> exy(m_e);
    for (int c = 0; c < m_e; ++c) { for (int r = 0; r <= c; ++r) exx(r, c) = dot(&X[(size_t)r * L], &X[(size_t)c * L]); exy[c] = dot(&X[(size_t)c * L], y.data()); }
    Model m; m.init(m_e, exx, exy, &prior);
    for (int t = 0; t < k; ++t) {
      std::vector<double> col(m.cols() + 1);
      const double* x = &X[(size_t)(m_e + t) * L];
      for (int c = 0; c < m.cols(); ++c) col[c] = dot(&X[(size_t)c * L], x);
      col[m.cols()] = dot(x, x);
      m.add_term(t, dot(x, y.data()), col.data(), 1.0 + 0.1 * t);
    }
    SubmodelEnumerator exh;
    std::vector<double> P((size_t)1 << ms), qa(ms, 0.01), qr(ms, 0.9);
    std::vector<unsigned char> order(ms);
    for (int i = 0; i < ms; ++i) order[i] = (unsigned char)i;
    const int reps = 20000 >> (ms > 6 ? ms - 6 : 0);
    double mx = 0, sink = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) { exh.run(m, k0 + ms - 1, 1, P.data(), mx); sink += P[1]; }   // fixed cost: shared part + block set-up
    auto t1 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) { exh.run(m, k0, ms, P.data(), mx); sink += P[1]; }
    auto t2 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) { std::fill(P.begin(), P.end(), 0.0); compute_proposal_probs_for_exh_modelset(ms, order.data(), qa.data(), qr.data(), 30.0, 18.0, k0, 100000, P.data()); sink += P[1]; }
    auto t3b = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) { std::fill(P.begin(), P.end(), 0.0); compute_proposal_probs_stepwise(ms, order.data(), qa.data(), qr.data(), 30.0, 18.0, k0, 100000, P.data()); sink += P[1]; }
    auto t4 = std::chrono::steady_clock::now();
    auto t3 = std::chrono::steady_clock::now();
    auto us = [&](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count() / reps; };
    const double a = us(t0, t1), b = us(t1, t2), c = us(t2, t3), c2 = us(t3b, t4);
    printf("ms %2d  set-up %5.2f us  enumeration %7.2f us (%5.1f ns/sub-model)  proposal probs %7.2f us (%5.1f ns/sub-model; stepwise %5.1f)  [%g]\n", ms, a, b,
           1e3 * b / (1 << ms), c, 1e3 * c / (1 << ms), 1e3 * c2 / (1 << ms), sink);
  }
  return 0;
}
