// tests/harness/host_harness.cpp -- TEST INFRASTRUCTURE.  A C window onto the pure host pieces of the sampler
// (bmagwa_b200/csrc/host/missing.hpp: no device dependency) so that they can be checked on the CPU against the
// unmodified reference (oracle/_ref).  Compiled by tests/test_cpu_missing_gibbs.py with g++; not part of the product.
#include <cstring>
#include "missing.hpp"

using namespace bmg;

extern "C" {

// One Gibbs update of the in-model SNPs' missing cells (Sampler::sample_missing, sampler.cpp:264-453).
//   xx: cols x cols column-major (upper triangle used), in/out;  xy: cols, in/out;  val: in/out (CSR over all SNPs)
//   rows_out must hold n entries; returns the number of rows used (q).  cells is filled by cell_value(l, row) callbacks
//   replaced here by a dense n x k matrix `xcols` (column-major, the overlay-applied genotype columns BEFORE the update).
int harness_gibbs(int m_e, int k, const unsigned* loci, double* xx, double* xy, const double* beta, double sigma2,
                  long m_g, const long* off, const int* idx, signed char* val, const double* prior3, const double* xcols,
                  const double* y, const double* e, long n, double yy, unsigned seed, double nu, int skip_draws)
{
  Model cur;
  cur.m_e = m_e;
  cur.loci.assign(loci, loci + k);
  const int cols = m_e + k;
  cur.xx.resize(cols);
  for (int c = 0; c < cols; ++c)
    for (int r = 0; r <= c; ++r) cur.xx(r, c) = xx[(size_t)c * cols + r];
  cur.xy.assign(xy, xy + cols);
  cur.beta.assign(beta, beta + cols);
  cur.sigma2 = sigma2;
  MissingCells mc;
  mc.off.assign(off, off + m_g + 1);
  mc.idx.assign(idx, idx + off[m_g]);
  mc.val.assign(val, val + off[m_g]);
  mc.prior3.assign(prior3, prior3 + 3 * m_g);
  std::vector<int32_t> rows;
  rows_missing_in_model(mc, cur.loci, rows);
  const size_t q = rows.size();
  std::vector<int8_t> cells((size_t)k * q);
  for (int l = 0; l < k; ++l)
    for (size_t u = 0; u < q; ++u)   // value in bits 0-1, bit 2 = the cell is a missing call (as bmg_chain_get_cells returns)
      cells[(size_t)l * q + u] = (int8_t)((int)xcols[(size_t)l * n + rows[u]] | (mc.is_missing(rows[u], loci[l]) ? 4 : 0));
  ChainRng rng(seed, nu);
  for (int i = 0; i < skip_draws; ++i) rng.u01();
  GibbsScratch ws;
  gibbs_missing_in_model(cur, mc, rows, cells.data(), y, e, (size_t)n, yy, rng, ws);
  for (int32_t v : ws.slot) if (v != -1) return -1;   // the scratch must come back clean
  for (int c = 0; c < cols; ++c)
    for (int r = 0; r <= c; ++r) xx[(size_t)c * cols + r] = cur.xx(r, c);
  std::memcpy(xy, cur.xy.data(), sizeof(double) * cols);
  std::memcpy(val, mc.val.data(), mc.val.size());
  return (int)q;
}

// DataModel::sample_missing (data_model.cpp:78-90): all SNPs not flagged in_model, from the prior
void harness_draw_all_from_prior(long m_g, const long* off, signed char* val, const double* prior3,
                                 const unsigned char* in_model, unsigned seed, double nu)
{
  MissingCells mc;
  mc.off.assign(off, off + m_g + 1);
  mc.idx.assign((size_t)off[m_g], 0);
  mc.val.assign(val, val + off[m_g]);
  mc.prior3.assign(prior3, prior3 + 3 * m_g);
  ChainRng rng(seed, nu);
  mc.draw_all_from_prior([&](size_t s) { return in_model[s] != 0; }, rng);
  std::memcpy(val, mc.val.data(), mc.val.size());
}

}  // extern "C"

#include <time.h>
// Wall seconds per call of gibbs_missing_in_model over `reps` calls on the same starting state (development timing aid).
extern "C" double harness_gibbs_seconds(int m_e, int k, const unsigned* loci, const double* xx, const double* xy, const double* beta,
                                        double sigma2, long m_g, const long* off, const int* idx, const signed char* val,
                                        const double* prior3, const double* xcols, const double* y, const double* e, long n,
                                        double yy, int reps)
{
  Model base;
  base.m_e = m_e;
  base.loci.assign(loci, loci + k);
  const int cols = m_e + k;
  base.xx.resize(cols);
  for (int c = 0; c < cols; ++c)
    for (int r = 0; r <= c; ++r) base.xx(r, c) = xx[(size_t)c * cols + r];
  base.xy.assign(xy, xy + cols);
  base.beta.assign(beta, beta + cols);
  base.sigma2 = sigma2;
  MissingCells mc0;
  mc0.off.assign(off, off + m_g + 1);
  mc0.idx.assign(idx, idx + off[m_g]);
  mc0.val.assign(val, val + off[m_g]);
  mc0.prior3.assign(prior3, prior3 + 3 * m_g);
  std::vector<int32_t> rows;
  GibbsScratch slot;
  ChainRng rng(1u, 1.0);
  double total = 0.0;
  for (int it = 0; it < reps; ++it) {
    Model cur;
    cur.assign(base);
    MissingCells mc = mc0;
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    rows_missing_in_model(mc, cur.loci, rows);
    const size_t q = rows.size();
    std::vector<int8_t> cells((size_t)k * q);
    for (int l = 0; l < k; ++l)
      for (size_t u = 0; u < q; ++u)
        cells[(size_t)l * q + u] = (int8_t)((int)xcols[(size_t)l * n + rows[u]] | (mc.is_missing(rows[u], loci[l]) ? 4 : 0));
    struct timespec a2;
    clock_gettime(CLOCK_MONOTONIC, &a2);
    gibbs_missing_in_model(cur, mc, rows, cells.data(), y, e, (size_t)n, yy, rng, slot);
    clock_gettime(CLOCK_MONOTONIC, &b);
    (void)a;
    total += (b.tv_sec - a2.tv_sec) + 1e-9 * (b.tv_nsec - a2.tv_nsec);
  }
  return total / reps;
}
