// missing.hpp -- missing genotypes of one chain on the host side: the index, the imputed values and the Gibbs step.
//
// What the reference keeps in Data::miss_loc_/miss_prior_ (src/data.cpp:341-376) and DataModel::miss_val_
// (src/data_model.hpp:75-101), and what Sampler::sample_missing (src/sampler.cpp:264-453),
// DataModel::sample_missing and sample_missing_single (src/data_model.cpp:78-103) do with them.
//
// The reference's Gibbs step walks columns of the n x k design matrix; no such matrix exists here (the packed device
// store is the column cache).  The step only ever touches the rows of individuals that have a missing call in some
// in-model SNP, so the sampler gathers exactly those cells from the device (bmg_chain_get_cells) and this file does the
// reference's arithmetic on them, in the reference's order and with the reference's random-number consumption (one
// uniform per missing cell).  Everything here is plain host code with no device dependency, so it is unit-tested on the
// CPU against the unmodified reference (tests/test_cpu_missing_gibbs.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>
#include "model.hpp"
#include "rng.hpp"

namespace bmg {

// Host mirror of the store's missing index plus the chain's imputed values (CSR over SNPs, rows ascending per SNP)
struct MissingCells {
  std::vector<int64_t> off;     // m_g + 1
  std::vector<int32_t> idx;     // individual of each missing cell
  std::vector<int8_t> val;      // imputed value 0/1/2, initially 0 (data_model.hpp:80-84)
  std::vector<double> prior3;   // per SNP: cumulative counts of 0/1/2 among the observed cells (data.cpp:357-372)

  int64_t total() const { return off.empty() ? 0 : off.back(); }
  int64_t count(size_t snp) const { return off[snp + 1] - off[snp]; }
  // DataModel::genotype_missing (data_model.hpp:104-117)
  bool is_missing(int32_t individual, size_t snp) const
  {
    const int32_t* b = idx.data() + off[snp];
    const int32_t* e = idx.data() + off[snp + 1];
    return std::binary_search(b, e, individual);
  }
  // DataModel::sample_missing_single (data_model.cpp:95-103): one uniform per cell through Utils::sample_discrete_naive
  void draw_from_prior(size_t snp, ChainRng& rng)
  {
    const double* cum = &prior3[3 * snp];
    for (int64_t q = off[snp]; q < off[snp + 1]; ++q) val[q] = (int8_t)draw3(cum, rng);
  }
  // DataModel::sample_missing (data_model.cpp:78-90): every SNP that is not in the model, in SNP order
  template <class InModel>
  void draw_all_from_prior(InModel in_model, ChainRng& rng)
  {
    const size_t m = off.size() - 1;
    for (size_t snp = 0; snp < m; ++snp) {
      if (off[snp + 1] == off[snp] || in_model(snp)) continue;
      draw_from_prior(snp, rng);
    }
  }
  // Utils::sample_discrete_naive (utils.cpp:46-57) for three classes
  static int draw3(const double* cumsum, ChainRng& rng)
  {
    const double r = rng.u01() * cumsum[2];
    for (int i = 0; i < 3; ++i)
      if (r < cumsum[i]) return i;
    throw std::logic_error("sample_discrete_naive reached end, exiting");
  }
};

// Individuals (ascending, distinct) that have a missing call in at least one SNP of the model: the rows the Gibbs
// step needs from the device.
inline void rows_missing_in_model(const MissingCells& mc, const std::vector<uint32_t>& loci, std::vector<int32_t>& rows)
{
  rows.clear();
  for (uint32_t snp : loci) rows.insert(rows.end(), mc.idx.begin() + mc.off[snp], mc.idx.begin() + mc.off[snp + 1]);
  std::sort(rows.begin(), rows.end());
  rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
}

// Buffers of the Gibbs step, kept between calls (the step runs every n_sample_tau2_and_missing iterations)
// A SNP of the model as the Gibbs step sees it: effect type 0 A, 1 H, 2 D, 3 R, 4 AH (data_model.hpp:41) and the design-matrix
// column(s) of its term(s); col2 is the heterozygous term of an AH SNP (model.hpp: x_ind1 / x_ind2)
struct GibbsSnp {
  uint32_t snp;
  int type, col1, col2;
};

struct GibbsScratch {
  std::vector<double> xold, xnew, y_hat, residual;
  std::vector<int32_t> slot;      // individual -> position in rows; all -1 between calls
  std::vector<uint8_t> changed;   // per row: did any of its cells get a new value
  std::vector<int32_t> col_snp;   // design-matrix column -> position of its SNP in the model
  std::vector<GibbsSnp> snps;
};

// value of a term of type t (0..3) on additive genotype g (DataModel::get_genotypes_*, data_model.cpp:41-72)
inline double typed_genotype(int t, int g)
{
  return t == 0 ? (double)g : (t == 1 ? (double)(g == 1) : (t == 2 ? (double)(g > 0) : (double)(g == 2)));
}

// Sampler::sample_missing (sampler.cpp:264-453) for SNPs of any effect type.
//   cur    current model; xx and xy are patched in place (sampler.cpp:393-450), mu_beta_computed is cleared
//   snps   the model's SNPs in model order with their types and columns
//   rows   individuals (ascending, distinct) with a missing call in some SNP of the model (q of them);
//          cells = n_snps x q ADDITIVE genotype values with the chain's imputed values applied, as they are BEFORE this
//          update, bit 2 set where the cell is a missing call (bmg_chain_get_cells over the SNPs)
//   y, e   phenotype (n) and covariates (n x m_e, column-major, ones column included)
//   yy     y'y.  The reference sums the squared residual over all n individuals; here r'r comes from the Gram
//          matrix, r'r = y'y - 2 b'X'y + b'X'X b (it only enters through differences in which it cancels).
// On return mc.val holds the new imputed values of the model's SNPs.
inline void gibbs_missing_typed(Model& cur, const std::vector<GibbsSnp>& snps, MissingCells& mc, const std::vector<int32_t>& rows,
                                const int8_t* cells, const double* y, const double* e, size_t n, double yy, ChainRng& rng,
                                GibbsScratch& ws)
{
  const int m_e = cur.m_e, cols = cur.cols(), k = (int)snps.size();
  const size_t q = rows.size();
  cur.mu_beta_computed = false;   // sampler.cpp:452, unconditional
  if (q == 0) return;
  const std::vector<double>& beta = cur.beta;
  const double sigma2_times_2 = cur.sigma2 * 2;

  // the touched rows of the design matrix, before (xold: the reference's new_model->x) and after (xnew) the update
  ws.xold.resize(q * (size_t)cols);
  ws.xnew.resize(q * (size_t)cols);
  ws.y_hat.resize(q);
  ws.residual.resize(q);
  ws.changed.assign(q, 0);
  ws.col_snp.assign(cols, -1);   // design-matrix column -> position of its SNP in snps (covariates: -1)
  double* const xold = ws.xold.data();
  double* const xnew = ws.xnew.data();
  double* const y_hat = ws.y_hat.data();
  double* const residual = ws.residual.data();
  for (int l = 0; l < k; ++l) {
    ws.col_snp[snps[l].col1] = l;
    if (snps[l].type == 4) ws.col_snp[snps[l].col2] = l;
  }
  for (size_t u0 = 0; u0 < q; u0 += 128) {   // blocks of rows, column by column inside: e and cells are column-major
    const size_t u1 = std::min(q, u0 + 128);
    for (int c = 0; c < m_e; ++c) {
      const double* ec = e + (size_t)c * n;
      for (size_t u = u0; u < u1; ++u) xold[u * cols + c] = ec[rows[u]];
    }
    for (int l = 0; l < k; ++l) {
      const int8_t* cl = cells + (size_t)l * q;
      const int ty = snps[l].type, c1 = snps[l].col1, c2 = snps[l].col2;
      if (ty == 0) {
        for (size_t u = u0; u < u1; ++u) xold[u * cols + c1] = (double)(cl[u] & 3);
      } else {
        const int t1 = ty == 4 ? 0 : ty;
        for (size_t u = u0; u < u1; ++u) xold[u * cols + c1] = typed_genotype(t1, cl[u] & 3);
        if (ty == 4)
          for (size_t u = u0; u < u1; ++u) xold[u * cols + c2] = typed_genotype(1, cl[u] & 3);
      }
    }
  }
  for (size_t u = 0; u < q; ++u) {
    const double* row = xold + u * cols;
    double s = 0.0;   // y_hat = X beta, accumulated column by column like the dgemv of Vector::set_to_product
    for (int c = 0; c < cols; ++c) s += beta[c] * row[c];
    y_hat[u] = s;
    residual[u] = y[rows[u]] - s;
  }
  std::copy(xold, xold + q * (size_t)cols, xnew);
  double r2 = yy;
  for (int c = 0; c < cols; ++c) r2 -= 2.0 * beta[c] * cur.xy[c];
  r2 += cur.quad(0, cols);

  if (ws.slot.size() != n) ws.slot.assign(n, -1);
  int32_t* const slot = ws.slot.data();
  for (size_t u = 0; u < q; ++u) slot[rows[u]] = (int32_t)u;

  double lprior[3], likelihood[3];
  for (int l = 0; l < k; ++l) {
    const size_t snp = snps[l].snp;
    if (mc.count(snp) == 0) continue;
    const int ty = snps[l].type, t1 = ty == 4 ? 0 : ty, x1 = snps[l].col1, x2 = snps[l].col2;
    const double* p3 = &mc.prior3[3 * snp];
    lprior[0] = std::log(p3[0]);
    lprior[1] = std::log(p3[1] - p3[0]);
    lprior[2] = std::log(p3[2] - p3[1]);
    for (int64_t c = mc.off[snp]; c < mc.off[snp + 1]; ++c) {
      const size_t u = (size_t)slot[mc.idx[c]];
      double* row = xnew + u * cols;
      const double old_term2 = residual[u] * residual[u];
      for (int g = 0; g < 3; ++g) {
        double new_term = residual[u] - beta[x1] * (typed_genotype(t1, g) - row[x1]);
        if (ty == 4) new_term -= beta[x2] * ((double)(g == 1) - row[x2]);
        likelihood[g] = lprior[g] - (r2 + new_term * new_term - old_term2) / sigma2_times_2;
      }
      likelihood[1] -= likelihood[0];
      likelihood[2] -= likelihood[0];
      likelihood[0] = 1;
      likelihood[1] = std::exp(likelihood[1]) + likelihood[0];
      likelihood[2] = std::exp(likelihood[2]) + likelihood[1];
      const int g = MissingCells::draw3(likelihood, rng);
      mc.val[c] = (int8_t)g;
      const double v1 = typed_genotype(t1, g);
      if (v1 != row[x1]) ws.changed[u] = 1;
      y_hat[u] += beta[x1] * (v1 - row[x1]);
      row[x1] = v1;
      if (ty == 4) {
        const double v2 = (double)(g == 1);
        if (v2 != row[x2]) ws.changed[u] = 1;
        y_hat[u] += beta[x2] * (v2 - row[x2]);
        row[x2] = v2;
      }
      residual[u] = y[mc.idx[c]] - y_hat[u];
      r2 += residual[u] * residual[u] - old_term2;
    }
  }

  // X'y and the upper triangle of X'X follow the changed cells (sampler.cpp:393-450): the term of a non-AH SNP or the
  // first term of an AH SNP, then the second term of an AH SNP.  A row in which no cell got a new value contributes
  // x*x' - x*x' = 0 exactly to every entry, so it is skipped.
  for (int l = 0; l < k; ++l) {
    const size_t snp = snps[l].snp;
    if (mc.count(snp) == 0) continue;
    for (int pass = 0; pass < (snps[l].type == 4 ? 2 : 1); ++pass) {
      const int x_ind = pass == 0 ? snps[l].col1 : snps[l].col2;
      for (int64_t c = mc.off[snp]; c < mc.off[snp + 1]; ++c) {
        const int32_t i_miss = mc.idx[c];
        const size_t u = (size_t)slot[i_miss];
        if (!ws.changed[u]) continue;
        const double* xn = xnew + u * cols;
        const double* xo = xold + u * cols;
        const double xnx = xn[x_ind], xox = xo[x_ind];
        cur.xy[x_ind] += y[i_miss] * (xnx - xox);
        double* up = cur.xx.col(x_ind);   // xx(j, x_ind), j < x_ind
        int j = 0;
        for (; j < m_e; ++j) up[j] += xnx * xn[j] - xox * xo[j];
        for (; j < x_ind; ++j)   // a column whose SNP is itself missing here is patched when its own cell comes up
          if (!(cells[(size_t)ws.col_snp[j] * q + u] & 4)) up[j] += xnx * xn[j] - xox * xo[j];
        for (; j < cols; ++j) cur.xx(x_ind, j) += xnx * xn[j] - xox * xo[j];
      }
    }
  }
  for (size_t u = 0; u < q; ++u) slot[rows[u]] = -1;
}

// The same for a model whose SNPs are all additive (model.types = A): SNP t of cur.loci owns column m_e + t
inline void gibbs_missing_in_model(Model& cur, MissingCells& mc, const std::vector<int32_t>& rows, const int8_t* cells,
                                   const double* y, const double* e, size_t n, double yy, ChainRng& rng, GibbsScratch& ws)
{
  ws.snps.resize(cur.size());
  for (size_t t = 0; t < cur.size(); ++t) ws.snps[t] = GibbsSnp{cur.loci[t], 0, cur.m_e + (int)t, -1};
  gibbs_missing_typed(cur, ws.snps, mc, rows, cells, y, e, n, yy, rng, ws);
}

}  // namespace bmg
