// tests/harness/host_harness.cpp -- TEST INFRASTRUCTURE.  A C window onto the pure host pieces of the sampler
// (bmagwa_b200/csrc/host/missing.hpp: no device dependency) so that they can be checked on the CPU against the
// unmodified reference (oracle/_ref).  Compiled by tests/test_cpu_missing_gibbs.py with g++; not part of the product.
#include <cstring>
#include "exhaustive.hpp"
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

// ---------------------------------------------------------------------------------------------------------------
// The other pure host pieces of the sampler (rng.hpp, model.hpp), for tests/test_cpu_host_model.py
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

// ChainRng (rand.hpp:36-191 re-stated): the draw pattern of tests/golden/make_golden.py::run_rng
void harness_rng_sequence(unsigned seed, double nu, int rounds, double* out)
{
  ChainRng g(seed, nu);
  for (int i = 0; i < rounds; ++i) {
    *out++ = g.u01(); *out++ = g.normal(); *out++ = g.sinvchi2_fixed(0.7); *out++ = g.sinvchi2(5.0, 0.05);
    *out++ = g.sinvchi2(0.6, 1.3); *out++ = g.u01(); *out++ = g.normal(); *out++ = g.sinvchi2(2.0, 1.0);
  }
}

static Prior* make_prior(long n, long m_g, int m_e, double yy, double e_qg, double var_qg, double nu_sigma2, double s2_sigma2,
                         double nu_tau2, double s2_tau2, double mu_alpha, int individual, double inv_tau2_e_const, double inv_tau2_e_val)
{
  const double types_prior[5] = {1, 1, 1, 1, 1};
  std::vector<double> inv_tau2_e((size_t)m_e, inv_tau2_e_val);
  inv_tau2_e[0] = inv_tau2_e_const;
  return new Prior((size_t)n, (size_t)m_g, (size_t)m_e, yy, types_prior, e_qg, var_qg, inv_tau2_e, nu_sigma2, s2_sigma2, nu_tau2,
                   s2_tau2, mu_alpha, individual != 0);
}

// Prior (prior.hpp:38-290): out = {n_plus_nu, nus2_plus_yy, alpha, shared inv_tau2_alpha2, e_g}, then for L = 0..L_max-1
// log_change_on_add(L), log_change_on_rem(L + 1), log_model(L)
void harness_prior(long n, long m_g, int m_e, double yy, double e_qg, double var_qg, double nu_sigma2, double s2_sigma2,
                   double nu_tau2, double s2_tau2, double mu_alpha, int individual, int L_max, double* out5, double* add,
                   double* rem, double* model)
{
  Prior* p = make_prior(n, m_g, m_e, yy, e_qg, var_qg, nu_sigma2, s2_sigma2, nu_tau2, s2_tau2, mu_alpha, individual, 0.0, 1.0);
  out5[0] = p->n_plus_nu; out5[1] = p->nus2_plus_yy; out5[2] = p->alpha(); out5[3] = p->shared_inv_tau2_alpha2(); out5[4] = p->e_g();
  for (int L = 0; L < L_max; ++L) { add[L] = p->log_change_on_add(L); rem[L] = p->log_change_on_rem(L + 1); model[L] = p->log_model(L); }
  delete p;
}

// Model (model.hpp:199-312,432-556): a sequence of add_term / remove_term with the Gram entries taken from dense columns.
// ops[3 i..] = {0 add | 1 remove, SNP | model index, inv_tau2_alpha2}; G = n x m_g additive columns, E = n x m_e (ones first).
// trace[0] = log-likelihood of the covariate-only model, trace[i + 1] after op i; then the final xx / l (cols x cols,
// upper), xy, v and the log-likelihood of a full recomputation.
int harness_model_trace(long n, long m_g, int m_e, const double* G, const double* E, const double* y, double yy, double e_qg,
                        double var_qg, double nu_sigma2, double s2_sigma2, double nu_tau2, double s2_tau2, int n_ops,
                        const double* ops, double* trace, double* xx_out, double* l_out, double* xy_out, double* v_out,
                        double* full_out, unsigned* loci_out)
{
  Prior* p = make_prior(n, m_g, m_e, yy, e_qg, var_qg, nu_sigma2, s2_sigma2, nu_tau2, s2_tau2, 1.0, 0, 0.0, 1.0);
  auto dot = [&](const double* a, const double* b) { double s = 0.0; for (long i = 0; i < n; ++i) s += a[i] * b[i]; return s; };
  UpperMat exx;
  exx.resize(m_e);
  std::vector<double> exy(m_e);
  for (int c = 0; c < m_e; ++c) {
    for (int r = 0; r <= c; ++r) exx(r, c) = dot(E + (size_t)r * n, E + (size_t)c * n);
    exy[c] = dot(E + (size_t)c * n, y);
  }
  Model m;
  m.init(m_e, exx, exy, p);
  trace[0] = m.log_likelihood;
  for (int i = 0; i < n_ops; ++i) {
    if (ops[3 * i] == 0.0) {
      const unsigned snp = (unsigned)ops[3 * i + 1];
      const double* x = G + (size_t)snp * n;
      std::vector<double> col(m.cols() + 1);
      for (int c = 0; c < m_e; ++c) col[c] = dot(E + (size_t)c * n, x);
      for (size_t t = 0; t < m.size(); ++t) col[m_e + t] = dot(G + (size_t)m.loci[t] * n, x);
      col[m.cols()] = dot(x, x);
      m.add_term(snp, dot(x, y), col.data(), ops[3 * i + 2]);
    } else {
      m.remove_term((int)ops[3 * i + 1]);
    }
    trace[i + 1] = m.log_likelihood;
  }
  const int cols = m.cols();
  for (int c = 0; c < cols; ++c)
    for (int r = 0; r < cols; ++r) {
      xx_out[(size_t)c * cols + r] = r <= c ? m.xx(r, c) : 0.0;
      l_out[(size_t)c * cols + r] = r <= c ? m.l(r, c) : 0.0;
    }
  for (int c = 0; c < cols; ++c) { xy_out[c] = m.xy[c]; v_out[c] = m.v[c]; }
  for (size_t t = 0; t < m.size(); ++t) loci_out[t] = m.loci[t];
  m.compute_log_likelihood();
  *full_out = m.log_likelihood;
  delete p;
  return cols;
}

// ProposalCdf (discrete_distribution.hpp:64-330 semantics): a recorded sequence of zero / unzero / sample operations.
// rec[4 i..] = {0 zero | 1 unzero | 2 sample, item, -, u}; got[i] = sampled item (or -1), total[i] = total weight afterwards.
// order = the in-order permutation; weights are given by item and laid out in-order here, with per-block sums, as the
// device does it.
void harness_proposal_cdf(long m, const int* order, const double* w, int block, int n_rec, const double* rec, long* got,
                          double* total)
{
  std::vector<int32_t> ord(order, order + m);
  std::vector<double> w_io(m);
  for (long p = 0; p < m; ++p) w_io[p] = w[ord[p]];
  const long nb = (m + block - 1) / block;
  std::vector<double> sums(nb, 0.0);
  for (long p = 0; p < m; ++p) sums[p / block] += w_io[p];
  ProposalCdf dd;
  dd.init(&ord, block);
  dd.update(w_io.data(), sums.data(), true, std::vector<uint32_t>());
  for (int i = 0; i < n_rec; ++i) {
    const int kind = (int)rec[4 * i];
    const uint32_t item = (uint32_t)rec[4 * i + 1];
    got[i] = -1;
    if (kind == 0) dd.zero(item);
    else if (kind == 1) dd.unzero(item);
    else got[i] = (long)dd.sample(rec[4 * i + 3]);
    total[i] = dd.total();
  }
}

}  // extern "C"

// Delayed rejection (sampler.cpp:882-980): log probabilities of all 2^ms sub-models of the last ms SNPs of a model, from
// SubmodelEnumerator (include / delete-first-column walk over the trailing block of the factor) in P, and from a fresh Model per sub-model
// (full incremental build) + the model prior in B.  Both are relative to the sub-model without any of the ms SNPs.
extern "C" void harness_exhaustive(long n, long m_g, int m_e, const double* G, const double* E, const double* y, double yy,
                                   double e_qg, double var_qg, double nu_sigma2, double s2_sigma2, double nu_tau2, double s2_tau2,
                                   int const_loci, int ms, const unsigned* snps, const double* taus, double* P, double* B)
{
  Prior* p = make_prior(n, m_g, m_e, yy, e_qg, var_qg, nu_sigma2, s2_sigma2, nu_tau2, s2_tau2, 1.0, 0, 0.0, 1.0);
  auto dot = [&](const double* a, const double* b) { double s = 0.0; for (long i = 0; i < n; ++i) s += a[i] * b[i]; return s; };
  UpperMat exx;
  exx.resize(m_e);
  std::vector<double> exy(m_e);
  for (int c = 0; c < m_e; ++c) {
    for (int r = 0; r <= c; ++r) exx(r, c) = dot(E + (size_t)r * n, E + (size_t)c * n);
    exy[c] = dot(E + (size_t)c * n, y);
  }
  auto add = [&](Model& m, int which) {
    const double* x = G + (size_t)snps[which] * n;
    std::vector<double> col(m.cols() + 1);
    for (int c = 0; c < m_e; ++c) col[c] = dot(E + (size_t)c * n, x);
    for (size_t t = 0; t < m.size(); ++t) col[m_e + t] = dot(G + (size_t)m.loci[t] * n, x);
    col[m.cols()] = dot(x, x);
    m.add_term(snps[which], dot(x, y), col.data(), taus[which]);
  };
  Model full;
  full.init(m_e, exx, exy, p);
  for (int i = 0; i < const_loci + ms; ++i) add(full, i);
  SubmodelEnumerator exh;
  double mx;
  exh.run(full, const_loci, ms, P, mx);
  for (unsigned long mask = 0; mask < (1ul << ms); ++mask) {
    Model m;
    m.init(m_e, exx, exy, p);
    for (int i = 0; i < const_loci; ++i) add(m, i);
    for (int b = 0; b < ms; ++b)
      if ((mask >> b) & 1) add(m, const_loci + b);
    B[mask] = m.log_likelihood + p->log_model((int)m.size());
  }
  const double p0 = P[0], b0 = B[0];
  for (unsigned long mask = 0; mask < (1ul << ms); ++mask) { P[mask] -= p0; B[mask] -= b0; }
  delete p;
}

// Delayed rejection: proposal probabilities of the sub-models (sampler.cpp:982-1049), the product's restatement
extern "C" void harness_dr_proposal_probs(int n_inds, const unsigned char* bit_to_normalized_order, const double* q_add,
                                          const double* q_rem, double z_add, double z_rem, long const_loci, long m_g,
                                          const double* log_q_add_types, double* log_prop_probs)
{
  compute_proposal_probs_for_exh_modelset(n_inds, bit_to_normalized_order, q_add, q_rem, z_add, z_rem, (size_t)const_loci,
                                          (size_t)m_g, log_prop_probs, log_q_add_types);
}

// One round of the model-level Gibbs updates on a model holding the given SNPs: Model::sample_beta_sigma2
// (model.hpp:326-342), Prior::sample_alpha_and_tau2 (prior.cpp:30-141), a full likelihood recomputation and
// Model::compute_pve (model.hpp:345-392), with the chain's random stream freshly seeded.
//   out_beta / out_tau: cols entries; out3 = {sigma2, alpha, log likelihood}; pves: 3
extern "C" void harness_model_gibbs(long n, long m_g, int m_e, const double* G, const double* E, const double* y, double yy,
                                    double e_qg, double var_qg, double nu_sigma2, double s2_sigma2, double nu_tau2, double s2_tau2,
                                    double mu_alpha, int individual, double inv_tau2_e_val, int k, const unsigned* snps,
                                    const double* taus, unsigned seed, double* out_beta, double* out_tau, double* out3, double* pves)
{
  Prior* p = make_prior(n, m_g, m_e, yy, e_qg, var_qg, nu_sigma2, s2_sigma2, nu_tau2, s2_tau2, mu_alpha, individual, 0.0, inv_tau2_e_val);
  auto dot = [&](const double* a, const double* b) { double s = 0.0; for (long i = 0; i < n; ++i) s += a[i] * b[i]; return s; };
  UpperMat exx;
  exx.resize(m_e);
  std::vector<double> exy(m_e);
  for (int c = 0; c < m_e; ++c) {
    for (int r = 0; r <= c; ++r) exx(r, c) = dot(E + (size_t)r * n, E + (size_t)c * n);
    exy[c] = dot(E + (size_t)c * n, y);
  }
  Model m;
  m.init(m_e, exx, exy, p);
  for (int i = 0; i < k; ++i) {
    const double* x = G + (size_t)snps[i] * n;
    std::vector<double> col(m.cols() + 1);
    for (int c = 0; c < m_e; ++c) col[c] = dot(E + (size_t)c * n, x);
    for (size_t t = 0; t < m.size(); ++t) col[m_e + t] = dot(G + (size_t)m.loci[t] * n, x);
    col[m.cols()] = dot(x, x);
    m.add_term(snps[i], dot(x, y), col.data(), taus[i]);
  }
  ChainRng rng(seed, (double)n + nu_sigma2);
  m.sample_beta_sigma2(rng);
  const int cols = m.cols();
  for (int c = 0; c < cols; ++c) out_beta[c] = m.beta[c];
  out3[0] = m.sigma2;
  p->sample_alpha_and_tau2(&m, rng);
  for (int c = 0; c < cols; ++c) out_tau[c] = m.inv_tau2_alpha2[c];
  out3[1] = p->alpha();
  m.compute_log_likelihood();
  out3[2] = m.log_likelihood;
  m.compute_pve((size_t)n, pves);
  delete p;
}

// The Gibbs update for a model whose SNPs have effect types (0 A, 1 H, 2 D, 3 R, 4 AH = two columns): as harness_gibbs, with
// snp_type[l] per model SNP; columns are laid out in model order (AH: additive column, then heterozygous column).
// xcols: n x k ADDITIVE columns (overlay applied, before the update).
extern "C" int harness_gibbs_typed(int m_e, int k, const unsigned* loci, const int* snp_type, double* xx, double* xy,
                                   const double* beta, double sigma2, long m_g, const long* off, const int* idx, signed char* val,
                                   const double* prior3, const double* xcols, const double* y, const double* e, long n, double yy,
                                   unsigned seed, double nu)
{
  std::vector<GibbsSnp> snps(k);
  int cols = m_e;
  Model cur;
  cur.m_e = m_e;
  for (int l = 0; l < k; ++l) {
    snps[l].snp = loci[l];
    snps[l].type = snp_type[l];
    snps[l].col1 = cols++;
    cur.loci.push_back(loci[l]);
    snps[l].col2 = -1;
    if (snp_type[l] == 4) { snps[l].col2 = cols++; cur.loci.push_back(loci[l]); }
  }
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
  std::vector<uint32_t> distinct(loci, loci + k);
  std::vector<int32_t> rows;
  rows_missing_in_model(mc, distinct, rows);
  const size_t q = rows.size();
  std::vector<int8_t> cells((size_t)k * q);
  for (int l = 0; l < k; ++l)
    for (size_t u = 0; u < q; ++u)
      cells[(size_t)l * q + u] = (int8_t)((int)xcols[(size_t)l * n + rows[u]] | (mc.is_missing(rows[u], loci[l]) ? 4 : 0));
  ChainRng rng(seed, nu);
  GibbsScratch ws;
  gibbs_missing_typed(cur, snps, mc, rows, cells.data(), y, e, (size_t)n, yy, rng, ws);
  for (int c = 0; c < cols; ++c)
    for (int r = 0; r <= c; ++r) xx[(size_t)c * cols + r] = cur.xx(r, c);
  std::memcpy(xy, cur.xy.data(), sizeof(double) * cols);
  std::memcpy(val, mc.val.data(), mc.val.size());
  return cols;
}

// Scratch reuse: the sampler keeps one GibbsScratch for the whole chain while the model grows and shrinks.  Runs the update
// for the first k1 SNPs, then (same scratch) for all k SNPs, and compares the second result with a run on a fresh scratch.
// Returns 1 when both runs give identical values, X'X and X'y; 0 otherwise.
extern "C" int harness_gibbs_scratch_reuse(int m_e, int k1, int k, const unsigned* loci, const double* xx_k1, const double* xy_k1,
                                           const double* xx, const double* xy, const double* beta, double sigma2, long m_g,
                                           const long* off, const int* idx, const signed char* val, const double* prior3,
                                           const double* xcols, const double* y, const double* e, long n, double yy)
{
  auto make_model = [&](int kk, const double* gx, const double* gy) {
    Model m;
    m.m_e = m_e;
    m.loci.assign(loci, loci + kk);
    const int cols = m_e + kk;
    m.xx.resize(cols);
    for (int c = 0; c < cols; ++c)
      for (int r = 0; r <= c; ++r) m.xx(r, c) = gx[(size_t)c * cols + r];
    m.xy.assign(gy, gy + cols);
    m.beta.assign(beta, beta + cols);
    m.sigma2 = sigma2;
    return m;
  };
  auto run = [&](Model& m, MissingCells& mc, GibbsScratch& ws, unsigned seed) {
    std::vector<int32_t> rows;
    rows_missing_in_model(mc, m.loci, rows);
    const size_t q = rows.size();
    const int kk = (int)m.size();
    std::vector<int8_t> cells((size_t)kk * q);
    for (int l = 0; l < kk; ++l)
      for (size_t u = 0; u < q; ++u)
        cells[(size_t)l * q + u] = (int8_t)((int)xcols[(size_t)l * n + rows[u]] | (mc.is_missing(rows[u], loci[l]) ? 4 : 0));
    ChainRng rng(seed, 1.0);
    gibbs_missing_in_model(m, mc, rows, cells.data(), y, e, (size_t)n, yy, rng, ws);
  };
  MissingCells base;
  base.off.assign(off, off + m_g + 1);
  base.idx.assign(idx, idx + off[m_g]);
  base.val.assign(val, val + off[m_g]);
  base.prior3.assign(prior3, prior3 + 3 * m_g);

  GibbsScratch shared;
  { Model small = make_model(k1, xx_k1, xy_k1); MissingCells mc = base; run(small, mc, shared, 5u); }
  Model a = make_model(k, xx, xy);
  MissingCells mca = base;
  run(a, mca, shared, 9u);
  GibbsScratch fresh;
  Model b = make_model(k, xx, xy);
  MissingCells mcb = base;
  run(b, mcb, fresh, 9u);
  if (mca.val != mcb.val || a.xy != b.xy) return 0;
  for (int c = 0; c < a.cols(); ++c)
    for (int r = 0; r <= c; ++r)
      if (a.xx(r, c) != b.xx(r, c)) return 0;
  return 1;
}

// ---- several effect types: the typed side of Prior (prior.hpp:60-183, prior.cpp:71-141) ----------------------------------
static void configure(Prior* p, int n_types, const int* types, double nu_tau2, double s2_tau2)
{
  const double tp[5] = {1, 1, 1, 1, 1}, nu[4] = {nu_tau2, nu_tau2, nu_tau2, nu_tau2}, s2[4] = {s2_tau2, s2_tau2, s2_tau2, s2_tau2};
  p->configure_types(std::vector<int>(types, types + n_types), tp, nu, s2);
}

// queries[7 i..] = {Ns[0..4], L, type}: add[i] = log_change_on_add, rem[i] = log_change_on_rem (Ns, L are the OLD state),
// model[i] = log_model(Ns); swi[i] = log_change_on_swi(Ns, type, swi_rem[i])
extern "C" void harness_prior_typed(long n, long m_g, int m_e, double yy, double e_qg, double var_qg, double s2_sigma2,
                                    int n_types, const int* types, int n_q, const int* queries, const int* swi_rem, double* add,
                                    double* rem, double* model, double* swi, double* shared4)
{
  Prior* p = make_prior(n, m_g, m_e, yy, e_qg, var_qg, 1.0, s2_sigma2, 5.0, 0.05, 1.0, 0, 0.0, 0.001);
  configure(p, n_types, types, 5.0, 0.05);
  for (int i = 0; i < n_q; ++i) {
    const int* Ns = queries + 7 * i;
    const int L = queries[7 * i + 5], t = queries[7 * i + 6];
    add[i] = p->log_change_on_add(Ns, L, t);
    rem[i] = Ns[t] > 0 ? p->log_change_on_rem(Ns, L, t) : 0.0;
    model[i] = p->log_model(Ns);
    swi[i] = Ns[swi_rem[i]] > 0 ? p->log_change_on_swi(Ns, t, swi_rem[i]) : 0.0;
  }
  for (int t = 0; t < 4; ++t) shared4[t] = p->allow_term(t) ? p->shared_inv_tau2_alpha2(t) : -1.0;
  delete p;
}

// harness_model_gibbs for a model whose SNPs have effect types (snp_type 0..4; AH = additive + heterozygous column):
// sample_beta_sigma2, then Prior::sample_alpha_and_tau2_typed
extern "C" int harness_model_gibbs_typed(long n, long m_g, int m_e, const double* G, const double* E, const double* y, double yy,
                                         double e_qg, double var_qg, double s2_sigma2, int individual, int n_types, const int* types,
                                         int k, const unsigned* snps, const int* snp_type, const double* taus2, unsigned seed,
                                         double* out_beta, double* out_tau, double* out3)
{
  Prior* p = make_prior(n, m_g, m_e, yy, e_qg, var_qg, 1.0, s2_sigma2, 5.0, 0.05, 1.0, individual, 0.0, 0.001);
  configure(p, n_types, types, 5.0, 0.05);
  auto dot = [&](const double* a, const double* b) { double s = 0.0; for (long i = 0; i < n; ++i) s += a[i] * b[i]; return s; };
  UpperMat exx;
  exx.resize(m_e);
  std::vector<double> exy(m_e);
  for (int c = 0; c < m_e; ++c) {
    for (int r = 0; r <= c; ++r) exx(r, c) = dot(E + (size_t)r * n, E + (size_t)c * n);
    exy[c] = dot(E + (size_t)c * n, y);
  }
  Model m;
  m.init(m_e, exx, exy, p);
  std::vector<std::vector<double>> colsx;   // typed dense columns of the model, in column order
  auto add_col = [&](unsigned snp, int tt, double tau) {
    std::vector<double> x(n);
    for (long i = 0; i < n; ++i) x[i] = typed_genotype(tt, (int)G[(size_t)snp * n + i]);
    std::vector<double> col(m.cols() + 1);
    for (int c = 0; c < m_e; ++c) col[c] = dot(E + (size_t)c * n, x.data());
    for (size_t t = 0; t < colsx.size(); ++t) col[m_e + t] = dot(colsx[t].data(), x.data());
    col[m.cols()] = dot(x.data(), x.data());
    m.add_term(snp, dot(x.data(), y), col.data(), tau, tt);
    colsx.push_back(x);
  };
  for (int l = 0; l < k; ++l) {
    if (snp_type[l] == 4) { add_col(snps[l], 0, taus2[2 * l]); add_col(snps[l], 1, taus2[2 * l + 1]); }
    else add_col(snps[l], snp_type[l], taus2[2 * l]);
  }
  ChainRng rng(seed, (double)n + 1.0);
  m.sample_beta_sigma2(rng);
  const int cols = m.cols();
  for (int c = 0; c < cols; ++c) out_beta[c] = m.beta[c];
  out3[0] = m.sigma2;
  p->sample_alpha_and_tau2_typed(&m, rng);
  for (int c = 0; c < cols; ++c) out_tau[c] = m.inv_tau2_alpha2[c];
  out3[1] = p->alpha();
  m.compute_log_likelihood();
  out3[2] = m.log_likelihood;
  delete p;
  return cols;
}

// Add / remove trace of a model whose SNPs have effect types, through TypedTerms + Model (AH: two add_term calls; removal
// drops the larger column first).  ops[5 i..] = {0 add | 1 remove, SNP | model index, effect type, tau1, tau2}.
// trace[i] = log-likelihood after op i; the final upper triangle of X'X goes to xx_out (cols x cols); returns cols.
extern "C" int harness_typed_model_trace(long n, long m_g, int m_e, const double* G, const double* E, const double* y, double yy,
                                         double s2_sigma2, int n_types, const int* types, int n_ops, const double* ops,
                                         double* trace, double* xx_out, int* Ns_out)
{
  Prior* p = make_prior(n, m_g, m_e, yy, 5.0, 20.0, 1.0, s2_sigma2, 5.0, 0.05, 1.0, 1, 0.0, 0.001);
  configure(p, n_types, types, 5.0, 0.05);
  auto dot = [&](const double* a, const double* b) { double s = 0.0; for (long i = 0; i < n; ++i) s += a[i] * b[i]; return s; };
  UpperMat exx;
  exx.resize(m_e);
  std::vector<double> exy(m_e);
  for (int c = 0; c < m_e; ++c) {
    for (int r = 0; r <= c; ++r) exx(r, c) = dot(E + (size_t)r * n, E + (size_t)c * n);
    exy[c] = dot(E + (size_t)c * n, y);
  }
  Model m;
  m.init(m_e, exx, exy, p);
  TypedTerms terms;
  std::vector<std::vector<double>> colsx;   // typed dense columns in column order
  for (int i = 0; i < n_ops; ++i) {
    const double* op = ops + 5 * i;
    if (op[0] == 0.0) {
      const unsigned snp = (unsigned)op[1];
      const int ty = (int)op[2];
      terms.add(snp, ty, m.cols());
      for (int c = 0; c < TypedTerms::n_columns(ty); ++c) {
        const int tt = TypedTerms::term_type(ty, c);
        std::vector<double> x(n);
        for (long r = 0; r < n; ++r) x[r] = typed_genotype(tt, (int)G[(size_t)snp * n + r]);
        std::vector<double> col(m.cols() + 1);
        for (int e2 = 0; e2 < m_e; ++e2) col[e2] = dot(E + (size_t)e2 * n, x.data());
        for (size_t t = 0; t < colsx.size(); ++t) col[m_e + t] = dot(colsx[t].data(), x.data());
        col[m.cols()] = dot(x.data(), x.data());
        m.add_term(snp, dot(x.data(), y), col.data(), op[3 + c], tt);
        colsx.push_back(x);
      }
    } else {
      int rem[2];
      const int cnt = terms.remove((int)op[1], rem);
      for (int c = 0; c < cnt; ++c) {
        m.remove_term(rem[c] - m_e);
        colsx.erase(colsx.begin() + (rem[c] - m_e));
      }
    }
    trace[i] = m.log_likelihood;
  }
  const int cols = m.cols();
  for (int c = 0; c < cols; ++c)
    for (int r = 0; r < cols; ++r) xx_out[(size_t)c * cols + r] = r <= c ? m.xx(r, c) : 0.0;
  for (int t = 0; t < 5; ++t) Ns_out[t] = terms.Ns[t];
  delete p;
  return cols;
}

// Delayed-rejection enumeration for SNPs with effect types: SubmodelEnumerator::run_typed against a model built
// from scratch per sub-model (as harness_exhaustive, with snp_type per SNP; AH = two columns).
extern "C" void harness_exhaustive_typed(long n, long m_g, int m_e, const double* G, const double* E, const double* y, double yy,
                                         double s2_sigma2, int n_types, const int* types, int const_loci, int ms,
                                         const unsigned* snps, const int* snp_type, const double* taus2, double* P, double* B)
{
  Prior* p = make_prior(n, m_g, m_e, yy, 5.0, 20.0, 1.0, s2_sigma2, 5.0, 0.05, 1.0, 1, 0.0, 0.001);
  configure(p, n_types, types, 5.0, 0.05);
  auto dot = [&](const double* a, const double* b) { double s = 0.0; for (long i = 0; i < n; ++i) s += a[i] * b[i]; return s; };
  UpperMat exx;
  exx.resize(m_e);
  std::vector<double> exy(m_e);
  for (int c = 0; c < m_e; ++c) {
    for (int r = 0; r <= c; ++r) exx(r, c) = dot(E + (size_t)r * n, E + (size_t)c * n);
    exy[c] = dot(E + (size_t)c * n, y);
  }
  struct Built { Model m; TypedTerms terms; std::vector<std::vector<double>> colsx; };
  auto add = [&](Built& b, int which) {
    const unsigned snp = snps[which];
    const int ty = snp_type[which];
    b.terms.add(snp, ty, b.m.cols());
    for (int c = 0; c < TypedTerms::n_columns(ty); ++c) {
      const int tt = TypedTerms::term_type(ty, c);
      std::vector<double> x(n);
      for (long r = 0; r < n; ++r) x[r] = typed_genotype(tt, (int)G[(size_t)snp * n + r]);
      std::vector<double> col(b.m.cols() + 1);
      for (int e2 = 0; e2 < m_e; ++e2) col[e2] = dot(E + (size_t)e2 * n, x.data());
      for (size_t t = 0; t < b.colsx.size(); ++t) col[m_e + t] = dot(b.colsx[t].data(), x.data());
      col[b.m.cols()] = dot(x.data(), x.data());
      b.m.add_term(snp, dot(x.data(), y), col.data(), taus2[2 * which + c], tt);
      b.colsx.push_back(x);
    }
  };
  Built full;
  full.m.init(m_e, exx, exy, p);
  for (int i = 0; i < const_loci + ms; ++i) add(full, i);
  SubmodelEnumerator exh;
  double mx;
  exh.run_typed(full.m, full.terms, const_loci, ms, P, mx);
  for (unsigned long mask = 0; mask < (1ul << ms); ++mask) {
    Built b;
    b.m.init(m_e, exx, exy, p);
    for (int i = 0; i < const_loci; ++i) add(b, i);
    for (int bit = 0; bit < ms; ++bit)
      if ((mask >> bit) & 1) add(b, const_loci + bit);
    B[mask] = b.m.log_likelihood + p->log_model(b.terms.Ns);
  }
  const double p0 = P[0], b0 = B[0];
  for (unsigned long mask = 0; mask < (1ul << ms); ++mask) { P[mask] -= p0; B[mask] -= b0; }
  delete p;
}

// ---- the memo of column statistics (gramcache.hpp): a pair table driven through growth and restarts ----------------------
// Files n_pairs products v(a, b) = a * 1e6 + b for pseudo-random pairs; after every insertion checks a sample of earlier
// pairs: a pair must be either absent (the table started over since) or hold exactly its value.  Returns the number of
// violations; out3 = {restarts, pairs held at the end, pairs found with the right value in the final sweep}.
#include "gramcache.hpp"
extern "C" long harness_gramcache(long max_slots, long n_pairs, long m_g, double* out3)
{
  GramCache gc;
  gc.init((size_t)m_g, 1);
  if (max_slots > 0) gc.set_max_slots((size_t)max_slots);
  std::vector<std::pair<uint32_t, uint32_t>> filed;
  uint64_t st = 88172645463325252ull;
  auto rnd = [&st]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
  long bad = 0;
  for (long i = 0; i < n_pairs; ++i) {
    const uint32_t a = (uint32_t)(rnd() % (uint64_t)m_g), b = (uint32_t)(rnd() % (uint64_t)m_g);
    gc.put_pair(a, b, (double)std::min(a, b) * 1e6 + (double)std::max(a, b));
    filed.emplace_back(a, b);
    double v = 0.0;
    if (!gc.get_pair(b, a, &v) || v != (double)std::min(a, b) * 1e6 + (double)std::max(a, b)) ++bad;   // symmetric, just filed
    for (int probe = 0; probe < 4; ++probe) {
      const auto& p = filed[(size_t)(rnd() % filed.size())];
      if (gc.get_pair(p.first, p.second, &v) && v != (double)std::min(p.first, p.second) * 1e6 + (double)std::max(p.first, p.second)) ++bad;
    }
  }
  long found = 0;
  for (const auto& p : filed) {
    double v = 0.0;
    if (gc.get_pair(p.first, p.second, &v)) {
      if (v == (double)std::min(p.first, p.second) * 1e6 + (double)std::max(p.first, p.second)) ++found; else ++bad;
    }
  }
  out3[0] = (double)gc.restarts(); out3[1] = (double)gc.pairs(); out3[2] = (double)found;
  // per-SNP entries and the phenotype epoch
  const double xe[1] = {3.0};
  gc.put_snp(5, 1.5, xe, 7.0);
  if (!gc.have_snp(5) || gc.xy(5) != 1.5 || gc.xx(5) != 7.0 || gc.xe(5)[0] != 3.0 || gc.have_snp(6)) ++bad;
  gc.bump_phenotype();
  if (gc.have_snp(5)) ++bad;
  if (gc.repeat_visitor(9) || !gc.repeat_visitor(9)) ++bad;   // kept from the second proposal on
  return bad;
}
